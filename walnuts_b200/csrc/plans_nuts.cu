// Kernel family "nuts" (fixedLeapFrog only): see wn_dispatch.cuh.
#include "wn_dispatch.cuh"

bool wn_pick_plan_nuts(const wn_config& c, wn::LaunchPlan& p) { return wn::pick_plan_family<wn::FAM_NUTS>(c, p); }
