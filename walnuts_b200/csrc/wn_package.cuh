// placeholder until the package-mode kernel lands (next commit)
#pragma once
#include "wn_common.cuh"
#include "wn_targets.cuh"
namespace wn {
struct PkgParams {
  int n_chains, d, dg, max_depth, compat, n_iter;
  uint32_t iter0, seed_lo, seed_hi, chain_offset;
  double macro_step, max_error;
  const double* inv_mass;
  double* state; double* draws;
  unsigned long long* neval; unsigned long long* totals;
  double2* scratch; int nslot; unsigned int* queue; TargetParams tp;
};
__host__ __device__ inline int package_scratch_vectors(int M) { return 8 + 2 * (M + 1); }
template <template <int, int> class TargetTT, int G, int E2, int NT>
__global__ void __launch_bounds__(NT) package_kernel(const __grid_constant__ PkgParams P) {}
}
