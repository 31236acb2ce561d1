// User-defined CUDA targets: the role of the reference's arbitrary Python
// `lpFun(q) -> [lp, grad]` (WALNUTSpy/targetDistr.py:18) / `logp`, `grad` (walnuts/walnuts.py:296-297) and of
// walnuts_stan.py's compiled Stan model.  The user writes ONE device function over the whole coordinate vector
//
//     WN_TARGET_LP_GRAD(q, g, data, n_data) {
//       // q[0..WN_D-1] in, g[0..WN_D-1] out, data[0..n_data-1] = the array given to cuda_target(data=...)
//       return lp;      // the log density
//     }
//
// which walnuts_b200.targets.cuda_target() compiles with nvcc for sm_100a into a plug-in library next to the
// sampler kernels (WN_D <= 64: one thread per chain; 64 < WN_D <= 512: one warp per chain, the function is then called by
// every lane with q and g in shared memory).  This header is included BEFORE the user's source.
#pragma once
#include <cmath>
#include <cstdint>

#ifndef WN_USER_D
#error "WN_USER_D (the dimension) must be defined by the generated translation unit"
#endif
#define WN_D WN_USER_D

#define WN_TARGET_LP_GRAD(q, g, data, n_data)                                                              \
  __device__ __forceinline__ double wn_user_lp_grad(const double* __restrict__ q, double* __restrict__ g,   \
                                                    const double* __restrict__ data, int n_data)
