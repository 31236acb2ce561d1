// Second half of a user-target plug-in (included AFTER the user's WN_TARGET_LP_GRAD definition): wraps the
// user's function as a target of the persistent kernels (one thread per chain, the whole vector in registers /
// local memory) and exports the plug-in entry points that capi.cu binds with dlopen.  Kernel launches happen
// INSIDE the plug-in: it carries its own statically linked CUDA runtime, so its kernels are registered there.
#pragma once
#include "wn_dispatch.cuh"

namespace wn {

template <int G, int E2>
struct UserThreadT {
  static constexpr int E = 2 * E2;
  static constexpr bool PAIR_LAYOUT = true;
  static constexpr bool BLOCK_LOCKSTEP = false;
  static constexpr bool LAZY_ENERGY = false;
  static constexpr bool COOP = false;
  static_assert(G == 1 && E >= WN_USER_D, "user targets run one thread per chain");
  __host__ __device__ static constexpr int smem_doubles(int) { return 0; }
  const double* data;
  int nd;
  __device__ __forceinline__ int coord(int e, int) const { return e; }
  __device__ __forceinline__ void init(const TargetParams& tp, int, int, double*) { data = tp.p0; nd = tp.n0; }
  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double*, int&) const {
    double qq[WN_USER_D], gg[WN_USER_D];
#pragma unroll
    for (int e = 0; e < WN_USER_D; ++e) { qq[e] = q[e]; gg[e] = 0.0; }
    const double lp = wn_user_lp_grad(qq, gg, data, nd);
#pragma unroll
    for (int e = 0; e < E; ++e) g[e] = (e < WN_USER_D) ? gg[e < WN_USER_D ? e : 0] : 0.0;
    return lp;
  }
};

constexpr int WN_USER_E2 = (WN_USER_D + 1) / 2;

static LaunchPlan user_plan(int family) {
  // FAM_EXT kernels serve every WALNUTSpy integrator and the warm-up adaptation (one instantiation per plug-in)
  if (family == FAM_PKG) return plan_pkg<UserThreadT, 1, WN_USER_E2, 128>();
  return plan_wpy<UserThreadT, 1, WN_USER_E2, 128, 1, true, true>();
}

}  // namespace wn

extern "C" {

int wn_user_abi(void) { return WN_ABI_VERSION; }
int wn_user_dim(void) { return WN_USER_D; }

// fills the launch geometry for `family` (wn::FAM_PKG or anything else = WALNUTSpy driver)
int wn_user_plan(int family, int* G, int* E2, int* NT, size_t* smem, int* package) {
  const wn::LaunchPlan p = wn::user_plan(family);
  *G = p.G; *E2 = p.E2; *NT = p.NT; *smem = p.smem; *package = p.package ? 1 : 0;
  return 0;
}

// resident blocks per SM (after raising the dynamic shared memory limit); < 0: CUDA error code
int wn_user_occupancy(int family, int device) {
  const wn::LaunchPlan p = wn::user_plan(family);
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess) return -2;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p.fn, p.NT, p.smem) != cudaSuccess) return -3;
  return occ;
}

// launches the family's kernel on `stream`; params = wn::RunParams or wn::PkgParams of the caller
int wn_user_launch(int family, int device, const void* params, unsigned blocks, void* stream) {
  const wn::LaunchPlan p = wn::user_plan(family);
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  void* args[] = {const_cast<void*>(params)};
  const cudaError_t e = cudaLaunchKernel(p.fn, dim3(blocks), dim3(p.NT), args, p.smem, (cudaStream_t)stream);
  return e == cudaSuccess ? 0 : -(int)e - 100;
}

}  // extern "C"
