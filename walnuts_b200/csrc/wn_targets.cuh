// CUDA target registry: log density + gradient evaluated on the register-resident coordinates
// of a chain group.  Replaces reference WALNUTSpy/targetDistr.py and test/targets.py.
//
// Protocol: `init(tp, d, t)` loads per-thread constants; `lp_grad(q, g, red, parity)` fills the
// gradient of this thread's coordinates and returns this thread's PARTIAL of lp (the sum of the
// partials over the group is lp).  A target may reduce over the group internally; control flow
// is uniform within a group.
#pragma once
#include "wn_common.cuh"

namespace wn {

struct TargetParams {
  const double* p0;  // inv_var [d]      | X [N,P] row-major | y [T]
  const double* p1;  // -                | y [N]             | -
  int n0, n1;        // -                | N, P              | T
  double c0;         // -                | 1/tau^2           | -
};

// coordinate index of element e (= 2*e2 + h) held by thread t
template <int G>
__device__ __forceinline__ int coord_of(int e, int t) {
  return 2 * ((e >> 1) * G + t) + (e & 1);
}

// ---- T1/T2: standard normal and diagonal Gaussian -------------------------------------
// lp = -1/2 sum q_i^2 s_i, grad = -q_i s_i (s = 1: reference targetDistr.stdGauss :18-21,
// test/targets.py:4-7).  12 flop per coordinate per leapfrog step including the energy.
template <int G, int E2, bool UNIT>
struct DiagGaussT {
  static constexpr int E = 2 * E2;
  double s[UNIT ? 1 : E];
  __device__ __forceinline__ void init(const TargetParams& tp, int d, int t) {
    if constexpr (!UNIT) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = coord_of<G>(e, t);
        s[e] = (j < d) ? tp.p0[j] : 0.0;
      }
    }
  }
  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double*, int&) const {
    double acc0 = 0.0, acc1 = 0.0;   // two partial sums: halves the dependent-FMA chain
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if constexpr (UNIT) g[e] = -q[e];
      else g[e] = -(q[e] * s[e]);
      if (e & 1) acc1 = fma(q[e], g[e], acc1);
      else acc0 = fma(q[e], g[e], acc0);
    }
    return 0.5 * (acc0 + acc1);
  }
};

// ---- T3: Neal's funnel, reference targetDistr.funnel10 :74-78 (d = 1 + n) ---------------
// lp = logN(q0; 0, 3) + sum_i logN(q_i; 0, exp(q0/2));  G <= 32.
template <int G, int E2>
struct FunnelT {
  static constexpr int E = 2 * E2;
  static_assert(G <= 32, "funnel target needs the chain inside one warp");
  int d_;
  double nn;
  __device__ __forceinline__ void init(const TargetParams&, int d, int) {
    d_ = d;
    nn = (double)(d - 1);
  }
  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double* red, int& parity) const {
    const int t = threadIdx.x & (G - 1);
    double x[1] = {0.0};
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int j = coord_of<G>(e, t);
      if (j >= 1 && j < d_) x[0] = fma(q[e], q[e], x[0]);
    }
    Group<G>::template sum<1>(x, red, parity);
    double q0 = q[0];
    if constexpr (G > 1) q0 = __shfl_sync(Group<G>::mask(), q0, (threadIdx.x & 31u) & ~(unsigned)(G - 1));
    const double ss = x[0];
    const double ex = exp(-q0);
    const double LOG_SQRT_2PI = 0.91893853320467274178;
    const double LOG3 = 1.09861228866810969140;
    const double q03 = q0 / 3.0;
    const double lp = (-(q03 * q03) / 2.0 - LOG_SQRT_2PI - LOG3) +
                      (-0.5 * ex * ss - nn * LOG_SQRT_2PI - nn * 0.5 * q0);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int j = coord_of<G>(e, t);
      g[e] = (j >= 1 && j < d_) ? -q[e] * ex : 0.0;
    }
    if (t == 0) g[0] = -0.5 * nn - q0 / 9.0 + 0.5 * ex * ss;
    return (t == 0) ? lp : 0.0;
  }
};

// ---- test/targets.py:23-29 funnel (package protocol; a different density from funnel10) ---
template <int G, int E2>
struct FunnelPkgT {
  static constexpr int E = 2 * E2;
  static_assert(G <= 32, "funnel target needs the chain inside one warp");
  int d_;
  __device__ __forceinline__ void init(const TargetParams&, int d, int) { d_ = d; }
  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double* red, int& parity) const {
    const int t = threadIdx.x & (G - 1);
    double x[1] = {0.0};
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int j = coord_of<G>(e, t);
      if (j >= 1 && j < d_) x[0] = fma(q[e], q[e], x[0]);
    }
    Group<G>::template sum<1>(x, red, parity);
    double q0 = q[0];
    if constexpr (G > 1) q0 = __shfl_sync(Group<G>::mask(), q0, (threadIdx.x & 31u) & ~(unsigned)(G - 1));
    const double ss = x[0];
    const double eh = exp(0.5 * q0);
    const double lp = -0.5 * q0 * q0 / 9.0 - 0.5 * ss / eh;
    const double m = -1.0 / eh;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int j = coord_of<G>(e, t);
      g[e] = (j >= 1 && j < d_) ? m * q[e] : 0.0;
    }
    if (t == 0) g[0] = -(q0 / 9.0 - 0.25 * ss / eh);
    return (t == 0) ? lp : 0.0;
  }
};

// ---- reference targetDistr.corrGauss :25-31 (2-d, rho = 0.5); G = 1, E2 = 1 ---------------
template <int G, int E2>
struct CorrGaussT {
  static constexpr int E = 2 * E2;
  static_assert(G == 1 && E2 == 1, "corr_gauss is 2-d");
  __device__ __forceinline__ void init(const TargetParams&, int, int) {}
  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double*, int&) const {
    const double rho = 0.5, tmp = 1.0 - rho * rho;
    const double r = q[1] - rho * q[0];
    g[0] = -(q[0] - rho * q[1]) / tmp;
    g[1] = -(q[1] - rho * q[0]) / tmp;
    return -0.5 * q[0] * q[0] - (0.5 / tmp) * (r * r);
  }
};

}  // namespace wn
