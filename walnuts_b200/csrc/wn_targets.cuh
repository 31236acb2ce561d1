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
  const double* p2;  // -                | X^T [P,N]         | -
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
  static constexpr bool PAIR_LAYOUT = true;
  static constexpr bool BLOCK_LOCKSTEP = false;
  // Energies are only consumed at the end of a pass (adaptiveIntegrators.py:87) plus the
  // all(isfinite(Hams)) test (:92).  While every |q_i|, |v_i| stays below 2^480 (and inv_var <= 2^60) each
  // intermediate energy is provably finite, so the intermediate steps may skip the two energy FMAs per
  // coordinate; if the bound is ever violated the sampler re-runs the pass with per-step energies.
  static constexpr bool LAZY_ENERGY = true;
  __host__ __device__ static constexpr int smem_doubles(int) { return 0; }
  __device__ __forceinline__ int coord(int e, int t) const { return coord_of<G>(e, t); }
  double s[UNIT ? 1 : E];
  double smax;     // max inv_var over ALL coordinates (host-computed, TargetParams::c0)
  bool lazy_ok;
  __device__ __forceinline__ void init(const TargetParams& tp, int d, int t, double*) {
    lazy_ok = true;
    smax = UNIT ? 1.0 : tp.c0;
    lazy_ok = (smax <= 0x1p60);
    if constexpr (!UNIT) {
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int j = coord_of<G>(e, t);
        s[e] = (j < d) ? tp.p0[j] : 0.0;
        lazy_ok = lazy_ok && (fabs(s[e]) <= 0x1p60);
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
  // upper bound of max(|q'|,|v'|) / max(|q|,|v|) over one leapfrog step of size hh on this target:
  // v1 = v - a s q, q' = q + hh v1, v' = v1 - a s q'  with a = hh/2, s <= smax
  __device__ __forceinline__ double step_growth(double hh) const {
    const double A = 0.5 * hh * smax;
    return fmax(1.0 + hh + hh * A, 1.0 + 2.0 * A + hh * A + hh * A * A);
  }
  __device__ __forceinline__ void grad_only(const double (&q)[E], double (&g)[E]) const {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if constexpr (UNIT) g[e] = -q[e];
      else g[e] = -(q[e] * s[e]);
    }
  }
};

// ---- T3: Neal's funnel, reference targetDistr.funnel10 :74-78 (d = 1 + n) ---------------
// lp = logN(q0; 0, 3) + sum_i logN(q_i; 0, exp(q0/2));  G <= 32.
template <int G, int E2>
struct FunnelT {
  static constexpr int E = 2 * E2;
  static constexpr bool PAIR_LAYOUT = true;
  static constexpr bool BLOCK_LOCKSTEP = false;
  static constexpr bool LAZY_ENERGY = false;
  __host__ __device__ static constexpr int smem_doubles(int) { return 0; }
  __device__ __forceinline__ int coord(int e, int t) const { return coord_of<G>(e, t); }
  static_assert(G <= 32, "funnel target needs the chain inside one warp");
  int d_;
  double nn;
  __device__ __forceinline__ void init(const TargetParams&, int d, int, double*) {
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
  static constexpr bool PAIR_LAYOUT = true;
  static constexpr bool BLOCK_LOCKSTEP = false;
  static constexpr bool LAZY_ENERGY = false;
  __host__ __device__ static constexpr int smem_doubles(int) { return 0; }
  __device__ __forceinline__ int coord(int e, int t) const { return coord_of<G>(e, t); }
  static_assert(G <= 32, "funnel target needs the chain inside one warp");
  int d_;
  __device__ __forceinline__ void init(const TargetParams&, int d, int, double*) { d_ = d; }
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
  static constexpr bool PAIR_LAYOUT = true;
  static constexpr bool BLOCK_LOCKSTEP = false;
  static constexpr bool LAZY_ENERGY = false;
  __host__ __device__ static constexpr int smem_doubles(int) { return 0; }
  __device__ __forceinline__ int coord(int e, int t) const { return coord_of<G>(e, t); }
  static_assert(G == 1 && E2 == 1, "corr_gauss is 2-d");
  __device__ __forceinline__ void init(const TargetParams&, int, int, double*) {}
  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double*, int&) const {
    const double rho = 0.5, tmp = 1.0 - rho * rho;
    const double r = q[1] - rho * q[0];
    g[0] = -(q[0] - rho * q[1]) / tmp;
    g[1] = -(q[1] - rho * q[0]) / tmp;
    return -0.5 * q[0] * q[0] - (0.5 / tmp) * (r * r);
  }
};

}  // namespace wn

namespace wn {

// ---- T5: Stock-Watson stochastic volatility ------------------------------------------------------
// Unconstrained parameterisation of reference WALNUTSpy_examples/StockWatson/sw_innov.stan:7-52
// (bridgestan, propto = True); formulas: SURVEY.md appendix B, restated in oracle/targets.py
// (make_stock_watson).  theta = [tS, z1, zinn[T-2], x1, xinn[T-1], tau1, tauinn[T-1]], d = 3T.
//
// Layout: time-aligned.  Thread t owns the B consecutive time indices k = t*B .. t*B+B-1 of the three
// series Z = [z1, zinn...], X = [x1, xinn...], U = [tau1, tauinn...] (elements 0..3B-1) and thread 0
// additionally owns tS (element 3B).  One evaluation = 2 forward and 2 reverse block scans over the
// group (warp shuffles + one shared-memory exchange per scan when the chain spans several warps).
template <int G, int E2>
struct StockWatsonT {
  static constexpr int E = 2 * E2;
  static constexpr int B = (E - 2) / 3;
  static constexpr bool PAIR_LAYOUT = false;
  static constexpr bool BLOCK_LOCKSTEP = false;
  static constexpr bool LAZY_ENERGY = false;
  __host__ __device__ static constexpr int smem_doubles(int) { return 0; }
  static constexpr int WARPS = (G + 31) / 32;
  static_assert(3 * B + 2 == E, "E must be 3B + 2");
  static_assert(G % 32 == 0, "Stock-Watson target needs whole warps per chain");
  int T;
  double y[B];

  __device__ __forceinline__ void init(const TargetParams& tp, int, int t, double*) {
    T = tp.n0;
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = t * B + i;
      y[i] = (k < T) ? tp.p0[k] : 0.0;
    }
  }
  __device__ __forceinline__ int coord(int e, int t) const {
    const int BIG = 1 << 30;
    if (e < B) { const int k = t * B + e; return (k <= T - 2) ? k + 1 : BIG; }                 // z1, zinn
    if (e < 2 * B) { const int k = t * B + (e - B); return (k <= T - 1) ? T + k : BIG; }       // x1, xinn
    if (e < 3 * B) { const int k = t * B + (e - 2 * B); return (k <= T - 1) ? 2 * T + k : BIG; } // tau1, tauinn
    if (e == 3 * B) return (t == 0) ? 0 : BIG;                                                 // tS
    return BIG;
  }

  // inclusive scan over the threads of the group of NV channels; FWD: prefix, else suffix.
  // On return x = inclusive value, ex = exclusive value (sum over strictly preceding / following threads).
  template <int NV, bool FWD>
  __device__ __forceinline__ static void scan(double (&x)[NV], double (&ex)[NV], double* red, int& parity) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const double yv = FWD ? __shfl_up_sync(0xffffffffu, x[k], off) : __shfl_down_sync(0xffffffffu, x[k], off);
        const bool take = FWD ? (lane >= off) : (lane + off < 32);
        x[k] += take ? yv : 0.0;
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const double yv = FWD ? __shfl_up_sync(0xffffffffu, x[k], 1) : __shfl_down_sync(0xffffffffu, x[k], 1);
      ex[k] = (FWD ? (lane >= 1) : (lane < 31)) ? yv : 0.0;
    }
    if constexpr (G > 32) {
      const int w = (threadIdx.x % G) >> 5;
      double* buf = red + parity * (WARPS * 8);
      if (lane == (FWD ? 31 : 0)) {
#pragma unroll
        for (int k = 0; k < NV; ++k) buf[w * 8 + k] = x[k];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        double off = 0.0;
        if (FWD) { for (int ww = 0; ww < w; ++ww) off += buf[ww * 8 + k]; }
        else { for (int ww = WARPS - 1; ww > w; --ww) off += buf[ww * 8 + k]; }
        x[k] += off;
        ex[k] += off;
      }
      parity ^= 1;
    }
  }

  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double* red, int& parity) const {
    const int t = threadIdx.x % G;
    const int k0 = t * B;
    // ---- phase 1: prefix sums of the innovations; index-0 entries and tS travel as "base" channels ----
    double locZ[B], locX[B];
    double ch[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, ex1[5];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = k0 + i;
      const double zi = (k >= 1 && k <= T - 2) ? q[i] : 0.0;
      const double xi = (k >= 1 && k <= T - 1) ? q[B + i] : 0.0;
      ch[0] += zi;
      ch[1] += xi;
      locZ[i] = ch[0];      // local inclusive prefix
      locX[i] = ch[1];
    }
    if (t == 0) { ch[2] = q[0]; ch[3] = q[B]; ch[4] = q[3 * B]; }
    scan<5, true>(ch, ex1, red, parity);
    const double Z0 = ch[2], X0 = ch[3], tS = ch[4];
    const double sigma = exp(-0.5 * tS), etS = exp(tS);
    double z[B], ez[B], xx[B], w[B], c[B];
    double ch2[2] = {0.0, 0.0}, ex2[2];
    const double ez_prev = exp(0.5 * fma(sigma, ex1[0], Z0));   // exp(z_{k0-1}/2)
    double locC[B];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = k0 + i;
      z[i] = fma(sigma, ex1[0] + locZ[i], Z0);
      xx[i] = fma(sigma, ex1[1] + locX[i], X0);
      ez[i] = exp(0.5 * z[i]);
      w[i] = exp(-xx[i]);
      const double ezm = (i == 0) ? ez_prev : ez[(i > 0) ? i - 1 : 0];
      c[i] = (k >= 1 && k <= T - 1) ? ezm * q[2 * B + i] : 0.0;   // exp(z_{k-1}/2) * tauinn_{k-1}
      ch2[0] += c[i];
      locC[i] = ch2[0];
    }
    if (t == 0) ch2[1] = q[2 * B];
    // ---- phase 2: tau ----
    scan<2, true>(ch2, ex2, red, parity);
    const double U0 = ch2[1];
    double lp = 0.0;
    double r[B], a[B];
    double ch3[2] = {0.0, 0.0}, ex3[2];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = k0 + i;
      const bool obs = (k <= T - 1);
      const double tau = U0 + (ex2[0] + locC[i]);
      const double e = y[i] - tau;
      const double eew = e * e * w[i];
      r[i] = obs ? e * w[i] : 0.0;
      a[i] = obs ? (-0.5 + 0.5 * eew) : 0.0;
      if (obs) lp += -0.5 * xx[i] - 0.5 * eew;
      if (k >= 1 && k <= T - 2) lp -= 0.5 * q[i] * q[i];
      if (k >= 1 && k <= T - 1) lp -= 0.5 * (q[B + i] * q[B + i] + q[2 * B + i] * q[2 * B + i]);
    }
    // local suffix sums
    double sufR[B], sufA[B];
#pragma unroll
    for (int i = B - 1; i >= 0; --i) {
      ch3[0] += r[i];
      ch3[1] += a[i];
      sufR[i] = ch3[0];
      sufA[i] = ch3[1];
    }
    // ---- phase 3: R, A suffix sums ----
    scan<2, false>(ch3, ex3, red, parity);
    double bb[B], Rk[B], Ak[B];
    double ch4[2] = {0.0, 0.0}, ex4[2];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      Rk[i] = ex3[0] + sufR[i];
      Ak[i] = ex3[1] + sufA[i];
      bb[i] = 0.5 * c[i] * Rk[i];            // b_{k-1} of appendix B (c is already masked)
    }
    double sufB[B];
#pragma unroll
    for (int i = B - 1; i >= 0; --i) {
      const int k = k0 + i;
      sufB[i] = ch4[0];                       // exclusive local suffix
      ch4[0] += bb[i];
      // d/dtS partial: sum_k zinn_k Bz_{k+1} + sum_k xinn_k A_{k+1}, with the first sum re-ordered as
      // sum_j b_j * (prefix of zinn before j)
      const double pzex = (ex1[0] + locZ[i]) - ((k >= 1 && k <= T - 2) ? q[i] : 0.0);
      ch4[1] += bb[i] * pzex + ((k >= 1 && k <= T - 1) ? q[B + i] * Ak[i] : 0.0);
    }
    // ---- phase 4: Bz suffix sums and the tS total ----
    scan<2, false>(ch4, ex4, red, parity);
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = k0 + i;
      const double Sx = ex4[0] + sufB[i];     // sum_{j > k} bb_j = Bz of appendix B at the next index
      const double ezm = (i == 0) ? ez_prev : ez[(i > 0) ? i - 1 : 0];
      g[i] = (k == 0) ? Sx : ((k <= T - 2) ? fma(sigma, Sx, -q[i]) : 0.0);
      g[B + i] = (k == 0) ? Ak[i] : ((k <= T - 1) ? fma(sigma, Ak[i], -q[B + i]) : 0.0);
      g[2 * B + i] = (k == 0) ? Rk[i] : ((k <= T - 1) ? fma(ezm, Rk[i], -q[2 * B + i]) : 0.0);
    }
    g[3 * B] = (t == 0) ? (5.0 - 0.5 * etS - 0.5 * sigma * ch4[1]) : 0.0;
    g[3 * B + 1] = 0.0;
    if (t == 0) lp += 5.0 * tS - 0.5 * etS;
    return lp;
  }
};

}  // namespace wn

namespace wn {

// ---- T4: Bayesian logistic regression (not in the reference) -----------------------
// lp = sum_n [y_n eta_n - log(1 + exp(eta_n))] - |beta|^2 / (2 tau^2),  eta = X beta,
// grad = X^T (y - sigmoid(eta)) - beta / tau^2;  X [N,P] row-major plus its transpose X^T [P,N].
// One warp per chain (G = 32, P <= 64*E2).  Rows are processed in tiles of RT: phase 1 computes
// eta / residuals with one lane per row (X^T: coalesced over rows, beta broadcast from shared memory),
// phase 2 accumulates X^T r with one lane per coordinate pair (X: coalesced over coordinates, residuals
// broadcast).  The warps of a block run their evaluations in lock-step (one barrier per trip of the
// sampler's loop), so that the X tiles one warp pulls into L1 are reused by the block's other chains.
template <int G, int E2>
struct LogRegT {
  static constexpr int E = 2 * E2;
  static constexpr bool PAIR_LAYOUT = true;
  static constexpr bool BLOCK_LOCKSTEP = true;
  static constexpr bool LAZY_ENERGY = false;
  static constexpr int RT = 128, RJ = RT / 32, PMAX = 2 * G * E2;
  static_assert(G == 32, "logistic regression target: one warp per chain");
  __host__ __device__ static constexpr int smem_doubles(int NT) { return (NT / 32) * (PMAX + RT); }
  const double *X, *XT, *y;
  int N, P;
  double itau2;
  double *bs, *rs;
  __device__ __forceinline__ int coord(int e, int t) const { return coord_of<G>(e, t); }
  __device__ __forceinline__ void init(const TargetParams& tp, int d, int, double* tsm) {
    X = tp.p0; y = tp.p1; XT = tp.p2; N = tp.n0; P = d; itau2 = tp.c0;
    bs = tsm + (threadIdx.x >> 5) * (PMAX + RT);
    rs = bs + PMAX;
  }
  __device__ __forceinline__ double lp_grad(const double (&q)[E], double (&g)[E], double*, int&) const {
    const int t = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; ++e) bs[coord_of<G>(e, t)] = q[e];
    __syncwarp();
    double lp = 0.0;
    double acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.0;
    for (int n0 = 0; n0 < N; n0 += RT) {
      // phase 1: eta for rows n0 + t + 32 j
      double eta[RJ];
#pragma unroll
      for (int j = 0; j < RJ; ++j) eta[j] = 0.0;
      if (n0 + RT <= N) {
#pragma unroll 4
        for (int k = 0; k < P; ++k) {
          const double b = bs[k];
          const double* col = XT + (size_t)k * N + n0 + t;
#pragma unroll
          for (int j = 0; j < RJ; ++j) eta[j] = fma(__ldg(col + 32 * j), b, eta[j]);
        }
      } else {
        for (int k = 0; k < P; ++k) {
          const double b = bs[k];
          const double* col = XT + (size_t)k * N + n0 + t;
#pragma unroll
          for (int j = 0; j < RJ; ++j)
            if (n0 + t + 32 * j < N) eta[j] = fma(__ldg(col + 32 * j), b, eta[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < RJ; ++j) {
        const int n = n0 + t + 32 * j;
        double r = 0.0;
        if (n < N) {
          const double yy = __ldg(y + n), et = eta[j];
          const double ex = exp(-fabs(et));
          const double inv = 1.0 / (1.0 + ex);
          const double sig = (et >= 0.0) ? inv : ex * inv;
          r = yy - sig;
          lp += yy * et - (fmax(et, 0.0) + log1p(ex));
        }
        rs[t + 32 * j] = r;
      }
      __syncwarp();
      // phase 2: acc_k += X[n][k] * r_n for this lane's coordinates
      const int nend = min(RT, N - n0);
#pragma unroll 4
      for (int i = 0; i < nend; ++i) {
        const double r = rs[i];
        const double* row = X + (size_t)(n0 + i) * P;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int k = coord_of<G>(e, t);
          if (k < P) acc[e] = fma(__ldg(row + k), r, acc[e]);
        }
      }
      __syncwarp();
    }
    double qq = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int k = coord_of<G>(e, t);
      g[e] = (k < P) ? fma(-itau2, q[e], acc[e]) : 0.0;
      qq = fma(q[e], q[e], qq);
    }
    return lp - 0.5 * itau2 * qq;
  }
};

}  // namespace wn
