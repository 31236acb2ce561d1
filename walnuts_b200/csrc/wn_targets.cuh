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
  static constexpr bool COOP = false;
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
  static constexpr bool COOP = false;
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
  static constexpr bool COOP = false;
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
  static constexpr bool COOP = false;
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
  static constexpr bool COOP = false;
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
    // (temporaries are kept to 8 arrays of B doubles: the kernel is register-bound)
    // ---- phase 1: prefix sums of the innovations; index-0 entries and tS travel as "base" channels ----
    double locZ[B], locX[B];
    double ch[5] = {0.0, 0.0, 0.0, 0.0, 0.0}, ex1[5];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = k0 + i;
      ch[0] += (k >= 1 && k <= T - 2) ? q[i] : 0.0;
      ch[1] += (k >= 1 && k <= T - 1) ? q[B + i] : 0.0;
      locZ[i] = ch[0];      // local inclusive prefix
      locX[i] = ch[1];
    }
    if (t == 0) { ch[2] = q[0]; ch[3] = q[B]; ch[4] = q[3 * B]; }
    scan<5, true>(ch, ex1, red, parity);
    const double Z0 = ch[2], X0 = ch[3], tS = ch[4];
    const double sigma = exp(-0.5 * tS), etS = exp(tS);
    double lp = 0.0;
    double ez[B], w[B], c[B], locC[B];
    double ch2[2] = {0.0, 0.0}, ex2[2];
    const double ez_prev = exp(0.5 * fma(sigma, ex1[0], Z0));   // exp(z_{k0-1}/2)
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = k0 + i;
      ez[i] = exp(0.5 * fma(sigma, ex1[0] + locZ[i], Z0));
      const double xx = fma(sigma, ex1[1] + locX[i], X0);
      w[i] = exp(-xx);
      if (k <= T - 1) lp -= 0.5 * xx;
      if (k >= 1 && k <= T - 2) lp -= 0.5 * q[i] * q[i];
      if (k >= 1 && k <= T - 1) lp -= 0.5 * (q[B + i] * q[B + i] + q[2 * B + i] * q[2 * B + i]);
      const double ezm = (i == 0) ? ez_prev : ez[(i > 0) ? i - 1 : 0];
      c[i] = (k >= 1 && k <= T - 1) ? ezm * q[2 * B + i] : 0.0;   // exp(z_{k-1}/2) * tauinn_{k-1}
      ch2[0] += c[i];
      locC[i] = ch2[0];
    }
    if (t == 0) ch2[1] = q[2 * B];
    // ---- phase 2: tau ----
    scan<2, true>(ch2, ex2, red, parity);
    const double U0 = ch2[1];
    // residuals and their local suffix sums in one reverse sweep
    double sufR[B], sufA[B];
    double ch3[2] = {0.0, 0.0}, ex3[2];
#pragma unroll
    for (int i = B - 1; i >= 0; --i) {
      const int k = k0 + i;
      const bool obs = (k <= T - 1);
      const double e = y[i] - (U0 + (ex2[0] + locC[i]));
      const double eew = e * e * w[i];
      if (obs) lp -= 0.5 * eew;
      ch3[0] += obs ? e * w[i] : 0.0;
      ch3[1] += obs ? (-0.5 + 0.5 * eew) : 0.0;
      sufR[i] = ch3[0];
      sufA[i] = ch3[1];
    }
    // ---- phase 3: R, A suffix sums ----
    scan<2, false>(ch3, ex3, red, parity);
    double sufB[B];
    double ch4[2] = {0.0, 0.0}, ex4[2];
#pragma unroll
    for (int i = B - 1; i >= 0; --i) {
      const int k = k0 + i;
      const double bb = 0.5 * c[i] * (ex3[0] + sufR[i]);     // b_{k-1} of appendix B (c is already masked)
      sufB[i] = ch4[0];                                      // exclusive local suffix
      ch4[0] += bb;
      // d/dtS partial: sum_k zinn_k Bz_{k+1} + sum_k xinn_k A_{k+1}, with the first sum re-ordered as
      // sum_j b_j * (prefix of zinn before j)
      const double pzex = (ex1[0] + locZ[i]) - ((k >= 1 && k <= T - 2) ? q[i] : 0.0);
      ch4[1] += bb * pzex + ((k >= 1 && k <= T - 1) ? q[B + i] * (ex3[1] + sufA[i]) : 0.0);
    }
    // ---- phase 4: Bz suffix sums and the tS total ----
    scan<2, false>(ch4, ex4, red, parity);
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = k0 + i;
      const double Sx = ex4[0] + sufB[i];     // sum_{j > k} bb_j = Bz of appendix B at the next index
      const double Rk = ex3[0] + sufR[i], Ak = ex3[1] + sufA[i];
      const double ezm = (i == 0) ? ez_prev : ez[(i > 0) ? i - 1 : 0];
      g[i] = (k == 0) ? Sx : ((k <= T - 2) ? fma(sigma, Sx, -q[i]) : 0.0);
      g[B + i] = (k == 0) ? Ak : ((k <= T - 1) ? fma(sigma, Ak, -q[B + i]) : 0.0);
      g[2 * B + i] = (k == 0) ? Rk : ((k <= T - 1) ? fma(ezm, Rk, -q[2 * B + i]) : 0.0);
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
  static constexpr bool COOP = false;
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

namespace wn {

// ---- T4, block-cooperative version: the chains of a CTA evaluate their gradients TOGETHER --------------------
// 8 chains per CTA of 256 threads (one warp owns one chain's coordinates, as in LogRegT).  Every trip of the
// sampler's loop the active chains publish beta to shared memory; then ALL 256 threads stream the rows once:
//   phase 1 (4 rows per thread): eta_n^c = x_n . beta^c for the 8 chains  (X^T coalesced over rows; 32 FMA per
//                                4 loads + 4 shared broadcasts), r_n^c = y_n - sigmoid(eta_n^c),
//                                lp^c += y_n eta - log(1+e^eta) only for chains at the end of a pass -> shared r tile
//   phase 2 (lane = 4 coordinates, warp = row stream): acc[k][c] += X[n][k] r_n^c            (32 FMA per row)
// so every element of X is loaded once per CTA trip instead of once per chain.  Deterministic: partial sums are
// combined in a fixed order (no atomics).
template <int G, int E2>
struct LogRegCoopT {
  static constexpr int E = 2 * E2;
  static constexpr bool PAIR_LAYOUT = true;
  static constexpr bool BLOCK_LOCKSTEP = false;   // the cooperative evaluation has its own barriers
  static constexpr bool LAZY_ENERGY = false;
  static constexpr bool COOP = true;
  static constexpr int NTC = 256, C = 8, PMAX = 128, RPT = 4, RT = NTC * RPT;
  static_assert(G == 32 && E2 == 2, "cooperative logistic regression: one warp per chain, P <= 128");
  // Staging ring: S stages of STG doubles, filled by bulk async copies (TMA engine) that complete on a
  // "full" mbarrier per stage; consumers release a stage by arriving on its "empty" mbarrier, and the
  // producer (thread 0) refills it S-1 chunks ahead -- across the phase-1 / phase-2 / row-tile boundaries.
  // Chunk stream of one evaluation: per row tile, P chunks "coordinate k of X^T for the tile's rows"
  // followed by ceil(rows/8) chunks "8 rows of X".
  // Measured on B200: with only 32 FMA per thread between two barrier operations the
  // per-chunk mbarrier wait/arrive costs more than the L2 latency it hides (0.99e5 vs 1.37e5 evals/s with plain
  // read-only loads and 16 loads in flight per thread), so the staging ring is compiled out by default.
  static constexpr bool USE_BULK = false;
  static constexpr int S = 8, STG = 1024, RC = 8;
  // shared: bs[PMAX][C] | rs[RT][C] | gs[C][PMAX] | part[8][PMAX][4] | lps[8 warps][C] | act[C] | need[C]
  //         | full[S] | empty[S] | stage[S][STG]
  __host__ __device__ static constexpr int smem_doubles(int) {
    return PMAX * C + RT * C + C * PMAX + 8 * PMAX * 4 + 8 * C + 2 * C + 2 * S + (USE_BULK ? S * STG : 0);
  }
  const double *X, *XT, *y;
  int N, P;
  double itau2;
  double *bs, *rs, *gs, *part, *lps, *act, *need, *stage;
  uint64_t *full, *empty;
  bool bulk;      // 16-byte alignment of every staged chunk holds (N and P even)
  __device__ __forceinline__ int coord(int e, int t) const { return coord_of<G>(e, t); }
  __device__ __forceinline__ void init(const TargetParams& tp, int d, int, double* tsm) {
    X = tp.p0; y = tp.p1; XT = tp.p2; N = tp.n0; P = d; itau2 = tp.c0;
    bs = tsm; rs = bs + PMAX * C; gs = rs + RT * C; part = gs + C * PMAX; lps = part + 8 * PMAX * 4; act = lps + 8 * C;
    need = act + C;
    full = reinterpret_cast<uint64_t*>(need + C);
    empty = full + S;
    stage = need + C + 2 * S;
    bulk = USE_BULK && ((N & 1) == 0) && ((P & 1) == 0);
    if (USE_BULK && threadIdx.x == 0) {
      for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NTC); }
      mbar_fence_init();
    }
    __syncthreads();
  }
  // not used by COOP kernels (the generic protocol entry point)
  __device__ __forceinline__ double lp_grad(const double (&)[E], double (&)[E], double*, int&) const { return 0.0; }

  // need_lp: the energy of this step is consumed (last step of a pass).  For this target a non-finite
  // intermediate energy implies a non-finite state, which persists to the end of the pass, so skipping the
  // log-likelihood of intermediate steps does not change all(isfinite(Hams)) (adaptiveIntegrators.py:92).
  __device__ __forceinline__ void publish(const double (&q)[E], bool active, bool need_lp) const {
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < E; ++e) bs[coord_of<G>(e, t) * C + w] = active ? q[e] : 0.0;
    if (t == 0) { act[w] = active ? 1.0 : 0.0; need[w] = (active && need_lp) ? 1.0 : 0.0; }
  }

  // `gch`: number of chunks consumed so far by this CTA (kept by the caller across evaluations); chunk g
  // lives in stage g % S and completes phase (g / S) & 1 of that stage's barriers.
  __device__ __forceinline__ void coop_eval(uint32_t& gch) const {
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    bool any = false;
    bool nl[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { any = any || (act[c] != 0.0); nl[c] = need[c] != 0.0; }
    if (!any) return;
    double acc[4][C];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[j][c] = 0.0;
    double lp[C];
#pragma unroll
    for (int c = 0; c < C; ++c) lp[c] = 0.0;
    const int k4 = 4 * lane;

    // ---- producer cursor (thread 0 only): next chunk to issue ----
    uint32_t pg = gch;            // global index of the next chunk to issue
    int pn0 = 0, pi = 0;          // its row tile and index inside the tile's chunk list
    auto produce = [&]() {        // issue one chunk if any is left in this evaluation
      if (pn0 >= N) return;
      const int nrows = min(RT, N - pn0);
      const int st_ = pg % S;
      if (pg >= (uint32_t)S) mbar_wait(&empty[st_], ((pg / S) & 1u) ^ 1u);   // previous use released by all
      double* dst = stage + st_ * STG;
      if (pi < P) {
        mbar_expect_tx(&full[st_], (uint32_t)(nrows * 8));
        bulk_g2s(dst, XT + (size_t)pi * N + pn0, (uint32_t)(nrows * 8), &full[st_]);
      } else {
        const int r0 = (pi - P) * RC, rn = min(RC, nrows - r0);
        mbar_expect_tx(&full[st_], (uint32_t)(rn * P * 8));
        bulk_g2s(dst, X + (size_t)(pn0 + r0) * P, (uint32_t)(rn * P * 8), &full[st_]);
      }
      ++pg;
      ++pi;
      if (pi >= P + (nrows + RC - 1) / RC) { pi = 0; pn0 += RT; }
    };
    if (bulk && tid == 0) {
      for (int i = 0; i < S - 1; ++i) produce();
    }
    auto consume_begin = [&]() -> const double* {
      if (tid == 0) produce();                       // keep S-1 chunks in flight
      const int st_ = gch % S;
      mbar_wait(&full[st_], (gch / S) & 1u);
      return stage + st_ * STG;
    };
    auto consume_end = [&]() {
      mbar_arrive(&empty[gch % S]);
      ++gch;
    };

    for (int n0 = 0; n0 < N; n0 += RT) {
      const int nrows = min(RT, N - n0);
      // ---- phase 1: rows n0 + tid + 256 j, j < 4 ----
      {
        double eta[RPT][C];
#pragma unroll
        for (int j = 0; j < RPT; ++j)
#pragma unroll
          for (int c = 0; c < C; ++c) eta[j][c] = 0.0;
        bool in[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) in[j] = tid + NTC * j < nrows;
#pragma unroll 2
        for (int k = 0; k < P; ++k) {
          double x[RPT];
          if (bulk) {
            const double* sb = consume_begin();
#pragma unroll
            for (int j = 0; j < RPT; ++j) x[j] = in[j] ? sb[tid + NTC * j] : 0.0;
            consume_end();
          } else {
#pragma unroll
            for (int j = 0; j < RPT; ++j) x[j] = in[j] ? __ldg(XT + (size_t)k * N + n0 + tid + NTC * j) : 0.0;
          }
          const double2* b2 = reinterpret_cast<const double2*>(bs + k * C);
#pragma unroll
          for (int c2 = 0; c2 < C / 2; ++c2) {
            const double2 bb = b2[c2];
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              eta[j][2 * c2] = fma(x[j], bb.x, eta[j][2 * c2]);
              eta[j][2 * c2 + 1] = fma(x[j], bb.y, eta[j][2 * c2 + 1]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int n = n0 + tid + NTC * j;
          const double yy = in[j] ? __ldg(y + n) : 0.0;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const double et = eta[j][c];
            const double ex = exp(-fabs(et));
            const double inv = 1.0 / (1.0 + ex);
            const double sig = (et >= 0.0) ? inv : ex * inv;
            eta[j][c] = in[j] ? (yy - sig) : 0.0;                                  // residual
            if (nl[c] && in[j]) lp[c] += yy * et - (fmax(et, 0.0) + log1p(ex));
          }
          double2* r2 = reinterpret_cast<double2*>(rs + (tid + NTC * j) * C);
#pragma unroll
          for (int c2 = 0; c2 < C / 2; ++c2) r2[c2] = make_double2(eta[j][2 * c2], eta[j][2 * c2 + 1]);
        }
      }
      __syncthreads();
      // ---- phase 2: chunks of 8 rows; warp w takes row w of the chunk; lane owns coordinates 4*lane .. +3 ----
      {
        const int nchunk = (nrows + RC - 1) / RC;
#pragma unroll 4
        for (int ch = 0; ch < nchunk; ++ch) {
          const int i = ch * RC + w;                 // row inside the tile
          double x[4];
          if (bulk) {
            const double* sb = consume_begin();
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = (i < nrows && k4 + j < P) ? sb[w * P + k4 + j] : 0.0;
            consume_end();
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] = (i < nrows && k4 + j < P) ? __ldg(X + (size_t)(n0 + i) * P + k4 + j) : 0.0;
          }
          if (i < nrows) {
            const double2* r2 = reinterpret_cast<const double2*>(rs + i * C);
#pragma unroll
            for (int c2 = 0; c2 < C / 2; ++c2) {
              const double2 r = r2[c2];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc[j][2 * c2] = fma(x[j], r.x, acc[j][2 * c2]);
                acc[j][2 * c2 + 1] = fma(x[j], r.y, acc[j][2 * c2 + 1]);
              }
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- combine the 8 row streams (fixed order), 4 chains per round ----
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) part[(w * PMAX + k4 + j) * 4 + c] = acc[j][4 * half + c];
      __syncthreads();
      {
        const int k = tid & (PMAX - 1), cp = tid >> 7;          // 2 chains per thread
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * cp + cc;
          double s = 0.0;
#pragma unroll
          for (int ww = 0; ww < 8; ++ww) s += part[(ww * PMAX + k) * 4 + c];
          gs[(4 * half + c) * PMAX + k] = s;
        }
      }
      __syncthreads();
    }
    // ---- log-likelihood: lanes -> warp (butterfly), warps -> block (fixed order in collect) ----
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double v = lp[c];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) lps[w * C + c] = v;
    }
    __syncthreads();
  }

  // gradient of this warp's chain (after coop_eval); returns this thread's partial of lp
  __device__ __forceinline__ double collect(const double (&q)[E], double (&g)[E]) const {
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    double qq = 0.0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int k = coord_of<G>(e, t);
      g[e] = (k < P) ? fma(-itau2, q[e], gs[w * PMAX + k]) : 0.0;
      qq = fma(q[e], q[e], qq);
    }
    double lp = -0.5 * itau2 * qq;
    if (t == 0) {
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) lp += lps[ww * C + w];
    }
    return lp;
  }
};

}  // namespace wn
