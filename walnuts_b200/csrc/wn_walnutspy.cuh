// Persistent many-chain kernel for the WALNUTSpy transition (reference WALNUTSpy/WALNUTS.py:189-695
// driver + WALNUTSpy/adaptiveIntegrators.py:49-137,361-475 macro steps).
//
// One group of G threads owns one chain at a time and keeps (q, v, g) of the active orbit end in
// registers for the whole transition.  The kernel is ONE flat loop whose body is a single leapfrog
// micro-step; everything else (step-size search bookkeeping, tree logic, state selection, momentum
// refresh, output) is a small state machine entered only when a pass of 2^c micro-steps ends.  So
// chains that sit at different tree depths / different c never serialise each other's hot loop --
// the SIMT analogue of "lock-free" chains.  Groups pull chains from a global queue.
//
// Per-chain memory outside registers:
//   shared : checkpoint of the macro-step start state S, later the accepted out state O (3 vectors)
//   global : slot-indexed scratch (other orbit end, two proposal slots, left-end stack of <= M
//            pending dyadic levels) -- indexed by resident slot, not by chain, so it stays L2-sized.
#pragma once
#include "wn_common.cuh"
#include "wn_targets.cuh"

namespace wn {

struct RunParams {
  int n_chains, d, dg, M, kind, minC, maxC;
  int n_iter;
  uint32_t iter0;       // iteration number of the first transition of this call (1-based)
  uint32_t seed_lo, seed_hi;
  uint32_t chain_offset;
  double H0, delta0, jitter, p0, log_p0, log_1mp0;
  const double* Hstep;  // [n_chains] or null
  const double* delta;  // [n_chains] or null
  double* state;        // [n_chains, d]
  double* draws;        // [n_iter, n_chains, dg] or null
  double* diag;         // [n_iter, n_chains, 24] or null
  unsigned long long* nevalF;  // [n_chains] or null
  unsigned long long* nevalB;
  unsigned long long* totals;  // [2] grid totals (forward, backward)
  double2* scratch;
  int nslot;
  unsigned int* queue;
  TargetParams tp;
};

enum { KIND_FIXED = 0, KIND_D = 1, KIND_R2P = 2 };
enum { PH_FWD = 0, PH_REDO = 1, PH_BWD = 2 };
enum { ST_CHAIN = 0, ST_ITER, ST_LEVEL, ST_MACRO, ST_PASS_END, ST_LEAF, ST_LEVEL_END, ST_ITER_END, ST_RUN, ST_EXIT };

// scratch vector ids
enum { V_PARK_Q = 0, V_PARK_V = 1, V_PARK_G = 2, V_PROP0 = 3, V_PROP1 = 4, V_STACK = 5 };  // stack: 5 + 2*lvl (+1 for v)
__host__ __device__ inline int scratch_vectors(int M) { return V_STACK + 2 * (M + 1); }

#define WN_LOG_ZERO (-700.0)
#define WN_WT_SUM_THRESH 0x1.78694fe9f73ccp-1009 /* numpy exp(-699) = 2.680137958338607e-304, reference constants.py:14 */

template <template <int, int> class TargetTT, int G, int E2, int NT>
__global__ void __launch_bounds__(NT) walnutspy_kernel(const __grid_constant__ RunParams P) {
  constexpr int E = 2 * E2;
  constexpr int GPB = NT / G;  // groups per block
  static_assert(NT % G == 0 && (G <= 32 || NT == G), "block must hold whole groups");
  using Grp = Group<G>;
  using Target = TargetTT<G, E2>;

  extern __shared__ double smem[];
  double* ck = smem;                       // checkpoint: [3*E][NT]
  double* red = smem + 3 * E * NT;         // reduction scratch (G > 32)
  __shared__ uint32_t sh_bcast;

  const int tid = threadIdx.x;
  const int t = tid % G;
  const int slot = blockIdx.x * GPB + tid / G;
  int parity = 0;

  // scratch addressing: vector vi, pair e2 -> scratch[((vi*E2 + e2) * nslot + slot) * G + t]
  const size_t sc_stride = (size_t)P.nslot * G;
  const size_t sc_off = (size_t)slot * G + t;
  auto sc = [&](int vi, int e2) -> double2* { return P.scratch + ((size_t)(vi * E2 + e2)) * sc_stride + sc_off; };

  Target target;
  target.init(P.tp, P.d, t);

  // ---- register-resident state of the active end / working state -----------------------
  double q[E], v[E], g[E];
  // ---- per-chain control scalars (uniform across the group) ----------------------------
  RngKey key;
  key.k0 = P.seed_lo;
  key.k1 = P.seed_hi;
  key.chain = 0;
  key.iter = 0;
  uint32_t nseq = 0;
  int chain = -1, it = 0;
  double Hbig = 0, delta = 0, jlo = 0, jhi = 0;
  uint32_t dirbits = 0;
  int level = 0, side = -1;
  uint32_t nleaf = 0, n_new = 0;
  double xi = 1.0;
  double h = 0, h2 = 0;
  int phase = PH_FWD, c = 0, If = 0, Ib = 0, cSim = 0, maxTry = 0;
  uint32_t steps_left = 0;
  double hh = 0, ha = 0;
  double Ham0 = 0, Hfwd = 0, H0 = 0, lwtf = 0, lwt = 0;
  // two-element per-side state kept as scalar pairs (index 0 = forward end, 1 = backward end);
  // dynamic indexing would push them to local memory
  double endH0 = 0, endH1 = 0, lwtSum0 = 0, lwtSum1 = 0, timeLen0 = 0, timeLen1 = 0;
  int maxInt0 = 0, maxInt1 = 0;
  double WoldSum = 1.0, WnewSum = 0.0;
  int L_ = 0, Lold = 0;
  double indexStat = 0, indexStatOld = 0, orbitLen = 0, orbitLenSam = 0;
  unsigned long long nF = 0, nB = 0, chainF = 0, chainB = 0, totF = 0, totB = 0;
  int NdS = 0, NdC = 0, stopCode = 0;
  bool bothPassive = false, candValid = false, forced = false, wIntact = true;
  int propCur = 0;
  // diagnostics statistics over used steps (WALNUTS.py:660-692)
  int sN = 0, sMinIf = 0, sMaxIf = 0, sMinC = 0, sMaxC = 0, sNne = 0, sNz = 0;
  double sMinL = 0, sMaxL = 0, sHmax = 0, sHmin = 0;
  bool sHnan = false;
  // per-pass accumulators
  double hp = 0.0;
  bool bad = false;

  auto useq = [&]() -> double { return rng_uniform(key, STREAM_SEQ, nseq++); };
  auto jit = [&](double u) -> double { return __dadd_rn(jlo, __dmul_rn(__dadd_rn(jhi, -jlo), u)); };

  auto save_ck = [&]() {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      ck[(0 * E + e) * NT + tid] = q[e];
      ck[(1 * E + e) * NT + tid] = v[e];
      ck[(2 * E + e) * NT + tid] = g[e];
    }
  };
  auto load_ck = [&](double vsign) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      q[e] = ck[(0 * E + e) * NT + tid];
      v[e] = vsign * ck[(1 * E + e) * NT + tid];
      g[e] = ck[(2 * E + e) * NT + tid];
    }
  };
  auto start_pass = [&](int cc) {
    steps_left = 1u << cc;
    hh = ldexp(h, -cc);
    ha = 0.5 * hh;
    bad = false;
  };
  // U-turn criterion, reference WALNUTS.py:95-97; (ql, vl) read from scratch, the other state is the
  // register-resident end (q, xi*v).  Orientation: minus end = more backward state.
  auto uturn_vs = [&](int viq, int viv) -> bool {
    double x[2] = {0.0, 0.0};
#pragma unroll
    for (int e2 = 0; e2 < E2; ++e2) {
      const double2 ql = *sc(viq, e2), vl = *sc(viv, e2);
      // forward level: plus = current, minus = left;  backward level: minus = current, plus = left
      const double t0 = xi * (q[2 * e2] - ql.x), t1 = xi * (q[2 * e2 + 1] - ql.y);   // qp - qm
      x[0] = fma(xi * v[2 * e2], t0, x[0]);
      x[0] = fma(xi * v[2 * e2 + 1], t1, x[0]);  // v_cur(fwd time) . tmp
      x[1] = fma(vl.x, t0, x[1]);
      x[1] = fma(vl.y, t1, x[1]);  // v_left . tmp
    }
    Grp::template sum<2>(x, red, parity);
    return (x[0] < 0.0) || (x[1] < 0.0);
  };

  int st = ST_CHAIN;
  for (;;) {
    // =============================== hot: one leapfrog micro-step ===============================
    if (st == ST_RUN) {
      // reference adaptiveIntegrators.py:79-84 (and :50-55 for fixedLeapFrog)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        v[e] = fma(ha, g[e], v[e]);
        q[e] = fma(hh, v[e], q[e]);
      }
      const double lpp = target.lp_grad(q, g, red, parity);
      double ke = 0.0;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        v[e] = fma(ha, g[e], v[e]);
        ke = fma(v[e], v[e], ke);
      }
      hp = fma(0.5, ke, -lpp);
      bad |= !finite_d(hp);
      if (--steps_left != 0) continue;
      st = ST_PASS_END;
    }
    // =============================== cold: per-chain state machine ==============================
    while (st != ST_RUN && st != ST_EXIT) {
      switch (st) {
        case ST_CHAIN: {  // grab the next chain from the queue
          uint32_t cidx = 0;
          if (t == 0) cidx = atomicAdd(P.queue, 1u);
          cidx = Grp::bcast0(cidx, &sh_bcast);
          if (cidx >= (uint32_t)P.n_chains) {
            st = ST_EXIT;
            break;
          }
          chain = (int)cidx;
          key.chain = P.chain_offset + cidx;
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int j = coord_of<G>(e, t);
            q[e] = (j < P.d) ? P.state[(size_t)chain * P.d + j] : 0.0;
          }
          Hbig = P.Hstep ? P.Hstep[chain] : P.H0;
          delta = P.delta ? P.delta[chain] : P.delta0;
          jlo = __dmul_rn(Hbig, __dadd_rn(1.0, -P.jitter));   // WALNUTS.py:298
          jhi = __dmul_rn(Hbig, __dadd_rn(1.0, P.jitter));
          it = 0;
          chainF = chainB = 0;
          st = ST_ITER;
          break;
        }
        case ST_ITER: {  // per-iteration setup, WALNUTS.py:196-276
          key.iter = P.iter0 + (uint32_t)it;
          nseq = 0;
          dirbits = 0;
          for (int k = 0; k < P.M; ++k) {
            const double u = rng_uniform(key, STREAM_DIR, (uint32_t)k);   // B = floor(U(0,2)), :216
            dirbits |= (u >= 0.5 ? 1u : 0u) << k;
          }
          double x[1];
          {
            double ke = 0.0;
#pragma unroll
            for (int e2 = 0; e2 < E2; ++e2) {   // v ~ N(0, I), :236
              double z0, z1;
              const int p = e2 * G + t;
              rng_normal_pair(key, STREAM_MOM, (uint32_t)p, z0, z1);
              v[2 * e2] = (2 * p < P.d) ? z0 : 0.0;
              v[2 * e2 + 1] = (2 * p + 1 < P.d) ? z1 : 0.0;
              ke = fma(v[2 * e2], v[2 * e2], ke);
              ke = fma(v[2 * e2 + 1], v[2 * e2 + 1], ke);
            }
            const double lpp = target.lp_grad(q, g, red, parity);        // :249
            x[0] = fma(0.5, ke, -lpp);
          }
          Grp::template sum<1>(x, red, parity);
          H0 = x[0];                                                      // :256
          endH0 = endH1 = H0;
#pragma unroll
          for (int e2 = 0; e2 < E2; ++e2) {   // origin is both ends; it is also the first proposal
            const double2 qq = make_double2(q[2 * e2], q[2 * e2 + 1]);
            *sc(V_PARK_Q, e2) = qq;
            *sc(V_PARK_V, e2) = make_double2(v[2 * e2], v[2 * e2 + 1]);
            *sc(V_PARK_G, e2) = make_double2(g[2 * e2], g[2 * e2 + 1]);
            *sc(V_PROP0, e2) = qq;
          }
          propCur = 0;
          lwtSum0 = lwtSum1 = 0.0;
          timeLen0 = timeLen1 = 0.0;
          maxInt0 = maxInt1 = 0;
          WoldSum = 1.0;
          L_ = 0;
          indexStat = 0.0;
          orbitLen = orbitLenSam = 0.0;
          nF = nB = 0;
          NdS = NdC = 0;
          stopCode = 0;
          bothPassive = false;
          forced = false;
          sN = 0;
          sNne = sNz = 0;
          sHmax = sHmin = H0;
          sHnan = false;
          side = -1;
          xi = 1.0;
          level = 0;
          st = ST_LEVEL;
          break;
        }
        case ST_LEVEL: {  // start doubling `level`, WALNUTS.py:281-294
          const int ns = (dirbits >> level) & 1u;   // 0 forward, 1 backward
          const double nxi = ns ? -1.0 : 1.0;
          if (side < 0) {
            // registers hold the origin with forward-time v; switch to integration convention
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] *= nxi;
          } else if (ns != side) {
            // swap the active end with the parked one (stored in forward-time convention)
#pragma unroll
            for (int e2 = 0; e2 < E2; ++e2) {
              const double2 pq = *sc(V_PARK_Q, e2), pv = *sc(V_PARK_V, e2), pg = *sc(V_PARK_G, e2);
              *sc(V_PARK_Q, e2) = make_double2(q[2 * e2], q[2 * e2 + 1]);
              *sc(V_PARK_V, e2) = make_double2(xi * v[2 * e2], xi * v[2 * e2 + 1]);
              *sc(V_PARK_G, e2) = make_double2(g[2 * e2], g[2 * e2 + 1]);
              q[2 * e2] = pq.x; q[2 * e2 + 1] = pq.y;
              v[2 * e2] = nxi * pv.x; v[2 * e2 + 1] = nxi * pv.y;
              g[2 * e2] = pg.x; g[2 * e2 + 1] = pg.y;
            }
          }
          side = ns;
          xi = nxi;
          nleaf = 0;
          n_new = 1u << level;
          WnewSum = 0.0;
          Lold = L_;
          indexStatOld = indexStat;
          candValid = false;
          st = ST_MACRO;
          break;
        }
        case ST_MACRO: {  // start one macro step from the active end
          ++nleaf;
          if (level == 0) {
            h = jit(useq());              // :298
            orbitLen += h;                // :300
          } else if (nleaf & 1u) {
            h = jit(useq());              // :395 (two draws per leaf pair)
            h2 = jit(useq());
          } else {
            h = h2;
          }
          Ham0 = side ? endH1 : endH0;
          phase = PH_FWD;
          c = (P.kind == KIND_FIXED) ? 0 : P.minC;
          if (P.kind != KIND_FIXED) save_ck();   // S = start state (integration convention)
          wIntact = true;
          start_pass(c);
          st = ST_RUN;
          break;
        }
        case ST_PASS_END: {  // a pass of 2^c micro-steps finished
          double x[2] = {hp, bad ? 1.0 : 0.0};
          Grp::template sum<2>(x, red, parity);
          const double Hend = x[0];
          const bool anybad = x[1] != 0.0;
          if (phase == PH_FWD) {
            nF += 1ull << c;
            const bool ok = !anybad && fabs(Ham0 - Hend) < delta;   // adaptiveIntegrators.py:87-92
            if (!(P.kind == KIND_FIXED || ok || c == P.maxC)) {
              ++c;
              load_ck(1.0);
              start_pass(c);
              st = ST_RUN;
              break;
            }
            If = c;
            cSim = If;
            lwtf = 0.0;
            if (P.kind == KIND_R2P) {
              if (useq() < P.p0) {              // adaptiveIntegrators.py:392
                lwtf = P.log_p0;
              } else {                          // :400-424 redo at If+1
                cSim = If + 1;
                phase = PH_REDO;
                load_ck(1.0);
                start_pass(cSim);
                st = ST_RUN;
                break;
              }
            }
          } else if (phase == PH_REDO) {
            nF += 1ull << cSim;
            lwtf = P.log_1mp0;
          }
          if (phase != PH_BWD) {
            // forward simulation done: registers hold the out state O
            Hfwd = Hend;
            if (P.kind == KIND_FIXED) {
              Ib = 0;
              lwt = 0.0;
              st = ST_LEAF;
              break;
            }
            if (P.kind == KIND_D || cSim == If) { maxTry = If - 1; Ib = If; }   // :104-111 / :430-433
            else { maxTry = P.maxC; Ib = P.maxC; }                              // :434-437
            if (maxTry >= P.minC) {
              save_ck();             // O replaces S
              wIntact = false;
              phase = PH_BWD;
              c = P.minC;
#pragma unroll
              for (int e = 0; e < E; ++e) v[e] = -v[e];
              start_pass(c);
              st = ST_RUN;
              break;
            }
          } else {
            nB += 1ull << c;
            const bool ok = !anybad && fabs(Hfwd - Hend) < delta;   // :129-132 / :461-464
            if (ok) Ib = c;
            if (!ok && c < maxTry) {
              ++c;
              load_ck(-1.0);
              start_pass(c);
              st = ST_RUN;
              break;
            }
          }
          // macro step complete
          if (P.kind == KIND_D) {
            lwt = (If != Ib) ? WN_LOG_ZERO : 0.0;                    // :136
          } else {
            double lwtb = WN_LOG_ZERO;                               // :467-471
            if (cSim == Ib) lwtb = P.log_p0;
            else if (cSim == Ib + 1) lwtb = P.log_1mp0;
            lwt = lwtb - lwtf;
          }
          if (!wIntact) load_ck(1.0);
          st = ST_LEAF;
          break;
        }
        case ST_LEAF: {  // driver bookkeeping after a macro step, WALNUTS.py:302-368,398-570
          const int idx = (level == 0) ? (side ? -1 : 1) : (side ? maxInt1 - 1 : maxInt0 + 1);
          const double tl = (level == 0) ? h : (side ? timeLen1 : timeLen0) + h;
          if (side) { maxInt1 = idx; timeLen1 = tl; endH1 = Hfwd; }
          else { maxInt0 = idx; timeLen0 = tl; endH0 = Hfwd; }
          {  // running statistics over used steps
            const int cs = (P.kind == KIND_FIXED) ? 0 : cSim;
            if (sN == 0) {
              sMinIf = sMaxIf = If;
              sMinC = sMaxC = cs;
              sMinL = sMaxL = lwt;
            } else {
              sMinIf = min(sMinIf, If); sMaxIf = max(sMaxIf, If);
              sMinC = min(sMinC, cs); sMaxC = max(sMaxC, cs);
              sMinL = fmin(sMinL, lwt); sMaxL = fmax(sMaxL, lwt);
            }
            ++sN;
            sNne += (If != Ib);
            sNz += (If == 0);
            if (Hfwd != Hfwd) sHnan = true;
            else { sHmax = fmax(sHmax, Hfwd); sHmin = fmin(sHmin, Hfwd); }
          }
          if (!finite_d(Hfwd)) {   // forced reject, :316,350,414,457,501,544 (quirks A14 ii, iii)
            forced = true;
            if (level == 0 || (nleaf & 1u)) stopCode = 999;
            st = ST_ITER_END;
            break;
          }
          double ls = side ? lwtSum1 : lwtSum0;
          if (level == 0) ls = lwt;                                                  // :321,354
          else if (!(side == 1 && !(nleaf & 1u))) ls += lwt;                         // :420,507,550; quirk A14(i)
          if (side) lwtSum1 = ls; else lwtSum0 = ls;
          const double Wnew = exp(-Hfwd + H0 + ls);                                  // :322,...
          bool pick;
          if (level == 0) {
            WnewSum = Wnew;
            pick = true;                                                            // :326,359
          } else {
            WnewSum += Wnew;
            pick = false;
            if (WnewSum > WN_WT_SUM_THRESH) pick = useq() < Wnew / WnewSum;          // :426,464,512,554
            orbitLen += h;                                                          // :432,...
          }
          if (pick) {
#pragma unroll
            for (int e2 = 0; e2 < E2; ++e2) *sc(V_PROP0 + (propCur ^ 1), e2) = make_double2(q[2 * e2], q[2 * e2 + 1]);
            candValid = true;
            L_ = idx;
            indexStat = side ? -tl : tl;
          }
          bool sub = false;
          if (level > 0) {
            if (nleaf & 1u) {
              // left end of the pending dyadic levels 1..ctz(nleaf-1) (all when nleaf == 1)
              const int lvl = (nleaf == 1u) ? level : (__ffs(nleaf - 1u) - 1);
#pragma unroll
              for (int e2 = 0; e2 < E2; ++e2) {
                *sc(V_STACK + 2 * lvl, e2) = make_double2(q[2 * e2], q[2 * e2 + 1]);
                *sc(V_STACK + 2 * lvl + 1, e2) = make_double2(xi * v[2 * e2], xi * v[2 * e2 + 1]);
              }
            } else {
              // post-order sub-U-turn checks (WALNUTS.py:22-41 plan; :479,568,582)
              for (int s = 1; s <= level && (nleaf & ((1u << s) - 1u)) == 0u; ++s) {
                const uint32_t m = nleaf - (1u << s) + 1u;
                const int lvl = (m == 1u) ? level : (__ffs(m - 1u) - 1);
                if (uturn_vs(V_STACK + 2 * lvl, V_STACK + 2 * lvl + 1)) {
                  sub = true;
                  break;
                }
              }
            }
          }
          if (sub) {                         // :597-605
            indexStat = indexStat / (timeLen0 + timeLen1);
            candValid = false;
            L_ = Lold;
            indexStat = indexStatOld;
            NdS = level;
            NdC = level + 1;
            stopCode = 5;
            st = ST_ITER_END;
          } else {
            st = (nleaf == n_new) ? ST_LEVEL_END : ST_MACRO;
          }
          break;
        }
        case ST_LEVEL_END: {  // WALNUTS.py:595-648
          indexStat = indexStat / (timeLen0 + timeLen1);                         // :595
          if (!(useq() < WnewSum / WoldSum)) {                                       // :613
            L_ = Lold;
            indexStat = indexStatOld;
          } else if (candValid) {
            propCur ^= 1;
          }
          candValid = false;
          const bool joined = uturn_vs(V_PARK_Q, V_PARK_V);                          // :622
          bothPassive = (lwtSum1 < WN_LOG_ZERO + 1.0) && (lwtSum0 < WN_LOG_ZERO + 1.0);   // :624
          NdS = NdC = level + 1;
          orbitLenSam = orbitLen;
          if (joined || bothPassive) {
            stopCode = joined ? 4 : -4;
            st = ST_ITER_END;
            break;
          }
          WoldSum += WnewSum;                                                        // :641
          ++level;
          st = (level == P.M) ? ST_ITER_END : ST_LEVEL;
          break;
        }
        case ST_ITER_END: {  // qc = qProp; outputs, WALNUTS.py:653-695
          const int pv = V_PROP0 + ((forced && candValid) ? (propCur ^ 1) : propCur);
#pragma unroll
          for (int e2 = 0; e2 < E2; ++e2) {
            const double2 qq = *sc(pv, e2);
            q[2 * e2] = qq.x;
            q[2 * e2 + 1] = qq.y;
          }
          const size_t row = (size_t)it * P.n_chains + chain;
          if (P.draws) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
              const int j = coord_of<G>(e, t);
              if (j < P.dg) P.draws[row * P.dg + j] = q[e];
            }
          }
          if (P.diag && t == 0) {
            double* dg = P.diag + row * 24;
            const double n = (double)sN;
            dg[0] = L_; dg[1] = NdS; dg[2] = orbitLen; dg[3] = orbitLenSam;
            dg[4] = maxInt0; dg[5] = maxInt1; dg[6] = (double)nF; dg[7] = (double)nB;
            dg[8] = sMinIf; dg[9] = sMaxIf; dg[10] = sMinL; dg[11] = sMaxL;
            dg[12] = bothPassive ? 1.0 : 0.0;
            dg[13] = ((lwtSum1 < WN_LOG_ZERO + 1.0) || (lwtSum0 < WN_LOG_ZERO + 1.0)) ? 1.0 : 0.0;
            dg[14] = (double)sNne / n; dg[15] = Hbig; dg[16] = (double)sNz / n;
            dg[17] = sHnan ? __longlong_as_double(0x7ff8000000000000ll) : sHmax - sHmin;
            dg[18] = delta; dg[19] = stopCode; dg[20] = NdC; dg[21] = sMinC; dg[22] = sMaxC;
            dg[23] = indexStat;
          }
          chainF += nF;
          chainB += nB;
          ++it;
          if (it < P.n_iter) {
            st = ST_ITER;
            break;
          }
#pragma unroll
          for (int e = 0; e < E; ++e) {
            const int j = coord_of<G>(e, t);
            if (j < P.d) P.state[(size_t)chain * P.d + j] = q[e];
          }
          if (t == 0) {
            if (P.nevalF) P.nevalF[chain] = chainF;
            if (P.nevalB) P.nevalB[chain] = chainB;
          }
          totF += chainF;
          totB += chainB;
          st = ST_CHAIN;
          break;
        }
        default:
          st = ST_EXIT;
          break;
      }
    }
    if (st == ST_EXIT) break;
  }
  if (t == 0 && (totF | totB)) {
    atomicAdd(P.totals, totF);
    atomicAdd(P.totals + 1, totB);
  }
}

}  // namespace wn
