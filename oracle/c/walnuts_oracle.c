/*
 * C restatement of the WALNUTSpy transition (TEST INFRASTRUCTURE / CPU BASELINE ONLY; nothing in the
 * product package links or calls this file).
 *
 * Follows oracle/walnutspy_oracle.py line by line, which restates reference
 * WALNUTSpy/WALNUTS.py:111-727 (driver) and WALNUTSpy/adaptiveIntegrators.py:49-137,361-475
 * (fixedLeapFrog, adaptLeapFrogD, adaptLeapFrogR2P); targets: WALNUTSpy/targetDistr.py:18-21 (stdGauss),
 * :74-78 (funnel10) and the diagonal Gaussian of SURVEY.md row T2.  Philox4x32-10 streams as in
 * oracle/philox.py.  Pinned by tests/test_oracle_golden.py against the golden fixtures produced by the real
 * reference.  Stands in for the absent `walnuts_cpp` as the strong CPU baseline in bench.py.
 *
 * Build: see oracle/c/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, sums strictly left to right,
 * i.e. the arithmetic of the Python reference up to numpy's pairwise np.sum).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LOG_ZERO (-700.0)
enum { T_STD = 0, T_DIAG = 1, T_FUNNEL = 2 };
enum { FIXED = 0, ADAPT_D = 1, ADAPT_R2P = 2 };

/* ---------------------------------------------------------------- Philox4x32-10 (oracle/philox.py) */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1,
             n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
static double u53(uint32_t a, uint32_t b) { return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * 0x1p-53; }

typedef struct { uint32_t k0, k1, chain, iter, nseq; } rng_t;
static double rng_uniform(const rng_t* r, uint32_t stream, uint32_t idx) {
  uint32_t c[4] = {idx >> 1, r->iter, r->chain, stream};
  philox(c, r->k0, r->k1);
  return (idx & 1u) ? u53(c[2], c[3]) : u53(c[0], c[1]);
}
static double useq(rng_t* r) { return rng_uniform(r, 2, r->nseq++); }
static void rng_normals(const rng_t* r, int d, double* z) {
  for (int p = 0; 2 * p < d; ++p) {
    uint32_t c[4] = {(uint32_t)p, r->iter, r->chain, 1};
    philox(c, r->k0, r->k1);
    double u1 = u53(c[0], c[1]), u2 = u53(c[2], c[3]);
    double rad = sqrt(-2.0 * log(1.0 - u1)), ang = 2.0 * M_PI * u2;
    z[2 * p] = rad * cos(ang);
    if (2 * p + 1 < d) z[2 * p + 1] = rad * sin(ang);
  }
}

/* ---------------------------------------------------------------- targets: lp and gradient */
typedef struct { int kind, d; const double* s; } target_t;
static double lp_grad(const target_t* t, const double* q, double* g) {
  const int d = t->d;
  if (t->kind == T_FUNNEL) {
    const double LOG_SQRT_2PI = 0.91893853320467274178, n = d - 1;
    double e = exp(-q[0]), ss = 0.0;
    for (int i = 1; i < d; ++i) ss += q[i] * q[i];
    double q03 = q[0] / 3.0;
    double lp = (-(q03 * q03) / 2.0 - LOG_SQRT_2PI - log(3.0)) + (-0.5 * e * ss - n * LOG_SQRT_2PI - n * 0.5 * q[0]);
    g[0] = -0.5 * n - q[0] / 9.0 + 0.5 * e * ss;
    for (int i = 1; i < d; ++i) g[i] = -q[i] * e;
    return lp;
  }
  /* four partial sums (vectorisable): a fair CPU baseline should not be latency-bound on one accumulator;
     differs from the reference's left-to-right sum by ~1e-16 relative (tests allow 1e-10) */
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = 0;
  if (t->kind == T_STD) {
    for (; i + 4 <= d; i += 4) {
      g[i] = -q[i]; g[i + 1] = -q[i + 1]; g[i + 2] = -q[i + 2]; g[i + 3] = -q[i + 3];
      a0 += q[i] * g[i]; a1 += q[i + 1] * g[i + 1]; a2 += q[i + 2] * g[i + 2]; a3 += q[i + 3] * g[i + 3];
    }
    for (; i < d; ++i) { g[i] = -q[i]; a0 += q[i] * g[i]; }
  } else {
    const double* s = t->s;
    for (; i + 4 <= d; i += 4) {
      g[i] = -(q[i] * s[i]); g[i + 1] = -(q[i + 1] * s[i + 1]); g[i + 2] = -(q[i + 2] * s[i + 2]); g[i + 3] = -(q[i + 3] * s[i + 3]);
      a0 += q[i] * g[i]; a1 += q[i + 1] * g[i + 1]; a2 += q[i + 2] * g[i + 2]; a3 += q[i + 3] * g[i + 3];
    }
    for (; i < d; ++i) { g[i] = -(q[i] * s[i]); a0 += q[i] * g[i]; }
  }
  return 0.5 * ((a0 + a1) + (a2 + a3));
}
static double sumsq(const double* v, int d) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = 0;
  for (; i + 4 <= d; i += 4) { a0 += v[i] * v[i]; a1 += v[i + 1] * v[i + 1]; a2 += v[i + 2] * v[i + 2]; a3 += v[i + 3] * v[i + 3]; }
  for (; i < d; ++i) a0 += v[i] * v[i];
  return (a0 + a1) + (a2 + a3);
}

/* ---------------------------------------------------------------- one pass of 2^c micro-steps */
static double run_pass(const target_t* t, double* q, double* vv, double* g, double h, int c, int* all_finite) {
  const int d = t->d;
  const long nstep = 1L << c;
  const double hh = h / (double)nstep, a = 0.5 * hh;
  double Hk = 0.0;
  int ok = 1;
  for (long s = 0; s < nstep; ++s) {
    for (int i = 0; i < d; ++i) { vv[i] = vv[i] + a * g[i]; q[i] = q[i] + hh * vv[i]; }
    double f = lp_grad(t, q, g);
    for (int i = 0; i < d; ++i) vv[i] = vv[i] + a * g[i];
    Hk = -f + 0.5 * sumsq(vv, d);
    ok = ok && isfinite(Hk);
  }
  *all_finite = ok;
  return Hk;
}

typedef struct {
  double *q, *v, *g;   /* end state, v in forward-time convention */
  double H;
} end_t;
typedef struct { int nF, nB, If, Ib, c; double lwt, H; } mres_t;

/* adaptiveIntegrators.py:49-137,361-475; scratch: qs, vs, gs, qo, vo, go (d each) */
static void macro_step(const target_t* t, int kind, end_t* e, double h, double xi, double delta, int minC, int maxC,
                       double p0, rng_t* rng, double* w, mres_t* out) {
  const int d = t->d;
  double *qq = w, *vv = w + d, *gg = w + 2 * d, *qo = w + 3 * d, *vo = w + 4 * d, *go = w + 5 * d;
  const double Ham0 = e->H;
  int fin;
  long nF = 0, nB = 0;
  int If = maxC, cSim;
  double Hl = 0.0, lwtf = 0.0;
  if (kind == FIXED) {
    memcpy(qq, e->q, d * sizeof(double)); memcpy(gg, e->g, d * sizeof(double));
    for (int i = 0; i < d; ++i) vv[i] = xi * e->v[i];
    Hl = run_pass(t, qq, vv, gg, h, 0, &fin);
    memcpy(e->q, qq, d * sizeof(double)); memcpy(e->g, gg, d * sizeof(double));
    for (int i = 0; i < d; ++i) e->v[i] = xi * vv[i];
    e->H = Hl;
    out->nF = 1; out->nB = 0; out->If = out->Ib = out->c = 0; out->lwt = 0.0; out->H = Hl;
    return;
  }
  for (int c = minC; c <= maxC; ++c) {
    memcpy(qq, e->q, d * sizeof(double)); memcpy(gg, e->g, d * sizeof(double));
    for (int i = 0; i < d; ++i) vv[i] = xi * e->v[i];
    Hl = run_pass(t, qq, vv, gg, h, c, &fin);
    nF += 1L << c;
    if (fin && fabs(Ham0 - Hl) < delta) { If = c; break; }
  }
  cSim = If;
  if (kind == ADAPT_R2P) {
    if (useq(rng) < p0) lwtf = log(p0);
    else {
      cSim = If + 1;
      memcpy(qq, e->q, d * sizeof(double)); memcpy(gg, e->g, d * sizeof(double));
      for (int i = 0; i < d; ++i) vv[i] = xi * e->v[i];
      Hl = run_pass(t, qq, vv, gg, h, cSim, &fin);
      nF += 1L << cSim;
      lwtf = log(1.0 - p0);
    }
  }
  memcpy(qo, qq, d * sizeof(double)); memcpy(vo, vv, d * sizeof(double)); memcpy(go, gg, d * sizeof(double));
  const double HO = Hl;
  int maxTry, Ib;
  if (kind == ADAPT_D || cSim == If) { maxTry = If - 1; Ib = If; } else { maxTry = maxC; Ib = maxC; }
  for (int c = minC; c <= maxTry; ++c) {
    memcpy(qq, qo, d * sizeof(double)); memcpy(gg, go, d * sizeof(double));
    for (int i = 0; i < d; ++i) vv[i] = -vo[i];
    double Hb = run_pass(t, qq, vv, gg, h, c, &fin);
    nB += 1L << c;
    if (fin && fabs(HO - Hb) < delta) { Ib = c; break; }
  }
  double lwt;
  if (kind == ADAPT_D) lwt = (If != Ib) ? LOG_ZERO : 0.0;
  else {
    double lwtb = LOG_ZERO;
    if (cSim == Ib) lwtb = log(p0);
    else if (cSim == Ib + 1) lwtb = log(1.0 - p0);
    lwt = lwtb - lwtf;
  }
  memcpy(e->q, qo, d * sizeof(double)); memcpy(e->g, go, d * sizeof(double));
  for (int i = 0; i < d; ++i) e->v[i] = xi * vo[i];
  e->H = HO;
  out->nF = (int)nF; out->nB = (int)nB; out->If = If; out->Ib = Ib; out->c = cSim; out->lwt = lwt; out->H = HO;
}

static int stop_condition(const double* qm, const double* vm, const double* qp, const double* vp, int d) {
  double a = 0.0, b = 0.0;
  for (int i = 0; i < d; ++i) { double t = qp[i] - qm[i]; a += vp[i] * t; b += vm[i] * t; }
  return (a < 0.0) || (b < 0.0);
}
static int ctz32(uint32_t x) { return __builtin_ctz(x); }

/* One chain: n_iter transitions.  draws [n_iter, d], diag [n_iter, 24] (either may be NULL). */
int wno_run_chain(int target, int kind, int d, const double* inv_var, const double* q0, double H, double delta,
                  double jitter, int M, int minC, int maxC, double p0, uint64_t seed, uint32_t chain,
                  uint32_t first_iter, int n_iter, double* draws, double* diag, double* q_out, uint64_t* nevals,
                  int compat) {
  target_t t = {target, d, inv_var};
  rng_t rng = {(uint32_t)seed, (uint32_t)(seed >> 32), chain, 0, 0};
  double* buf = (double*)malloc(sizeof(double) * d * (size_t)(3 + 3 + 3 + 6 + 2 + 2 * (M + 2)));
  if (!buf) return -1;
  double* qc = buf;
  end_t ends[2];
  ends[0].q = buf + 1 * d; ends[0].v = buf + 2 * d; ends[0].g = buf + 3 * d;
  ends[1].q = buf + 4 * d; ends[1].v = buf + 5 * d; ends[1].g = buf + 6 * d;
  double *v0 = buf + 7 * d, *g0 = buf + 8 * d, *w = buf + 9 * d, *qProp = buf + 15 * d, *qPropLast = buf + 16 * d,
         *stack = buf + 17 * d;
  memcpy(qc, q0, d * sizeof(double));
  const double thresh = exp(LOG_ZERO + 1.0);
  const double lo = H * (1 - jitter), hi = H * (1 + jitter);
  uint64_t total = 0;
  for (int it = 0; it < n_iter; ++it) {
    rng.iter = first_iter + (uint32_t)it;
    rng.nseq = 0;
    uint32_t dirbits = 0;
    for (int k = 0; k < M; ++k) if (floor(2.0 * rng_uniform(&rng, 0, (uint32_t)k)) == 1.0) dirbits |= 1u << k;
    rng_normals(&rng, d, v0);
    double f0 = lp_grad(&t, qc, g0);
    const double H0 = -f0 + 0.5 * sumsq(v0, d);
    for (int s = 0; s < 2; ++s) {
      memcpy(ends[s].q, qc, d * sizeof(double)); memcpy(ends[s].v, v0, d * sizeof(double));
      memcpy(ends[s].g, g0, d * sizeof(double)); ends[s].H = H0;
    }
    double lwtSum[2] = {0, 0}, timeLen[2] = {0, 0}, WoldSum = 1.0, indexStat = 0, orbitLen = 0, orbitLenSam = 0;
    int maxInt[2] = {0, 0}, L_ = 0, NdS = 0, NdC = 0, stopCode = 0, bothPassive = 0, forced = 0;
    long nF = 0, nB = 0;
    int sN = 0, sMinIf = 0, sMaxIf = 0, sMinC = 0, sMaxC = 0, sNne = 0, sNz = 0, sHnan = 0;
    double sMinL = 0, sMaxL = 0, sHmax = H0, sHmin = H0;
    memcpy(qProp, qc, d * sizeof(double));
    for (int i = 0; i < M; ++i) {
      const int side = (dirbits >> i) & 1;
      const double xi = side ? -1.0 : 1.0;
      const uint32_t n_new = 1u << i;
      memcpy(qPropLast, qProp, d * sizeof(double));
      const int Lold = L_;
      const double indexStatOld = indexStat;
      double WnewSum = 0.0, h2 = 0.0;
      int expand = 1;
      for (uint32_t n = 1; n <= n_new; ++n) {
        double h;
        if (i == 0) { h = lo + (hi - lo) * useq(&rng); orbitLen += h; }
        else if (n & 1u) { h = lo + (hi - lo) * useq(&rng); h2 = lo + (hi - lo) * useq(&rng); }
        else h = h2;
        mres_t o;
        macro_step(&t, kind, &ends[side], h, xi, delta, minC, maxC, p0, &rng, w, &o);
        nF += o.nF; nB += o.nB;
        const int idx = (i > 0) ? maxInt[side] + (side ? -1 : 1) : (side ? -1 : 1);
        maxInt[side] = idx;
        timeLen[side] = (i == 0) ? h : timeLen[side] + h;
        if (sN == 0) { sMinIf = sMaxIf = o.If; sMinC = sMaxC = o.c; sMinL = sMaxL = o.lwt; }
        else {
          if (o.If < sMinIf) sMinIf = o.If; if (o.If > sMaxIf) sMaxIf = o.If;
          if (o.c < sMinC) sMinC = o.c; if (o.c > sMaxC) sMaxC = o.c;
          if (o.lwt < sMinL) sMinL = o.lwt; if (o.lwt > sMaxL) sMaxL = o.lwt;
        }
        ++sN; sNne += (o.If != o.Ib); sNz += (o.If == 0);
        if (o.H != o.H) sHnan = 1; else { if (o.H > sHmax) sHmax = o.H; if (o.H < sHmin) sHmin = o.H; }
        if (!isfinite(o.H)) { forced = 1; if (i == 0 || (n & 1u)) stopCode = 999; break; }
        if (i == 0) lwtSum[side] = o.lwt;
        else if (!compat || !(side == 1 && !(n & 1u))) lwtSum[side] += o.lwt;       /* quirk A14(i) */
        const double Wnew = exp(-o.H + H0 + lwtSum[side]);
        const double* eq = ends[side].q; const double* ev = ends[side].v;
        if (i == 0) {
          WnewSum = Wnew;
          memcpy(qProp, eq, d * sizeof(double)); L_ = idx; indexStat = xi * timeLen[side];
        } else {
          WnewSum += Wnew;
          if (WnewSum > thresh && useq(&rng) < Wnew / WnewSum) {
            memcpy(qProp, eq, d * sizeof(double)); L_ = idx; indexStat = xi * timeLen[side];
          }
          orbitLen += h;
          if (n & 1u) {
            const int lvl = (n == 1u) ? i : ctz32(n - 1u);
            memcpy(stack + (size_t)(2 * lvl) * d, eq, d * sizeof(double));
            memcpy(stack + (size_t)(2 * lvl + 1) * d, ev, d * sizeof(double));
          } else {
            for (int s = 1; s <= i && (n & ((1u << s) - 1u)) == 0u; ++s) {
              const uint32_t m = n - (1u << s) + 1u;
              const int lvl = (m == 1u) ? i : ctz32(m - 1u);
              const double *ql = stack + (size_t)(2 * lvl) * d, *vl = stack + (size_t)(2 * lvl + 1) * d;
              const int ut = (xi > 0) ? stop_condition(ql, vl, eq, ev, d) : stop_condition(eq, ev, ql, vl, d);
              if (ut) { expand = 0; break; }
            }
            if (!expand) break;
          }
        }
      }
      if (forced) break;
      indexStat = indexStat / (timeLen[0] + timeLen[1]);
      if (!expand) {
        memcpy(qProp, qPropLast, d * sizeof(double)); L_ = Lold; indexStat = indexStatOld;
        NdS = i; NdC = i + 1; stopCode = 5;
        break;
      }
      if (!(useq(&rng) < WnewSum / WoldSum)) { memcpy(qProp, qPropLast, d * sizeof(double)); L_ = Lold; indexStat = indexStatOld; }
      const int joined = stop_condition(ends[1].q, ends[1].v, ends[0].q, ends[0].v, d);
      bothPassive = (lwtSum[1] < LOG_ZERO + 1.0) && (lwtSum[0] < LOG_ZERO + 1.0);
      NdS = NdC = i + 1; orbitLenSam = orbitLen;
      if (joined || bothPassive) { stopCode = joined ? 4 : -4; break; }
      WoldSum += WnewSum;
    }
    memcpy(qc, qProp, d * sizeof(double));
    total += (uint64_t)(nF + nB);
    if (draws) memcpy(draws + (size_t)it * d, qc, d * sizeof(double));
    if (diag) {
      double* g = diag + (size_t)it * 24;
      g[0] = L_; g[1] = NdS; g[2] = orbitLen; g[3] = orbitLenSam; g[4] = maxInt[0]; g[5] = maxInt[1];
      g[6] = (double)nF; g[7] = (double)nB; g[8] = sMinIf; g[9] = sMaxIf; g[10] = sMinL; g[11] = sMaxL;
      g[12] = bothPassive; g[13] = (lwtSum[1] < LOG_ZERO + 1.0) || (lwtSum[0] < LOG_ZERO + 1.0);
      g[14] = (double)sNne / sN; g[15] = H; g[16] = (double)sNz / sN; g[17] = sHnan ? NAN : sHmax - sHmin;
      g[18] = delta; g[19] = stopCode; g[20] = NdC; g[21] = sMinC; g[22] = sMaxC; g[23] = indexStat;
    }
  }
  if (q_out) memcpy(q_out, qc, d * sizeof(double));
  if (nevals) *nevals = total;
  free(buf);
  return 0;
}

/* Many independent chains on `threads` host threads (pthreads, dynamic chain queue); q [n_chains, d] in/out. */
#include <pthread.h>
typedef struct {
  int target, kind, d, n_chains, M, minC, maxC, n_iter, compat;
  const double* inv_var;
  double* q;
  double H, delta, jitter, p0;
  uint64_t seed;
  uint32_t chain0, first_iter;
  volatile int next;
  uint64_t total;
  int err;
  pthread_mutex_t mu;
} many_t;

static void* many_worker(void* arg) {
  many_t* m = (many_t*)arg;
  uint64_t tot = 0;
  int err = 0;
  for (;;) {
    const int c = __atomic_fetch_add(&m->next, 1, __ATOMIC_RELAXED);
    if (c >= m->n_chains) break;
    uint64_t ne = 0;
    err |= wno_run_chain(m->target, m->kind, m->d, m->inv_var, m->q + (size_t)c * m->d, m->H, m->delta, m->jitter,
                         m->M, m->minC, m->maxC, m->p0, m->seed, m->chain0 + (uint32_t)c, m->first_iter, m->n_iter,
                         NULL, NULL, m->q + (size_t)c * m->d, &ne, m->compat);
    tot += ne;
  }
  pthread_mutex_lock(&m->mu);
  m->total += tot;
  m->err |= err;
  pthread_mutex_unlock(&m->mu);
  return NULL;
}

int wno_run_many(int target, int kind, int d, const double* inv_var, double* q, int n_chains, double H, double delta,
                 double jitter, int M, int minC, int maxC, double p0, uint64_t seed, uint32_t chain0,
                 uint32_t first_iter, int n_iter, int threads, uint64_t* nevals_total, int compat) {
  many_t m = {target, kind, d, n_chains, M, minC, maxC, n_iter, compat, inv_var, q, H, delta, jitter, p0, seed, chain0,
              first_iter, 0, 0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (threads < 1) threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  if (!th) return -1;
  for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, many_worker, &m);
  for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
  free(th);
  if (nevals_total) *nevals_total = m.total;
  return m.err;
}
