// Cross-chain convergence statistics on the device, across GPUs: bulk ESS and split-R-hat of the monitored draws
// (Vehtari, Gelman, Simpson, Carpenter, Buerkner 2021 -- the estimator behind arviz.ess, which the reference's
// experiment scripts use: WALNUTSpy_examples/gaussian/mainGaussESS.py:17,50-55) and cross-GPU moments.
//
// Multi-GPU: chains shard across GPUs with no exchange while sampling; the ONE collective is an NCCL all-gather of the
// monitored draws (wn_ess_rhat) or an all-reduce of (n, sum, sum of squares) (wn_moments_all).  NCCL is bound at run
// time with dlopen -- the library has no link-time dependency on it; a process that already loaded libnccl.so.2
// (e.g. through torch) shares that copy.
//
// Pipeline of wn_ess_rhat per coordinate: gather column -> radix sort (cub) -> ranks -> z = Phi^-1((r - 3/8) / (S + 1/4))
// -> split every chain in halves -> per split chain: mean, variance and ALL autocovariances in shared memory -> sums over
// chains in a fixed order -> Geyer's initial monotone sequence on the host (a few thousand numbers).
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every function is resolved with dlsym

#include <cmath>
#include <cstring>
#include <cub/cub.cuh>
#include <mutex>
#include <vector>

#include "wn_handle.hpp"

namespace {

struct NcclApi {
  void* dl = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string path;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

int nccl_bind(const char* path, std::string* err) {
  std::lock_guard<std::mutex> lock(g_nccl_mutex);
  if (g_nccl.dl) return WN_OK;
  void* dl = nullptr;
  if (path && *path) dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // already in the process (torch's bundled copy)
  if (!dl) dl = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!dl) dl = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!dl) {
    if (err) *err = std::string("libnccl.so.2 not found: ") + dlerror();
    return WN_EUNSUPPORTED;
  }
  NcclApi a;
  a.dl = dl;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(dl, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(dl, "ncclCommInitRank");
  a.CommInitAll = (decltype(a.CommInitAll))dlsym(dl, "ncclCommInitAll");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(dl, "ncclCommDestroy");
  a.AllGather = (decltype(a.AllGather))dlsym(dl, "ncclAllGather");
  a.AllReduce = (decltype(a.AllReduce))dlsym(dl, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(dl, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommInitAll || !a.CommDestroy || !a.AllGather || !a.AllReduce ||
      !a.GetErrorString) {
    if (err) *err = "libnccl.so.2 lacks a required symbol";
    return WN_EUNSUPPORTED;
  }
  g_nccl = a;
  return WN_OK;
}

#define NCCL_TRY(h, expr)                                                                         \
  do {                                                                                            \
    ncclResult_t _r = (expr);                                                                     \
    if (_r != ncclSuccess) return fail(h, WN_ECUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

// ---- kernels -------------------------------------------------------------------------------------------------------

// column j of draws [W][n_iter][n_chains][dg] -> x[(w * n_chains + c) * n_iter + t]  (chain-major), idx = identity
__global__ void gather_column(const double* __restrict__ draws, int W, int n_iter, int n_chains, int dg, int j,
                              double* __restrict__ x, unsigned* __restrict__ idx) {
  const size_t S = (size_t)W * n_iter * n_chains;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (size_t)gridDim.x * blockDim.x) {
    const size_t t = i % n_iter, wc = i / n_iter, c = wc % n_chains, w = wc / n_chains;
    x[i] = draws[(((size_t)w * n_iter + t) * n_chains + c) * dg + j];
    idx[i] = (unsigned)i;
  }
}

// z[idx_sorted[r]] = Phi^-1((r + 1 - 3/8) / (S + 1/4))   (ordinal ranks; ties have probability zero)
__global__ void rank_normalize(const unsigned* __restrict__ idx_sorted, size_t S, double* __restrict__ z) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < S; r += (size_t)gridDim.x * blockDim.x)
    z[idx_sorted[r]] = normcdfinv(((double)(r + 1) - 0.375) / ((double)S + 0.25));
}

// One block per split chain: chain c of length n is split into [0, h) and [n - h, n), h = n / 2 (split == 0: the
// whole chain, h = n).  Writes mean, variance (1 / (h - 1)) and the biased autocovariances acov[lag] (1 / h), lag < h.
__global__ void split_chain_stats(const double* __restrict__ z, int n, int h, int split, double* __restrict__ mean,
                                  double* __restrict__ var, double* __restrict__ acov /* [n_split][h] */) {
  extern __shared__ double xs[];
  __shared__ double red[32];
  const int sc = blockIdx.x;
  const int c = split ? (sc >> 1) : sc;
  const int off = (split && (sc & 1)) ? (n - h) : 0;
  const double* src = z + (size_t)c * n + off;
  double s = 0.0;
  for (int i = threadIdx.x; i < h; i += blockDim.x) {
    const double v = src[i];
    xs[i] = v;
    s += v;
  }
  // block sum (fixed order: deterministic)
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const double m = tot / h;
  __syncthreads();
  for (int i = threadIdx.x; i < h; i += blockDim.x) xs[i] -= m;
  __syncthreads();
  for (int lag = threadIdx.x; lag < h; lag += blockDim.x) {
    double a = 0.0;
    for (int i = 0; i + lag < h; ++i) a = fma(xs[i], xs[i + lag], a);
    acov[(size_t)sc * h + lag] = a / h;
    if (lag == 0) {
      mean[sc] = m;
      var[sc] = a / (h - 1);
    }
  }
}

// out[lag] = sum over split chains of acov[sc][lag] (fixed order); out[h] = sum mean, out[h+1] = sum mean^2, out[h+2] = sum var
__global__ void reduce_chains(const double* __restrict__ acov, const double* __restrict__ mean,
                              const double* __restrict__ var, int n_split, int h, double* __restrict__ out) {
  const int lag = blockIdx.x * blockDim.x + threadIdx.x;
  if (lag < h) {
    double s = 0.0;
    for (int c = 0; c < n_split; ++c) s += acov[(size_t)c * h + lag];
    out[lag] = s;
  } else if (lag == h) {
    double a = 0.0, b = 0.0, v = 0.0;
    for (int c = 0; c < n_split; ++c) { a += mean[c]; b += mean[c] * mean[c]; v += var[c]; }
    out[h] = a; out[h + 1] = b; out[h + 2] = v;
  }
}

// out[j] = sum over chains of state[c][j]  (pass 1)  /  of (state[c][j] - mean[j])^2  (pass 2, mean != null)
__global__ void moment_sums(const double* __restrict__ state, int n_chains, int d, const double* __restrict__ mean,
                            double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const double m = mean ? mean[j] : 0.0;
  double s = 0.0;
  for (int c = 0; c < n_chains; ++c) {
    const double x = state[(size_t)c * d + j] - m;
    s = mean ? fma(x, x, s) : s + x;
  }
  out[j] = s;
}

// Geyer's initial positive + monotone sequence on chain-averaged autocorrelations (the arviz / Stan rule);
// acov_sum[lag] = sum over the m chains of the biased lag autocovariance, n = draws per chain
void ess_from_sums(const double* acov_sum, int T, int m, int n, double sum_mean, double sum_mean2, double sum_var,
                   double* ess, double* rhat) {
  const double nan = std::nan("");
  *ess = nan;
  *rhat = nan;
  if (n < 4 || T < 2) return;
  const double mean_var = sum_var / m;
  const double mom = sum_mean / m;
  double var_plus = mean_var * (n - 1.0) / n;
  if (m > 1) var_plus += (sum_mean2 - m * mom * mom) / (m - 1);
  if (!(var_plus > 0)) return;
  auto rho = [&](int t) { return t < T ? 1.0 - (mean_var - acov_sum[t] / m) / var_plus : 0.0; };
  std::vector<double> rh((size_t)n + 2, 0.0);
  double even = 1.0, odd = rho(1);
  rh[0] = even;
  rh[1] = odd;
  int t = 1;
  while (t < n - 3 && (even + odd) > 0.0) {
    even = rho(t + 1);
    odd = rho(t + 2);
    if (even + odd >= 0) { rh[t + 1] = even; rh[t + 2] = odd; }
    t += 2;
  }
  const int max_t = t - 2;
  if (even > 0) rh[max_t + 1] = even;
  t = 1;
  while (t <= max_t - 2) {
    if (rh[t + 1] + rh[t + 2] > rh[t - 1] + rh[t]) {
      rh[t + 1] = (rh[t - 1] + rh[t]) / 2.0;
      rh[t + 2] = rh[t + 1];
    }
    t += 2;
  }
  const double total = (double)m * n;
  double tau = -1.0;
  for (int i = 0; i <= max_t; ++i) tau += 2.0 * rh[i];
  tau += rh[max_t + 1];
  tau = std::fmax(tau, 1.0 / std::log10(total));
  *ess = total / tau;
  *rhat = mean_var > 0 ? std::sqrt(var_plus / mean_var) : nan;
}

}  // namespace

extern "C" {

int wn_comm_load(const char* libnccl_path) { return nccl_bind(libnccl_path, nullptr); }

int wn_comm_unique_id(void* id128) {
  if (!id128) return WN_EINVAL;
  int rc = nccl_bind(nullptr, nullptr);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return WN_ECUDA;
  memcpy(id128, &id, sizeof(id));
  return WN_OK;
}

int wn_comm_init_rank(wn_handle* h, int nranks, int rank, const void* id128) {
  if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, WN_EINVAL, "wn_comm_init_rank: bad argument");
  if (h->comm) return fail(h, WN_ESTATE, "the handle already has a communicator");
  int rc = nccl_bind(nullptr, &h->err);
  if (rc) return rc;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  NCCL_TRY(h, g_nccl.CommInitRank(&comm, nranks, id, rank));
  h->comm = comm;
  h->comm_rank = rank;
  h->comm_size = nranks;
  return WN_OK;
}

int wn_comm_init_all(wn_handle** hs, int n) {
  if (!hs || n < 1) return WN_EINVAL;
  for (int i = 0; i < n; ++i)
    if (!hs[i] || hs[i]->comm) return fail(hs[i], WN_ESTATE, "wn_comm_init_all: null handle or communicator already set");
  int rc = nccl_bind(nullptr, &hs[0]->err);
  if (rc) return rc;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) devs[i] = hs[i]->cfg.device;
  std::vector<ncclComm_t> comms(n, nullptr);
  NCCL_TRY(hs[0], g_nccl.CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) {
    hs[i]->comm = comms[i];
    hs[i]->comm_rank = i;
    hs[i]->comm_size = n;
  }
  return WN_OK;
}

int wn_comm_destroy(wn_handle* h) {
  if (!h) return WN_EINVAL;
  if (h->comm) {
    cudaSetDevice(h->cfg.device);
    g_nccl.CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
    h->comm_size = 1;
    h->comm_rank = 0;
  }
  return WN_OK;
}

int wn_ess_rhat(wn_handle* h, const double* draws, int64_t n_iter, int64_t n_chains, int32_t dg, int on_device,
                int32_t split, double* ess, double* rhat) {
  if (!h || !draws || !ess || !rhat || n_iter < 4 || n_chains < 1 || dg < 1)
    return fail(h, WN_EINVAL, "wn_ess_rhat: need draws [n_iter >= 4, n_chains, dg] and ess / rhat [dg]");
  const int W = h->comm ? h->comm_size : 1;
  const int n = (int)n_iter;
  const int hlen = split ? n / 2 : n;
  if ((size_t)hlen * sizeof(double) > 200 * 1024) return fail(h, WN_EUNSUPPORTED, "wn_ess_rhat: more than 25 600 draws per (split) chain");
  const size_t per_rank = (size_t)n_iter * n_chains * dg;
  const size_t S = (size_t)W * n_iter * n_chains;
  if (S > 0x7ffffff0ull) return fail(h, WN_EUNSUPPORTED, "wn_ess_rhat: more than 2^31 draws per coordinate");
  const int n_split = (int)((size_t)W * n_chains * (split ? 2 : 1));
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  cudaStream_t st = h->stream;
  double *d_in = nullptr, *d_all = nullptr, *d_x = nullptr, *d_xs = nullptr, *d_z = nullptr, *d_mean = nullptr,
         *d_var = nullptr, *d_acov = nullptr, *d_out = nullptr;
  unsigned *d_idx = nullptr, *d_idxs = nullptr;
  void* d_tmp = nullptr;
  size_t tmp_bytes = 0;
  std::vector<double> out((size_t)hlen + 3);
  int rc = WN_OK;
  auto cleanup = [&]() {
    cudaFree(d_in); cudaFree(d_all); cudaFree(d_x); cudaFree(d_xs); cudaFree(d_z); cudaFree(d_mean); cudaFree(d_var);
    cudaFree(d_acov); cudaFree(d_out); cudaFree(d_idx); cudaFree(d_idxs); cudaFree(d_tmp);
  };
#define ST_TRY(expr)                                                                    \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      cleanup();                                                                        \
      return fail(h, WN_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    }                                                                                   \
  } while (0)
  const double* src = draws;
  if (!on_device) {
    ST_TRY(cudaMalloc(&d_in, per_rank * sizeof(double)));
    ST_TRY(cudaMemcpyAsync(d_in, draws, per_rank * sizeof(double), cudaMemcpyHostToDevice, st));
    src = d_in;
  }
  if (W > 1) {
    // the one collective of a multi-GPU run: every rank receives the monitored draws of all chains
    ST_TRY(cudaMalloc(&d_all, per_rank * W * sizeof(double)));
    ncclResult_t r = g_nccl.AllGather(src, d_all, per_rank, ncclDouble, (ncclComm_t)h->comm, st);
    if (r != ncclSuccess) {
      cleanup();
      return fail(h, WN_ECUDA, std::string("ncclAllGather: ") + g_nccl.GetErrorString(r));
    }
    src = d_all;
  }
  ST_TRY(cudaMalloc(&d_x, S * sizeof(double)));
  ST_TRY(cudaMalloc(&d_xs, S * sizeof(double)));
  ST_TRY(cudaMalloc(&d_z, S * sizeof(double)));
  ST_TRY(cudaMalloc(&d_idx, S * sizeof(unsigned)));
  ST_TRY(cudaMalloc(&d_idxs, S * sizeof(unsigned)));
  ST_TRY(cudaMalloc(&d_mean, (size_t)n_split * sizeof(double)));
  ST_TRY(cudaMalloc(&d_var, (size_t)n_split * sizeof(double)));
  ST_TRY(cudaMalloc(&d_acov, (size_t)n_split * hlen * sizeof(double)));
  ST_TRY(cudaMalloc(&d_out, ((size_t)hlen + 3) * sizeof(double)));
  ST_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_x, d_xs, d_idx, d_idxs, (int)S, 0, 64, st));
  ST_TRY(cudaMalloc(&d_tmp, tmp_bytes));
  ST_TRY(cudaFuncSetAttribute(split_chain_stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(hlen * sizeof(double))));
  const int nb = (int)((S + 255) / 256 < 4096 ? (S + 255) / 256 : 4096);
  for (int j = 0; j < dg && rc == WN_OK; ++j) {
    gather_column<<<nb, 256, 0, st>>>(src, W, n, (int)n_chains, dg, j, d_x, d_idx);
    ST_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_x, d_xs, d_idx, d_idxs, (int)S, 0, 64, st));
    rank_normalize<<<nb, 256, 0, st>>>(d_idxs, S, d_z);
    split_chain_stats<<<n_split, 256, hlen * sizeof(double), st>>>(d_z, n, hlen, split ? 1 : 0, d_mean, d_var, d_acov);
    reduce_chains<<<(hlen + 1 + 255) / 256, 256, 0, st>>>(d_acov, d_mean, d_var, n_split, hlen, d_out);
    ST_TRY(cudaMemcpyAsync(out.data(), d_out, ((size_t)hlen + 3) * sizeof(double), cudaMemcpyDeviceToHost, st));
    ST_TRY(cudaStreamSynchronize(st));
    ess_from_sums(out.data(), hlen, n_split, hlen, out[hlen], out[hlen + 1], out[hlen + 2], &ess[j], &rhat[j]);
  }
#undef ST_TRY
  cleanup();
  return rc;
}

int wn_moments_all(wn_handle* h, double* mean, double* var) {
  if (!h || !mean || !var) return fail(h, WN_EINVAL, "wn_moments_all: bad argument");
  if (!h->have_state) return fail(h, WN_ESTATE, "no state set");
  const wn_config& c = h->cfg;
  CUDA_TRY(h, cudaSetDevice(c.device));
  const int d = c.d;
  double* buf = nullptr;   // [d + 1] sums (+ chain count) | [d] means
  CUDA_TRY(h, cudaMalloc(&buf, (2 * (size_t)d + 1) * sizeof(double)));
  std::vector<double> host((size_t)d + 1);
  auto bail = [&](const std::string& msg) {
    cudaFree(buf);
    return fail(h, WN_ECUDA, msg);
  };
  // two passes (sum -> mean, then centred sum of squares), each followed by ONE all-reduce when a communicator is set
  for (int pass = 0; pass < 2; ++pass) {
    moment_sums<<<(d + 127) / 128, 128, 0, h->stream>>>(h->d_state, c.n_chains, d, pass ? buf + d + 1 : nullptr, buf);
    const double nloc = (double)c.n_chains;
    cudaError_t e = cudaMemcpyAsync(buf + d, &nloc, sizeof(double), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);      // nloc is a stack variable
    if (e != cudaSuccess) return bail(cudaGetErrorString(e));
    if (h->comm) {
      ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)d + 1, ncclDouble, ncclSum, (ncclComm_t)h->comm, h->stream);
      if (r != ncclSuccess) return bail(std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
    }
    e = cudaMemcpyAsync(host.data(), buf, host.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return bail(cudaGetErrorString(e));
    const double n = host[d];
    if (pass == 0) {
      for (int j = 0; j < d; ++j) mean[j] = host[j] / n;
      e = cudaMemcpyAsync(buf + d + 1, mean, (size_t)d * sizeof(double), cudaMemcpyHostToDevice, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) return bail(cudaGetErrorString(e));
    } else {
      for (int j = 0; j < d; ++j) var[j] = n > 1 ? host[j] / (n - 1) : 0.0;
    }
  }
  cudaFree(buf);
  return WN_OK;
}

}  // extern "C"
