// Chain scheduling across calls: longest chains first.
//
// The persistent kernels hand chains to thread groups from a global queue; a call ends when the LAST chain ends.  The cost
// of a transition varies by orders of magnitude between chains (funnel neck: up to 2^10 micro-steps per macro step and
// 2^12 leaves, reference WALNUTSpy_examples/funnel/mainFunnel.py:24-32), so a slow chain that leaves the queue late
// keeps a few thread groups busy long after every other SM has drained.  Successive calls on the same handle continue the same chains, and a
// chain that was slow in the previous call (still in the neck) is very likely slow in the next one: every kernel
// records the gradient evaluations of each chain (`cost`), and the next call serves the queue in descending order of
// that count (longest-processing-time-first), so the slow chains start at once and the short ones fill in behind them.
//
// Chains that share a warp (G < 32 threads per chain, W = 32 / G chains per warp) diverge: W slow chains in one warp run
// one after the other, and the warp holding the W slowest chains becomes the new tail (measured: the plain descending
// order is 1.5x SLOWER than the natural order at config 3).  The lanes of a warp take W consecutive queue entries, so
// the sorted list is dealt out in W columns: queue entry k W + j is the k-th longest chain of the j-th W-quantile --
// every warp starts with one chain of the slowest class and W - 1 shorter ones.
//
// Measured at BASELINE config 3 (262 144 chains x 10 transitions, R2P): 257 -> 250 ms per call (65 536 chains: 152 ->
// 146 ms; fixedLeapFrog 112 -> 108 ms).  What remains of the ~95 ms that the call exceeds its balanced time (1 048 576
// chains: 743 ms, i.e. 0.62 ms per 1000 chains) is the slowest chain ITSELF: started first, its ten transitions in the
// neck are one sequential dependence chain of ~10^6 micro-steps -- the call cannot end before it does.
//
// The order is a scheduling hint only: a chain's random numbers are keyed by (seed, chain id, iteration), its state is
// its own, so the draws are bit-identical with and without it (tests/test_gpu_stats.py).  WN_SCHED=0 in the environment
// turns it off (A/B measurement).  The sort (cub radix sort of n_chains 32-bit keys, ~20 us at 262 144 chains) runs on the
// handle's stream inside the timed region of the call.
#include <cstdlib>
#include <cub/cub.cuh>

#include "wn_handle.hpp"

namespace {
__global__ void iota_kernel(unsigned int* x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = (unsigned int)i;
}
// rank r of the descending list -> queue position: column r / m, row r % m of an [m, W] table (the n - m W lightest
// chains keep their place at the end)
__global__ void deal_kernel(const unsigned int* sorted, unsigned int* order, int n, int W) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int m = n / W;
  const int p = (r < m * W) ? (r % m) * W + r / m : r;
  order[p] = sorted[r];
}
}  // namespace

int wn_sched_prepare(wn_handle* h, int nslot, int chains_per_warp, const unsigned int** order) {
  *order = nullptr;
  const int n = h->cfg.n_chains;
  static const bool enabled = [] { const char* v = getenv("WN_SCHED"); return !(v && atoi(v) == 0); }();
  // every chain has its own slot from the start: the order cannot matter
  // ... and a call of a few milliseconds is not worth the sort (the previous call's duration is known after wn_sync)
  if (!enabled || n <= nslot || (h->last_ms > 0.f && h->last_ms < 5.f)) return WN_OK;
  if (!h->d_cost) {
    CUDA_TRY(h, cudaMalloc(&h->d_cost, (size_t)n * sizeof(unsigned int)));
    CUDA_TRY(h, cudaMalloc(&h->d_cost_sorted, (size_t)n * sizeof(unsigned int)));
    CUDA_TRY(h, cudaMalloc(&h->d_iota, (size_t)n * sizeof(unsigned int)));
    CUDA_TRY(h, cudaMalloc(&h->d_order, (size_t)n * sizeof(unsigned int)));
    CUDA_TRY(h, cudaMalloc(&h->d_order_sorted, (size_t)n * sizeof(unsigned int)));
    iota_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_iota, n);
    CUDA_TRY(h, cudaGetLastError());
    size_t bytes = 0;
    CUDA_TRY(h, cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, h->d_cost, h->d_cost_sorted, h->d_iota,
                                                          h->d_order_sorted, n, 0, 32, h->stream));
    CUDA_TRY(h, cudaMalloc(&h->d_sort_tmp, bytes));
    h->sort_tmp_bytes = bytes;
    h->have_cost = false;
  }
  if (!h->have_cost) return WN_OK;   // first call: queue order = chain order; the kernel records the costs
  size_t bytes = h->sort_tmp_bytes;
  CUDA_TRY(h, cub::DeviceRadixSort::SortPairsDescending(h->d_sort_tmp, bytes, h->d_cost, h->d_cost_sorted, h->d_iota,
                                                        h->d_order_sorted, n, 0, 32, h->stream));
  const int W = chains_per_warp < 1 ? 1 : chains_per_warp;
  if (W == 1) {
    *order = h->d_order_sorted;
    return WN_OK;
  }
  deal_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_order_sorted, h->d_order, n, W);
  CUDA_TRY(h, cudaGetLastError());
  *order = h->d_order;
  return WN_OK;
}

void wn_sched_free(wn_handle* h) {
  cudaFree(h->d_cost); cudaFree(h->d_cost_sorted); cudaFree(h->d_iota); cudaFree(h->d_order); cudaFree(h->d_order_sorted);
  cudaFree(h->d_sort_tmp);
  h->d_cost = h->d_cost_sorted = h->d_iota = h->d_order = h->d_order_sorted = nullptr;
  h->d_sort_tmp = nullptr;
  h->have_cost = false;
}
