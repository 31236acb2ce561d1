// FP64 pipe micro-benchmarks (tuning aid, not part of the product): what DFMA rate is attainable
// for (A) constant-operand FMAs, (B) 3-register-operand FMAs, (D) the leapfrog instruction mix.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int E>
__global__ void k(double* out, int iters, double a, double b) {
  double x[E], y[E], z[E], s[E];
  for (int e = 0; e < E; ++e) { x[e] = threadIdx.x + e; y[e] = 1.0 + 1e-9 * (threadIdx.x + e); z[e] = 1e-7 * e; s[e] = 1.0 + 0.01 * e; }
  double acc = 0, ke = 0;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = fma(x[e], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = fma(x[e], y[e], z[e]);
    } else if (MODE == 2) {
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = fma(y[e], z[e], x[e]);
    } else {
      // leapfrog mix: x=q, y=v, z=g
      double a0 = 0, a1 = 0, k0 = 0, k1 = 0;
#pragma unroll
      for (int e = 0; e < E; ++e) { y[e] = fma(a, z[e], y[e]); x[e] = fma(b, y[e], x[e]); }
#pragma unroll
      for (int e = 0; e < E; ++e) { z[e] = -(x[e] * s[e]); if (e & 1) a1 = fma(x[e], z[e], a1); else a0 = fma(x[e], z[e], a0); }
#pragma unroll
      for (int e = 0; e < E; ++e) { y[e] = fma(a, z[e], y[e]); if (e & 1) k1 = fma(y[e], y[e], k1); else k0 = fma(y[e], y[e], k0); }
      if (MODE == 3) { acc += a0 + a1; ke += k0 + k1; }
      else { acc = fmax(acc, a0 + a1 + k0 + k1); }
    }
  }
  double r = acc + ke;
  for (int e = 0; e < E; ++e) r += x[e] + y[e] + z[e];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE, int E>
void run(const char* name, int threads, int blocks_per_sm, int fp64_per_iter) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int blocks = sms * blocks_per_sm, iters = 1 << 15;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); k<MODE, E><<<blocks, threads>>>(out, iters, 1e-3, 2e-3); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  double ops = (double)fp64_per_iter * iters * blocks * threads;   // FP64 instructions (per thread)
  printf("%-34s E=%2d thr=%3d blk/SM=%d : %.2f T FP64-instr/s  (x2 = %.2f TFLOP/s if all FMA)\n", name, E, threads, blocks_per_sm, ops / best / 1e9, 2 * ops / best / 1e9);
  cudaFree(out);
}
int main() {
  run<0, 8>("A fma(x,const,const)", 256, 8, 8);
  run<1, 8>("B x=fma(x,y,z) 3 regs", 256, 4, 8);
  run<2, 8>("C x=fma(y,z,x) 3 regs", 256, 4, 8);
  run<1, 8>("B x=fma(x,y,z) 3 regs", 128, 2, 8);
  run<3, 8>("D leapfrog mix (6/elem + 4)", 128, 2, 6 * 8 + 4);
  run<3, 8>("D leapfrog mix", 128, 3, 6 * 8 + 4);
  run<3, 8>("D leapfrog mix", 128, 4, 6 * 8 + 4);
  run<3, 16>("D leapfrog mix", 64, 4, 6 * 16 + 4);
  run<3, 16>("D leapfrog mix", 64, 8, 6 * 16 + 4);
  run<3, 4>("D leapfrog mix", 256, 4, 6 * 4 + 4);
  return 0;
}
