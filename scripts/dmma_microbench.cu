// FP64 tensor-core (DMMA) throughput on this GPU next to the FP64 FMA pipe: decides whether the logistic-regression
// gradient should be written with mma.sync f64 fragments.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dmma884(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma16816(double* out, int iters) {
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
  double c[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fma(double* out, int iters, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: DMMA and DFMA in the same warp (do they share the pipe?)
__global__ void k_mixed(double* out, int iters, double a2, double b2) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c[4][2], x[8];
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a2, b2);
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int wps : {4, 8, 16, 32}) {     // warps per SM
    const int nt = 32 * (wps > 32 ? 32 : wps), blocks = sms;
    float ms = timeit([&] { k_dmma884<<<blocks, nt>>>(out, iters); });
    double fl = 2.0 * 256 * 8 * (double)iters * (nt / 32) * blocks;
    printf("dmma m8n8k4    warps/SM %2d: %.2f TFLOP/s\n", wps, fl / ms * 1e-9);
    ms = timeit([&] { k_dmma16816<<<blocks, nt>>>(out, iters / 4); });
    fl = 2.0 * 16 * 8 * 16 * 4 * (double)(iters / 4) * (nt / 32) * blocks;
    printf("dmma m16n8k16  warps/SM %2d: %.2f TFLOP/s\n", wps, fl / ms * 1e-9);
    ms = timeit([&] { k_fma<<<blocks, nt>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 8 * (double)iters * nt * blocks;
    printf("dfma           warps/SM %2d: %.2f TFLOP/s\n", wps, fl / ms * 1e-9);
    ms = timeit([&] { k_mixed<<<blocks, nt>>>(out, iters, 1.0000001, 1e-9); });
    fl = (2.0 * 256 * 4 / 32 + 2.0 * 8) * (double)iters * nt * blocks;
    printf("mixed 4dmma+8dfma warps/SM %2d: %.2f TFLOP/s (sum)\n", wps, fl / ms * 1e-9);
  }
  return 0;
}
