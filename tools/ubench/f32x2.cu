// Microbenchmark (sm_100a): issue/throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2), alone and interleaved with
// integer ALU work.  Prints G-FMA/s and the ratio; used to decide whether packing two call variants into f32x2 pays.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int MODE>
__global__ void __launch_bounds__(256, 4) k(float *out, int iters, float s, uint32_t m) {
  float2 a[8];
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i); u[i] = threadIdx.x * 17 + i; }
  const float2 b = make_float2(s, s * 0.5f), c = make_float2(0.25f, 0.125f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0 || MODE == 2) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }       // 2 FFMA
      if (MODE == 1 || MODE == 3) a[i] = __ffma2_rn(a[i], b, c);                                              // 1 FFMA2
      if (MODE == 2 || MODE == 3) { u[i] = (u[i] ^ m) + (u[i] >> 3); u[i] = (u[i] & m) | (u[i] << 1); }       // ~4 ALU ops
    }
  }
  float r = 0; uint32_t q = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) { r += a[i].x + a[i].y; q ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + (float)q;
}

template <int MODE> double run(float *d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 4, 256>>>(d, 16, 0.999f, 0x5555u);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 256>>>(d, iters, 0.999f, 0x5555u);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  float *d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  const int iters = 20000;
  const double fma = 148.0 * 4 * 256 * (double)iters * 16;   // FMAs per launch
  double t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
  printf("scalar FFMA          : %8.3f ms  %7.2f TFLOP/s\n", t0, 2 * fma / t0 * 1e-9);
  printf("packed FFMA2         : %8.3f ms  %7.2f TFLOP/s  (x%.2f)\n", t1, 2 * fma / t1 * 1e-9, t0 / t1);
  printf("scalar FFMA + 2xALU4 : %8.3f ms  %7.2f TFLOP/s\n", t2, 2 * fma / t2 * 1e-9);
  printf("packed FFMA2 + ALU   : %8.3f ms  %7.2f TFLOP/s  (x%.2f)\n", t3, 2 * fma / t3 * 1e-9, t2 / t3);
  return 0;
}
