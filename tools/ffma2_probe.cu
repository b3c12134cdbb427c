// Microbenchmark: does FFMA2 (packed fp32x2 FMA, sm_100) raise FP32 FMA throughput or only save issue slots?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
      else a[i] = __ffma2_rn(a[i], m, c);
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters); else k<1><<<148 * 8, 256>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * 8 * 256 * iters * 16.0;
    printf("%s: %.3f ms, %.1f TFMA/s (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
  }
  return 0;
}
