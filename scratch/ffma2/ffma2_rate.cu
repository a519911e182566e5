// B200: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a2.x, b2.x); acc[i].y = fmaf(acc[i].y, a2.y, b2.y); }
        else acc[i] = __ffma2_rn(acc[i], a2, b2);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters); else k<1><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 8 * 256 * (double)iters * 8 * 8 * 2;
      if (rep) printf("%s: %.3f ms, %.1f TFLOP/s (fp32 FMA = 2 flop)\n", mode ? "FFMA2 (packed f32x2)" : "FFMA (scalar)", ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
