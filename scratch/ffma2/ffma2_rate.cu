// B200: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) rate.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
//   mode 0/1: accumulator chains with two constant operands (operand reuse cache always hits)
//   mode 2/3: outer-product form acc[i][j] += x[i] * y[j] (8 x 8 accumulators, the shape of a register-tiled convolution): every
//             instruction reads three distinct register (pairs)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float s = 0.f;
  if (MODE < 2) {
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
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  } else {
    float2 acc[8][4], x[8], y[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(a + i * 1e-3f, a - i * 1e-3f);
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(threadIdx.x * 1e-3f + i, j * 0.5f); }
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = make_float2(b + j * 1e-3f, b - j * 1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (MODE == 2) { acc[i][j].x = fmaf(x[i].x, y[j].x, acc[i][j].x); acc[i][j].y = fmaf(x[i].y, y[j].y, acc[i][j].y); }
            else acc[i][j] = __ffma2_rn(x[i], y[j], acc[i][j]);
          }
        // perturb the operands so that nothing is loop-invariant
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i].x = x[i].x * 0.999f;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s += acc[i][j].x + acc[i][j].y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 4096;
  const char* names[4] = {"FFMA  chains, constant operands", "FFMA2 chains, constant operands", "FFMA  outer product 8 x 4 pairs", "FFMA2 outer product 8 x 4 pairs"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      else if (mode == 1) k<1><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      else if (mode == 2) k<2><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      else k<3><<<148 * 8, 256>>>(out, 0.999f, 0.001f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 8 * 256 * (double)iters * 64 * 2;
      if (rep) printf("%s: %.3f ms, %.1f TFLOP/s (fp32 FMA = 2 flop)\n", names[mode], ms, 2 * fma / ms / 1e9);
    }
  }
  return 0;
}
