// Probe: achievable HBM read bandwidth of the conv_h converter's access pattern (148 CTAs x 256 threads, per thread batches of 4 positions x 8
// channel planes, scalar 4-byte loads, channel stride 4 KB) against float4 loads and a deeper batch.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(const float* __restrict__ x, float* out, int N, int HW, int cin) {
  // item = (sample, 8 rows of 32); 64 channels; positions 256 per item
  const int items = N * 4, tid = threadIdx.x;
  float acc = 0.f;
  for (int it = blockIdx.x * items / gridDim.x; it < (blockIdx.x + 1) * items / gridDim.x; ++it) {
    const int n = it >> 2, b = it & 3;
    const float* xn = x + (long long)n * cin * HW + b * 256;
    if (MODE == 0) {            // 2 batches of 4 octets: lane = position, 8 channel loads per octet
      for (int h = 0; h < 2; ++h) {
        float v[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) v[i][c] = __ldcg(xn + (long long)((h * 4 + i) * 8 + c) * HW + tid);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc += v[i][c];
      }
    } else if (MODE == 1) {     // one batch of 8 octets (64 loads in flight per thread)
      float v[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) v[i][c] = __ldcg(xn + (long long)(i * 8 + c) * HW + tid);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc += v[i][c];
    } else {                    // float4: thread = (channel group, 4 positions): 64 threads cover a channel's 256 positions, 4 channels per pass, 16 passes
      float4 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(xn + (long long)(i * 4 + (tid >> 6)) * HW) + (tid & 63));
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
int main() {
  const int N = 1920, HW = 1024, cin = 64;
  float *x, *out; size_t n = (size_t)N * cin * HW;
  cudaMalloc(&x, n * 4 * 2); cudaMalloc(&out, 4); cudaMemset(x, 0, n * 4 * 2);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode)
    for (int grid : {148, 296, 592}) {
      float ms = 0;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) {
          const float* xi = x + (i & 1) * n;
          if (mode == 0) probe<0><<<grid, 256>>>(xi, out, N, HW, cin);
          else if (mode == 1) probe<1><<<grid, 256>>>(xi, out, N, HW, cin);
          else probe<2><<<grid, 256>>>(xi, out, N, HW, cin);
        }
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      }
      printf("mode %d grid %d: %.3f ms per launch, %.0f GB/s\n", mode, grid, ms / 10, n * 4 / (ms / 10) / 1e6);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
