// umma_probe.cu -- single-CTA tcgen05 TF32 GEMM probe: validates the no-swizzle K-major smem descriptor
// (LBO/SBO meaning), arbitrary 16-byte row shifts of the A start address, the instruction descriptor and the
// TMEM 32x32b load mapping used by conv_tc.cu.   usage: umma_probe <mode 0|1> <shift rows> <N> <K>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int N, int K, int rowsA, int shift, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  float* sA = reinterpret_cast<float*>(smem);                 // [K/4][rowsA][4]
  float* sB = sA + (size_t)(K / 4) * rowsA * 4;              // [K/4][N][4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (K / 4) * rowsA * 4; i += 128) {
    int c = i & 3, r = (i >> 2) % rowsA, p = (i >> 2) / rowsA;
    sA[i] = A[(size_t)r * K + p * 4 + c];
  }
  for (int i = tid; i < (K / 4) * N * 4; i += 128) {
    int c = i & 3, r = (i >> 2) % N, p = (i >> 2) / N;
    sB[i] = B[(size_t)r * K + p * 4 + c];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t planeA = rowsA * 16, planeB = N * 16;
    for (int k8 = 0; k8 < K / 8; ++k8) {
      uint32_t a_addr = smem_u32(sA) + (2 * k8) * planeA + shift * 16;
      uint32_t b_addr = smem_u32(sB) + (2 * k8) * planeB;
      uint64_t da = mode == 0 ? make_desc(a_addr, planeA, 128) : make_desc(a_addr, 128, planeA);
      uint64_t db = mode == 0 ? make_desc(b_addr, planeB, 128) : make_desc(b_addr, 128, planeB);
      uint32_t acc = k8 > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  {
    uint32_t done = 0, addr = smem_u32(&mbar);
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(addr), "r"(0) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0, shift = argc > 2 ? atoi(argv[2]) : 0, N = argc > 3 ? atoi(argv[3]) : 32,
      K = argc > 4 ? atoi(argv[4]) : 16;
  const int rowsA = 128 + 40;
  std::vector<float> A((size_t)rowsA * K), B((size_t)N * K), D(128 * (size_t)N, -1.f);
  srand(1);
  for (auto& v : A) v = (float)((rand() % 17) - 8) * 0.25f;     // exactly representable in tf32
  for (auto& v : B) v = (float)((rand() % 13) - 6) * 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  size_t smem = ((size_t)(K / 4) * rowsA * 4 + (size_t)(K / 4) * N * 4) * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, N, K, rowsA, shift, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("probe mode=%d shift=%d N=%d K=%d: CUDA error %s\n", mode, shift, N, K, cudaGetErrorString(e)); return 2; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)(r + shift) * K + k] * B[(size_t)n * K + k];
      double err = fabs(ref - D[(size_t)r * N + n]);
      if (err > 1e-3) ++bad;
      if (err > maxerr) maxerr = err;
    }
  printf("probe mode=%d shift=%d N=%d K=%d: max err %.4g, mismatches %d / %d  -> %s\n", mode, shift, N, K, maxerr, bad, 128 * N,
         bad == 0 ? "OK" : "FAIL");
  if (bad) { printf("  D[0][0..3] = %g %g %g %g ; D[1][0] = %g ; D[32][0] = %g\n", D[0], D[1], D[2], D[3], D[N], D[32 * (size_t)N]); }
  return bad ? 1 : 0;
}
