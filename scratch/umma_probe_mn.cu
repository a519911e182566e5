// umma_probe_mn.cu -- tcgen05 kind::tf32 with an MN-major B operand, as the tensor-core weight gradient needs (K = positions).
// Findings on B200: for tf32, MN-major operands only work with descriptor layout type 1
// (SWIZZLE_128B_BASE32B); the other layout types silently produce zeros.  Layout validated here:
//   B element (k, n): byte = (n/32)*LBO + (k/4)*SBO + (k%4)*128 + (n%32)*4, then bits[5,7) ^= bits[7,9) (32-byte chunks XOR row&3)
//   A stays K-major, no swizzle: [K/4][M][4 floats] (rows 16 B apart, SBO = 128, LBO = plane).
// D[M x N] = sum_k A[m][k] * B[k + shiftB][n].   usage: umma_probe_mn <M> <N> <K> <shift> <which: 0 = B MN-major, 1 = A MN-major (B K-major)>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t boff) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)(boff & 7) << 49) | ((uint64_t)layout << 61);
}
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int M, int N, int K, int KA, int shiftB, int which) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  // operand X = the MN-major one (extent XN: N if which == 0, M if which == 1), operand Y = the K-major one (extent YN)
  const int XN = which ? M : N, YN = which ? N : M;
  const float* X = which ? A : B; const float* Y = which ? B : A;
  const int NB = (XN + 31) / 32;
  float* sB = reinterpret_cast<float*>(smem);                       // X: [NB][KA][32] swizzled, 1024-aligned
  float* sA = sB + (size_t)NB * KA * 32;                            // Y: [K/4][YN][4]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < NB * KA * 32; i += 128) {
    const int n32 = i & 31, k = (i >> 5) % KA, nb = (i >> 5) / KA, n = nb * 32 + n32;
    uint32_t byte = (uint32_t)nb * KA * 128 + k * 128 + n32 * 4;
    byte ^= ((byte >> 7) & 3) << 5;
    sB[byte >> 2] = n < XN ? (which ? X[(size_t)n * KA + k] : X[(size_t)k * XN + n]) : 0.f;      // A given as [M][KA], B as [KA][N]
  }
  for (int i = tid; i < (K / 4) * YN * 4; i += 128) { const int c = i & 3, m = (i >> 2) % YN, k4 = (i >> 2) / YN; sA[i] = which ? Y[(size_t)(k4 * 4 + c) * YN + m] : Y[(size_t)m * K + k4 * 4 + c]; }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((which ? 1u : 0u) << 15) | ((which ? 0u : 1u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int k8 = 0; k8 < K / 8; ++k8) {
      const uint32_t y_addr = smem_u32(sA) + k8 * 2 * (YN * 16);
      const uint32_t x_addr = smem_u32(sB) + (k8 * 8 + shiftB) * 128;
      const uint64_t dy = make_desc(y_addr, YN * 16, 128, 0, 0);
      const uint64_t dx = make_desc(x_addr, KA * 128, 512, 1, 0);
      const uint64_t da = which ? dx : dy, db = which ? dy : dx;
      const uint32_t acc = k8 > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  { uint32_t done = 0, addr = smem_u32(&mbar);
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(addr), "r"(0) : "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {       // dump all 128 lanes x N columns
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
int main(int argc, char** argv) {
  const int M = atoi(argv[1]), N = atoi(argv[2]), K = atoi(argv[3]), shiftB = argc > 4 ? atoi(argv[4]) : 0, which = argc > 5 ? atoi(argv[5]) : 0;
  const int KA = K + 24;
  std::vector<float> A((size_t)M * (which ? KA : K)), B((size_t)(which ? K : KA) * N), D(128 * (size_t)N, -7.f);
  srand(2);
  for (auto& v : A) v = (float)((rand() % 17) - 8) * 0.25f;
  for (auto& v : B) v = (float)((rand() % 13) - 6) * 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = ((size_t)((which ? M : N) + 31) / 32 * KA * 32 + (size_t)K * (which ? N : M)) * 4 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, M, N, K, KA, shiftB, which);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d N=%d K=%d shiftB=%d A-MN-major=%d: CUDA error %s\n", M, N, K, shiftB, which, cudaGetErrorString(e)); return 2; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  std::vector<double> ref((size_t)M * N);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += which ? (double)A[(size_t)m * KA + k + shiftB] * B[(size_t)k * N + n] : (double)A[(size_t)m * K + k] * B[(size_t)(k + shiftB) * N + n]; ref[(size_t)m * N + n] = s; }
  int found = 0, identity = 1; int map[128];
  for (int m = 0; m < M; ++m) {
    map[m] = -1;
    for (int l = 0; l < 128; ++l) { bool ok = true; for (int n = 0; n < N && ok; ++n) ok = fabs(D[(size_t)l * N + n] - ref[(size_t)m * N + n]) < 1e-3; if (ok) { map[m] = l; break; } }
    if (map[m] >= 0) ++found;
    if (map[m] != m) identity = 0;
  }
  // column-wise diagnosis when rows do not match: how many individual entries of lane m agree
  int good = 0; for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) good += fabs(D[(size_t)m * N + n] - ref[(size_t)m * N + n]) < 1e-3;
  printf("M=%d N=%d K=%d shiftB=%d A-MN-major=%d: rows matched %d / %d, identity lane map: %s, entries ok %d / %d", M, N, K, shiftB, which, found, M, identity ? "yes" : "no", good, M * N);
  if (!identity && found == M) { printf("  map[0,1,15,16,17,31,32,48,63] = %d %d %d %d %d %d %d %d %d", map[0], map[1], map[15], map[16], map[17], map[31], map[32], map[48], map[63]); }
  printf("\n");
  return found == M ? 0 : 1;
}
