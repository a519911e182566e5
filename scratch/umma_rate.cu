// umma_rate.cu -- tcgen05.mma issue-rate microbenchmark: cycles per MMA (M=128, K=8 tf32) for several N and smem layouts.
// usage: umma_rate <layout 0=none 1=sw32 2=sw64 3=sw128> <N> <n_mma> <distinct 0|1>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512) rate(long long* out, int N, int n_mma, int layout, int distinct, int shiftA, int lboA, int boff) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ __align__(8) uint64_t mbar2;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 1023);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar2))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    // layout_type field (bits 61-63): none 0, sw128 2, sw64 4, sw32 6
    const uint32_t lt = layout == 0 ? 0u : (layout == 1 ? 6u : (layout == 2 ? 4u : 2u));
    uint32_t lbo, sbo;
    if (layout == 0) { lbo = 4096; sbo = 128; }             // [k chunk][row][16B], plane = 256 rows
    else if (layout == 1) { lbo = 16; sbo = 256; }          // rows of 32 B
    else if (layout == 2) { lbo = 16; sbo = 512; }
    else { lbo = 16; sbo = 1024; }
    const uint32_t hi = ((sbo >> 4) & 0x3FFF) | (1u << 14) | (lt << 29);
    const uint32_t a0 = smem_u32(smem) + (uint32_t)shiftA, b0 = smem_u32(smem) + (uint32_t)boff;
    const uint64_t da = ((uint64_t)hi << 32) | ((a0 >> 4) & 0x3FFF) | ((((lboA ? (uint32_t)lboA : lbo) >> 4) & 0x3FFF) << 16);
    const uint64_t db = ((uint64_t)hi << 32) | ((b0 >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16);
    const uint64_t da2 = da + (distinct ? 512 : 0), db2 = db + (distinct ? 512 : 0);
#define MMA(DA, DB) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(DA), "l"(DB), "r"(idesc), "r"(1u) : "memory")
    t0 = clock64();
    if (distinct < 2) {
      for (int i = 0; i < n_mma; i += 8) {
        if (pred) { MMA(da, db); MMA(da2, db); MMA(da, db2); MMA(da2, db2); MMA(da, db); MMA(da2, db); MMA(da, db2); MMA(da2, db2); }
      }
    } else if (distinct == 2) {        // dependent in-place descriptor updates, as the conv kernel does
      if (pred) {
        uint64_t xa = da, xb = db;
        for (int i = 0; i < n_mma; i += 2) {
          MMA(xa, xb); MMA(xa + 8192, xb);
          xa += 32; xb += 64;
          if ((i & 63) == 62) { xa = da; xb = db; }
        }
      }
    } else if (distinct == 4) {         // alternate MMA shapes like the conv kernel: N' = 2N into [0,2N), then N into [N,2N)
      const uint32_t idesc_h = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((N / 2) >> 3) << 17) | ((128u >> 4) << 24);
#define MMAX(T, DA, DB, ID) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(T), "l"(DA), "l"(DB), "r"(ID), "r"(1u) : "memory")
      if (pred) {
        for (int i = 0; i < n_mma; i += 2) { MMAX(tmem, da, db, idesc); MMAX(tmem + N / 2, da2, db, idesc_h); }
      }
    } else if (distinct == 6) {         // truly distinct operands: A walks 16 x 4 KB blocks, B walks blocks of N*32 B (no reuse between MMAs)
      if (pred) {
        const uint32_t bstep = (uint32_t)(N * 32) >> 4;
        const int nb = (90 * 1024) / (N * 32) < 16 ? (90 * 1024) / (N * 32) : 16;
        for (int i = 0; i < n_mma; i += 16) {
          uint64_t xa = da, xb = db;
          for (int j = 0; j < 16; ++j) { MMA(xa, xb); xa += 256; xb = (j % nb == nb - 1) ? db : xb + bstep; }
        }
      }
    } else if (distinct == 5) {         // same shape, alternating accumulator column ranges
      if (pred) {
        for (int i = 0; i < n_mma; i += 2) { MMAX(tmem, da, db, idesc); MMAX(tmem + N, da2, db, idesc); }
      }
    } else {                            // whole loop under the elected lane, constant descriptors
      if (pred) {
        for (int i = 0; i < n_mma; i += 2) { MMA(da, db); MMA(da2, db2); }
      }
    }
    if (pred) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0, addr = smem_u32(&mbar);
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(addr), "r"(0) : "memory");
    t1 = clock64();
    if (pred && blockIdx.x == 0) out[0] = t1 - t0;
    if (pred) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&mbar2)) : "memory");
  } else if (warp >= 4) {
    uint32_t done = 0, addr = smem_u32(&mbar2);
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(addr), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}
int main(int argc, char** argv) {
  int nthreads = getenv("NT") ? atoi(getenv("NT")) : 128; int layout = atoi(argv[1]), N = atoi(argv[2]), n = atoi(argv[3]), distinct = argc > 4 ? atoi(argv[4]) : 1, shiftA = argc > 5 ? atoi(argv[5]) : 0, lboA = argc > 6 ? atoi(argv[6]) : 0;
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) rate<<<148, nthreads, 200 * 1024>>>(d, N, n, layout, distinct, shiftA, lboA, getenv("BOFF") ? atoi(getenv("BOFF")) : 96 * 1024);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("shiftA=%d lboA=%d layout=%d N=%3d distinct=%d: %s  %.1f clk/MMA (floor %d)\n", shiftA, lboA, layout, N, distinct, cudaGetErrorString(e), (double)h / n, 128 * N / 256);
  return 0;
}
