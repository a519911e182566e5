// umma_loop.cu -- the conv_tc issuer loop nest in isolation (no staging / epilogue warps): cycles per MMA pair.
// usage: umma_loop <ntile> <tchunk> <nc8> <coutP> <PP> <nchunks> <mode>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
struct P { int ntile, tchunk, nc8, coutP, PP, nchunks, mode, pad; int tap_aoff[49]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t idesc_tf32(int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }
__global__ void __launch_bounds__(128) loopk(long long* out, const __grid_constant__ P a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 1023);
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred));
    const bool leader = pred != 0;
    const int N2 = 2 * a.coutP, nc8 = a.nc8, nc4 = 2 * a.nc8;
    const uint32_t idesc1 = idesc_tf32(N2), idesc2 = idesc_tf32(a.coutP);
    const uint32_t PP = (uint32_t)a.PP;
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t pa_hi = (smem_u32(smem) >> 4) | (PP << 16), pa_lo = pa_hi + (uint32_t)(nc4 * a.PP);
    const uint32_t wb0 = ((smem_u32(smem) + 40 * 1024) >> 4) | ((uint32_t)N2 << 16);
    long long t0 = clock64();
    long long npairs = 0;
    for (int chunk = 0; chunk < a.nchunks; ++chunk) {
      const uint32_t wb = wb0;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        const uint64_t hi64 = (uint64_t)desc_hi << 32;
        const uint32_t first_chunk = chunk == 0 ? 0u : 1u;
        if (a.mode <= 1) {
        for (int t = 0; t < a.ntile; ++t) {
          const uint32_t d_tmem = tmem + (uint32_t)(t * N2);
          uint64_t db = hi64 | wb;
          for (int tl = 0; tl < a.tchunk; ++tl) {
            const uint32_t al = (uint32_t)a.tap_aoff[tl] + (uint32_t)(t * 128);
            uint64_t da_h = hi64 | (pa_hi + al), da_l = hi64 | (pa_lo + al);
            uint32_t acc = (tl == 0) ? first_chunk : 1u;
            if (a.mode == 0) {
#pragma unroll 2
              for (int c8 = 0; c8 < nc8; ++c8) {
                umma_tf32(d_tmem, da_h, db, idesc1, acc);
                umma_tf32(d_tmem + (uint32_t)a.coutP, da_l, db, idesc2, 1u);
                acc = 1u;
                da_h += 2 * PP; da_l += 2 * PP; db += 2 * (uint32_t)N2;
              }
            } else {       // mode 1: only the first MMA of each pair (N' = 2N)
              for (int c8 = 0; c8 < nc8; ++c8) {
                umma_tf32(d_tmem, da_h, db, idesc1, acc);
                acc = 1u;
                da_h += 2 * PP; db += 2 * (uint32_t)N2;
              }
            }
          }
        }
        } else if (a.mode == 2) {          // prefetch the next tap offset one iteration ahead
        for (int t = 0; t < a.ntile; ++t) {
          const uint32_t d_tmem = tmem + (uint32_t)(t * N2);
          uint64_t db = hi64 | wb;
          uint32_t al_next = (uint32_t)a.tap_aoff[0];
          for (int tl = 0; tl < a.tchunk; ++tl) {
            const uint32_t al = al_next + (uint32_t)(t * 128);
            al_next = (uint32_t)a.tap_aoff[tl + 1 < a.tchunk ? tl + 1 : 0];
            uint64_t da_h = hi64 | (pa_hi + al), da_l = hi64 | (pa_lo + al);
            uint32_t acc = (tl == 0) ? first_chunk : 1u;
            for (int c8 = 0; c8 < nc8; ++c8) {
              umma_tf32(d_tmem, da_h, db, idesc1, acc);
              umma_tf32(d_tmem + (uint32_t)a.coutP, da_l, db, idesc2, 1u);
              acc = 1u;
              da_h += 2 * PP; da_l += 2 * PP; db += 2 * (uint32_t)N2;
            }
          }
        }
        } else if (a.mode == 3) {          // arithmetic tap offsets (kw x kw taps, pitch), no table
        const int kw = a.tchunk == 9 ? 3 : (a.tchunk == 25 ? 5 : a.tchunk);
        const int kh = a.tchunk / kw;
        for (int t = 0; t < a.ntile; ++t) {
          const uint32_t d_tmem = tmem + (uint32_t)(t * N2);
          uint64_t db = hi64 | wb;
          uint32_t acc = first_chunk;
          uint32_t rowoff = (uint32_t)(t * 128);
          for (int ky = 0; ky < kh; ++ky, rowoff += 34) {
            uint64_t da_h = hi64 | (pa_hi + rowoff), da_l = hi64 | (pa_lo + rowoff);
            for (int kx = 0; kx < kw; ++kx) {
              uint64_t xh = da_h, xl = da_l;
              for (int c8 = 0; c8 < nc8; ++c8) {
                umma_tf32(d_tmem, xh, db, idesc1, acc);
                umma_tf32(d_tmem + (uint32_t)a.coutP, xl, db, idesc2, 1u);
                acc = 1u;
                xh += 2 * PP; xl += 2 * PP; db += 2 * (uint32_t)N2;
              }
              da_h += 1; da_l += 1;
            }
          }
        }
        }
      }
      npairs += (long long)a.ntile * a.tchunk * nc8;
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    uint32_t done = 0, addr = smem_u32(&mbar);
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(addr), "r"(0) : "memory");
    long long t1 = clock64();
    if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = npairs; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main(int argc, char** argv) {
  P p; p.ntile = atoi(argv[1]); p.tchunk = atoi(argv[2]); p.nc8 = atoi(argv[3]); p.coutP = atoi(argv[4]); p.PP = atoi(argv[5]); p.nchunks = atoi(argv[6]); p.mode = argc > 7 ? atoi(argv[7]) : 0;
  const int pitch = 34;
  for (int t = 0; t < 49; ++t) p.tap_aoff[t] = (t / 3) * pitch + (t % 3);
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(loopk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) loopk<<<148, 128, 200 * 1024>>>(d, p);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("ntile=%d tchunk=%d nc8=%d coutP=%d PP=%d nchunks=%d mode=%d: %s  %.1f clk/pair (%lld pairs)\n", p.ntile, p.tchunk, p.nc8, p.coutP, p.PP, p.nchunks, p.mode,
         cudaGetErrorString(e), (double)h[0] / (double)h[1], h[1]);
  return 0;
}
