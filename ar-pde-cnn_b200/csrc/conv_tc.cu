// conv_tc.cu -- circular convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05 / TMEM),
// fp32-grade accuracy through an error-compensated TF32 split.
//
// Replaces the same reference code as conv_fwd.cu (F.pad(circular) + aten::convolution of every
// nn.Conv{1,2}d(padding_mode='circular') of 2D-Burgers-SWAG/nn/denseEDcirc2d.py:34-190,:242-302 and
// 1D-*/nn/denseEDcirc.py, with the BatchNorm+ReLU prologue, nearest x2 upsampling, dense-block concat offset,
// output activation and train-mode batch statistics) for the layers whose channel counts make them a real
// contraction.  conv_fwd.cu (fp32 FFMA) stays the path for tiny-N layers and is the in-library cross-check.
//
// Formulation ("shifted window" implicit GEMM, no im2col):
//   * the (wrapped, BN+ReLU'd, optionally upsampled) input patch of a CTA is staged ONCE per channel chunk in
//     shared memory in the canonical no-swizzle K-major UMMA layout [channel/4][position][4 floats]; "position"
//     is the flattened index of a padded image whose row pitch is Wo + (kernel span), so that
//         A_tap[row q][k = channel] = patch[q + ky'*pitch + kx'][channel]
//     is the SAME buffer with the descriptor start address advanced by (ky'*pitch + kx') * 16 bytes (validated
//     by scratch/umma_probe.cu: arbitrary 16-byte row shifts are legal for SWIZZLE_NONE descriptors);
//     stride-2 convolutions keep one such plane per input parity (2x2 planes) so that taps stay unit-stride;
//   * one tcgen05.mma (M=128 output positions, K=8 channels) per (M tile, tap, 8-channel group) and split term;
//     the accumulators of up to `ntile` M tiles live in TMEM for the whole K loop;
//   * 3xTF32: x = x_hi + x_lo, w = w_hi + w_lo (both parts rounded to TF32 with cvt.rna); per K step
//         D[:, 0:2N] += A_hi * [W_hi ; W_lo]^T   (one MMA with N' = 2N: columns [0,N) hi*hi, [N,2N) hi*lo)
//         D[:, 0:N]  += A_lo * W_hi^T
//     and the epilogue adds the two column halves (lo*lo ~ 2^-22 relative is dropped);
//   * warp roles: 8 staging/epilogue warps + 1 MMA-issuing warp, mbarrier rings (2 patch slots, 2 weight
//     slots) between them; weights arrive pre-split and pre-arranged by conv_tc_wprep_kernel.
// M tiles that straddle image rows waste only the (pitch - Wo)/pitch halo positions.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"
#include <stdlib.h>

#define TC_MAX_TAPS 49
#define TC_STAGER_WARPS 8
#define TC_STAGERS (32 * TC_STAGER_WARPS)
#define TC_THREADS (TC_STAGERS + 32)

struct ConvTcArgs {
  const float* x; long long x_bs;
  const float* scale; const float* shift; long long ss_gs;
  const float* wprep; long long wprep_gs;
  float* y; double* stats;
  int cin, H, W, up_h, up_w, Hc, Wc, sy, sx;
  int relu_in, act;
  int y_ctot, y_coff, cout, Ho, Wo;
  int n_per_group;
  int pitch, min_ry, min_rx, ntile, ctas_per_sample, PP, PHP, nplanes;
  int cchunk, ncc, tchunk, ntc, ntaps, coutP, tmem_cols;
  int srcoff_bytes, patch_slot_floats, w_slot_floats, nslot_p;
  int tap_plane[TC_MAX_TAPS];
  int tap_off[TC_MAX_TAPS];
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: rows 16 B apart inside an 8-row core matrix, SBO between
// 8-row groups, LBO between the two 16-byte K chunks of one K=8 (tf32) step; version field = 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (bits 7-9, 10-12 = 2), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_rna(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------ weight preparation
// w [groups][cout][cin][ntaps]  ->  out [groups][ncc][ntc][tchunk][cchunk/4][2*coutP][4]: the exact shared-memory image
// of every (channel chunk, tap chunk) weight slot: rows [0,coutP) = tf32 hi part, rows [coutP,2coutP) = lo part.
__global__ void __launch_bounds__(256) conv_tc_wprep_kernel(const float* __restrict__ w, long long w_gs, float* __restrict__ out,
                                                            long long out_gs, int groups, int cout, int cin, int ntaps,
                                                            int cchunk, int ncc, int tchunk, int ntc, int coutP) {
  const int nc4 = cchunk >> 2, rows = 2 * coutP;
  const long long per_group = (long long)ncc * ntc * tchunk * nc4 * rows * 4;
  const long long total = per_group * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / per_group);
    long long r = idx - (long long)g * per_group;
    const int e = (int)(r & 3); r >>= 2;
    const int row = (int)(r % rows); r /= rows;
    const int c4 = (int)(r % nc4); r /= nc4;
    const int tl = (int)(r % tchunk); r /= tchunk;
    const int tc = (int)(r % ntc); const int cc = (int)(r / ntc);
    const int tap = tc * tchunk + tl, c = cc * cchunk + c4 * 4 + e, co = row < coutP ? row : row - coutP;
    float v = 0.f;
    if (co < cout && c < cin) v = __ldg(w + (long long)g * w_gs + ((long long)co * cin + c) * ntaps + tap);
    const float hi = tf32_rna(v);
    out[(long long)g * out_gs + (idx - (long long)g * per_group)] = row < coutP ? hi : tf32_rna(v - hi);
  }
}

// ------------------------------------------------------------------------------------------ main kernel
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvTcArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full_p = reinterpret_cast<uint64_t*>(smem_raw);        // [2]
  uint64_t* empty_p = full_p + 2;                                    // [2]
  uint64_t* full_w = full_p + 4;                                     // [2]
  uint64_t* empty_w = full_p + 6;                                    // [2]
  uint64_t* acc_bar = full_p + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 96);
  int* srcoff = reinterpret_cast<int*>(smem_raw + 128);
  float* patch = reinterpret_cast<float*>(smem_raw + 128 + a.srcoff_bytes);
  float* wbuf = patch + (size_t)a.nslot_p * a.patch_slot_floats;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.x / a.ctas_per_sample, b = blockIdx.x - n * a.ctas_per_sample;
  const int g = n / a.n_per_group;
  const int q0 = b * a.ntile * 128;
  const int r0 = q0 / a.pitch, qlocal0 = q0 - r0 * a.pitch;
  const int nc4 = a.cchunk >> 2, nc8 = a.cchunk >> 3, N2 = 2 * a.coutP;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(full_p + s, TC_STAGERS); mbar_init(empty_p + s, 1); mbar_init(full_w + s, TC_STAGERS); mbar_init(empty_w + s, 1); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_STAGER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // source offsets of every patch position (shared by all channels): wrap + nearest-upsample folded in
  {
    const int sxs = a.sx, PHP = a.PHP;
    for (int e = tid; e < a.nplanes * PHP; e += TC_THREADS) {
      const int pl = e / PHP, pos = e - pl * PHP;
      const int i = pos / a.pitch, j = pos - i * a.pitch;
      const int py = pl / sxs, px = pl - py * sxs;
      int cy = ((r0 + a.min_ry + i) * a.sy + py) % a.Hc; if (cy < 0) cy += a.Hc;
      int cx = ((a.min_rx + j) * a.sx + px) % a.Wc; if (cx < 0) cx += a.Wc;
      srcoff[e] = (cy >> a.up_h) * a.W + (cx >> a.up_w);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < TC_STAGER_WARPS) {
    // =============================== staging warps ===============================
    const float* xn = a.x + (long long)n * a.x_bs;
    const float* scale = a.scale ? a.scale + (long long)g * a.ss_gs : nullptr;
    const float* shift = a.scale ? a.shift + (long long)g * a.ss_gs : nullptr;
    const float4* wsrc = reinterpret_cast<const float4*>(a.wprep + (long long)g * a.wprep_gs);
    const long long HW = (long long)a.H * a.W;
    const int w_slot_f4 = a.w_slot_floats >> 2;
    int chunk = 0;
    for (int cc = 0; cc < a.ncc; ++cc) {
      const int ps = cc % a.nslot_p, pk = cc / a.nslot_p;
      if (pk > 0) mbar_wait(empty_p + ps, (pk - 1) & 1);
      float4* ph = reinterpret_cast<float4*>(patch + (size_t)ps * a.patch_slot_floats);
      float4* pl = ph + (a.patch_slot_floats >> 3);
      for (int pc = 0; pc < a.nplanes * nc4; ++pc) {
        const int plane = pc / nc4, c4 = pc - plane * nc4;
        const int c0 = cc * a.cchunk + c4 * 4;
        float sc[4], sh[4];
        const float* src[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + e;
          const bool ok = c < a.cin;
          src[e] = ok ? xn + c * HW : nullptr;
          sc[e] = (ok && scale) ? __ldg(scale + c) : 1.f;
          sh[e] = (ok && shift) ? __ldg(shift + c) : 0.f;
        }
        const int* so = srcoff + plane * a.PHP;
        float4* dh = ph + (size_t)pc * a.PP;
        float4* dl = pl + (size_t)pc * a.PP;
        for (int pos = tid; pos < a.PHP; pos += TC_STAGERS) {
          const int off = so[pos];
          float v[4], hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[e] = src[e] ? fmaf(__ldg(src[e] + off), sc[e], sh[e]) : 0.f;
            if (a.relu_in) v[e] = fmaxf(v[e], 0.f);
            hi[e] = tf32_rna(v[e]);
            lo[e] = tf32_rna(v[e] - hi[e]);
          }
          dh[pos] = make_float4(hi[0], hi[1], hi[2], hi[3]);
          dl[pos] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_async_smem();
      mbar_arrive(full_p + ps);
      for (int tc = 0; tc < a.ntc; ++tc, ++chunk) {
        const int ws = chunk & 1, wk = chunk >> 1;
        if (wk > 0) mbar_wait(empty_w + ws, (wk - 1) & 1);
        float4* dst = reinterpret_cast<float4*>(wbuf + (size_t)ws * a.w_slot_floats);
        const float4* s4 = wsrc + (size_t)chunk * w_slot_f4;
        for (int i = tid; i < w_slot_f4; i += TC_STAGERS) dst[i] = __ldg(s4 + i);
        fence_async_smem();
        mbar_arrive(full_w + ws);
      }
    }
    // =============================== epilogue ===============================
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int lanebase = (warp & 3) * 32, half = warp >> 2, cols_half = a.coutP >> 1;
    const long long HoWo = (long long)a.Ho * a.Wo;
    for (int t = 0; t < a.ntile; ++t) {
      const int q = q0 + t * 128 + lanebase + lane;
      const int oy = q / a.pitch, ox = q - oy * a.pitch;
      const bool valid = oy < a.Ho && ox < a.Wo;
      float* yp = a.y + ((long long)n * a.y_ctot + a.y_coff) * HoWo + (long long)oy * a.Wo + ox;
      const uint32_t trow = tmem + ((uint32_t)lanebase << 16) + (uint32_t)(t * N2);
      for (int c0 = half * cols_half; c0 < (half + 1) * cols_half && c0 < a.cout; c0 += 8) {
        float v1[8], v2[8];
        tmem_ld8(trow + c0, v1);
        tmem_ld8(trow + a.coutP + c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int co = c0 + j;
          float val = apply_act(v1[j] + v2[j], a.act);
          const bool st = valid && co < a.cout;
          if (st) yp[co * HoWo] = val;
          if (a.stats) {
            if (!st) val = 0.f;
            const float s1 = warp_sum(val), s2 = warp_sum(val * val);
            if (lane == 0 && co < a.cout) {
              double* sp = a.stats + 2 * (a.y_coff + co);
              atomicAdd(sp, (double)s1); atomicAdd(sp + 1, (double)s2);
            }
          }
        }
      }
    }
  } else {
    // =============================== MMA issuing warp ===============================
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_tf32(128, N2), idesc2 = umma_idesc_tf32(128, a.coutP);
      const uint32_t lbo_a = (uint32_t)a.PP * 16u, lbo_b = (uint32_t)N2 * 16u;
      const uint32_t patch_base = smem_u32(patch), w_base = smem_u32(wbuf);
      const uint32_t slot_bytes = (uint32_t)a.patch_slot_floats * 4u;
      int chunk = 0;
      for (int cc = 0; cc < a.ncc; ++cc) {
        const int ps = cc % a.nslot_p, pk = cc / a.nslot_p;
        mbar_wait(full_p + ps, pk & 1);
        const uint32_t pa_hi = patch_base + (uint32_t)ps * slot_bytes, pa_lo = pa_hi + (slot_bytes >> 1);
        for (int tc = 0; tc < a.ntc; ++tc, ++chunk) {
          const int ws = chunk & 1, wk = chunk >> 1;
          mbar_wait(full_w + ws, wk & 1);
          tc_fence_after();
          const uint32_t wb = w_base + (uint32_t)ws * (uint32_t)a.w_slot_floats * 4u;
          for (int t = 0; t < a.ntile; ++t) {
            const uint32_t d_tmem = tmem + (uint32_t)(t * N2);
            for (int tl = 0; tl < a.tchunk; ++tl) {
              const int tap = tc * a.tchunk + tl;
              const uint32_t pos0 = (uint32_t)(qlocal0 + t * 128 + a.tap_off[tap]);
              const uint32_t plane = (uint32_t)a.tap_plane[tap];
              for (int c8 = 0; c8 < nc8; ++c8) {
                const uint32_t a_off = ((plane * nc4 + 2 * c8) * (uint32_t)a.PP + pos0) * 16u;
                const uint32_t b_off = (uint32_t)((tl * nc4 + 2 * c8) * N2) * 16u;
                const uint64_t db = umma_desc(wb + b_off, lbo_b, 128);
                const uint32_t first = (chunk == 0 && tl == 0 && c8 == 0) ? 0u : 1u;
                umma_tf32(d_tmem, umma_desc(pa_hi + a_off, lbo_a, 128), db, idesc1, first);
                umma_tf32(d_tmem, umma_desc(pa_lo + a_off, lbo_a, 128), db, idesc2, 1u);
              }
            }
          }
          umma_commit(empty_w + ws);
        }
        umma_commit(empty_p + ps);
      }
      umma_commit(acc_bar);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_STAGER_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------ host side
static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static int floor_div(int a, int b) { int q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }

int arpde_tc_enabled() {
  static int cached = -1;
  if (cached < 0) { const char* e = getenv("ARPDE_TC"); cached = (e && e[0] == '0') ? 0 : 1; }
  return cached;
}

// Plan the tiling of one convolution for conv_tc_kernel.  Returns false if the layer is outside what the kernel covers.
static bool tc_plan(const ConvDesc& d, ConvTcArgs& a, size_t* smem_out) {
  if (d.dilate || d.pad_override >= 0) return false;
  if (d.stride != 1 && d.stride != 2) return false;
  const int one_d = (d.H == 1 && d.kh == 1);
  memset(&a, 0, sizeof(a));
  a.cin = d.cin; a.H = d.H; a.W = d.W; a.up_w = d.up ? 1 : 0; a.up_h = one_d ? 0 : a.up_w;
  a.Hc = d.H << a.up_h; a.Wc = d.W << a.up_w; a.sy = one_d ? 1 : d.stride; a.sx = d.stride;
  if (a.Hc % a.sy || a.Wc % a.sx) return false;
  a.Ho = a.Hc / a.sy; a.Wo = a.Wc / a.sx;
  a.cout = d.cout; a.coutP = round_up(d.cout, 16);
  if (a.coutP > 128) return false;
  a.ntaps = d.kh * d.kw;
  if (a.ntaps > TC_MAX_TAPS) return false;
  const int pad_y = d.kh == 1 ? 0 : (d.kh == 4 ? 1 : (d.kh - 1) / 2), pad_x = d.kw == 1 ? 0 : (d.kw == 4 ? 1 : (d.kw - 1) / 2);
  int min_ry = 1 << 30, max_ry = -(1 << 30), min_rx = 1 << 30, max_rx = -(1 << 30);
  int ry[8], py[8], rx[8], px[8];
  for (int k = 0; k < d.kh; ++k) { const int dy = k - pad_y; py[k] = ((dy % a.sy) + a.sy) % a.sy; ry[k] = floor_div(dy - py[k], a.sy); if (ry[k] < min_ry) min_ry = ry[k]; if (ry[k] > max_ry) max_ry = ry[k]; }
  for (int k = 0; k < d.kw; ++k) { const int dx = k - pad_x; px[k] = ((dx % a.sx) + a.sx) % a.sx; rx[k] = floor_div(dx - px[k], a.sx); if (rx[k] < min_rx) min_rx = rx[k]; if (rx[k] > max_rx) max_rx = rx[k]; }
  a.min_ry = min_ry; a.min_rx = min_rx;
  const int span_y = max_ry - min_ry, span_x = max_rx - min_rx;
  a.pitch = a.Wo + span_x;
  a.nplanes = a.sy * a.sx;
  for (int ky = 0; ky < d.kh; ++ky)
    for (int kx = 0; kx < d.kw; ++kx) {
      a.tap_plane[ky * d.kw + kx] = py[ky] * a.sx + px[kx];
      a.tap_off[ky * d.kw + kx] = (ry[ky] - min_ry) * a.pitch + (rx[kx] - min_rx);
    }
  const int max_off = span_y * a.pitch + span_x;
  const int Q = a.Ho * a.pitch, total_tiles = (Q + 127) / 128;
  const int cin8 = round_up(d.cin, 8);
  const int N2 = 2 * a.coutP;
  int ntile_max = 256 / N2; if (ntile_max < 1) ntile_max = 1; if (ntile_max > 4) ntile_max = 4;
  if (ntile_max > total_tiles) ntile_max = total_tiles;
  // search: prefer two resident CTAs per SM (<= 112 KB each), then the largest tile count / channel chunk
  const size_t budgets[2] = {112 * 1024, 220 * 1024};
  for (int bi = 0; bi < 2; ++bi) {
    for (int ntile = ntile_max; ntile >= 1; --ntile) {
      const int cands[4] = {cin8 <= 32 ? cin8 : 32, 16, 8, 0};
      for (int ci = 0; cands[ci]; ++ci) {
        const int cchunk = cands[ci];
        if (cchunk > cin8 || (ci > 0 && cchunk >= cands[0])) continue;
        const int ncc = (d.cin + cchunk - 1) / cchunk;
        const int rows = (a.pitch - 1 + ntile * 128 - 1) / a.pitch + 1 + span_y;
        const int PHP = rows * a.pitch;
        int PP = a.pitch + ntile * 128 + max_off; if (PP < PHP) PP = PHP; PP = round_up(PP, 8);
        if (PP > 16383) continue;
        int tchunk = 1;
        for (int tcand = a.ntaps; tcand >= 1; --tcand)
          if (a.ntaps % tcand == 0 && (size_t)tcand * (cchunk / 4) * N2 * 16 <= 32 * 1024) { tchunk = tcand; break; }
        const size_t w_slot = (size_t)tchunk * (cchunk / 4) * N2 * 16;
        const size_t p_slot = (size_t)2 * a.nplanes * (cchunk / 4) * PP * 16;
        const int nslot_p = ncc > 1 ? 2 : 1;
        const size_t srcoff_bytes = (size_t)round_up(a.nplanes * PHP * 4, 128);
        const size_t smem = 128 + srcoff_bytes + nslot_p * p_slot + 2 * w_slot;
        if (smem > budgets[bi]) continue;
        a.ntile = ntile; a.cchunk = cchunk; a.ncc = ncc; a.tchunk = tchunk; a.ntc = a.ntaps / tchunk;
        a.PP = PP; a.PHP = PHP; a.nslot_p = nslot_p; a.srcoff_bytes = (int)srcoff_bytes;
        a.patch_slot_floats = (int)(p_slot / 4); a.w_slot_floats = (int)(w_slot / 4);
        a.ctas_per_sample = (total_tiles + ntile - 1) / ntile;
        int cols = ntile * N2, p2 = 32; while (p2 < cols) p2 <<= 1;
        a.tmem_cols = p2;
        *smem_out = smem;
        return true;
      }
    }
  }
  return false;
}

// floats of prepared weights one group needs for this layer (upper bound over every plan tc_plan may pick)
long long arpde_conv_tc_wprep_floats(int cin, int cout, int kh, int kw) {
  return (long long)kh * kw * round_up(cin, 32) * 2 * round_up(cout, 16);
}

// heuristic: which layers go to the tensor cores (the rest stay on the fp32 FFMA kernel)
bool arpde_conv_tc_wanted(const ConvDesc& d) {
  if (!arpde_tc_enabled()) return false;
  if (d.H == 1 && d.kh == 1) return false;                 // 1D nets: tiles of 128 positions are mostly halo
  if (d.cout < 4 || d.cin * d.kh * d.kw < 32) return false;
  ConvTcArgs a; size_t smem;
  return tc_plan(d, a, &smem);
}

int arpde_conv_tc_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs,
                              cudaStream_t stream) {
  ConvTcArgs a; size_t smem;
  if (!tc_plan(d, a, &smem)) { arpde_set_error("conv_tc: layer not supported by the tensor-core kernel"); return ARPDE_ERR_UNSUPPORTED; }
  const long long per_group = (long long)a.ncc * a.ntc * a.w_slot_floats;
  ARPDE_CHECK_ARG(groups == 1 || wprep_gs >= per_group, "conv_tc: wprep group stride too small");
  const long long total = per_group * groups;
  const int blocks = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  conv_tc_wprep_kernel<<<blocks, 256, 0, stream>>>(w, w_gs, wprep, wprep_gs, groups, d.cout, d.cin, a.ntaps, a.cchunk, a.ncc, a.tchunk,
                                                   a.ntc, a.coutP);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

int arpde_conv_tc_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const float* scale,
                         const float* shift, long long ss_gs, float* y, double* stats, cudaStream_t stream) {
  ARPDE_CHECK_ARG(x && wprep && y, "conv_tc: null pointer");
  if (d.N == 0) return ARPDE_OK;
  ConvTcArgs a; size_t smem;
  if (!tc_plan(d, a, &smem)) { arpde_set_error("conv_tc: layer not supported by the tensor-core kernel"); return ARPDE_ERR_UNSUPPORTED; }
  a.x = x; a.x_bs = x_bs; a.scale = scale; a.shift = shift; a.ss_gs = ss_gs; a.wprep = wprep; a.wprep_gs = wprep_gs;
  a.y = y; a.stats = stats; a.relu_in = d.relu_in; a.act = d.act; a.y_ctot = d.y_ctot; a.y_coff = d.y_coff;
  a.n_per_group = d.n_per_group;
  static thread_local size_t smem_set = 0;
  if (smem > smem_set) {
    ARPDE_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const long long blocks = (long long)d.N * a.ctas_per_sample;
  ARPDE_CHECK_ARG(blocks < (1ll << 31), "conv_tc: grid too large");
  conv_tc_kernel<<<(unsigned)blocks, TC_THREADS, smem, stream>>>(a);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// C ABI: one circular convolution on the tensor cores (same contract as arpde_conv_fwd plus the weight scratch).
extern "C" int64_t arpde_conv_fwd_tc_scratch_floats(int cin, int cout, int kh, int kw, int groups) {
  return arpde_conv_tc_wprep_floats(cin, cout, kh, kw) * (groups > 0 ? groups : 1);
}

extern "C" int arpde_conv_fwd_tc(const float* x, int64_t x_bstride, const float* w, int64_t w_gstride, const float* scale,
                                 const float* shift, int64_t ss_gstride, float* y, double* stats, int N, int cin, int H, int W,
                                 int cout, int kh, int kw, int stride, int up, int relu_in, int act, int y_ctot, int y_coff,
                                 int n_per_group, float* scratch, int64_t scratch_floats, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1;
  d.N = N; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride; d.up = up;
  d.relu_in = relu_in; d.act = act; d.y_ctot = y_ctot; d.y_coff = y_coff; d.n_per_group = n_per_group > 0 ? n_per_group : (N > 0 ? N : 1);
  ARPDE_CHECK_ARG(x && w && y && scratch, "conv_fwd_tc: null pointer");
  ARPDE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_fwd_tc: scale/shift must both be given or both be null");
  ARPDE_CHECK_ARG(N >= 0 && cin > 0 && cout > 0 && H > 0 && W > 0, "conv_fwd_tc: bad dims");
  ARPDE_CHECK_ARG(y_coff >= 0 && y_coff + cout <= y_ctot, "conv_fwd_tc: channel window [%d,%d) outside y_ctot=%d", y_coff, y_coff + cout, y_ctot);
  if (N == 0) return ARPDE_OK;
  ARPDE_CHECK_ARG(N % d.n_per_group == 0, "conv_fwd_tc: N=%d not a multiple of n_per_group=%d", N, d.n_per_group);
  const int groups = N / d.n_per_group;
  const long long per_group = arpde_conv_tc_wprep_floats(cin, cout, kh, kw);
  ARPDE_CHECK_ARG(scratch_floats >= per_group * groups, "conv_fwd_tc: scratch too small (%lld < %lld floats)", (long long)scratch_floats,
                  per_group * groups);
  int rc = arpde_conv_tc_prep_launch(d, w, w_gstride, groups, scratch, per_group, stream);
  if (rc) return rc;
  return arpde_conv_tc_launch(d, x, x_bstride, scratch, per_group, scale, shift, ss_gstride, y, stats, stream);
}

// Diagnostic (host only, no GPU needed): the tiling conv_tc_kernel would use for a layer.  out[0..23] = pitch, min_ry, min_rx, ntile,
// ctas_per_sample, PP, PHP, nplanes, cchunk, ncc, tchunk, ntc, ntaps, coutP, tmem_cols, nslot_p, smem bytes, Ho, Wo, Hc, Wc, sy, sx,
// wanted-by-heuristic; out[32+t] = tap plane, out[32+49+t] = tap offset.  Returns ARPDE_ERR_UNSUPPORTED if not plannable.
extern "C" int arpde_conv_tc_plan(int cin, int H, int W, int cout, int kh, int kw, int stride, int up, int32_t* out) {
  ARPDE_CHECK_ARG(out, "conv_tc_plan: null output");
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1; d.N = 1; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride;
  d.up = up; d.relu_in = 0; d.act = 0; d.y_ctot = cout; d.y_coff = 0; d.n_per_group = 1;
  ConvTcArgs a; size_t smem = 0;
  if (!tc_plan(d, a, &smem)) { arpde_set_error("conv_tc_plan: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  const int v[24] = {a.pitch, a.min_ry, a.min_rx, a.ntile, a.ctas_per_sample, a.PP, a.PHP, a.nplanes, a.cchunk, a.ncc, a.tchunk, a.ntc,
                     a.ntaps, a.coutP, a.tmem_cols, a.nslot_p, (int)smem, a.Ho, a.Wo, a.Hc, a.Wc, a.sy, a.sx, arpde_conv_tc_wanted(d) ? 1 : 0};
  for (int i = 0; i < 24; ++i) out[i] = v[i];
  for (int t = 0; t < TC_MAX_TAPS; ++t) { out[32 + t] = t < a.ntaps ? a.tap_plane[t] : 0; out[32 + TC_MAX_TAPS + t] = t < a.ntaps ? a.tap_off[t] : 0; }
  return ARPDE_OK;
}
