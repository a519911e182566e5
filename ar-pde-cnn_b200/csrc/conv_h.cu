// conv_h.cu -- eval-mode stride-1 convolutions of the decoder on the tensor cores with fp16 two-term operands: the 1x1 convolution of a
// transition / LastTransUp (BatchNorm, ReLU, conv1x1: 2D-Burgers-SWAG/nn/denseEDcirc2d.py:103-108,170-175) and the 3x3 convolution that
// follows the nearest x2 upsampling (BatchNorm, ReLU, UpsamplingNearest2d, circular conv3x3: :176-183), the latter folded to ONE 3x3
// convolution on the low-resolution grid with 4 * cout parity-major output channels (same fold as conv_tc.cu: an output pixel of parity
// (py, px) sees a 2x2 block of low-resolution pixels, its weights are sums of the original taps).
//
// Why not conv_tc.cu: the 3xTF32 form moves every activation through shared memory three times (raw ring, hi / lo patch, K = 8 operand
// reads at 4 KB per MMA) and is bound by that, not by the tensor pipe (0.24 + 0.43 ms at 1920 samples for 0.115 + 0.115 ms of HBM traffic).
// Here, as in conv_z.cu / conv_stem.cu: x = x_hi + 2^-11 x_lo' with x_hi = fp16(x), x_lo' = fp16((x - x_hi) * 2^11) (22 mantissa bits, exact
// products, fp32 accumulation in TMEM), per K = 16 step
//     D[:, 0:2N] += A_hi * [W_hi ; W_lo']^T        D[:, N:2N] += A_lo' * W_hi^T          y = D[:, 0:N] + 2^-11 D[:, N:2N]
// i.e. half the MMAs and half the shared-memory bytes per MAC, and no raw ring: 8 converter warps read the fp32 NCHW input straight
// from global memory (a lane = one position, 8 channel planes -> one 16-byte K unit), apply the eval-mode BatchNorm affine + ReLU, split
// and store the K-major no-swizzle operand image [hi|lo][channel octet][padded position][16 B] of the NEXT item into the second patch
// buffer while the MMAs of the current item run.  A filter tap is the same buffer with the descriptor start advanced by
// (ky * pitch + kx) * 16 B (the shifted-window formulation of conv_tc.cu), the second K unit of an MMA is the next channel octet (LBO = one
// octet plane).  The prepared weights of a SWAG sample stay resident in shared memory.
//
// Persistent, 13 warps, one CTA per SM.  Work item = (sample, RI output rows); M tiles of 128 padded positions; accumulators: a ring of
// 2-4 tile slots in TMEM drained by 4 epilogue warps (folded layer: both x parities of a channel are one float2 store, a warp writes 256
// contiguous bytes per instruction).  13 warps: with the 4-warp register allocation granularity 16 x 32 x 128 registers fit the file.
// Roofline: HBM (conv1x1 64->32 at 32x32: 393 KB per sample in + out; up-conv 32->16: 131 KB in, 262 KB out).
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"
#include <cuda_fp16.h>
#include <mutex>
#include "tc_ptx.cuh"

#define CH_CONV_WARPS 8
#define CH_EPI_WARPS 4
#define CH_WARP_EPI0 CH_CONV_WARPS                        // a multiple of 4: epilogue warp e reads TMEM lane quarter e % 4
#define CH_WARP_MMA (CH_CONV_WARPS + CH_EPI_WARPS)
#define CH_THREADS (32 * (CH_WARP_MMA + 1))
#define CH_CONV_THREADS (32 * CH_CONV_WARPS)
#ifndef CH_PIPELINED
#define CH_PIPELINED 0
#endif
#define CH_KP 8                                           // (position, channel octet) units a converter thread holds in registers
#define CH_MAX_TSLOTS 4
#define CH_HEADER 256
#ifndef CH_SLEEP
#define CH_SLEEP 0                                        // nanosleep back-off of the mbarrier waits (0 = none: 13 warps do not crowd the shared-memory pipe)
#endif

#ifdef ARPDE_DEV
#define CH_TL(role, cond) do { if (a.tl && blockIdx.x == 0 && (cond) && tlc < 500) a.tl[(role) * 512 + tlc++] = clock64(); } while (0)
#else
#define CH_TL(role, cond) do { } while (0)
#endif

namespace {

struct ConvHArgs {
  const float* x; long long x_bs;
  const float* scale; const float* shift; long long ss_gs; int relu_in;
  const uint4* wh; long long wh_gs16;
  float* y; int y_ctot, y_coff;
  int N, n_per_group, cin, noct, ksteps, H, W, cout, coutP, coutE, k, pad, up;
  int pitch, RI, items_per_sample, P, nts, tile_cols;
  uint32_t inv_pitch;                                     // ceil(2^32 / pitch): p / pitch = umulhi(p, inv_pitch) for p < 2^16 (checked by the planner)
  long long* tl;                                          // development build: clock64 timeline of CTA 0 (arpde_conv_tc_set_debug buffer)
  int dbg;                                                // development build only (ARPDE_H_DBG): 1 no loads, 2 no conversion, 4 no MMAs, 8 no output stores
  int w16, ss_bytes;                                      // 16-byte units of the resident weight image; bytes of the scale / shift staging area
};

__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

// fp16 two-term split of two values: hi = fp16(v), lo' = fp16((v - hi) * 2^11), packed as half2 words.  Both conversions saturate to the
// largest finite fp16 value (F2FP.SATFINITE: one instruction, no clamps), so an out-of-range activation gives a wrong but finite result.
__device__ __forceinline__ uint32_t pack_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void split2(float e0, float e1, uint32_t* hw, uint32_t* lw) {
  const uint32_t h = pack_sat(e0, e1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
  *hw = h; *lw = pack_sat((e0 - hf.x) * 2048.f, (e1 - hf.y) * 2048.f);
}

template <int HWC>                                        // channel stride of the input in floats when known at compile time (0: H * W at run time)
__global__ void __launch_bounds__(CH_THREADS, 1) conv_h_kernel(const __grid_constant__ ConvHArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* patch_full = reinterpret_cast<uint64_t*>(smem_raw);      // [2] converted patch in place (one arrival per converter warp)
  uint64_t* patch_empty = patch_full + 2;                             // [2] all MMAs of the item complete (tcgen05.commit)
  uint64_t* w_full = patch_full + 4;                                  // resident weight image (re)loaded (bulk-copy tx bytes)
  uint64_t* w_free = patch_full + 5;                                  // all MMAs issued so far complete (before the image is replaced)
  uint64_t* tile_full = patch_full + 6;                               // [CH_MAX_TSLOTS] accumulator slot complete (tcgen05.commit)
  uint64_t* tile_empty = tile_full + CH_MAX_TSLOTS;                   // [CH_MAX_TSLOTS] slot drained (one arrival per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 192);
  float* ss = reinterpret_cast<float*>(smem_raw + CH_HEADER);        // [2][cin] scale | shift of the current weight group
  uint4* wbuf = reinterpret_cast<uint4*>(smem_raw + CH_HEADER + a.ss_bytes);   // [tap][k step][unit 2][2 coutE rows][16 B]
  uint4* patch = wbuf + a.w16;                                        // [buffer 2][hi|lo][octet][P positions][16 B]
  const int patch16 = 2 * a.noct * a.P;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long items_total = (long long)a.N * a.items_per_sample;
  const int it_begin = (int)((long long)blockIdx.x * items_total / gridDim.x);
  const int it_end = (int)((long long)(blockIdx.x + 1) * items_total / gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(patch_full + s, CH_CONV_WARPS); mbar_init(patch_empty + s, 1); }
    mbar_init(w_full, 1); mbar_init(w_free, 1);
    for (int s = 0; s < CH_MAX_TSLOTS; ++s) { mbar_init(tile_full + s, 1); mbar_init(tile_empty + s, CH_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CH_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp < CH_CONV_WARPS) {
    // =============================== converter warps: fp32 NCHW -> BN + ReLU -> fp16 hi / lo' operand image ===============================
    // A unit = (position, channel octet) = one 16-byte K unit of the hi and of the lo' image.  A thread owns the positions tid + 256 pb of
    // an item; the work is a flat sequence of batches of CH_KP (octet, position block) pairs over the items, and the loads of batch j + 1
    // are issued before batch j is converted and stored (two register buffers), so global-load latency overlaps the split arithmetic,
    // the shared-memory stores and the barrier hand-offs.  History (ncu, 1920 samples): v1 ran load -> wait -> convert per item with the
    // units of an item flattened over (octet, position): 250 instructions per unit (two integer divisions, 64-bit address arithmetic per
    // load, clamps around both conversions), converter warps at 1 IPC per SM while the MMA warp polled patch_full 64 times per item.
    const int HW = HWC ? HWC : a.H * a.W;
    int g_loaded = -1;
    int tlc = 0; (void)tlc;
    struct Cur { int it, v0; };
    auto item_geom = [&](int it, int* Prows, int* nposb) {
      const int b = it % a.items_per_sample;
      *Prows = (min(a.RI, a.H - b * a.RI) + a.k - 1) * a.pitch;
      *nposb = (*Prows + CH_CONV_THREADS - 1) / CH_CONV_THREADS;
    };
    auto advance = [&](Cur& c) {
      int Prows, nposb; item_geom(c.it, &Prows, &nposb);
      c.v0 += CH_KP;
      if (c.v0 >= a.noct * nposb) { c.v0 = 0; ++c.it; }
    };
    auto load = [&](const Cur& c, float (&xr)[CH_KP][8], int (&po)[CH_KP]) {
      if (c.it >= it_end) return;
      CH_TL(0, tid == 0);                                     // [0] load issue begins
      const int n = c.it / a.items_per_sample, b = c.it - n * a.items_per_sample;
      int Prows, nposb; item_geom(c.it, &Prows, &nposb);
      const int nv = a.noct * nposb;
      const float* xn = a.x + (long long)n * a.x_bs;
      const int y0 = b * a.RI - a.pad;
#pragma unroll
      for (int i = 0; i < CH_KP; ++i) {
        const int v = c.v0 + i;
        const int o = nposb == 1 ? v : (nposb == 2 ? v >> 1 : v / nposb), pb = v - o * nposb;
        const int p = pb * CH_CONV_THREADS + tid;
        po[i] = -1;
        if (v < nv && p < Prows) {
          const int r = (int)__umulhi((uint32_t)p, a.inv_pitch), cx = p - r * a.pitch;
          const int sy = wrap(y0 + r, a.H), sx = wrap(cx - a.pad, a.W);
          const float* src = xn + (long long)(o * 8) * HW + (sy * a.W + sx);
          if (!(a.dbg & 1)) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) xr[i][ch] = __ldcg(src + ch * HW);      // L2 only: the 4 KB channel stride of a 32x32 map puts all 64 planes of an item on the same few L1 sets
          } else {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) xr[i][ch] = 0.f;
          }
          po[i] = (o << 16) | p;
        }
      }
    };
    auto store = [&](const Cur& c, float (&xr)[CH_KP][8], int (&po)[CH_KP]) {
      if (c.it >= it_end) return;
      CH_TL(0, tid == 0);                                     // [1] store phase begins (loads of the next batch issued)
      const int lit = c.it - it_begin, buf = lit & 1;
      int Prows, nposb; item_geom(c.it, &Prows, &nposb);
      if (c.v0 == 0) {
        const int g = (c.it / a.items_per_sample) / a.n_per_group;
        if (a.scale && g != g_loaded) {
          bar_sync_named(1, CH_CONV_THREADS);                 // nobody still reads the previous group's values
          for (int ch = tid; ch < a.cin; ch += CH_CONV_THREADS) {
            ss[ch] = __ldg(a.scale + (long long)g * a.ss_gs + ch); ss[a.cin + ch] = __ldg(a.shift + (long long)g * a.ss_gs + ch);
          }
          bar_sync_named(1, CH_CONV_THREADS);
          g_loaded = g;
        }
        // the MMAs of the item that used this patch buffer two items ago no longer read it
        if (lit >= 2) mbar_wait<CH_SLEEP>(patch_empty + buf, ((lit >> 1) - 1) & 1);
      }
      CH_TL(0, tid == 0);                                     // [2] patch buffer free
      uint4* ph = patch + (size_t)buf * patch16;
#pragma unroll
      for (int i = 0; i < CH_KP; ++i) {
        if (po[i] >= 0 && !(a.dbg & 2)) {
          const int o = po[i] >> 16, p = po[i] & 0xffff;
          float e[8];
          if (a.scale) {
            const float4 s0 = *reinterpret_cast<const float4*>(ss + o * 8), s1 = *reinterpret_cast<const float4*>(ss + o * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(ss + a.cin + o * 8), b1 = *reinterpret_cast<const float4*>(ss + a.cin + o * 8 + 4);
            e[0] = fmaf(xr[i][0], s0.x, b0.x); e[1] = fmaf(xr[i][1], s0.y, b0.y); e[2] = fmaf(xr[i][2], s0.z, b0.z); e[3] = fmaf(xr[i][3], s0.w, b0.w);
            e[4] = fmaf(xr[i][4], s1.x, b1.x); e[5] = fmaf(xr[i][5], s1.y, b1.y); e[6] = fmaf(xr[i][6], s1.z, b1.z); e[7] = fmaf(xr[i][7], s1.w, b1.w);
          } else {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) e[ch] = xr[i][ch];
          }
          if (a.relu_in) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) e[ch] = fmaxf(e[ch], 0.f);
          }
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split2(e[2 * j], e[2 * j + 1], hw + j, lw + j);
          uint4* dst = ph + (size_t)o * a.P + p;
          dst[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          dst[(size_t)a.noct * a.P] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      CH_TL(0, tid == 0);                                     // [3] converted and stored
      if (c.v0 + CH_KP >= a.noct * nposb) {                   // last batch of the item
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(patch_full + buf);
      }
      CH_TL(0, tid == 0);                                     // [4] arrived
    };
#if CH_PIPELINED
    float xa[CH_KP][8], xb[CH_KP][8];
    int pa[CH_KP], pb_[CH_KP];
    Cur ca; ca.it = it_begin; ca.v0 = 0;
    load(ca, xa, pa);
    while (ca.it < it_end) {
      Cur cb = ca; advance(cb);
      load(cb, xb, pb_);
      store(ca, xa, pa);
      ca = cb; advance(ca);
      load(ca, xa, pa);
      store(cb, xb, pb_);
    }
#else
    float xa[CH_KP][8];
    int pa[CH_KP];
    Cur ca; ca.it = it_begin; ca.v0 = 0;
    while (ca.it < it_end) {
      load(ca, xa, pa);
      store(ca, xa, pa);
      advance(ca);
    }
#endif
  } else if (warp < CH_WARP_MMA) {
    // =============================== epilogue warps ===============================
    const int e = warp - CH_WARP_EPI0;
    const int lanebase = (e & 3) * 32;
    const long long HW = HWC ? HWC : (long long)a.H * a.W;
    const float lo_scale = 1.f / 2048.f;
    const int nts = a.nts;
    int gt = 0;
    int tlc = 0; (void)tlc;
    for (int it = it_begin; it < it_end; ++it) {
      const int n = it / a.items_per_sample, b = it - n * a.items_per_sample;
      const int r_item = min(a.RI, a.H - b * a.RI);
      const int nt = (r_item * a.pitch + 127) >> 7;
      for (int t = 0; t < nt; ++t, ++gt) {
        const int slot = gt % nts, kk = gt / nts;
        CH_TL(1, e == 0 && lane == 0);                        // [0] waiting for the tile
        mbar_wait<CH_SLEEP>(tile_full + slot, kk & 1);
        CH_TL(1, e == 0 && lane == 0);                        // [1] tile complete
        tc_fence_after();
        const uint32_t trow = tmem + ((uint32_t)lanebase << 16) + (uint32_t)(slot * a.tile_cols);
        const int q = t * 128 + lanebase + lane;
        const int oyl = q / a.pitch, ox = q - oyl * a.pitch;
        const bool valid = oyl < r_item && ox < a.W && !(a.dbg & 8);
        const int oy = b * a.RI + oyl;
        if (!a.up) {
          float* yp = a.y + ((long long)n * a.y_ctot + a.y_coff) * HW + (long long)oy * a.W + ox;
          for (int c0 = 0; c0 < a.coutP; c0 += 16) {
            float v1[16], v2[16];
            tmem_ld16(trow + c0, v1); tmem_ld16(trow + a.coutE + c0, v2);
            tmem_ld_wait();
            if (valid) {
              float* pc = yp + (long long)c0 * HW;
#pragma unroll
              for (int j = 0; j < 16; ++j) { if (c0 + j < a.cout) *pc = fmaf(v2[j], lo_scale, v1[j]); pc += HW; }
            }
          }
        } else {
          // folded upsampling: column (py * 2 + px) * coutP + co -> pixel (2 oy + py, 2 ox + px); both px of a channel are one float2 store
          const long long HWo = 4 * HW;
          float* yp = a.y + ((long long)n * a.y_ctot + a.y_coff) * HWo + (long long)(2 * oy) * (2 * a.W) + 2 * ox;
          for (int py = 0; py < 2; ++py)
            for (int c0 = 0; c0 < a.coutP; c0 += 16) {
              const int colA = (py * 2) * a.coutP + c0, colB = colA + a.coutP;
              float a1[16], a2[16], b1[16], b2[16];
              tmem_ld16(trow + colA, a1); tmem_ld16(trow + a.coutE + colA, a2);
              tmem_ld16(trow + colB, b1); tmem_ld16(trow + a.coutE + colB, b2);
              tmem_ld_wait();
              if (valid) {
                float* pc = yp + (long long)c0 * HWo + py * (2 * a.W);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (c0 + j < a.cout) *reinterpret_cast<float2*>(pc) = make_float2(fmaf(a2[j], lo_scale, a1[j]), fmaf(b2[j], lo_scale, b1[j]));
                  pc += HWo;
                }
              }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tile_empty + slot);
        CH_TL(1, e == 0 && lane == 0);                        // [2] drained
      }
    }
  } else {
    // =============================== MMA issuing warp (also (re)loads the resident weight image) ===============================
    const bool leader = elect_one();
    const uint32_t idesc1 = umma_idesc_f16(128, 2 * a.coutE), idesc2 = umma_idesc_f16(128, a.coutE);
    const uint64_t hi64 = (uint64_t)((128u >> 4) | (1u << 14)) << 32;          // SBO = 128 B, descriptor version 1
    const uint32_t P = (uint32_t)a.P, coutE = (uint32_t)a.coutE;
    const uint32_t a_lbo = P << 16;                                             // second K unit = the next channel octet
    const uint32_t b_base = (smem_u32(wbuf) >> 4) | ((2u * coutE) << 16);
    const uint32_t kstep16 = 4u * coutE;                                        // 16-byte units of one (tap, k step) weight piece
    const uint32_t lo_off = (uint32_t)a.noct * P;
    const int nts = a.nts, k = a.k, pitch = a.pitch, ksteps = a.ksteps;
    int g_loaded = -1, wgen = 0, wfree = 0, gt = 0;
    int tlc = 0; (void)tlc;
    for (int it = it_begin; it < it_end; ++it) {
      const int lit = it - it_begin, buf = lit & 1;
      const int n = it / a.items_per_sample, b = it - n * a.items_per_sample, g = n / a.n_per_group;
      const int r_item = min(a.RI, a.H - b * a.RI);
      const int nt = (r_item * pitch + 127) >> 7;
      if (g != g_loaded) {
        if (leader) {
          if (g_loaded >= 0) { umma_commit(w_free); mbar_wait<CH_SLEEP>(w_free, wfree & 1); }    // every MMA that reads the old image is complete
          const uint32_t bytes = (uint32_t)a.w16 * 16u;
          mbar_arrive_expect_tx(w_full, bytes);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(a.wh + (long long)g * a.wh_gs16);
          for (uint32_t o = 0; o < bytes; o += 32768u) {
            const uint32_t nb = min(32768u, bytes - o);
            bulk_g2s(reinterpret_cast<unsigned char*>(wbuf) + o, src + o, nb, w_full);
          }
        }
        if (g_loaded >= 0) ++wfree;
        __syncwarp();
        mbar_wait<CH_SLEEP>(w_full, wgen & 1); ++wgen; g_loaded = g;
      }
      CH_TL(2, leader);                                       // [0] waiting for the patch
      mbar_wait<CH_SLEEP>(patch_full + buf, (lit >> 1) & 1);
      tc_fence_after();
      CH_TL(2, leader);                                       // [1] patch in place
      const uint32_t a_hi = (smem_u32(patch + (size_t)buf * patch16) >> 4);
      for (int t = 0; t < nt; ++t, ++gt) {
        const int slot = gt % nts, kk = gt / nts;
        if (kk > 0) mbar_wait<CH_SLEEP>(tile_empty + slot, (kk - 1) & 1);
        tc_fence_after();
        if (leader && !(a.dbg & 4)) {
          const uint32_t d = tmem + (uint32_t)slot * (uint32_t)a.tile_cols;
          uint32_t acc = 0u, wb = b_base;
          for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) {
              uint32_t ah = a_hi + (uint32_t)(t * 128 + ky * pitch + kx);
              for (int s = 0; s < ksteps; ++s) {
                const uint64_t db = hi64 | wb;
                umma_f16(d, hi64 | (uint64_t)(ah | a_lbo), db, idesc1, acc);                    // hi * [hi ; lo']
                umma_f16(d + coutE, hi64 | (uint64_t)((ah + lo_off) | a_lbo), db, idesc2, 1u);  // lo' * hi -> correction half
                acc = 1u; ah += 2u * P; wb += kstep16;
              }
            }
        }
        __syncwarp();
        if (leader) umma_commit(tile_full + slot);
        CH_TL(2, leader);                                     // [2 + t] tile issued
      }
      if (leader) umma_commit(patch_empty + buf);
    }
    __syncwarp();
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == CH_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// w [groups][cout][cin][k][k] -> out [groups][tap][k step][unit 2][row 2 coutE][8 halfs]: unit u of step s holds channels (2 s + u) * 8 .. + 8,
// rows [0, coutE) the fp16 hi part of output column `row`, rows [coutE, 2 coutE) fp16((w - hi) * 2^11).  Column = output channel, or -- for
// the folded upsampling layer -- parity * coutP + channel with the low-resolution tap (ty, tx) = sum of the 3x3 taps that land on it.
__global__ void conv_h_wprep_kernel(const float* __restrict__ w, long long w_gs, __half* __restrict__ out, long long out_gs, int cin, int cout,
                                    int coutP, int coutE, int k, int up, int groups) {
  const int ksteps = cin / 16, ntaps = k * k;
  const long long per_group = (long long)ntaps * ksteps * 2 * (2 * coutE) * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_group * groups; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i / per_group);
    long long r = i - (long long)g * per_group;
    const int j = (int)(r & 7); r >>= 3;
    const int row = (int)(r % (2 * coutE)); r /= 2 * coutE;
    const int u = (int)(r & 1); r >>= 1;
    const int s = (int)(r % ksteps), tap = (int)(r / ksteps);
    const int ci = (2 * s + u) * 8 + j, col = row < coutE ? row : row - coutE;
    float v = 0.f;
    if (!up) {
      if (col < cout) v = __ldg(w + (long long)g * w_gs + ((long long)col * cin + ci) * ntaps + tap);
    } else {
      const int par = col / coutP, co = col - par * coutP, py = par >> 1, px = par & 1, ty = tap / 3, tx = tap - ty * 3;
      if (co < cout) {
        const float* wp = w + (long long)g * w_gs + ((long long)co * cin + ci) * 9;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          if (((py + ky + 1) >> 1) != ty) continue;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            if (((px + kx + 1) >> 1) == tx) v += __ldg(wp + ky * 3 + kx);
        }
      }
    }
    const __half hi = __float2half_rn(v);
    out[(long long)g * out_gs + (i - (long long)g * per_group)] = row < coutE ? hi : __float2half_rn((v - __half2float(hi)) * 2048.f);
  }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

bool ch_plan(const ConvDesc& d, ConvHArgs& a, size_t* smem_out) {
  if (d.kh != d.kw || (d.kh != 1 && d.kh != 3) || d.stride != 1 || d.dilate || d.pad_override >= 0 || d.act != 0) return false;
  if (d.up && d.kh != 3) return false;
  if (d.cin % 16 || d.cin < 16 || d.cin > 256 || d.cout < 1 || d.H < 2 || d.W < 4 || d.W < d.kh || d.H < d.kh) return false;
  memset(&a, 0, sizeof(a));
  a.cin = d.cin; a.noct = d.cin / 8; a.ksteps = d.cin / 16; a.H = d.H; a.W = d.W; a.cout = d.cout; a.k = d.kh; a.pad = (d.kh - 1) / 2; a.up = d.up ? 1 : 0;
  a.coutP = round_up(d.cout, 16); a.coutE = a.up ? 4 * a.coutP : a.coutP;
  if (a.coutE > 128) return false;                        // N' = 2 coutE <= 256
  a.tile_cols = 2 * a.coutE; a.nts = a.tile_cols <= 128 ? 4 : 2;
  a.pitch = d.W + d.kh - 1;
  if (a.pitch >= 65536) return false;
  a.inv_pitch = (uint32_t)(((1ull << 32) + a.pitch - 1) / a.pitch);     // exact for p, pitch < 2^16
  a.w16 = a.k * a.k * a.ksteps * 2 * 2 * a.coutE;
  a.ss_bytes = round_up(2 * d.cin * 4, 128);
  const size_t budget = 226 * 1024;
  long long best_tiles = -1; int best_RI = 0; double best_ratio = -1.;
  for (int RI = d.H; RI >= 1; --RI) {
    const int items = (d.H + RI - 1) / RI, rows_last = d.H - (items - 1) * RI;
    const int nt_full = (RI * a.pitch + 127) / 128, nt_last = (rows_last * a.pitch + 127) / 128;
    int P = nt_full * 128 + (a.k - 1) * a.pitch + a.k - 1;
    if (P < (RI + a.k - 1) * a.pitch) P = (RI + a.k - 1) * a.pitch;
    P = round_up(P, 8);
    if (P > 16000) continue;
    const size_t smem = CH_HEADER + a.ss_bytes + (size_t)a.w16 * 16 + (size_t)2 * 2 * a.noct * P * 16;
    if (smem > budget) continue;
    const long long tiles = (long long)(items - 1) * nt_full + nt_last;
    const double ratio = (double)rows_last / RI;
    if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && ratio > best_ratio + 1e-9)) {
      best_tiles = tiles; best_RI = RI; best_ratio = ratio;
      a.RI = RI; a.items_per_sample = items; a.P = P;
      *smem_out = smem;
    }
  }
  return best_RI > 0;
}

}  // namespace

// eval-mode layers this kernel takes from conv_tc.cu (d: the layer as the program runner describes it)
bool arpde_conv_h_wanted(const ConvDesc& d) {
  if (!arpde_tc_enabled() || arpde_dev_env("ARPDE_NO_H")) return false;
  if (d.cout < 8 || d.cin * d.kh * d.kw < 32) return false;
  ConvHArgs a; size_t smem;
  return ch_plan(d, a, &smem);
}

// prepared fp16 weights per group, in fp32 floats of scratch (<= arpde_conv_tc_wprep_floats of the same layer)
long long arpde_conv_h_wprep_floats(int cin, int cout, int k, int up) {
  const int coutP = round_up(cout, 16), coutE = up ? 4 * coutP : coutP;
  return (long long)k * k * (cin / 16) * 2 * (2 * coutE) * 8 / 2;
}

int arpde_conv_h_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs, cudaStream_t stream) {
  ConvHArgs a; size_t smem;
  if (!ch_plan(d, a, &smem)) { arpde_set_error("conv_h: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  const long long n = (long long)a.w16 * 8 * groups;
  long long blocks = (n + 255) / 256; if (blocks > 4096) blocks = 4096;
  conv_h_wprep_kernel<<<(unsigned)blocks, 256, 0, stream>>>(w, w_gs, reinterpret_cast<__half*>(wprep), 2 * wprep_gs, d.cin, d.cout, a.coutP, a.coutE, a.k,
                                                            a.up, groups);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

int arpde_conv_h_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const float* scale,
                        const float* shift, long long ss_gs, float* y, cudaStream_t stream) {
  ARPDE_CHECK_ARG(x && wprep && y, "conv_h: null pointer");
  if (d.N == 0) return ARPDE_OK;
  ConvHArgs a; size_t smem;
  if (!ch_plan(d, a, &smem)) { arpde_set_error("conv_h: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  ARPDE_CHECK_ARG((((uintptr_t)wprep) & 15) == 0 && (wprep_gs & 3) == 0, "conv_h: prepared weights must be 16-byte aligned");
  ARPDE_CHECK_ARG(!a.up || (((uintptr_t)y) & 7) == 0, "conv_h: output of the folded upsampling layer must be 8-byte aligned");
  a.x = x; a.x_bs = x_bs; a.scale = scale; a.shift = scale ? shift : nullptr; a.ss_gs = ss_gs; a.relu_in = d.relu_in;
  a.wh = reinterpret_cast<const uint4*>(wprep); a.wh_gs16 = wprep_gs / 4;
  if (const char* e = arpde_dev_env("ARPDE_H_DBG")) a.dbg = atoi(e);
  a.tl = arpde_tc_debug_buffer();
  a.y = y; a.y_ctot = d.y_ctot; a.y_coff = d.y_coff; a.N = d.N; a.n_per_group = d.n_per_group;
  static std::mutex attr_mutex;
  static bool attr_done[64] = {false};
  {
    int dev = 0, optin = 0;
    ARPDE_CUDA(cudaGetDevice(&dev));
    ARPDE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    ARPDE_CHECK_ARG(smem <= (size_t)optin, "conv_h: plan needs %zu bytes of shared memory, device allows %d", smem, optin);
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      ARPDE_CUDA(cudaFuncSetAttribute(conv_h_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
      ARPDE_CUDA(cudaFuncSetAttribute(conv_h_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  const long long items = (long long)d.N * a.items_per_sample;
  ARPDE_CHECK_ARG(items < (1ll << 31), "conv_h: too many work items");
  long long blocks = (long long)arpde_num_sms();
  if (blocks > items) blocks = items;
  if (d.H * d.W == 1024) ARPDE_CUDA(arpde_launch1(conv_h_kernel<1024>, (unsigned)blocks, CH_THREADS, smem, stream, d.pdl != 0, a));   // 32x32 maps
  else ARPDE_CUDA(arpde_launch1(conv_h_kernel<0>, (unsigned)blocks, CH_THREADS, smem, stream, d.pdl != 0, a));
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// C ABI: y = conv_k(up2?(relu?(scale * x + shift))), circular, stride 1, k = 1 or 3 (up: k = 3), on the fp16 two-term tensor-core kernel.
// scratch: arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, groups) floats.  ARPDE_ERR_UNSUPPORTED for shapes the kernel does not cover
// (16 | cin <= 256; cout <= 128, folded: <= 32).
extern "C" int64_t arpde_conv_fwd_h_scratch_floats(int cin, int cout, int k, int up, int groups) {
  if (cin <= 0 || cout <= 0 || k <= 0) return 0;
  long long per_group = arpde_conv_tc_wprep_floats(cin, cout, k, k);
  const long long need_h = (arpde_conv_h_wprep_floats(cin, cout, k, up) + 3) & ~3ll;
  if (per_group < need_h) per_group = need_h;
  return per_group * (groups > 0 ? groups : 1);
}

extern "C" int arpde_conv_fwd_h(const float* x, int64_t x_bstride, const float* w, int64_t w_gstride, const float* scale, const float* shift,
                                int64_t ss_gstride, float* y, int N, int cin, int H, int W, int cout, int k, int up, int relu_in, int y_ctot,
                                int y_coff, int n_per_group, float* scratch, int64_t scratch_floats, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(x && w && y && scratch, "conv_fwd_h: null pointer");
  ARPDE_CHECK_ARG(N >= 0 && cin > 0 && cout > 0 && H > 0 && W > 0, "conv_fwd_h: bad dims");
  ARPDE_CHECK_ARG(y_coff >= 0 && y_coff + cout <= y_ctot, "conv_fwd_h: channel window [%d,%d) outside y_ctot=%d", y_coff, y_coff + cout, y_ctot);
  ARPDE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_fwd_h: scale and shift go together");
  if (N == 0) return ARPDE_OK;
  ConvDesc d;
  d.N = N; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = k; d.kw = k; d.stride = 1; d.up = up; d.relu_in = relu_in; d.act = 0;
  d.y_ctot = y_ctot; d.y_coff = y_coff; d.n_per_group = n_per_group > 0 ? n_per_group : N;
  ARPDE_CHECK_ARG(N % d.n_per_group == 0, "conv_fwd_h: N=%d not a multiple of n_per_group=%d", N, d.n_per_group);
  ConvHArgs a; size_t smem;
  if (!ch_plan(d, a, &smem)) { arpde_set_error("conv_fwd_h: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  const int groups = N / d.n_per_group;
  long long per_group = arpde_conv_tc_wprep_floats(cin, cout, k, k);
  const long long need_h = (arpde_conv_h_wprep_floats(cin, cout, k, up) + 3) & ~3ll;
  if (per_group < need_h) per_group = need_h;
  ARPDE_CHECK_ARG(scratch_floats >= per_group * groups, "conv_fwd_h: scratch too small (%lld < %lld floats)", (long long)scratch_floats, per_group * groups);
  int rc = arpde_conv_h_prep_launch(d, w, w_gstride, groups, scratch, per_group, stream);
  if (rc) return rc;
  return arpde_conv_h_launch(d, x, x_bstride, scratch, per_group, scale, shift, ss_gstride, y, stream);
}
