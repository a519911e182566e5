// conv_wgrad_tc.cu -- weight gradient of the circular convolution on the tcgen05 tensor cores, 3xTF32.
//
// Replaces, for the wide layers, the same autograd node as wgrad_kernel (conv_bwd.cu): the dW of
// F.pad(circular) + aten::convolution that loss.backward() derives at 2D-Burgers-SWAG/main.py:81 for the convolutions of
// 2D-Burgers-SWAG/nn/denseEDcirc2d.py:34-190,:242-302, with the BatchNorm+ReLU prologue, nearest x2 upsampling and
// stride 2 folded into the operand staging.
//
//   dW[co][ci][ky][kx] = sum_{n, oy, ox} dO[n][co][oy][ox] * a[n][ci][oy*s + ky - p][ox*s + kx - p]      (a = up2?(relu(bn(x))), wrapped)
//
// As a GEMM the contraction index K is the output POSITION, so with NCHW data the input operand is "MN-major"
// (channels contiguous per position).  Measured on B200 (scratch/umma_probe_mn.cu): for kind::tf32 an MN-major operand
// only works with descriptor layout type 1 (128-byte swizzle, 32-byte base: rows of 32 channels = 128 B per position,
// 32-byte chunks XORed with (position & 3)); every other layout type silently yields zeros.  With that layout a filter
// tap is -- exactly as in conv_tc.cu -- a shift of the descriptor start address by (ky'*pitch + kx') positions = rows
// of 128 B, which the hardware accepts at any row (the swizzle is a function of the absolute address).
//
//   D[co][(tap, ci)] += A[co][q] * B[q + off(tap)][ci]        one accumulator column range per tap, resident in TMEM
//     A = dO, K-major no-swizzle [q/4][co][4 floats], zero in the pad columns of the flattened padded row (pitch = Wo + span)
//     B = input patch, MN-major [position][32 channels], kept in a RING of 32-position chunks that a persistent CTA
//         streams through whole row blocks -- the halo (span_y rows) is paid once per row block, not per K tile; the
//         first 8 rows are mirrored behind the ring end so that an 8-position K step never wraps;
//     stride 2: one ring per input parity plane (tap -> plane, offset), as conv_tc.cu;
//     narrow-cin layers (cin * kh <= 32, e.g. the 6-channel stem): the 32 columns of a row hold (ky, ci) pairs, i.e.
//     the y taps are unrolled into the MMA's N and only the kw x-taps remain (offsets < kw: a one-chunk halo) -- N = 32
//     instead of 8 per instruction; with swapped roles up to 128 (ky, ci) rows (LastTransUp's 16 -> 2 5x5 conv).
//     small-cout layers (dense layers, cout = growth rate <= 8): roles swapped -- the patch ring is the M operand (MN-major A: up to
//     128 channels = four 32-channel blocks at LBO stride), dO the K-major N operand (N = 8), D[ci][(tap, co)]; otherwise 60 of the
//     64 accumulator rows would be padding (an MMA costs ~40 cycles whatever M and N <= 64 are).
//     nearest x2 + 3x3 (LastTransUp conv2): folded like the forward (conv_tc.cu "upfold") -- the gradient of the equivalent
//     low-resolution conv with 4*cout parity channels is accumulated (dO gathered with stride 2 at staging time, 4x fewer positions,
//     all 64 accumulator rows real) and the drain adds every folded tap into the original taps it is the sum of.
//   3xTF32: hi*hi + hi*lo + lo*hi into the same fp32 accumulator (lo*lo ~ 2^-22 dropped).
//   The tensor core TRUNCATES when it adds into the fp32 accumulator (measured: one 1100-step accumulation chain over a
//   training batch was off by 2.3e-5 relative, always towards zero), so a chain is cut after WG_FC chunks (96 adds):
//   four drain warps add the segment into a CTA-private fp32 slab in global memory (L2 resident, round-to-nearest) while
//   the MMAs continue in a second accumulator set when TMEM has room for two; the last segment goes, together with the
//   slab, into the fp64 dW64 scratch the FFMA kernel uses (atomics), from which dW is rounded once.
// Warp roles: 8 converter warps (LDG -> BN/ReLU -> split -> STS), 1 MMA-issuing warp, 4 drain warps.
#include "common.cuh"
#include "conv_common.cuh"
#include "tc_ptx.cuh"
#include "../../include/arpde.h"
#include <stdlib.h>
#include <mutex>

#define WG_MAX_TAPS 49
#define WG_CONV_WARPS 8
#define WG_CONVERTERS (32 * WG_CONV_WARPS)
#define WG_WARP_MMA WG_CONV_WARPS
#define WG_WARP_DRAIN0 (WG_CONV_WARPS + 1)
#define WG_THREADS (32 * (WG_CONV_WARPS + 1 + 4))
#define WG_FC 8                          // MMA chunks per accumulation segment (8 x 4 K steps x 3 terms = 96 truncating adds)
#define WG_CK 32                         // positions per chunk (4 K steps of 8)
#define WG_NG 3                          // dO chunk slots
#define WG_MAX_CIN 256
#define WG_HEADER 4096                   // barriers, tmem slot, scale/shift table; keeps the rings 1024-byte aligned
#define WG_MAX_NR 20

struct WgTcArgs {
  const float* x; long long x_bs;
  const float* scale; const float* shift; int relu_in;
  const float* dO; long long dO_bs;
  double* dW;
  float* slab;                                 // [CTA][cols][128 TMEM lanes] fp32 partial sums of the finished segments
  int N, cin, H, W, up_h, up_w, sy, sx, Ho, Wo, cout, kh, kw;
  int pitch, min_ry, min_rx;
  int RB, RBH, nch_u, lead, pcu, NR, ring;     // row blocks per sample, rows per block, MMA chunks / patch chunks per unit, ring chunks / rows
  int npl, NC, ncc, cchunk, packed, ntaps;     // planes, columns per tap (MMA N), channel chunks (grid split), channels per chunk, (kx,ci) packing, MMA taps
  int fold, cout_fold;                         // folded upsampling: cout = 4 * cout_fold parity-major channels on the low-resolution grid
  int swap, nblk;                              // swap = 1: ring is the M operand (nblk blocks of 32 channels), dO the N operand (NC = cout padded)
  int M, mrows, tmem_cols, slices, units, cols, nacc;   // cols = ntaps * NC accumulator columns per set, nacc sets
  long long* dbg;                              // cycle counters of CTA 0 (arpde_conv_tc_set_debug), else nullptr
  int dbg_flags;                               // experiments (ARPDE_WG_DBG): 1 = converters skip loads and stores, 4 = skip the MMAs
  int plane_floats, gy_slot_floats;            // one ring plane (hi or lo) incl. mirror; one dO slot (hi + lo)
  int tap_plane[WG_MAX_TAPS];
  int tap_off[WG_MAX_TAPS];
  signed char col_ci[128];                     // ring column (32 per channel block) -> channel inside the chunk (-1: padding)
  signed char col_dy[128];                     // ring column -> y-tap index folded into the column (packed layers), else 0
};

__device__ __forceinline__ uint64_t wg_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

__device__ __forceinline__ void wg_split(float t, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(t) + 0x1000u) & 0xffffe000u);      // tf32 round-to-nearest in two integer ops (conv_tc.cu)
  lo = t - hi;
}

// MMAs of one 32-position chunk, issued by the elected lane: 4 K steps x taps x 3 split terms.  Descriptor high words are constants
// (dO: K-major no-swizzle, SBO 128 B; ring: MN-major layout type 1, SBO 512 B, LBO = one ring plane between 32-channel blocks); only the
// 14-bit start-address field of the low word changes.  (The first version rebuilt four 64-bit descriptors per tap and branched on the
// role inside the loop: 60-80 cycles per MMA instead of the tensor core's ~40.)
template <bool SWAP>
__device__ __forceinline__ void wg_issue_chunk(const WgTcArgs& a, uint32_t gy_lo, uint32_t ring_lo, int base_row, uint32_t acc_base,
                                               uint32_t idesc, bool first) {
  const uint64_t gy_hi64 = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
  const uint64_t ring_hi64 = (uint64_t)((512u >> 4) | (1u << 14) | (1u << 29)) << 32;
  const uint32_t gy_part16 = (uint32_t)a.gy_slot_floats >> 3;                           // hi -> lo part of a dO slot, 16-byte units
  const uint32_t ring_part16 = (uint32_t)(a.npl * a.nblk) * ((uint32_t)a.plane_floats >> 2);
  const uint32_t plane16 = (uint32_t)a.nblk * ((uint32_t)a.plane_floats >> 2);        // parity planes are nblk ring planes apart
  const uint32_t k8_step16 = 2u * (uint32_t)a.mrows;                                    // two 16-byte K planes per K = 8 step
  const int ntaps = a.ntaps, ring_rows = a.ring;
  const uint32_t NC = (uint32_t)a.NC;
#pragma unroll 1
  for (int k8 = 0; k8 < 4; ++k8) {
    const uint64_t dg_hi = gy_hi64 | (gy_lo + (uint32_t)k8 * k8_step16), dg_lo = dg_hi + gy_part16;
    const uint32_t acc0 = (first && k8 == 0) ? 0u : 1u;
    const int row0 = base_row + k8 * 8;
    uint32_t d = acc_base;
#pragma unroll 1
    for (int t = 0; t < ntaps; ++t, d += NC) {
      int rr = row0 + a.tap_off[t];
      if (rr >= ring_rows) rr -= ring_rows;
      const uint64_t dr_hi = ring_hi64 | (ring_lo + (uint32_t)a.tap_plane[t] * plane16 + (uint32_t)rr * 8u), dr_lo = dr_hi + ring_part16;
      if (!SWAP) {
        umma_tf32(d, dg_hi, dr_hi, idesc, acc0);
        umma_tf32(d, dg_hi, dr_lo, idesc, 1u);
        umma_tf32(d, dg_lo, dr_hi, idesc, 1u);
      } else {
        umma_tf32(d, dr_hi, dg_hi, idesc, acc0);
        umma_tf32(d, dr_lo, dg_hi, idesc, 1u);
        umma_tf32(d, dr_hi, dg_lo, idesc, 1u);
      }
    }
  }
}

// FOLD (folded upsampling) is a template parameter: as a run-time branch in the dO staging it cost the other layers 5-35 %
template <bool FOLD>
__global__ void __launch_bounds__(WG_THREADS, 1) conv_wgrad_tc_kernel(const __grid_constant__ WgTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* patch_full = reinterpret_cast<uint64_t*>(smem_raw);        // [NR] ring chunk converted (one arrival per converter warp)
  uint64_t* patch_empty = patch_full + WG_MAX_NR;                       // [NR] ring chunk no longer read (tcgen05.commit)
  uint64_t* gy_full = patch_empty + WG_MAX_NR;                          // [NG]
  uint64_t* gy_empty = gy_full + WG_NG;                                 // [NG]
  uint64_t* acc_full = gy_empty + WG_NG;                                // [2] accumulation segment complete (tcgen05.commit)
  uint64_t* acc_free = acc_full + 2;                                    // [2] accumulator set drained (one arrival per drain warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 512);
  float* ss_tab = reinterpret_cast<float*>(smem_raw + 1024);            // [scale | shift][WG_MAX_CIN]
  float* ring = reinterpret_cast<float*>(smem_raw + WG_HEADER);         // [hi | lo][npl * nblk][ring + 8 rows][32]
  float* gyb = ring + (size_t)2 * a.npl * a.nblk * a.plane_floats;      // [NG][hi | lo][8][mrows][4]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cc = (int)blockIdx.x % a.ncc, slice = (int)blockIdx.x / a.ncc;
  const int u_begin = (int)((long long)a.units * slice / a.slices), u_end = (int)((long long)a.units * (slice + 1) / a.slices);
  const int nunits = u_end - u_begin;
  const int nchunks_cta = nunits * a.nch_u, nseg = (nchunks_cta + WG_FC - 1) / WG_FC;

  if (tid == 0) {
    for (int s = 0; s < a.NR; ++s) { mbar_init(patch_full + s, WG_CONV_WARPS); mbar_init(patch_empty + s, 1); }
    for (int s = 0; s < WG_NG; ++s) { mbar_init(gy_full + s, WG_CONV_WARPS); mbar_init(gy_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_free + s, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < WG_CONVERTERS) {
    for (int c = tid; c < a.cin; c += WG_CONVERTERS) {
      ss_tab[c] = a.scale ? __ldg(a.scale + c) : 1.f;
      ss_tab[WG_MAX_CIN + c] = a.scale ? __ldg(a.shift + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < WG_CONV_WARPS) {
    // =============================== converter warps ===============================
    // the four columns this thread fills in every patch row (one 16-byte store per plane and part)
    const int e0 = 4 * warp;
    int cch[4], cdy[4]; float csc[4], csh[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ci = a.col_ci[e0 + k];
      const int c = ci < 0 ? -1 : cc * a.cchunk + ci;
      cch[k] = (c >= 0 && c < a.cin) ? c : -1;
      cdy[k] = a.col_dy[e0 + k];
      csc[k] = cch[k] >= 0 ? ss_tab[cch[k]] : 0.f;
      csh[k] = cch[k] >= 0 ? ss_tab[WG_MAX_CIN + cch[k]] : 0.f;
    }
    const bool any_col = cch[0] >= 0 || cch[1] >= 0 || cch[2] >= 0 || cch[3] >= 0 || a.nblk > 1;
    const int gy_units = a.cout * 8;                   // (co, group of 4 positions) per dO chunk
    int pg = 0, hg = 0;                                // running patch-chunk / dO-chunk counters of this CTA
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
    long long d_wait_p = 0, d_wait_g = 0, d_load = 0;
    const long long d_t0 = dbg_on ? clock64() : 0;
    // loop-invariant geometry in registers; every per-step index is advanced incrementally (the first version recomputed two integer
    // divisions and 64-bit address products per load and spent ~1500 cycles per 32-position step on index arithmetic alone)
    const int pitch = a.pitch, Ho = a.Ho, Wo = a.Wo, Wsrc = a.W, sy = a.sy, sx = a.sx, up_h = a.up_h, up_w = a.up_w, npl = a.npl;
    const int HWi = a.H * a.W, HoWoi = a.Ho * a.Wo, nblk = a.nblk, nslot = a.npl * a.nblk;
    const int climit = min(a.cin, (cc + 1) * a.cchunk);           // channels of this CTA's chunk end here
    int coff[4];                                       // channel offset of each column inside the sample
#pragma unroll
    for (int k = 0; k < 4; ++k) coff[k] = (cch[k] < 0 ? 0 : cch[k]) * HWi;
    int gco[4], gq4[4], gpx[4];                                // dO unit of this thread: channel offset, first position inside the chunk
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int uu = tid + k * WG_CONVERTERS, ch = uu >> 3;
      gq4[k] = (uu & 7) << 2;
      if (FOLD) {      // channel ch = parity * cout_fold + co reads dO[co][2 oy + py][2 ox + px]: offset of (co, row parity), px kept apart
        const int par = ch / a.cout_fold, co = ch - par * a.cout_fold;
        gco[k] = co * 4 * HoWoi + (par >> 1) * 2 * a.Wo; gpx[k] = par & 1;
      } else { gco[k] = ch * HoWoi; gpx[k] = 0; }
    }
    const bool no_load = (a.dbg_flags & 1) != 0;
    for (int u = u_begin; u < u_end; ++u) {
      const int n = u / a.RB, rb = u - n * a.RB;
      const int R0 = rb * a.RBH;
      const float* xn = a.x + (long long)n * a.x_bs;
      const float* gn = a.dO + (long long)n * a.dO_bs + (long long)R0 * a.Wo * (FOLD ? 4 : 1);     // folded: row 2 R0 of the 2Wo-wide gradient
      int pi = 0, pj = lane;                           // (row, column) of this thread's patch position in the padded row pitch
      while (pj >= pitch) { pj -= pitch; ++pi; }
      int goy[4], gox[4];                              // (row, column) of the dO positions
#pragma unroll
      for (int k = 0; k < 4; ++k) { goy[k] = 0; gox[k] = gq4[k]; while (gox[k] >= pitch) { gox[k] -= pitch; ++goy[k]; } }
      for (int t = 0; t < a.pcu; ++t, ++pg) {
        // ---- loads of patch chunk t (32 positions, this thread: position `lane`, 4 columns, every plane)
        const long long d_t1 = dbg_on ? clock64() : 0;
        float xv[4][4];                                  // [column][slot], slot = plane (stride 2) or 32-channel block (swapped roles)
        {
          int C = a.min_rx + pj; if (C < 0) C += Wo; else if (C >= Wo) C -= Wo;
          const int col_off = (C * sx) >> up_w;
          const int Rb = R0 + a.min_ry + pi;             // row before the folded y tap; in [-Ho, 2 Ho)
          if (nblk == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              int R = Rb + cdy[k]; if (R < 0) R += Ho; else if (R >= Ho) R -= Ho; if (R >= Ho) R -= Ho;
              const float* src = xn + (coff[k] + ((R * sy) >> up_h) * Wsrc + col_off);
              const bool ok = cch[k] >= 0 && !no_load;
#pragma unroll
              for (int pl = 0; pl < 4; ++pl) xv[k][pl] = (pl < npl && ok) ? __ldg(src + (pl >> 1) * Wsrc + (pl & 1)) : 0.f;
            }
          } else {                                       // swapped roles: up to four 32-row blocks, rows looked up in the tables
#pragma unroll
            for (int b = 0; b < 4; ++b) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                xv[k][b] = 0.f;
                if (b < nblk) {
                  const int r = 32 * b + e0 + k;
                  const int ci = a.col_ci[r], c = cc * a.cchunk + ci;
                  if (ci >= 0 && c < climit && !no_load) {
                    int R = Rb + a.col_dy[r]; if (R < 0) R += Ho; else if (R >= Ho) R -= Ho; if (R >= Ho) R -= Ho;
                    xv[k][b] = __ldg(xn + (c * HWi + ((R * sy) >> up_h) * Wsrc + col_off));
                  }
                }
              }
            }
          }
        }
        pj += WG_CK;
        while (pj >= pitch) { pj -= pitch; ++pi; }
        // ---- loads of dO chunk t - lead
        const int jg = t - a.lead;
        float4 gv[4];
        if (jg >= 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            gv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tid + k * WG_CONVERTERS < gy_units) {
              if (gox[k] < Wo && !no_load) {
                if (!FOLD) gv[k] = __ldg(reinterpret_cast<const float4*>(gn + (gco[k] + goy[k] * Wo + gox[k])));
                else {     // 8 consecutive full-resolution columns from 2 ox: every other one belongs to this x parity
                  const float4* p8 = reinterpret_cast<const float4*>(gn + (gco[k] + goy[k] * 4 * Wo + 2 * gox[k]));
                  const float4 lo4 = __ldg(p8), hi4 = __ldg(p8 + 1);
                  gv[k] = gpx[k] ? make_float4(lo4.y, lo4.w, hi4.y, hi4.w) : make_float4(lo4.x, lo4.z, hi4.x, hi4.z);
                }
              }
              gox[k] += WG_CK;
              while (gox[k] >= pitch) { gox[k] -= pitch; ++goy[k]; }
            }
          }
        }
        // ---- patch chunk: convert + store
        {
          const int s = pg % a.NR, pk = pg / a.NR;
          const long long d_t2 = dbg_on ? clock64() : 0;
          if (pk > 0) mbar_wait(patch_empty + s, (pk - 1) & 1);
          if (dbg_on) { d_load += d_t2 - d_t1; d_wait_p += clock64() - d_t2; }
          if (any_col && !(a.dbg_flags & 2)) {
            const int rr = s * WG_CK + lane;
            const int fo = rr * 32 + ((((e0 >> 3) ^ (rr & 3)) << 3) | (e0 & 7));
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              if (sl < nslot) {
                float hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  float sc_ = csc[k], sh_ = csh[k];
                  if (nblk > 1 && sl > 0) {              // row of block sl: its own channel
                    const int ci = a.col_ci[32 * sl + e0 + k], c = cc * a.cchunk + ci;
                    const bool okc = ci >= 0 && c < climit;
                    sc_ = okc ? ss_tab[c] : 0.f; sh_ = okc ? ss_tab[WG_MAX_CIN + c] : 0.f;
                  }
                  float v = fmaf(xv[k][sl], sc_, sh_);
                  if (a.relu_in) v = fmaxf(v, 0.f);
                  wg_split(v, hi[k], lo[k]);
                }
                float* ph = ring + (size_t)sl * a.plane_floats + fo;
                float* plo = ph + (size_t)nslot * a.plane_floats;
                *reinterpret_cast<float4*>(ph) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(plo) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                if (s == 0 && lane < 8) {              // mirror of the first 8 ring rows behind the ring end
                  *reinterpret_cast<float4*>(ph + a.ring * 32) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                  *reinterpret_cast<float4*>(plo + a.ring * 32) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
              }
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(patch_full + s);
        }
        // ---- dO chunk: split + store
        if (jg >= 0) {
          const int s = hg % WG_NG, hk = hg / WG_NG;
          const long long d_t3 = dbg_on ? clock64() : 0;
          if (hk > 0) mbar_wait(gy_empty + s, (hk - 1) & 1);
          if (dbg_on) d_wait_g += clock64() - d_t3;
          float* gh = gyb + (size_t)s * a.gy_slot_floats;
          float* gl = gh + (a.gy_slot_floats >> 1);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int uu = tid + k * WG_CONVERTERS;
            if (uu < gy_units) {
              const int co = uu >> 3, q4 = uu & 7;
              float4 h, l;
              wg_split(gv[k].x, h.x, l.x); wg_split(gv[k].y, h.y, l.y); wg_split(gv[k].z, h.z, l.z); wg_split(gv[k].w, h.w, l.w);
              const int fo = (q4 * a.mrows + co) * 4;
              *reinterpret_cast<float4*>(gh + fo) = h;
              *reinterpret_cast<float4*>(gl + fo) = l;
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(gy_full + s);
          ++hg;
        }
      }
    }
    if (dbg_on) { a.dbg[0] = clock64() - d_t0; a.dbg[1] = d_wait_p; a.dbg[2] = d_wait_g; a.dbg[3] = d_load; a.dbg[4] = pg; }
  } else if (warp == WG_WARP_MMA) {
    // =============================== MMA issuing warp ===============================
    const bool leader = elect_one();
    // D = f32, A = B = tf32; the patch ring is the MN-major operand: B (bit 16) normally, A (bit 15) with swapped roles
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a.swap ? (1u << 15) : (1u << 16)) | ((uint32_t)(a.NC >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);
    const int swap = a.swap;
    const uint32_t ring_base = smem_u32(ring), gy_base = smem_u32(gyb);
    const uint32_t plane_bytes = (uint32_t)a.plane_floats * 4u, lo_bytes = (uint32_t)(a.npl * a.nblk) * plane_bytes;
    const uint32_t tap_plane_bytes = (uint32_t)a.nblk * plane_bytes;      // parity planes are nblk slots apart
    const uint32_t gy_slot_bytes = (uint32_t)a.gy_slot_floats * 4u, gy_lo_bytes = gy_slot_bytes >> 1;
    // low descriptor words (start address >> 4 | LBO >> 4 << 16) of dO slot 0 / ring plane 0; every operand lies below 256 KB, so adding
    // 16-byte offsets to the low word never carries into the LBO field
    const uint32_t gy_lo0 = (gy_base >> 4) | ((uint32_t)a.mrows << 16);
    const uint32_t ring_lo0 = (ring_base >> 4) | ((plane_bytes >> 4) << 16);
    const int ring_rows = a.ring, ntaps = a.ntaps, NC = a.NC, NR = a.NR, lead = a.lead, nch_u = a.nch_u, pcu = a.pcu;
    const int nacc = a.nacc, cols = a.cols;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && leader;
    long long d_wait_f = 0, d_wait_a = 0, d_iss = 0;
    const long long d_t0 = dbg_on ? clock64() : 0;
    int pg = 0, hg = 0;
    for (int u = u_begin; u < u_end; ++u) {
      for (int j = 0; j < nch_u; ++j, ++hg) {
        const int sg = hg / WG_FC, sc = hg - sg * WG_FC;          // accumulation segment and chunk inside it
        const int as = nacc == 2 ? (sg & 1) : 0, ak = nacc == 2 ? (sg >> 1) : sg;
        const long long d_t1 = dbg_on ? clock64() : 0;
        if (sc == 0 && ak > 0) mbar_wait(acc_free + as, (ak - 1) & 1);
        const long long d_t2 = dbg_on ? clock64() : 0;
        const bool first = sc == 0;
        const uint32_t acc_base = tmem + (uint32_t)(as * cols);
        const int pneed = pg + j + lead;
        mbar_wait(patch_full + pneed % NR, (pneed / NR) & 1);
        const int gs = hg % WG_NG;
        mbar_wait(gy_full + gs, (hg / WG_NG) & 1);
        tc_fence_after();
        const long long d_t3 = dbg_on ? clock64() : 0;
        const int base_row = ((pg + j) % NR) * WG_CK;
        if (leader && !(a.dbg_flags & 4)) {
          if (swap) wg_issue_chunk<true>(a, gy_lo0 + (uint32_t)gs * (gy_slot_bytes >> 4), ring_lo0, base_row, acc_base, idesc, first);
          else wg_issue_chunk<false>(a, gy_lo0 + (uint32_t)gs * (gy_slot_bytes >> 4), ring_lo0, base_row, acc_base, idesc, first);
        }
        if (leader) {
          umma_commit(gy_empty + gs);
          umma_commit(patch_empty + (pg + j) % NR);
          if (j == nch_u - 1)
            for (int t = 1; t <= lead; ++t) umma_commit(patch_empty + (pg + j + t) % NR);
          if (sc == WG_FC - 1 || hg == nchunks_cta - 1) umma_commit(acc_full + as);
        }
        if (dbg_on) { d_wait_a += d_t2 - d_t1; d_wait_f += d_t3 - d_t2; d_iss += clock64() - d_t3; }
        __syncwarp();
      }
      pg += pcu;
    }
    if (dbg_on) { a.dbg[8] = clock64() - d_t0; a.dbg[9] = d_wait_a; a.dbg[10] = d_wait_f; a.dbg[11] = d_iss; a.dbg[12] = hg; }
    __syncwarp();
  } else {
    // =============================== drain warps ===============================
    // TMEM lane quarter = warp % 4.  Segment sg: slab (+)= accumulators; the last segment adds slab + accumulators into dW64.
    const int q = warp & 3;
    const int row = a.M == 128 ? q * 32 + lane : (lane < 16 ? q * 16 + lane : -1);     // accumulator row of this TMEM lane
    // normal roles: row = output channel, column = (tap, ring column); swapped: row = input channel of the chunk, column = (tap, co)
    const int co = row;
    const int ci_row = (a.swap && row >= 0 && a.col_ci[row] >= 0) ? cc * a.cchunk + a.col_ci[row] : -1, ky_row = (a.swap && row >= 0) ? a.col_dy[row] : 0;
    const bool row_ok = row >= 0 && (a.swap ? (ci_row >= 0 && ci_row < a.cin) : row < a.cout);
    float* slab = a.slab + (size_t)blockIdx.x * a.cols * 128 + q * 32 + lane;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && q == 0 && lane == 0;
    long long d_wait = 0;
    const long long d_t0 = dbg_on ? clock64() : 0;
    for (int sg = 0; sg < nseg; ++sg) {
      const int as = a.nacc == 2 ? (sg & 1) : 0, ak = a.nacc == 2 ? (sg >> 1) : sg;
      const long long d_t1 = dbg_on ? clock64() : 0;
      mbar_wait<128>(acc_full + as, ak & 1);
      if (dbg_on) d_wait += clock64() - d_t1;
      tc_fence_after();
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * a.cols);
      const bool last = sg == nseg - 1;
      for (int c0 = 0; c0 < a.cols; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!last) {
          // fire-and-forget: plain stores for the first segment, RED.ADD.F32 (round-to-nearest, no load round trip) afterwards
          if (sg == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) slab[(size_t)(c0 + k) * 128] = v[k];
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) atomicAdd(slab + (size_t)(c0 + k) * 128, v[k]);
          }
        } else if (row_ok) {
          const int t = c0 / a.NC, eb = c0 - t * a.NC;           // NC is a multiple of 8; a 16-column batch may span two taps
          float o[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) o[k] = sg > 0 ? __ldcg(slab + (size_t)(c0 + k) * 128) : 0.f;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            int tt = t, e = eb + k;
            if (e >= a.NC) { e -= a.NC; ++tt; }
            if (a.swap) {
              if (e < a.cout && tt < a.ntaps)
                atomicAdd(a.dW + ((long long)e * a.cin + ci_row) * (a.kh * a.kw) + (a.packed ? ky_row * a.kw + tt : tt), (double)v[k] + (double)o[k]);
            } else {
              const int cil = a.col_ci[e];
              const int ci = cil < 0 ? -1 : cc * a.cchunk + cil;
              if (ci >= 0 && ci < a.cin && tt < a.ntaps) {
                if (FOLD) {
                  // folded tap (ty, tx) of parity (py, px) is the sum of the original taps with (py + ky + 1) >> 1 == ty, (px + kx + 1) >> 1 == tx
                  const int par = co / a.cout_fold, cr = co - par * a.cout_fold, py = par >> 1, px = par & 1, ty = tt / 3, tx = tt - ty * 3;
                  const double val = (double)v[k] + (double)o[k];
                  for (int ky = 0; ky < 3; ++ky) {
                    if (((py + ky + 1) >> 1) != ty) continue;
                    for (int kx = 0; kx < 3; ++kx)
                      if (((px + kx + 1) >> 1) == tx) atomicAdd(a.dW + ((long long)cr * a.cin + ci) * 9 + ky * 3 + kx, val);
                  }
                } else {
                  const int tap = a.packed ? a.col_dy[e] * a.kw + tt : tt;
                  atomicAdd(a.dW + ((long long)co * a.cin + ci) * (a.kh * a.kw) + tap, (double)v[k] + (double)o[k]);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free + as);
    }
    if (dbg_on) { a.dbg[5] = clock64() - d_t0; a.dbg[6] = d_wait; a.dbg[7] = nseg; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols));
}

// ------------------------------------------------------------------------------------------ host side
static int wg_round_up(int v, int m) { return (v + m - 1) / m * m; }
static int wg_floor_div(int a, int b) { int q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }

static bool wg_plan(const ConvDesc& d, WgTcArgs& a, size_t* smem_out) {
  if (d.dilate || d.pad_override >= 0) return false;
  if (d.up && d.stride == 1 && d.kh == 3 && d.kw == 3 && d.H > 1 && 4 * d.cout <= 128 && 4 * d.cout > 8 && !arpde_dev_env("ARPDE_WG_NO_FOLD")) {
    ConvDesc e = d; e.up = 0; e.cout = 4 * d.cout;         // equivalent low-resolution conv, parity-major channels (conv_tc.cu upfold)
    // (ky, ci)-packed rings (cin * kh <= 32) are not folded: the fold's drain maps accumulator columns to taps as one column per
    // channel (measured wrong for cin = 9, 10; found by the (2,3,2)-topology backward test) -- those layers take the unfolded paths below
    if (wg_plan(e, a, smem_out) && !a.swap && !a.packed) { a.fold = 1; a.cout_fold = d.cout; return true; }
  }
  if (d.H == 1 && d.kh == 1) return false;                       // 1D nets stay on the FFMA kernel
  if (d.stride != 1 && d.stride != 2) return false;
  if (d.up && d.stride != 1) return false;
  if (d.cout > 128 || d.cin > WG_MAX_CIN) return false;
  if (d.kh > 7 || d.kw > 7) return false;
  memset(&a, 0, sizeof(a));
  a.N = d.N; a.cin = d.cin; a.H = d.H; a.W = d.W; a.up_w = d.up ? 1 : 0; a.up_h = a.up_w; a.sy = a.sx = d.stride;
  a.cout = d.cout; a.kh = d.kh; a.kw = d.kw;
  const int Hc = d.H << a.up_h, Wc = d.W << a.up_w;
  if (Hc % a.sy || Wc % a.sx) return false;
  a.Ho = Hc / a.sy; a.Wo = Wc / a.sx;
  if (a.Wo % 4 || a.Wo < 16 || a.Ho < 8) return false;
  const int pad_y = d.kh == 1 ? 0 : (d.kh == 4 ? 1 : (d.kh - 1) / 2), pad_x = d.kw == 1 ? 0 : (d.kw == 4 ? 1 : (d.kw - 1) / 2);
  int min_ry = 1 << 30, max_ry = -(1 << 30), min_rx = 1 << 30, max_rx = -(1 << 30);
  int ry[8], py[8], rx[8], px[8];
  for (int k = 0; k < d.kh; ++k) { const int dy = k - pad_y; py[k] = ((dy % a.sy) + a.sy) % a.sy; ry[k] = wg_floor_div(dy - py[k], a.sy); if (ry[k] < min_ry) min_ry = ry[k]; if (ry[k] > max_ry) max_ry = ry[k]; }
  for (int k = 0; k < d.kw; ++k) { const int dx = k - pad_x; px[k] = ((dx % a.sx) + a.sx) % a.sx; rx[k] = wg_floor_div(dx - px[k], a.sx); if (rx[k] < min_rx) min_rx = rx[k]; if (rx[k] > max_rx) max_rx = rx[k]; }
  a.min_ry = min_ry; a.min_rx = min_rx;
  const int span_y = max_ry - min_ry, span_x = max_rx - min_rx;
  (void)span_y;
  a.pitch = wg_round_up(a.Wo + span_x, 4);
  a.npl = a.sy * a.sx;
  a.nblk = 1;
  // roles: cout on the accumulator rows (M) unless it is tiny -- then the input channels go there and cout becomes the MMA's N
  for (int e = 0; e < 128; ++e) { a.col_ci[e] = -1; a.col_dy[e] = 0; }
  const bool can_pack = a.sx == 1 && d.kh > 1;
  a.swap = (d.cout <= 8 && a.sx == 1 && (d.cin >= 32 || (can_pack && d.cin * d.kh >= 32 && d.cin * d.kh <= 128))) ? 1 : 0;
  if (!a.swap && d.cout < 16) return false;
  if (a.swap) {
    a.packed = (d.cin < 32) ? 1 : 0;
    int rows;
    if (a.packed) { a.cchunk = d.cin; a.ncc = 1; rows = d.cin * d.kh; a.ntaps = d.kw; }
    else { a.cchunk = d.cin < 128 ? d.cin : 128; a.ncc = (d.cin + a.cchunk - 1) / a.cchunk; rows = a.cchunk; a.ntaps = d.kh * d.kw; }
    a.M = rows <= 64 ? 64 : 128; a.nblk = a.M / 32;
    a.NC = wg_round_up(d.cout, a.M == 128 ? 16 : 8);
    for (int r = 0; r < rows; ++r) { a.col_ci[r] = (signed char)(a.packed ? r % d.cin : r); a.col_dy[r] = (signed char)(a.packed ? r / d.cin : 0); }
  } else {
    a.M = d.cout <= 64 ? 64 : 128;
    const int nmult = a.M == 128 ? 16 : 8;
    a.packed = (can_pack && d.cin * d.kh <= 32) ? 1 : 0;
    if (a.packed) {
      a.ncc = 1; a.cchunk = d.cin; a.NC = wg_round_up(d.cin * d.kh, nmult); a.ntaps = d.kw;
      for (int e = 0; e < d.cin * d.kh; ++e) { a.col_ci[e] = (signed char)(e % d.cin); a.col_dy[e] = (signed char)(e / d.cin); }
    } else {
      a.ntaps = d.kh * d.kw;
      a.cchunk = d.cin < 32 ? d.cin : 32;
      while (a.cchunk > 8 && a.ntaps * wg_round_up(a.cchunk, nmult) > 512) a.cchunk = wg_round_up(a.cchunk, 16) / 2;   // every tap keeps its own TMEM columns
      a.ncc = (d.cin + a.cchunk - 1) / a.cchunk; a.NC = wg_round_up(a.cchunk, nmult);
      for (int e = 0; e < a.cchunk; ++e) a.col_ci[e] = (signed char)e;
    }
  }
  // MMA taps: packed layers keep the x taps only (the y taps are rows/columns of the ring), the others all kh*kw
  if (a.packed) {
    for (int kx = 0; kx < d.kw; ++kx) { a.tap_plane[kx] = 0; a.tap_off[kx] = rx[kx] - min_rx; }
  } else {
    for (int ky = 0; ky < d.kh; ++ky)
      for (int kx = 0; kx < d.kw; ++kx) {
        a.tap_plane[ky * d.kw + kx] = py[ky] * a.sx + px[kx];
        a.tap_off[ky * d.kw + kx] = (ry[ky] - min_ry) * a.pitch + (rx[kx] - min_rx);
      }
  }
  if (a.ntaps > WG_MAX_TAPS || a.ntaps * a.NC > 512) return false;
  int max_off = 0;
  for (int t = 0; t < a.ntaps; ++t) if (a.tap_off[t] > max_off) max_off = a.tap_off[t];
  a.lead = (WG_CK - 1 + max_off) / WG_CK;
  // row blocks: the finest split of a sample whose blocks are whole chunks of 32 positions and at least 3 x the halo
  a.RB = 1;
  for (int rb = a.Ho; rb >= 1; --rb) {
    if (a.Ho % rb) continue;
    const int rbh = a.Ho / rb;
    if ((rbh * a.pitch) % WG_CK) continue;
    if (rbh * a.pitch / WG_CK < 3 * a.lead || rbh * a.pitch / WG_CK < 8) continue;
    a.RB = rb; break;
  }
  a.RBH = a.Ho / a.RB;
  if ((a.RBH * a.pitch) % WG_CK) return false;
  a.nch_u = a.RBH * a.pitch / WG_CK; a.pcu = a.nch_u + a.lead;
  a.mrows = (a.swap ? a.NC : d.cout) | 1;                                          // odd row count: the 8 position groups of a dO chunk hit distinct banks
  a.gy_slot_floats = 2 * 8 * a.mrows * 4;
  // an M-row MMA reads rows [0, M) of every 16-byte K plane: the over-read of the last plane must stay inside the slot ring
  const size_t gy_bytes = (size_t)WG_NG * a.gy_slot_floats * 4 + (size_t)a.M * 16;
  const size_t budget = 226 * 1024;
  a.NR = 0;
  for (int nr = a.lead + 6; nr >= a.lead + 2; --nr) {
    if (nr > WG_MAX_NR) continue;
    const size_t plane = (size_t)(nr * WG_CK + 8) * 128;
    if (WG_HEADER + 2 * a.npl * a.nblk * plane + gy_bytes <= budget) { a.NR = nr; break; }
  }
  if (!a.NR) return false;
  a.ring = a.NR * WG_CK;
  a.plane_floats = (a.ring + 8) * 32;
  a.cols = wg_round_up(a.ntaps * a.NC, 16);
  if (a.cols > 512) return false;
  a.nacc = 2 * a.cols <= 512 ? 2 : 1;
  int p2 = 32; while (p2 < a.nacc * a.cols) p2 <<= 1;
  a.tmem_cols = p2;
  a.units = d.N * a.RB;
  int slices = arpde_num_sms() / a.ncc; if (slices < 1) slices = 1; if (slices > a.units) slices = a.units;
  a.slices = slices;
  *smem_out = WG_HEADER + (size_t)2 * a.npl * a.nblk * a.plane_floats * 4 + gy_bytes;
  return true;
}

bool arpde_conv_wgrad_tc_wanted(const ConvDesc& d) {
  if (!arpde_tc_enabled()) return false;
  if (const char* e = arpde_dev_env("ARPDE_WGRAD_TC")) if (e[0] == '0') return false;
  // an MMA costs ~40 cycles whatever M is: with cout < 32 most of the 64 accumulator rows are padding and the FFMA kernel wins
  // (measured, N = 128: LastTransUp conv3x3 32->16 at 64x64: tensor core 0.41 ms, FFMA 0.375 ms)
  WgTcArgs a; size_t smem;
  if (!wg_plan(d, a, &smem)) return false;
  if (!a.fold && d.cout < 32 && d.cout > 8) return false;      // (a folded upsampling layer has 4 * cout real accumulator rows)
  // swapped + packed with four row blocks (LastTransUp's 16 -> 2 5x5 conv: 80 (ky, ci) rows): the converters stage 16 table-driven
  // rows per thread and step and become the bottleneck -- measured 0.43 ms against 0.37 ms for the FFMA kernel (N = 128)
  if (a.swap && a.packed && a.nblk > 2) return false;
  return true;
}

long long arpde_conv_wgrad_tc_scratch_floats(const ConvDesc& d) {
  WgTcArgs a; size_t smem;
  if (!wg_plan(d, a, &smem)) return 0;
  return (long long)a.ncc * a.slices * a.cols * 128;
}

// same contract as arpde_conv_wgrad_launch (conv_bwd.cu): dW (float) = sum over batch and positions; dW64 = scratch of cout*cin*kh*kw
// doubles; scratch = arpde_conv_wgrad_tc_scratch_floats(d) floats (uninitialised is fine)
int arpde_conv_wgrad_tc_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                               long long dO_bs, double* dW64, float* dW, float* scratch, long long scratch_floats, cudaStream_t stream) {
  ARPDE_CHECK_ARG(x && dO && dW64 && dW && scratch, "conv_wgrad_tc: null pointer");
  WgTcArgs a; size_t smem = 0;
  if (!wg_plan(d, a, &smem)) { arpde_set_error("conv_wgrad_tc: layer not supported by the tensor-core kernel"); return ARPDE_ERR_UNSUPPORTED; }
  ARPDE_CHECK_ARG((((uintptr_t)dO) & 15) == 0 && (dO_bs & 3) == 0, "conv_wgrad_tc: output gradient must be 16-byte aligned");
  const long long need = (long long)a.ncc * a.slices * a.cols * 128;
  ARPDE_CHECK_ARG(scratch_floats >= need, "conv_wgrad_tc: scratch too small (%lld < %lld floats)", scratch_floats, need);
  a.slab = scratch; a.dbg = arpde_tc_debug_buffer();
  if (const char* e = arpde_dev_env("ARPDE_WG_DBG")) a.dbg_flags = atoi(e);
  a.x = x; a.x_bs = x_bs; a.scale = scale; a.shift = shift; a.relu_in = d.relu_in; a.dO = dO; a.dO_bs = dO_bs; a.dW = dW64;
  const long long wn = (long long)d.cout * d.cin * d.kh * d.kw;
  ARPDE_CUDA(cudaMemsetAsync(dW64, 0, sizeof(double) * wn, stream));
  if (d.N > 0) {
    static std::mutex attr_mutex;
    static bool attr_done[64] = {false};
    int dev = 0, optin = 0;
    ARPDE_CUDA(cudaGetDevice(&dev));
    ARPDE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    ARPDE_CHECK_ARG(smem <= (size_t)optin, "conv_wgrad_tc: plan needs %zu bytes of shared memory, device allows %d", smem, optin);
    {
      std::lock_guard<std::mutex> lock(attr_mutex);
      if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        ARPDE_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        ARPDE_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
      }
    }
    if (a.fold) conv_wgrad_tc_kernel<true><<<(unsigned)(a.ncc * a.slices), WG_THREADS, smem, stream>>>(a);
    else conv_wgrad_tc_kernel<false><<<(unsigned)(a.ncc * a.slices), WG_THREADS, smem, stream>>>(a);
    ARPDE_LAUNCH_CHECK();
  }
  return arpde_d2f_launch(dW64, dW, wn, stream);
}

// Diagnostic (host only): the tiling conv_wgrad_tc_kernel would use.  out[0..25] = pitch, min_ry, min_rx, RB, RBH, nch_u, lead, pcu, NR,
// npl, NC, ncc, cchunk, packed, ntaps, M, mrows, tmem_cols, slices(N=1), smem bytes, Ho, Wo, sy, sx, swap, nblk, fold, cout_fold; out[32+t] = tap plane,
// out[81+t] = tap offset, out[130+e] = column channel, out[258+e] = column y-tap (e < 128).
extern "C" int arpde_conv_wgrad_tc_plan(int cin, int H, int W, int cout, int kh, int kw, int stride, int up, int32_t* out) {
  ARPDE_CHECK_ARG(out, "conv_wgrad_tc_plan: null output");
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1; d.N = 1; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride;
  d.up = up; d.relu_in = 0; d.act = 0; d.y_ctot = cout; d.y_coff = 0; d.n_per_group = 1;
  WgTcArgs a; size_t smem = 0;
  if (!wg_plan(d, a, &smem)) { arpde_set_error("conv_wgrad_tc_plan: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  out[26] = a.fold; out[27] = a.cout_fold;
  const int v[26] = {a.pitch, a.min_ry, a.min_rx, a.RB, a.RBH, a.nch_u, a.lead, a.pcu, a.NR, a.npl, a.NC, a.ncc, a.cchunk, a.packed, a.ntaps,
                     a.M, a.mrows, a.tmem_cols, a.slices, (int)smem, a.Ho, a.Wo, a.sy, a.sx, a.swap, a.nblk};
  for (int i = 0; i < 26; ++i) out[i] = v[i];
  for (int t = 0; t < WG_MAX_TAPS; ++t) { out[32 + t] = t < a.ntaps ? a.tap_plane[t] : 0; out[81 + t] = t < a.ntaps ? a.tap_off[t] : 0; }
  for (int e = 0; e < 128; ++e) { out[130 + e] = a.col_ci[e]; out[258 + e] = a.col_dy[e]; }
  return ARPDE_OK;
}

// C ABI: weight gradient of one circular convolution on the tensor cores (stand-alone entry point, used by the parity tests).
extern "C" int64_t arpde_conv_wgrad_tc_scratch(int N, int cin, int H, int W, int cout, int kh, int kw, int stride, int up) {
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1; d.N = N; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride;
  d.up = up; d.relu_in = 0; d.act = 0; d.y_ctot = cout; d.y_coff = 0; d.n_per_group = N > 0 ? N : 1;
  return arpde_conv_wgrad_tc_scratch_floats(d);
}

extern "C" int arpde_conv_wgrad_tc(const float* x, int64_t x_bstride, const float* scale, const float* shift, const float* dO,
                                   int64_t dO_bstride, double* dW64, float* dW, int N, int cin, int H, int W, int cout, int kh, int kw,
                                   int stride, int up, int relu_in, float* scratch, int64_t scratch_floats, arpde_stream_t stream_) {
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1; d.N = N; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride;
  d.up = up; d.relu_in = relu_in; d.act = 0; d.y_ctot = cout; d.y_coff = 0; d.n_per_group = N > 0 ? N : 1;
  ARPDE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_wgrad_tc: scale/shift must both be given or both be null");
  return arpde_conv_wgrad_tc_launch(d, x, x_bstride, scale, shift, dO, dO_bstride, dW64, dW, scratch, scratch_floats, (cudaStream_t)stream_);
}
