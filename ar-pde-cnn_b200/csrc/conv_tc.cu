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
//         D[:, N:2N] += A_lo * W_hi^T
//     where the hi*lo and lo*hi corrections accumulate in the SECOND column half, so that the main hi*hi chain is not
//     perturbed by them; the epilogue adds the two halves in fp32 (lo*lo ~ 2^-22 relative is dropped);
//   * persistent CTAs (one or two per SM) walk a contiguous range of work items (sample, group of `ntile` M tiles);
//     warp roles: 8 staging warps, 4 epilogue warps, 1 MMA-issuing warp, 1 weight-copy warp, connected by mbarrier
//     rings: 2 patch slots, 2 accumulator slots in TMEM (the epilogue of item i overlaps the MMAs of item i+1) and
//     either a 2-slot weight ring fed by cp.async.bulk (large layers) or the whole split weight image resident in
//     shared memory, reloaded only when the weight group (SWAG sample) changes;
//   * weights arrive pre-split (hi/lo) and pre-arranged in the UMMA layout by conv_tc_wprep_kernel.
// M tiles that straddle image rows waste only the (pitch - Wo)/pitch halo positions.
// Nearest x2 upsampling followed by a 3x3 conv is folded ("upfold"): an output pixel of parity (py, px) only sees a 2x2 block of
// distinct low-resolution pixels, so the layer is computed as ONE 3x3 conv on the low-resolution grid with 4*cout output channels
// (parity-major) whose weights are sums of the original taps (row offsets: py = 0: {w0 | w1+w2 | -}, py = 1: {- | w0+w1 | w2}); the
// epilogue scatters channel (parity, co) to pixel (2y+py, 2x+px).  4x fewer M tiles, N 4x larger: 2.7x fewer tensor-core cycles for
// LastTransUp's 32->16 layer and a 4x smaller patch to stage.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"
#include <stdlib.h>
#include <mutex>
#include <cuda_fp16.h>
#include <cuda.h>          // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint

#define TC_MAX_TAPS 49
#define TC_STAGER_WARPS 8
#define TC_STAGERS (32 * TC_STAGER_WARPS)
#define TC_EPI_WARPS 4
#define TC_WARP_EPI0 TC_STAGER_WARPS
#define TC_WARP_ISSUER (TC_STAGER_WARPS + TC_EPI_WARPS)
#define TC_ISSUERS 2                   // MMA-issuing warps: issuer i owns the M tiles t = i (mod TC_ISSUERS) of every item
#define TC_WARP_WPROD (TC_WARP_ISSUER + TC_ISSUERS)     // weight copy warp
#define TC_WARP_IPROD (TC_WARP_WPROD + 1)              // input copy warp (TMA rows of the input tensor)
#define TC_THREADS (32 * (TC_WARP_IPROD + 1))
#define TC_KP_MAX 8                    // a staging thread owns at most 8 patch positions (PHP <= 2048)
#define TC_MAX_CIN 256                 // input channels (padded to the channel chunk) the BatchNorm scale/shift table holds
#define TC_SMEM_HEADER (256 + 4 * TC_MAX_CIN * 4)   // barriers + tmem slot, then the two scale/shift tables

struct ConvTcArgs {
  const float* x; long long x_bs;
  const float* scale; const float* shift; long long ss_gs;
  const float* wprep; long long wprep_gs;
  float* y; double* stats;
  int cin, H, W, up_h, up_w, Hc, Wc, sy, sx;
  int relu_in, act;
  int y_ctot, y_coff, cout, Ho, Wo;
  int n_per_group, N;
  int pitch, min_ry, min_rx, ntile, total_tiles, ctas_per_sample, PP, PHP, nplanes;
  int cchunk, nc4_shift, ncc, tchunk, ntc, ntaps, coutP, tmem_cols;
  int patch_slot_floats, w_slot_floats, nslot_p;
  int resident;              // 1: the whole split weight image of a group stays in shared memory
  int nacc, acc_cols;        // accumulator ring in TMEM: nacc slots of acc_cols = ntile * tile_cols columns
  int split_terms, tile_cols; // 1: issuer 0 issues hi*[hi;lo], issuer 1 issues lo*hi into a third column range (tile_cols = 3*coutP)
  int kp;                    // patch positions per staging thread = ceil(PHP / 256)
  int pair;                  // stride-2 layer staged with float2 loads (both x parities at once)
  int tma;                   // 1: input rows arrive by TMA into the raw ring and the staging warps convert smem -> smem
  int row_mode;              // 0: stride 1, 1: nearest x2 upsampled input, 2: stride 2 (parity planes)
  int dilate;                // input is zero-inserted x2 (data gradient of a stride-2 layer): addressed like row_mode 1, odd rows/columns are 0
  int RS, raw_slot_floats;   // source rows per (item, channel chunk) and floats per raw ring slot [RS][cchunk][W]
  int upfold, cout_fold;     // upfold: `cout` = 4 * cout_fold parity-major channels on the low-resolution grid, scattered by the epilogue
  uint4* z; long long z_bs16; const float* z_scale; const float* z_shift; long long z_ss_gs;   // operand-format output (ZOut), else z == nullptr
  long long* dbg;            // optional cycle counters of CTA 0 (arpde_conv_tc_set_debug), else nullptr
  int dbg_flags;             // experiments: 1 = skip patch staging work, 2 = skip epilogue loads/stores, 4 = skip MMAs
  int tap_plane[TC_MAX_TAPS];
  int tap_off[TC_MAX_TAPS];
  int tap_aoff[TC_MAX_TAPS];  // (tap_plane * cchunk/4 * PP + tap_off): A-operand offset of the tap in 16-byte units
};

#include "tc_ptx.cuh"

// debug cycle accounting (CTA 0 only, one lane per role): DBG_T(slot, statement) adds the cycles `statement` took to dbg[slot]
#define DBG_T(slot, stmt)                                                         \
  do {                                                                            \
    const long long t__ = (a.dbg != nullptr) ? clock64() : 0;                     \
    stmt; /* always executed in warp-uniform control flow (tcgen05.ld is .sync.aligned) */ \
    if (dbg_on) dbg_acc[slot] += clock64() - t__;                                 \
  } while (0)

// ------------------------------------------------------------------------------------------ weight preparation
// w [groups][cout][cin][ntaps]  ->  out [groups][ncc][ntc][tchunk][cchunk/4][2*coutP][4]: the exact shared-memory image
// of every (channel chunk, tap chunk) weight slot: rows [0,coutP) = tf32 hi part, rows [coutP,2coutP) = lo part.
__global__ void __launch_bounds__(256) conv_tc_wprep_kernel(const float* __restrict__ w, long long w_gs, float* __restrict__ out,
                                                            long long out_gs, int groups, int cout, int cin, int ntaps,
                                                            int cchunk, int ncc, int tchunk, int ntc, int coutP, int cout_fold) {
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
    if (co < cout && c < cin) {
      if (cout_fold == 0) {
        v = __ldg(w + (long long)g * w_gs + ((long long)co * cin + c) * ntaps + tap);
      } else {
        // upfold: channel co = parity * cout_fold + cr, low-resolution tap (ty, tx) = sum of the 3x3 taps that land on it
        const int par = co / cout_fold, cr = co - par * cout_fold, py = par >> 1, px = par & 1, ty = tap / 3, tx = tap - ty * 3;
        const float* wp = w + (long long)g * w_gs + ((long long)cr * cin + c) * 9;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          if (((py + ky + 1) >> 1) != ty) continue;           // low-res row offset of original tap ky: floor((py + ky - 1) / 2) + 1
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            if (((px + kx + 1) >> 1) == tx) v += __ldg(wp + ky * 3 + kx);
        }
      }
    }
    const float hi = tf32_rna(v);
    out[(long long)g * out_gs + (idx - (long long)g * per_group)] = row < coutP ? hi : tf32_rna(v - hi);
  }
}

// ------------------------------------------------------------------------------------------ patch staging
// Input patch -> shared memory (hi and lo planes), one channel chunk at a time.  Staging thread `stid` owns the patch
// positions stid + 256*k (k < KP) of every "unit" = (plane, 4-channel group) -- or, for stride-2 layers (PAIR), of every
// (pair of x-parity planes, 4-channel group), loaded as one float2 per channel so that every fetched sector is used.
// The work is a flat sequence of batches of NB units over (item, channel chunk); the loads of batch i+1 are issued
// before batch i is converted and stored (two register buffers), and they do not wait for the patch slot to be free,
// so global-load latency overlaps the split arithmetic, the shared-memory stores and the barrier hand-off.
// (History: v1 consumed each load right away -> 60% long_scoreboard stalls; v2 batched ~32 loads per thread but ran
//  load -> wait -> convert -> store -> fence strictly in sequence: 2.9k cycles per batch, every layer staging-bound.)
struct StageCursor { int it, cc, ub; };

template <int KP, int NB, bool PAIR>
__device__ __forceinline__ void sb_load(const ConvTcArgs& a, const float* __restrict__ xn, long long HW, int cc, int ub,
                                        const int (&off0)[KP], uint32_t kmask, float (&v)[NB][KP][PAIR ? 8 : 4]) {
  const int nc4 = a.cchunk >> 2;
  const int nunits = (PAIR ? (a.nplanes >> 1) : a.nplanes) * nc4;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int u = ub * NB + i;
    const int pl = u >> a.nc4_shift, c4 = u & (nc4 - 1);
    const int c0 = cc * a.cchunk + c4 * 4;
    const int poff = PAIR ? pl * a.W : (pl >> 1) * a.W + (pl & 1);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c0 + e;
      const bool cok = u < nunits && c < a.cin;
      const float* src = xn + (long long)(cok ? c : 0) * HW + poff;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        const bool ok = cok && ((kmask >> k) & 1u);
        if (PAIR) {
          float2 t = make_float2(0.f, 0.f);
          if (ok) t = __ldg(reinterpret_cast<const float2*>(src + off0[k]));
          v[i][k][e] = t.x; v[i][k][4 + e] = t.y;
        } else {
          v[i][k][e] = ok ? __ldg(src + off0[k]) : 0.f;
        }
      }
    }
  }
}

template <bool RELU>
__device__ __forceinline__ void split_store(const float (&x)[4], const float4 sc, const float4 sh, float4* dh, float4* dl) {
  const float s[4] = {sc.x, sc.y, sc.z, sc.w}, b[4] = {sh.x, sh.y, sh.z, sh.w};
  float hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float t = fmaf(x[e], s[e], b[e]);
    if (RELU) t = fmaxf(t, 0.f);
    // tf32 round-to-nearest (ties away, == cvt.rna for finite values) in two integer ops; the residual t - hi is exact in
    // fp32 and is left unrounded: the tensor core drops its low 13 bits (< 2^-22 relative to t)
    hi[e] = __uint_as_float((__float_as_uint(t) + 0x1000u) & 0xffffe000u);
    lo[e] = t - hi[e];
  }
  *dh = make_float4(hi[0], hi[1], hi[2], hi[3]);
  *dl = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

template <int KP, int NB, bool RELU, bool PAIR>
__device__ __forceinline__ void sb_store(const ConvTcArgs& a, float4* __restrict__ ph, float4* __restrict__ pl, const float* ss_tab,
                                         int cc, int ub, uint32_t kmask, int stid, const float (&v)[NB][KP][PAIR ? 8 : 4]) {
  const int nc4 = a.cchunk >> 2;
  const int nunits = (PAIR ? (a.nplanes >> 1) : a.nplanes) * nc4;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int u = ub * NB + i;
    if (u < nunits) {
      const int plx = u >> a.nc4_shift, c4 = u & (nc4 - 1);
      const int c0 = cc * a.cchunk + c4 * 4;
      const float4 sc = *reinterpret_cast<const float4*>(ss_tab + c0);
      const float4 sh = *reinterpret_cast<const float4*>(ss_tab + TC_MAX_CIN + c0);
      const int pc = (PAIR ? 2 * plx : plx) * nc4 + c4;              // (plane, c4) index of the first plane written
      float4* dh = ph + (size_t)pc * a.PP + stid;
      float4* dl = pl + (size_t)pc * a.PP + stid;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if ((kmask >> k) & 1u) {
          const float x0[4] = {v[i][k][0], v[i][k][1], v[i][k][2], v[i][k][3]};
          split_store<RELU>(x0, sc, sh, dh + k * TC_STAGERS, dl + k * TC_STAGERS);
          if (PAIR) {
            const float x1[4] = {v[i][k][4], v[i][k][5], v[i][k][6], v[i][k][7]};
            split_store<RELU>(x1, sc, sh, dh + (size_t)nc4 * a.PP + k * TC_STAGERS, dl + (size_t)nc4 * a.PP + k * TC_STAGERS);
          }
        }
      }
    }
  }
}

template <int KP, int NB, bool RELU, bool PAIR>
__device__ __noinline__ void stager_loop(const ConvTcArgs& a, float* patch, float* ss_tab, uint64_t* full_p, uint64_t* empty_p,
                                         int it_begin, int it_end, int tid, int lane) {
  constexpr int NV = PAIR ? 8 : 4;
  const long long HW = (long long)a.H * a.W;
  const int nc4 = a.cchunk >> 2;
  const int nunits = (PAIR ? (a.nplanes >> 1) : a.nplanes) * nc4;
  const int nbat = (nunits + NB - 1) / NB;
  const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  long long dbg_acc[4] = {0, 0, 0, 0}, dbg_ld = 0;
  const long long dbg_t0 = dbg_on ? clock64() : 0;
  uint32_t kmask = 0;
#pragma unroll
  for (int k = 0; k < KP; ++k)
    if (tid + k * TC_STAGERS < a.PHP) kmask |= 1u << k;

  // ---- load-side context (item of the load cursor)
  const float* xn = nullptr;
  int off0[KP];
  int g_tab = -1;
  auto enter_item = [&](int it) {
    const int n = it / a.ctas_per_sample, b = it - n * a.ctas_per_sample;
    const int g = n / a.n_per_group;
    const int r0 = (b * a.ntile * 128) / a.pitch;
    xn = a.x + (long long)n * a.x_bs;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      off0[k] = 0;
      const int pos = tid + k * TC_STAGERS;
      if (pos < a.PHP) {
        const int i = pos / a.pitch, j = pos - i * a.pitch;
        int R = (r0 + a.min_ry + i) % a.Ho; if (R < 0) R += a.Ho;
        int C = (a.min_rx + j) % a.Wo; if (C < 0) C += a.Wo;
        off0[k] = ((R * a.sy) >> a.up_h) * a.W + ((C * a.sx) >> a.up_w);
      }
    }
    if (g != g_tab) {
      // BatchNorm scale/shift of this weight group -> shared table (two tables, alternating with the group, because the
      // store side may still be one batch behind in the previous group)
      float* tab = ss_tab + (g & 1) * 2 * TC_MAX_CIN;
      const int cpad = a.ncc * a.cchunk;
      for (int c = tid; c < cpad; c += TC_STAGERS) {
        const bool ok = c < a.cin;
        tab[c] = ok ? (a.scale ? __ldg(a.scale + (long long)g * a.ss_gs + c) : 1.f) : 0.f;
        tab[TC_MAX_CIN + c] = (ok && a.scale) ? __ldg(a.shift + (long long)g * a.ss_gs + c) : 0.f;
      }
      bar_sync_named(1, TC_STAGERS);
      g_tab = g;
    }
  };
  auto advance = [&](StageCursor c) {
    if (++c.ub == nbat) { c.ub = 0; if (++c.cc == a.ncc) { c.cc = 0; ++c.it; } }
    return c;
  };
  // ---- store side of one batch (+ the barrier hand-off at chunk boundaries)
  auto finish = [&](const StageCursor& c, const float (&v)[NB][KP][NV]) {
    const int pseq = (c.it - it_begin) * a.ncc + c.cc;
    const int ps = pseq % a.nslot_p, pk = pseq / a.nslot_p;
    if (c.ub == 0 && pk > 0) DBG_T(0, mbar_wait(empty_p + ps, (pk - 1) & 1));
    const long long t1 = dbg_on ? clock64() : 0;
    float4* ph = reinterpret_cast<float4*>(patch + (size_t)ps * a.patch_slot_floats);
    float4* pl = ph + (a.patch_slot_floats >> 3);
    const int n = c.it / a.ctas_per_sample;
    const float* tab = ss_tab + ((n / a.n_per_group) & 1) * 2 * TC_MAX_CIN;
    if (!(a.dbg_flags & 1)) sb_store<KP, NB, RELU, PAIR>(a, ph, pl, tab, c.cc, c.ub, kmask, tid, v);
    if (dbg_on) dbg_acc[1] += clock64() - t1;
    if (c.ub == nbat - 1) DBG_T(2, fence_async_smem(); __syncwarp(); if (lane == 0) mbar_arrive(full_p + ps));
  };

  float va[NB][KP][NV], vb[NB][KP][NV];
  StageCursor cur = {it_begin, 0, 0};
  if (cur.it < it_end) {
    enter_item(cur.it);
    sb_load<KP, NB, PAIR>(a, xn, HW, cur.cc, cur.ub, off0, kmask, va);
    while (true) {
      StageCursor nxt = advance(cur);
      bool more = nxt.it < it_end;
      if (more) {
        if (nxt.it != cur.it) DBG_T(3, enter_item(nxt.it));
        { const long long tl__ = dbg_on ? clock64() : 0; sb_load<KP, NB, PAIR>(a, xn, HW, nxt.cc, nxt.ub, off0, kmask, vb); if (dbg_on) dbg_ld += clock64() - tl__; }
      }
      finish(cur, va);
      if (!more) break;
      cur = nxt;
      nxt = advance(cur);
      more = nxt.it < it_end;
      if (more) {
        if (nxt.it != cur.it) DBG_T(3, enter_item(nxt.it));
        { const long long tl__ = dbg_on ? clock64() : 0; sb_load<KP, NB, PAIR>(a, xn, HW, nxt.cc, nxt.ub, off0, kmask, va); if (dbg_on) dbg_ld += clock64() - tl__; }
      }
      finish(cur, vb);
      if (!more) break;
      cur = nxt;
    }
  }
  if (dbg_on) { a.dbg[0] = clock64() - dbg_t0; a.dbg[1] = dbg_acc[0]; a.dbg[2] = dbg_acc[1]; a.dbg[3] = dbg_acc[2]; a.dbg[13] = dbg_ld; a.dbg[14] = dbg_acc[3]; }
}

// ------------------------------------------------------------------------------------------ patch staging, TMA path
// The input copy warp brings the source rows of an (item, channel chunk) into a raw ring slot [RS rows][cchunk][W] by
// TMA (one box per row, vertical wrap = row index mod H, channels beyond cin zero-filled by the TMA unit); the staging
// warps only convert shared -> shared: thread `stid` owns patch positions stid + 256*k and, per 4-channel group, reads
// 4 floats (conflict-free: consecutive lanes = consecutive columns), applies BN/ReLU, splits and writes hi/lo float4.
// Horizontal wrap, nearest-upsampling and the stride-2 parity split are constant per-thread index arithmetic.
template <int KP, bool RELU, bool PAIR>
__device__ __noinline__ void converter_loop(const ConvTcArgs& a, const float* raw, float* patch, float* ss_tab, uint64_t* full_p,
                                            uint64_t* empty_p, uint64_t* raw_full, uint64_t* raw_empty, int it_begin, int it_end, int tid,
                                            int lane) {
  const int nc4 = a.cchunk >> 2;
  const int nunits = (PAIR ? (a.nplanes >> 1) : a.nplanes) * nc4;
  const int rowf = a.cchunk * a.W;                     // floats per raw row
  const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  long long dbg_acc[4] = {0, 0, 0, 0};
  const long long dbg_t0 = dbg_on ? clock64() : 0;
  // constant part of the raw offset of every owned position (plane (0,0), channel 0)
  int irow[KP], coff[KP];
  uint32_t kmask = 0, codd = 0;            // codd bit k: the position's column is odd in the zero-inserted (dilated) input space
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    irow[k] = 0; coff[k] = 0;
    const int pos = tid + k * TC_STAGERS;
    if (pos < a.PHP) {
      kmask |= 1u << k;
      const int i = pos / a.pitch, j = pos - i * a.pitch;
      int C = (a.min_rx + j) % a.Wo; if (C < 0) C += a.Wo;
      irow[k] = i;
      coff[k] = (C * a.sx) >> a.up_w;
      if (a.dilate && (C & 1)) codd |= 1u << k;
    }
  }
  int g_tab = -1;
  for (int it = it_begin; it < it_end; ++it) {
    const int lit = it - it_begin;
    const int n = it / a.ctas_per_sample, b = it - n * a.ctas_per_sample;
    const int g = n / a.n_per_group;
    const int rbase = (b * a.ntile * 128) / a.pitch + a.min_ry;
    if (g != g_tab) {        // BatchNorm scale/shift of this weight group -> shared table (padded channels: 0, 0)
      bar_sync_named(1, TC_STAGERS);                    // every staging thread is done with the previous group's table
      const int cpad = a.ncc * a.cchunk;
      for (int c = tid; c < cpad; c += TC_STAGERS) {
        const bool ok = c < a.cin;
        ss_tab[c] = ok ? (a.scale ? __ldg(a.scale + (long long)g * a.ss_gs + c) : 1.f) : 0.f;
        ss_tab[TC_MAX_CIN + c] = (ok && a.scale) ? __ldg(a.shift + (long long)g * a.ss_gs + c) : 0.f;
      }
      bar_sync_named(1, TC_STAGERS);
      g_tab = g;
    }
    int roff[KP];
    uint32_t live = kmask;                 // positions that carry data (dilated input: even row and even column only)
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const int rl = a.row_mode == 1 ? ((irow[k] + (rbase & 1)) >> 1) : (a.row_mode == 2 ? 2 * irow[k] : irow[k]);
      roff[k] = rl * rowf + coff[k];
      if (a.dilate && (((rbase + irow[k]) & 1) || ((codd >> k) & 1u))) live &= ~(1u << k);
    }
    for (int cc = 0; cc < a.ncc; ++cc) {
      const int pseq = lit * a.ncc + cc;
      const int ps = pseq % a.nslot_p, pk = pseq / a.nslot_p;
      const int rs = pseq & 1, rk = pseq >> 1;
      DBG_T(3, mbar_wait(raw_full + rs, rk & 1));
      if (pk > 0) DBG_T(0, mbar_wait(empty_p + ps, (pk - 1) & 1));
      const long long t1 = dbg_on ? clock64() : 0;
      const float* rw = raw + (size_t)rs * a.raw_slot_floats;
      float4* ph = reinterpret_cast<float4*>(patch + (size_t)ps * a.patch_slot_floats) + tid;
      float4* pl = ph + (a.patch_slot_floats >> 3);
      if (!(a.dbg_flags & 1)) {
        for (int u = 0; u < nunits; ++u) {
          const int plx = u >> a.nc4_shift, c4 = u & (nc4 - 1);
          const float4 sc = *reinterpret_cast<const float4*>(ss_tab + cc * a.cchunk + c4 * 4);
          const float4 sh = *reinterpret_cast<const float4*>(ss_tab + TC_MAX_CIN + cc * a.cchunk + c4 * 4);
          // raw offset of (unit, channel 0): 4-channel group, and for stride 2 the parity plane (py -> next source row, px -> next column)
          const int uoff = c4 * 4 * a.W + (PAIR ? plx * rowf : (plx >> 1) * rowf + (plx & 1));
          const int pc = (PAIR ? 2 * plx : plx) * nc4 + c4;
          float v[KP][PAIR ? 8 : 4];
#pragma unroll
          for (int k = 0; k < KP; ++k) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (PAIR) {
                const float2 t2 = ((kmask >> k) & 1u) ? *reinterpret_cast<const float2*>(rw + roff[k] + uoff + e * a.W) : make_float2(0.f, 0.f);
                v[k][e] = t2.x; v[k][4 + e] = t2.y;
              } else {
                v[k][e] = ((live >> k) & 1u) ? rw[roff[k] + uoff + e * a.W] : 0.f;
              }
            }
          }
          float4* dh = ph + (size_t)pc * a.PP;
          float4* dl = pl + (size_t)pc * a.PP;
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            if ((kmask >> k) & 1u) {
              const float x0[4] = {v[k][0], v[k][1], v[k][2], v[k][3]};
              split_store<RELU>(x0, sc, sh, dh + k * TC_STAGERS, dl + k * TC_STAGERS);
              if (PAIR) {
                const float x1[4] = {v[k][4], v[k][5], v[k][6], v[k][7]};
                split_store<RELU>(x1, sc, sh, dh + (size_t)nc4 * a.PP + k * TC_STAGERS, dl + (size_t)nc4 * a.PP + k * TC_STAGERS);
              }
            }
          }
        }
      }
      if (dbg_on) dbg_acc[1] += clock64() - t1;
      DBG_T(2, fence_async_smem(); __syncwarp(); if (lane == 0) { mbar_arrive(full_p + ps); mbar_arrive(raw_empty + rs); });
    }
  }
  if (dbg_on) { a.dbg[0] = clock64() - dbg_t0; a.dbg[1] = dbg_acc[0]; a.dbg[2] = dbg_acc[1]; a.dbg[3] = dbg_acc[2]; a.dbg[13] = 0; a.dbg[14] = dbg_acc[3]; }
}

// ------------------------------------------------------------------------------------------ main kernel
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvTcArgs a, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full_p = reinterpret_cast<uint64_t*>(smem_raw);        // [2] patch slot filled (one arrival per staging warp)
  uint64_t* empty_p = full_p + 2;                                    // [2] patch slot consumed (tcgen05.commit)
  uint64_t* full_w = full_p + 4;                                     // [2] weight slot filled (bulk-copy tx bytes)
  uint64_t* empty_w = full_p + 6;                                    // [2] weight slot consumed (tcgen05.commit)
  uint64_t* tmem_full = full_p + 8;                                  // [2] accumulators of an item complete (tcgen05.commit)
  uint64_t* tmem_empty = full_p + 10;                                // [2] accumulators drained (one arrival per epilogue warp)
  uint64_t* wres_full = full_p + 12;                                 // resident weight image (re)loaded
  uint64_t* raw_full = full_p + 14;                                  // [2] raw input rows landed (TMA tx bytes)
  uint64_t* raw_empty = full_p + 16;                                 // [2] raw slot converted (one arrival per staging warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 192);
  float* ss_tab = reinterpret_cast<float*>(smem_raw + 256);          // [2 groups][scale | shift][TC_MAX_CIN]
  float* raw = reinterpret_cast<float*>(smem_raw + TC_SMEM_HEADER);  // [2 slots][RS][cchunk][W] (TMA path only)
  float* patch = raw + (a.tma ? 2 * (size_t)a.raw_slot_floats : 0);
  float* wbuf = patch + (size_t)a.nslot_p * a.patch_slot_floats;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int items_total = a.N * a.ctas_per_sample;
  const int per_cta = (items_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int it_begin = (int)blockIdx.x * per_cta;
  const int it_end = min(items_total, it_begin + per_cta);
  const int nc4 = a.cchunk >> 2, nc8 = a.cchunk >> 3, N2 = 2 * a.coutP;
  const int nchunks = a.ncc * a.ntc;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_p + s, TC_STAGER_WARPS); mbar_init(empty_p + s, TC_ISSUERS); mbar_init(full_w + s, 1); mbar_init(empty_w + s, TC_ISSUERS);
      mbar_init(tmem_full + s, TC_ISSUERS); mbar_init(tmem_empty + s, TC_EPI_WARPS);
    }
    mbar_init(wres_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, TC_STAGER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_WARP_ISSUER) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < TC_STAGER_WARPS) {
    // =============================== staging warps ===============================
#define TC_STAGE(KP_, NB_, PAIR_)                                                                                        \
  do {                                                                                                                   \
    if (a.relu_in) stager_loop<KP_, NB_, true, PAIR_>(a, patch, ss_tab, full_p, empty_p, it_begin, it_end, tid, lane);   \
    else stager_loop<KP_, NB_, false, PAIR_>(a, patch, ss_tab, full_p, empty_p, it_begin, it_end, tid, lane);            \
  } while (0)
#define TC_CONVERT(KP_, PAIR_)                                                                                                     \
  do {                                                                                                                             \
    if (a.relu_in) converter_loop<KP_, true, PAIR_>(a, raw, patch, ss_tab, full_p, empty_p, raw_full, raw_empty, it_begin, it_end, tid, lane);  \
    else converter_loop<KP_, false, PAIR_>(a, raw, patch, ss_tab, full_p, empty_p, raw_full, raw_empty, it_begin, it_end, tid, lane);           \
  } while (0)
    if (a.tma) {
      if (a.row_mode == 2 && a.kp <= 4) { if (a.kp <= 2) TC_CONVERT(2, true); else TC_CONVERT(4, true); }
      else if (a.kp <= 2) TC_CONVERT(2, false);
      else if (a.kp <= 4) TC_CONVERT(4, false);
      else TC_CONVERT(8, false);
    } else if (a.pair) {
      if (a.kp <= 2) TC_STAGE(2, 2, true);
      else TC_STAGE(4, 1, true);
    } else {
      if (a.kp <= 2) TC_STAGE(2, 4, false);
      else if (a.kp <= 4) TC_STAGE(4, 2, false);
      else TC_STAGE(8, 1, false);
    }
#undef TC_STAGE
#undef TC_CONVERT
  } else if (warp < TC_WARP_ISSUER) {
    // =============================== epilogue warps ===============================
    const int lanebase = (warp - TC_WARP_EPI0) * 32;           // == (warp % 4) * 32: the TMEM lane quarter this warp may read
    const long long HoWo = (long long)a.Ho * a.Wo;
    const bool plain = a.act == 0 && a.stats == nullptr;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 32 * TC_WARP_EPI0;
    long long dbg_acc[4] = {0, 0, 0, 0};
    const long long dbg_t0 = dbg_on ? clock64() : 0;
    for (int it = it_begin; it < it_end; ++it) {
      const int lit = it - it_begin;
      const int n = it / a.ctas_per_sample, b = it - n * a.ctas_per_sample;
      const int q0 = b * a.ntile * 128;
      const int as = lit % a.nacc, ak = lit / a.nacc;
      DBG_T(0, mbar_wait<128>(tmem_full + as, ak & 1));
      tc_fence_after();
      const int nt_item = min(a.ntile, a.total_tiles - b * a.ntile);        // the last item of a sample may hold fewer M tiles
      for (int t = 0; t < ((a.dbg_flags & 2) ? 0 : nt_item); ++t) {
        const int q = q0 + t * 128 + lanebase + lane;
        const int oy = q / a.pitch, ox = q - oy * a.pitch;
        const bool valid = oy < a.Ho && ox < a.Wo;
        float* yp = a.y + ((long long)n * a.y_ctot + a.y_coff) * HoWo + (long long)oy * a.Wo + ox;
        // upfold: channel c = parity * cout_fold + co goes to pixel (2 oy + py, 2 ox + px) of the 2Ho x 2Wo output
        const long long HoWo4 = 4 * HoWo;
        float* ypf = a.y + ((long long)n * a.y_ctot + a.y_coff) * HoWo4 + (long long)(2 * oy) * (2 * a.Wo) + 2 * ox;
        const uint32_t trow = tmem + ((uint32_t)lanebase << 16) + (uint32_t)(as * a.acc_cols + t * a.tile_cols);
        if (a.z != nullptr) {
          // operand-format output for the stride-2 consumer (conv_z.cu): BatchNorm affine + ReLU of the consumer, fp16 hi / scaled-lo split,
          // one 16-byte store per channel octet and term; lanes are consecutive x, i.e. alternate between the two column-parity planes
          const int g = n / a.n_per_group;
          const float* zs = a.z_scale + (long long)g * a.z_ss_gs;
          const float* zb = a.z_shift + (long long)g * a.z_ss_gs;
          const int PWz = (a.Wo >> 1) + 1, PHz = (a.Ho >> 1) + 1;
          const long long plane16 = (long long)PHz * PWz;
          const int yy = (oy >> 1) + 1, xx = (ox >> 1) + 1;
          uint4* zp0 = a.z + (long long)n * a.z_bs16 + (long long)((oy & 1) * 2 + (ox & 1)) * plane16;
          for (int c0 = 0; c0 < a.cout; c0 += 16) {
            float v1[16], v2[16];
            tmem_ld16(trow + c0, v1); tmem_ld16(trow + a.coutP + c0, v2);
            tmem_ld_wait();
            if (a.split_terms) {
              float v3[16];
              tmem_ld16(trow + 2 * a.coutP + c0, v3);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) v2[j] += v3[j];
            }
            if (valid) {
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                if (c0 + 8 * u < a.cout) {
                  uint32_t hw[4], lw[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const int c = c0 + 8 * u + 2 * j;
                    const float e0 = fmaxf(fmaf(v1[8 * u + 2 * j] + v2[8 * u + 2 * j], __ldg(zs + c), __ldg(zb + c)), 0.f);
                    const float e1 = fmaxf(fmaf(v1[8 * u + 2 * j + 1] + v2[8 * u + 2 * j + 1], __ldg(zs + c + 1), __ldg(zb + c + 1)), 0.f);
                    const float h0 = fminf(e0, 65504.f), h1 = fminf(e1, 65504.f);
                    const __half2 hh = __floats2half2_rn(h0, h1);
                    const float2 hf = __half22float2(hh);
                    const __half2 ll = __floats2half2_rn(fminf((e0 - hf.x) * 2048.f, 65504.f), fminf((e1 - hf.y) * 2048.f, 65504.f));
                    hw[j] = *reinterpret_cast<const uint32_t*>(&hh); lw[j] = *reinterpret_cast<const uint32_t*>(&ll);
                  }
                  const uint4 hv = make_uint4(hw[0], hw[1], hw[2], hw[3]), lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                  uint4* zp = zp0 + (long long)(((c0 >> 3) + u) * 2) * 4 * plane16;       // [c8][hl][plane]: hi at hl = 0, lo one hl stride further
                  const long long hl = 4 * plane16;
                  zp[yy * PWz + xx] = hv; zp[hl + yy * PWz + xx] = lv;
                  // circular halo of the parity plane: its last row / column is also stored as row 0 / column 0
                  const bool hx = xx == PWz - 1, hy = yy == PHz - 1;
                  if (hx) { zp[yy * PWz] = hv; zp[hl + yy * PWz] = lv; }
                  if (hy) { zp[xx] = hv; zp[hl + xx] = lv; }
                  if (hx && hy) { zp[0] = hv; zp[hl] = lv; }
                }
              }
            }
          }
        } else if (plain) {
          // common case (no output activation, no batch statistics): 16 columns per TMEM round trip, straight stores.
          // (ncu on the first version: the per-element activation switch + statistics branches made the 4 epilogue warps
          //  the critical path -- ~95% busy, half of their stalls branch_resolving / instruction-cache misses.)
          if (a.upfold && (a.cout_fold & 15) == 0) {
            // folded upsampling, 16 | cout_fold: the two x-parities of a (row parity, channel) are adjacent pixels -> one float2 store each,
            // a warp writes 256 contiguous bytes per instruction.  (First version: 64 scalar stride-2 stores per thread with a runtime
            // division per column -- the epilogue took 0.7 of the layer's 1.1 ms.)
            const int cf = a.cout_fold;
            for (int py = 0; py < 2; ++py) {
              for (int cr0 = 0; cr0 < cf; cr0 += 16) {
                const int ca = 2 * py * cf + cr0, cb = ca + cf;
                float a1[16], a2[16], b1[16], b2[16];
                tmem_ld16(trow + ca, a1); tmem_ld16(trow + a.coutP + ca, a2); tmem_ld16(trow + cb, b1); tmem_ld16(trow + a.coutP + cb, b2);
                tmem_ld_wait();
                if (a.split_terms) {
                  float a3[16], b3[16];
                  tmem_ld16(trow + 2 * a.coutP + ca, a3); tmem_ld16(trow + 2 * a.coutP + cb, b3);
                  tmem_ld_wait();
#pragma unroll
                  for (int j = 0; j < 16; ++j) { a2[j] += a3[j]; b2[j] += b3[j]; }
                }
                if (valid) {
                  float* d = ypf + (long long)py * (2 * a.Wo) + (long long)cr0 * HoWo4;
#pragma unroll
                  for (int j = 0; j < 16; ++j) { *reinterpret_cast<float2*>(d) = make_float2(a1[j] + a2[j], b1[j] + b2[j]); d += HoWo4; }
                }
              }
            }
          } else
          for (int c0 = 0; c0 < a.cout; c0 += 16) {
            float v1[16], v2[16];
            DBG_T(1, tmem_ld16(trow + c0, v1); tmem_ld16(trow + a.coutP + c0, v2); tmem_ld_wait());
            if (a.split_terms) {          // third column range: the lo*hi products of the second issuing warp
              float v3[16];
              tmem_ld16(trow + 2 * a.coutP + c0, v3);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) v2[j] += v3[j];
            }
            float* pc = yp + c0 * HoWo;
            if (valid && a.upfold) {       // generic folded store (cout_fold not a multiple of 16): parity / channel advanced incrementally
              int par = c0 / a.cout_fold, cr = c0 - par * a.cout_fold;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < a.cout) ypf[cr * HoWo4 + (par >> 1) * (2 * a.Wo) + (par & 1)] = v1[j] + v2[j];
                if (++cr == a.cout_fold) { cr = 0; ++par; }
              }
            } else if (valid) {
              if (c0 + 16 <= a.cout) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { *pc = v1[j] + v2[j]; pc += HoWo; }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (c0 + j < a.cout) *pc = v1[j] + v2[j];
                  pc += HoWo;
                }
              }
            }
          }
        } else {
          for (int c0 = 0; c0 < a.cout; c0 += 8) {
            float v1[8], v2[8];
            tmem_ld8(trow + c0, v1);
            tmem_ld8(trow + a.coutP + c0, v2);
            tmem_ld_wait();
            if (a.split_terms) {
              float v3[8];
              tmem_ld8(trow + 2 * a.coutP + c0, v3);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) v2[j] += v3[j];
            }
            for (int j = 0; j < 8; ++j) {
              const int co = c0 + j;
              float val = apply_act(v1[j] + v2[j], a.act);
              const bool st = valid && co < a.cout;
              const int par = a.upfold ? co / a.cout_fold : 0, cr = a.upfold ? co - par * a.cout_fold : co;     // real output channel
              if (st) { if (a.upfold) ypf[cr * HoWo4 + (par >> 1) * (2 * a.Wo) + (par & 1)] = val; else yp[co * HoWo] = val; }
              if (a.stats) {
                if (!st) val = 0.f;
                const float s1 = warp_sum(val), s2 = warp_sum(val * val);
                if (lane == 0 && co < a.cout) {
                  double* sp = a.stats + 2 * (a.y_coff + cr);
                  atomicAdd(sp, (double)s1); atomicAdd(sp + 1, (double)s2);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);
    }
    if (dbg_on) { a.dbg[4] = clock64() - dbg_t0; a.dbg[5] = dbg_acc[0]; a.dbg[6] = dbg_acc[1]; }
  } else if (warp < TC_WARP_WPROD) {
    // =============================== MMA issuing warp ===============================
    // All 32 lanes run the (warp-uniform) loops so that the address arithmetic stays on the uniform datapath; one elected
    // lane issues the tcgen05 instructions.  (First version: a single lane computed every descriptor in vector registers
    // and needed ~400 cycles per MMA pair -- the tensor pipe was idle 80-95% of the time.)
    const bool leader = elect_one();
    const int issuer = warp - TC_WARP_ISSUER;          // this warp issues the MMAs of tiles issuer, issuer + TC_ISSUERS, ...
    const uint32_t idesc1 = umma_idesc_tf32(128, N2), idesc2 = umma_idesc_tf32(128, a.coutP);
    const uint32_t PP = (uint32_t)a.PP;
    // descriptor words: hi = SBO (128 B) | version 1; lo = start address >> 4 | LBO << 16.  Offsets (in 16-byte units) are added
    // to the low word: every operand lies below 256 KB, so the 14-bit address field cannot carry into the LBO field.
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_base = (smem_u32(patch) >> 4) | (PP << 16);
    const uint32_t b_lo_base = (smem_u32(wbuf) >> 4) | ((uint32_t)N2 << 16);
    const uint32_t slot16 = (uint32_t)a.patch_slot_floats >> 2, w_slot16 = (uint32_t)a.w_slot_floats >> 2;
    // loop-invariant parameters in registers (ncu: the per-tap loop spent most of its ~290 cycles re-reading them from the constant bank)
    const int tchunk = a.tchunk, ntc = a.ntc, ncc = a.ncc, nslot_p = a.nslot_p, nacc = a.nacc, resident = a.resident, split_terms = a.split_terms;
    const uint32_t coutP = (uint32_t)a.coutP, tile_cols = (uint32_t)a.tile_cols, acc_cols = (uint32_t)a.acc_cols, b_step = 2u * (uint32_t)N2;
    int g_loaded = -1, wgen = 0;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && leader && issuer == 0;
    long long dbg_acc[4] = {0, 0, 0, 0};
    const long long dbg_t0 = dbg_on ? clock64() : 0;
    // item coordinates are advanced incrementally (five runtime integer divisions per item cost the single issuing lane ~10% of In_conv1)
    const int cps = a.ctas_per_sample, npg = a.n_per_group, ntile = a.ntile, total_tiles = a.total_tiles, pitch = a.pitch;
    int n = it_begin / cps, b = it_begin - n * cps, g = n / npg, n_in_g = n - g * npg;
    int pseq = 0;
    for (int it = it_begin; it < it_end; ++it) {
      const int lit = it - it_begin;
      const int q0 = b * ntile * 128;
      const uint32_t qlocal0 = (uint32_t)(q0 % pitch);
      const int nt_item = min(ntile, total_tiles - b * ntile);
      const int as = nacc == 2 ? (lit & 1) : 0, ak = nacc == 2 ? (lit >> 1) : lit;
      if (ak > 0) DBG_T(0, mbar_wait(tmem_empty + as, (ak - 1) & 1));
      if (resident && g != g_loaded) { DBG_T(2, mbar_wait(wres_full, wgen & 1)); ++wgen; g_loaded = g; }
      tc_fence_after();
      const uint32_t acc_base = tmem + (uint32_t)as * acc_cols;
      for (int cc = 0; cc < ncc; ++cc, ++pseq) {
        const int ps = nslot_p == 2 ? (pseq & 1) : 0, pk = nslot_p == 2 ? (pseq >> 1) : pseq;
        DBG_T(1, mbar_wait(full_p + ps, pk & 1));
        const uint32_t pa_hi = a_lo_base + (uint32_t)ps * slot16 + qlocal0, pa_lo = pa_hi + (slot16 >> 1);
        for (int tc = 0; tc < ntc; ++tc) {
          const int chunk = cc * ntc + tc;
          uint32_t wb;
          int ws = 0;
          if (resident) {
            wb = b_lo_base + (uint32_t)chunk * w_slot16;
          } else {
            const int wseq = lit * nchunks + chunk;
            ws = wseq & 1;
            DBG_T(2, mbar_wait(full_w + ws, (wseq >> 1) & 1));
            wb = b_lo_base + (uint32_t)ws * w_slot16;
          }
          tc_fence_after();
          const long long t_iss = dbg_on ? clock64() : 0;
          if (leader && !(a.dbg_flags & 4)) {
            // one elected lane walks (tile, tap, 8-channel group); 64-bit descriptors are advanced in place (the low word only carries
            // the 14-bit address field, which never overflows); the next tap's offset is fetched one iteration ahead
            const uint64_t hi64 = (uint64_t)desc_hi << 32;
            const uint32_t first_chunk = chunk == 0 ? 0u : 1u;
            const int tap0 = tc * tchunk;
            if (split_terms) {
              // issuer 0: A_hi * [W_hi;W_lo] into columns [0,2N); issuer 1: A_lo * W_hi into its own range [2N,3N) -- two
              // independent accumulation chains, so the two warps never touch the same TMEM columns
              const uint32_t idesc = issuer ? idesc2 : idesc1;
              const uint32_t pa = issuer ? pa_lo : pa_hi;
              for (int t = 0; t < nt_item; ++t) {
                const uint32_t d_tmem = acc_base + (uint32_t)t * tile_cols + (issuer ? 2u * coutP : 0u);
                uint64_t db = hi64 | wb;
                uint32_t al_next = (uint32_t)a.tap_aoff[tap0];
                for (int tl = 0; tl < tchunk; ++tl) {
                  uint64_t da = hi64 | (pa + al_next + (uint32_t)(t * 128));
                  if (tl + 1 < tchunk) al_next = (uint32_t)a.tap_aoff[tap0 + tl + 1];
                  uint32_t acc = (tl == 0) ? first_chunk : 1u;
#pragma unroll 4
                  for (int c8 = 0; c8 < nc8; ++c8) {
                    umma_tf32(d_tmem, da, db, idesc, acc);
                    acc = 1u;
                    da += 2 * PP; db += b_step;
                  }
                }
              }
            } else {
              for (int t = issuer; t < nt_item; t += TC_ISSUERS) {
                const uint32_t d_tmem = acc_base + (uint32_t)t * tile_cols;
                uint64_t db = hi64 | wb;
                uint32_t al_next = (uint32_t)a.tap_aoff[tap0];
                for (int tl = 0; tl < tchunk; ++tl) {
                  const uint32_t al = al_next + (uint32_t)(t * 128);
                  if (tl + 1 < tchunk) al_next = (uint32_t)a.tap_aoff[tap0 + tl + 1];
                  uint64_t da_h = hi64 | (pa_hi + al), da_l = hi64 | (pa_lo + al);
                  uint32_t acc = (tl == 0) ? first_chunk : 1u;
#pragma unroll 2
                  for (int c8 = 0; c8 < nc8; ++c8) {
                    umma_tf32(d_tmem, da_h, db, idesc1, acc);                  // hi * [hi;lo]
                    umma_tf32(d_tmem + coutP, da_l, db, idesc2, 1u);           // lo * hi -> correction half
                    acc = 1u;
                    da_h += 2 * PP; da_l += 2 * PP; db += b_step;
                  }
                }
              }
            }
          }
          if (dbg_on) dbg_acc[3] += clock64() - t_iss;
          if (!resident && leader) umma_commit(empty_w + ws);
        }
        if (leader) umma_commit(empty_p + ps);
      }
      if (leader) umma_commit(tmem_full + as);
      if (++b == cps) { b = 0; ++n; if (++n_in_g == npg) { n_in_g = 0; ++g; } }
    }
    if (dbg_on) { a.dbg[8] = clock64() - dbg_t0; a.dbg[9] = dbg_acc[0]; a.dbg[10] = dbg_acc[1]; a.dbg[11] = dbg_acc[2]; a.dbg[12] = it_end - it_begin; a.dbg[15] = dbg_acc[3]; }
    __syncwarp();
  } else if (warp == TC_WARP_WPROD) {
    // =============================== weight copy warp ===============================
    if (lane == 0) {
      const uint32_t w_slot_bytes = (uint32_t)a.w_slot_floats * 4u;
      int g_loaded = -1;
      for (int it = it_begin; it < it_end; ++it) {
        const int lit = it - it_begin;
        const int n = it / a.ctas_per_sample;
        const int g = n / a.n_per_group;
        const float* wsrc = a.wprep + (long long)g * a.wprep_gs;
        if (a.resident) {
          // follow every accumulator hand-off in order (an mbarrier parity wait is only meaningful one phase behind), so that at a
          // group change all MMAs that read the old weight image are known to be complete
          if (lit > 0) { const int pl_ = lit - 1; mbar_wait<256>(tmem_full + (pl_ % a.nacc), (pl_ / a.nacc) & 1); }
          if (g != g_loaded) {
            mbar_arrive_expect_tx(wres_full, (uint32_t)nchunks * w_slot_bytes);
            for (int c = 0; c < nchunks; ++c)
              bulk_g2s(wbuf + (size_t)c * a.w_slot_floats, wsrc + (size_t)c * a.w_slot_floats, w_slot_bytes, wres_full);
            g_loaded = g;
          }
        } else {
          for (int c = 0; c < nchunks; ++c) {
            const int wseq = lit * nchunks + c;
            const int ws = wseq & 1, wk = wseq >> 1;
            if (wk > 0) mbar_wait<128>(empty_w + ws, (wk - 1) & 1);
            mbar_arrive_expect_tx(full_w + ws, w_slot_bytes);
            bulk_g2s(wbuf + (size_t)ws * a.w_slot_floats, wsrc + (size_t)c * a.w_slot_floats, w_slot_bytes, full_w + ws);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== input copy warp (TMA) ===============================
    if (a.tma) {
      const uint32_t row_bytes = (uint32_t)(a.cchunk * a.W) * 4u;
      for (int it = it_begin; it < it_end; ++it) {
        const int lit = it - it_begin;
        const int n = it / a.ctas_per_sample, b = it - n * a.ctas_per_sample;
        const int rbase = (b * a.ntile * 128) / a.pitch + a.min_ry;
        const int row0 = a.row_mode == 1 ? (rbase >> 1) : (a.row_mode == 2 ? 2 * rbase : rbase);   // first source row (may be negative)
        for (int cc = 0; cc < a.ncc; ++cc) {
          const int rseq = lit * a.ncc + cc;
          const int rs = rseq & 1, rk = rseq >> 1;
          if (rk > 0) mbar_wait<128>(raw_empty + rs, (rk - 1) & 1);
          if (lane == 0) mbar_arrive_expect_tx(raw_full + rs, (uint32_t)a.RS * row_bytes);
          __syncwarp();
          float* dst = raw + (size_t)rs * a.raw_slot_floats;
          for (int r = lane; r < a.RS; r += 32) {                    // one lane per source row; vertical wrap = row index mod H
            int row = (row0 + r) % a.H; if (row < 0) row += a.H;
            tma_load_4d(dst + (size_t)r * (a.cchunk * a.W), &tmap, 0, row, cc * a.cchunk, n, raw_full + rs);
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == TC_WARP_ISSUER) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------ host side
static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static int floor_div(int a, int b) { int q = a / b, r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q; }

int arpde_tc_enabled() {
  static int cached = -1;
  if (cached < 0) { const char* e = arpde_dev_env("ARPDE_TC"); cached = (e && e[0] == '0') ? 0 : 1; }
  return cached;
}

// Plan the tiling of one convolution for conv_tc_kernel.  Returns false if the layer is outside what the kernel covers.
static bool tc_plan(const ConvDesc& d, ConvTcArgs& a, size_t* smem_out) {
  if (d.pad_override >= 0) return false;
  // nearest x2 + 3x3: plan the equivalent low-resolution conv with 4*cout parity-major channels (see the header comment)
  if (d.up && !d.dilate && d.stride == 1 && d.kh == 3 && d.kw == 3 && d.H > 1 && 4 * d.cout <= 128 && !arpde_dev_env("ARPDE_TC_NO_UPFOLD")) {
    ConvDesc e = d; e.up = 0; e.cout = 4 * d.cout;
    if (tc_plan(e, a, smem_out)) { a.upfold = 1; a.cout_fold = d.cout; return true; }
  }
  if (d.dilate && (d.up || d.stride != 1 || (d.H == 1 && d.kh == 1))) return false;
  if (d.stride != 1 && d.stride != 2) return false;
  if (d.up && d.stride != 1) return false;
  const int one_d = (d.H == 1 && d.kh == 1);
  memset(&a, 0, sizeof(a));
  a.cin = d.cin; a.H = d.H; a.W = d.W; a.up_w = (d.up || d.dilate) ? 1 : 0; a.up_h = one_d ? 0 : a.up_w; a.dilate = d.dilate ? 1 : 0;
  a.Hc = d.H << a.up_h; a.Wc = d.W << a.up_w; a.sy = one_d ? 1 : d.stride; a.sx = d.stride;
  if (a.Hc % a.sy || a.Wc % a.sx) return false;
  a.Ho = a.Hc / a.sy; a.Wo = a.Wc / a.sx;
  a.cout = d.cout; a.coutP = round_up(d.cout, 16);
  if (a.coutP > 128 || round_up(d.cin, 32) > TC_MAX_CIN) return false;
  a.ntaps = d.kh * d.kw;
  if (a.ntaps > TC_MAX_TAPS || d.kh > 7 || d.kw > 7) return false;
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
  int ntile_max = 512 / N2; if (ntile_max > 4) ntile_max = 4;
  if (ntile_max > total_tiles) ntile_max = total_tiles;
  const size_t budget = 226 * 1024;
  a.row_mode = a.up_w ? 1 : (a.sx == 2 ? 2 : 0);
  const bool tma_shape = (d.W % 4 == 0) && d.W <= 256 && ((long long)d.H * d.W) % 4 == 0;
  // preference order: weights resident in shared memory (no re-streaming per item), input rows by TMA, double-buffered patch
  // slots (staging overlaps the MMAs), then many M tiles per item (less halo, fewer weight passes), wide channel chunks (fewer
  // pipeline hand-offs), wide tap chunks
  // experiment hook: ARPDE_TC_FORCE="nslot_p,ntile,cchunk" restricts the search (0 = free) -- used for A/B timing only
  int f_slot = 0, f_ntile = 0, f_cchunk = 0;
  if (const char* e = arpde_dev_env("ARPDE_TC_FORCE")) sscanf(e, "%d,%d,%d", &f_slot, &f_ntile, &f_cchunk);
  // (streamed weights, stride 2, all taps in one weight slot -- pass 0: two M tiles sharing one patch slot halve the weight streaming
  //  and the halo per tile; measured on In_conv3 at N = 1920: 2.10 ms against 2.22 ms for one tile with two patch slots, 2.31 / 3.75 ms
  //  for three / four tiles.  Stride-1 layers and split tap chunks lose with it: 144->32 3x3 at 16x16 0.70 vs 0.47 ms.)
  for (int resident = 1; resident >= 0; --resident)
   for (int pass = resident ? 1 : 0; pass <= 1; ++pass)
    for (int use_tma = tma_shape ? 1 : 0; use_tma >= (d.dilate ? 1 : 0); --use_tma)
      for (int nslot_p = 2; nslot_p >= 1; --nslot_p)
        for (int ntile = ntile_max; ntile >= 1; --ntile)
          for (int cchunk = 32; cchunk >= 8; cchunk >>= 1) {
            if (pass == 0 && (!use_tma || a.nplanes != 4 || nslot_p != 1 || ntile != 2 || 2 * 2 * N2 > 512 || f_slot || f_ntile || f_cchunk)) continue;
            if ((f_slot && nslot_p != f_slot) || (f_ntile && ntile != f_ntile) || (f_cchunk && cchunk != f_cchunk)) continue;
            if (cchunk > 8 && cchunk > cin8) continue;
            const int ncc = (d.cin + cchunk - 1) / cchunk;
            if (nslot_p == 1 && ncc == 1 && ntile_max > 1) continue;      // a single slot only pays off for multi-chunk layers
            const int rows = (a.pitch - 1 + ntile * 128 - 1) / a.pitch + 1 + span_y;
            const int PHP = rows * a.pitch;
            if (PHP > TC_STAGERS * TC_KP_MAX) continue;
            int PP = a.pitch + ntile * 128 + max_off; if (PP < PHP) PP = PHP; PP = round_up(PP, 8);
            if (PP > 16383) continue;
            const size_t p_slot = (size_t)2 * a.nplanes * (cchunk / 4) * PP * 16;
            const int RS = a.row_mode == 1 ? (rows >> 1) + 1 : (a.row_mode == 2 ? 2 * rows : rows);
            const size_t raw_slot = use_tma ? (size_t)RS * cchunk * d.W * 4 : 0;
            for (int tchunk = a.ntaps; tchunk >= 1; --tchunk) {
              if (a.ntaps % tchunk) continue;
              if (pass == 0 && tchunk != a.ntaps) continue;
              const size_t w_slot = (size_t)tchunk * (cchunk / 4) * N2 * 16;
              if (w_slot > 48 * 1024) continue;
              const int ntc = a.ntaps / tchunk;
              const size_t w_bytes = resident ? w_slot * ncc * ntc : 2 * w_slot;
              const size_t smem = TC_SMEM_HEADER + 2 * raw_slot + nslot_p * p_slot + w_bytes;
              if (smem > budget) continue;
              a.resident = resident; a.ntile = ntile; a.cchunk = cchunk; a.nc4_shift = cchunk == 32 ? 3 : (cchunk == 16 ? 2 : 1);
              a.ncc = ncc; a.tchunk = tchunk; a.ntc = ntc; a.PP = PP; a.PHP = PHP; a.nslot_p = nslot_p;
              a.kp = (PHP + TC_STAGERS - 1) / TC_STAGERS;
              a.patch_slot_floats = (int)(p_slot / 4); a.w_slot_floats = (int)(w_slot / 4);
              a.tma = use_tma; a.RS = RS; a.raw_slot_floats = (int)(raw_slot / 4);
              a.ctas_per_sample = (total_tiles + ntile - 1) / ntile; a.total_tiles = total_tiles;
              for (int t = 0; t < a.ntaps; ++t) a.tap_aoff[t] = a.tap_plane[t] * (cchunk / 4) * PP + a.tap_off[t];
              // two issuing warps: split by M tile; single-tile items split by product term (separate column range for lo*hi) when a
              // double-buffered accumulator of 3*coutP columns fits in TMEM (measured: splitting multi-tile items by term is slower)
              a.split_terms = (ntile == 1 && 2 * 3 * a.coutP <= 512) ? 1 : 0;
              a.tile_cols = a.split_terms ? 3 * a.coutP : N2;
              a.acc_cols = ntile * a.tile_cols;
              a.nacc = 2 * a.acc_cols <= 512 ? 2 : 1;
              int p2 = 32; while (p2 < a.nacc * a.acc_cols) p2 <<= 1;
              a.tmem_cols = p2;
              *smem_out = smem;
              return true;
            }
          }
  return false;
}

// floats of prepared weights one group needs for this layer (upper bound over every plan tc_plan may pick)
long long arpde_conv_tc_wprep_floats(int cin, int cout, int kh, int kw) {
  const int cout_eff = (kh == 3 && kw == 3 && 4 * cout <= 128) ? 4 * cout : cout;      // the layer may be an upsampling one (upfold)
  return (long long)kh * kw * round_up(cin, 32) * 2 * round_up(cout_eff, 16);
}

// heuristic: which layers go to the tensor cores (the rest stay on the fp32 FFMA kernel)
bool arpde_conv_tc_wanted(const ConvDesc& d) {
  if (!arpde_tc_enabled()) return false;
  if (d.H == 1 && d.kh == 1) return false;                 // 1D nets: tiles of 128 positions are mostly halo
  if (d.cout < 4 || d.cin * d.kh * d.kw < 32) return false;
  if (arpde_conv_fwd_rt_shape_ok(d) && !arpde_dev_env("ARPDE_TC_DENSE")) return false;     // narrow layers: the row-tiled FFMA kernel is faster
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
  conv_tc_wprep_kernel<<<blocks, 256, 0, stream>>>(w, w_gs, wprep, wprep_gs, groups, a.cout, d.cin, a.ntaps, a.cchunk, a.ncc, a.tchunk,
                                                   a.ntc, a.coutP, a.upfold ? a.cout_fold : 0);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// Tensor map of the input view [N][cin][H][W] (batch stride x_bs floats, channel stride H*W): box = one full-width row of
// `cchunk` channels.  cuTensorMapEncodeTiled comes from the driver through the runtime (no link-time libcuda dependency).
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int tc_encode_input_map(CUtensorMap* map, const float* x, long long x_bs, int N, int cin, int H, int W, int cchunk) {
  static tc_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ARPDE_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    ARPDE_CHECK_ARG(p != nullptr && qres == cudaDriverEntryPointSuccess, "conv_tc: cuTensorMapEncodeTiled not available from the driver");
    fn = (tc_encode_fn)p;
  }
  const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)cin, (cuuint64_t)N};
  const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)x_bs * 4};
  const cuuint32_t box[4] = {(cuuint32_t)W, 1u, (cuuint32_t)cchunk, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { arpde_set_error("conv_tc: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d] stride %lld", (int)r, N, cin, H, W, x_bs); return ARPDE_ERR_CUDA; }
  return ARPDE_OK;
}

// Profiling hooks: only a development build (-DARPDE_DEV) keeps this process-global state and exports the two setters; in the
// shipped library the buffer is always null and the flags always 0.
#ifdef ARPDE_DEV
static long long* g_tc_dbg = nullptr;
static int g_tc_dbg_flags = 0;
extern "C" int arpde_conv_tc_set_debug_flags(int flags) { g_tc_dbg_flags = flags; return ARPDE_OK; }
// debug hook: cycle counters of CTA 0 of the following conv_tc launches are written to dev_buf (16 int64):
// [0..3] stager: total, wait empty_p, staging, fence+arrive; [4,5] epilogue: total, wait tmem_full;
// [8..12] issuer: total, wait tmem_empty, wait full_p, wait weights, items
extern "C" int arpde_conv_tc_set_debug(long long* dev_buf) { g_tc_dbg = dev_buf; return ARPDE_OK; }
long long* arpde_tc_debug_buffer() { return g_tc_dbg; }
#else
static constexpr long long* g_tc_dbg = nullptr;
static constexpr int g_tc_dbg_flags = 0;
long long* arpde_tc_debug_buffer() { return nullptr; }
#endif

int arpde_conv_tc_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const float* scale,
                         const float* shift, long long ss_gs, float* y, double* stats, cudaStream_t stream, const ZOut* zo) {
  ARPDE_CHECK_ARG(x && wprep && (y || zo), "conv_tc: null pointer");
  if (d.N == 0) return ARPDE_OK;
  ConvTcArgs a; size_t smem;
  if (!tc_plan(d, a, &smem)) { arpde_set_error("conv_tc: layer not supported by the tensor-core kernel"); return ARPDE_ERR_UNSUPPORTED; }
  a.x = x; a.x_bs = x_bs; a.scale = scale; a.shift = shift; a.ss_gs = ss_gs; a.wprep = wprep; a.wprep_gs = wprep_gs;
  a.y = y; a.stats = stats; a.relu_in = d.relu_in; a.act = d.act; a.y_ctot = d.y_ctot; a.y_coff = d.y_coff;
  a.n_per_group = d.n_per_group; a.dbg = g_tc_dbg; a.dbg_flags = g_tc_dbg_flags;
  a.N = d.N;
  if (zo) {
    ARPDE_CHECK_ARG(zo->z && zo->scale && zo->shift && !a.upfold && d.act == 0 && !stats && d.cout % 16 == 0 && !(a.Ho & 1) && !(a.Wo & 1) &&
                    (((uintptr_t)zo->z) & 15) == 0, "conv_tc: operand-format output needs 16 | cout, even output dims, no activation / statistics");
    a.z = reinterpret_cast<uint4*>(zo->z); a.z_bs16 = arpde_conv_z_sample16(d.cout, a.Ho, a.Wo);
    a.z_scale = zo->scale; a.z_shift = zo->shift; a.z_ss_gs = zo->ss_gs;
  }
  // TMA needs a 16-byte aligned base and 16-byte multiples for every global stride; otherwise the staging warps load with LDG
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (a.tma && ((((uintptr_t)x) & 15) != 0 || (x_bs & 3) != 0 || arpde_dev_env("ARPDE_TC_NO_TMA"))) a.tma = 0;
  if (a.dilate && scale) { arpde_set_error("conv_tc: a zero-inserted input cannot carry a BatchNorm prologue"); return ARPDE_ERR_UNSUPPORTED; }
  if (!a.tma && a.dilate) { arpde_set_error("conv_tc: zero-inserted input needs the TMA path (16-byte aligned input)"); return ARPDE_ERR_UNSUPPORTED; }
  if (a.tma) {
    int rc = tc_encode_input_map(&tmap, x, x_bs, d.N, d.cin, d.H, d.W, a.cchunk);
    if (rc) return rc;
  }
  a.pair = (!a.tma && a.sx == 2 && !a.up_w && a.kp <= 4 && (a.W & 1) == 0 && (x_bs & 1) == 0 && (((uintptr_t)x) & 7) == 0 &&
            (((long long)a.H * a.W) & 1) == 0) ? 1 : 0;
  // opt in to the device's full shared memory once per device.  (Not a thread_local high-water mark: autograd runs backward() on its
  // own thread, whose first smaller request lowered the attribute under the forward thread's feet -> "invalid argument" launches.)
  static std::mutex attr_mutex;
  static bool attr_done[64] = {false};
  size_t smem_set = 0;
  {
    int dev = 0;
    ARPDE_CUDA(cudaGetDevice(&dev));
    int optin = 0;
    ARPDE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    smem_set = (size_t)optin;
    ARPDE_CHECK_ARG(smem <= smem_set, "conv_tc: plan needs %zu bytes of shared memory, device allows %zu", smem, smem_set);
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      ARPDE_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  const long long items = (long long)d.N * a.ctas_per_sample;
  ARPDE_CHECK_ARG(items < (1ll << 31), "conv_tc: too many work items");
  long long blocks = (long long)arpde_num_sms();       // persistent CTAs, one per SM
  if (blocks > items) blocks = items;
  conv_tc_kernel<<<(unsigned)blocks, TC_THREADS, smem, stream>>>(a, tmap);
  {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      arpde_set_error("conv_tc launch failed: %s (grid %lld, block %d, smem %zu (attribute set to %zu), N=%d cin=%d %dx%d cout=%d k=%dx%d s=%d up=%d tma=%d)",
                      cudaGetErrorString(e), blocks, TC_THREADS, smem, smem_set, d.N, d.cin, d.H, d.W, d.cout, d.kh, d.kw, d.stride, d.up, a.tma);
      return ARPDE_ERR_CUDA;
    }
    arpde_count_launch();
  }
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

// Diagnostic (host only, no GPU needed): the tiling conv_tc_kernel would use for a layer.  out[0..26] = pitch, min_ry, min_rx, ntile,
// ctas_per_sample, PP, PHP, nplanes, cchunk, ncc, tchunk, ntc, ntaps, coutP, tmem_cols, nslot_p, smem bytes, Ho, Wo, Hc, Wc, sy, sx,
// wanted-by-heuristic, resident, nacc, kp, upfold, cout_fold; out[32+t] = tap plane, out[32+49+t] = tap offset.  Returns ARPDE_ERR_UNSUPPORTED if not plannable.
extern "C" int arpde_conv_tc_plan(int cin, int H, int W, int cout, int kh, int kw, int stride, int up, int32_t* out) {
  ARPDE_CHECK_ARG(out, "conv_tc_plan: null output");
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1; d.N = 1; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride;
  d.up = up; d.relu_in = 0; d.act = 0; d.y_ctot = cout; d.y_coff = 0; d.n_per_group = 1;
  ConvTcArgs a; size_t smem = 0;
  if (!tc_plan(d, a, &smem)) { arpde_set_error("conv_tc_plan: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  out[27] = a.upfold; out[28] = a.cout_fold;
  const int v[27] = {a.pitch, a.min_ry, a.min_rx, a.ntile, a.ctas_per_sample, a.PP, a.PHP, a.nplanes, a.cchunk, a.ncc, a.tchunk, a.ntc,
                     a.ntaps, a.coutP, a.tmem_cols, a.nslot_p, (int)smem, a.Ho, a.Wo, a.Hc, a.Wc, a.sy, a.sx, arpde_conv_tc_wanted(d) ? 1 : 0,
                     a.resident, a.nacc, a.kp};
  for (int i = 0; i < 27; ++i) out[i] = v[i];
  for (int t = 0; t < TC_MAX_TAPS; ++t) { out[32 + t] = t < a.ntaps ? a.tap_plane[t] : 0; out[32 + TC_MAX_TAPS + t] = t < a.ntaps ? a.tap_off[t] : 0; }
  return ARPDE_OK;
}
