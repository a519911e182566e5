// conv_stem.cu -- producer side of the eval-mode stem pair (conv_z.cu): the first convolution of the network, few input channels
// (cin <= 7), k x k stride 1 circular (In_conv1: 6 -> 96, 5x5; 2D-Burgers-SWAG/nn/denseEDcirc2d.py:247-249), followed by the consumer's
// eval-mode BatchNorm affine + ReLU (In_conv_norm1 / In_conv_relu1, :250-251), written in the consumer's fp16 two-term operand format Z.
//
// conv_tc.cu runs this layer as 25 taps x (K = 8 TF32 MMAs with 2 of 8 channels padding) and, with the operand-format epilogue, its four
// epilogue warps are the critical path (2.44 ms at 1920 samples against 1.8 ms with the plain epilogue).  This kernel is built around the
// layer instead:
//   * operands in the fp16 two-term split of conv_z.cu (x = x_hi + 2^-11 x_lo', products exact, fp32 accumulation in TMEM);
//   * the whole input octet (cin <= 7 channels + the constant below) of a position is ONE 16-byte K unit, so the second K unit of a K = 16 MMA can be the
//     SAME buffer at another position -- the descriptor's LBO is free: 16 bytes = the next column, pitch x 16 bytes = the next row.  Taps
//     are paired vertically, (ky, kx) with (ky + 1, kx), the last filter row horizontally: 13 MMAs per accumulator half instead of 25
//     taps for 5x5 (pairing along the rows only needed 15: every third pair carried a zero tap);
//   * the prepared weights of a SWAG sample (13 pairs x 2 units x 2 cout rows x 16 B = 80 KB) stay resident in shared memory, the input patch of
//     half a sample (36 rows x 68 positions, hi + lo = 78 KB) is loaded AND split by 8 warps into registers while the current item's MMAs
//     run, so that only the 16-byte stores are left after the hand-over;
//   * the consumer's eval-mode BatchNorm is folded into the producer's GEMM: scale * conv(x, w) + shift = conv([x, 1], [scale * w, shift at
//     the centre tap]) -- the scale goes into the prepared weights (one set per SWAG sample anyway), the shift rides on the unused lane
//     `cin` of the input octet, which the converter sets to 1: the epilogue needs no per-channel constants at all;
//   * 8 epilogue warps (two per TMEM lane quarter, each half of the channels) drain a double-buffered accumulator: ReLU, split (saturating
//     F2FP), 16-byte stores into the four parity planes of Z (lanes = consecutive x alternate between two planes: 2 x 256 contiguous
//     bytes per store instruction), plus the circular halo copies the consumer expects; a slot is handed back as soon as a warp's last
//     block is in registers.
// Roofline: nominally the tensor pipe -- per 128 positions 13 x (N' = 2 cout = 192: 96 cycles + N = 96: 56 cycles) = 1980 cycles against
// 25 x 152 = 3800 for the TF32 form (measured: 2005-2060, scratch/stem_timeline.py, development build) -- but the launch is bound by its
// epilogue and the Z stores: 0.59 ms of MMAs, 0.17 ms of ReLU + split that the two-slot accumulator ring cannot hide, 0.20 ms of DRAM write
// pattern (ablations ARPDE_STEM_DBG = 1 / 2 / 4).  History: with the BatchNorm in the epilogue (8 vector loads of
// scale / shift + 250 instructions per 16-channel block) the epilogue set the pace at 3500 cycles per tile and 6 k cycles of idle tensor
// pipe at every item boundary while the patch was split; folding it into the GEMM gave 2700 per tile, 4 k per boundary.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"
#include <cuda_fp16.h>
#include <stdlib.h>
#include <mutex>
#include "tc_ptx.cuh"

#define CS_CONV_WARPS 8
#define CS_EPI_WARPS 8
#define CS_WARP_EPI0 CS_CONV_WARPS                       // a multiple of 4: epilogue warp e reads TMEM lane quarter e % 4
#define CS_WARP_MMA (CS_CONV_WARPS + CS_EPI_WARPS)
#define CS_WARP_W (CS_WARP_MMA + 1)
#define CS_THREADS (32 * (CS_WARP_W + 1))
#define CS_CONV_THREADS (32 * CS_CONV_WARPS)
#define CS_KP 10                                          // patch positions per converter thread
#define CS_HEADER 128

#ifdef ARPDE_DEV
#define CS_TL(role, cond) do { if (a.tl && blockIdx.x == 0 && (cond) && tlc < 500) a.tl[(role) * 512 + tlc++] = clock64(); } while (0)
#else
#define CS_TL(role, cond) do { } while (0)
#endif

namespace {

struct ConvStemArgs {
  const float* x; long long x_bs;
  const uint4* wz; long long wz_gs16;
  uint4* z; long long z_bs16;
  // lane `cin` (< 8) of every input octet is the constant 1, its centre-tap weight the consumer's BatchNorm shift; the scale is folded into the weights
  int N, n_per_group, cin, H, W, cout, k, pad, npairs, nvert;     // tap pairs: nvert vertical ones (ky even, every kx), then the last row in horizontal pairs
  int pitch, RI, items_per_sample, nt, P;                          // rows per item, M tiles per item, patch positions (padded)
  int w16;                                                          // 16-byte units of the resident weight image
  long long* tl;                                                    // development build: clock64 timeline of CTA 0 (arpde_conv_tc_set_debug buffer)
  int dbg;                                                          // development build only (ARPDE_STEM_DBG): 1 = the epilogue skips its stores
};

__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

// fp16 two-term split of two values: hi = fp16(v), lo' = fp16((v - hi) * 2^11), packed as half2 words.  Both conversions saturate to the
// largest finite fp16 value (F2FP.SATFINITE: one instruction instead of a clamp pair around each conversion -- the epilogue warps of this
// kernel are co-critical with the tensor pipe), so an out-of-range activation gives a wrong but finite result.
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

// tap pair p of a k x k filter: unit 0 = tap (ky, kx), unit 1 = the tap below it (vertical pairs, p < nvert) or to its right (the last
// filter row; kx + 1 == k: no second tap).  Shared by the MMA issue loop (patch offset, LBO) and the weight preparation.
__host__ __device__ __forceinline__ void stem_pair_taps(int p, int k, int nvert, int* ky, int* kx, int* vertical) {
  if (p < nvert) { *ky = 2 * (p / k); *kx = p - (p / k) * k; *vertical = 1; }
  else { *ky = k - 1; *kx = 2 * (p - nvert); *vertical = 0; }
}

__global__ void __launch_bounds__(CS_THREADS, 1) conv_stem_kernel(const __grid_constant__ ConvStemArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* patch_full = reinterpret_cast<uint64_t*>(smem_raw);      // converted patch of an item in place (one arrival per converter warp)
  uint64_t* patch_empty = patch_full + 1;                             // all MMAs of an item complete (tcgen05.commit)
  uint64_t* w_full = patch_full + 2;                                  // resident weight image (re)loaded (bulk-copy tx bytes)
  uint64_t* tile_full = patch_full + 3;                               // [2] accumulator slot complete (tcgen05.commit)
  uint64_t* tile_empty = patch_full + 5;                              // [2] accumulator slot drained (one arrival per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 96);
  uint4* wbuf = reinterpret_cast<uint4*>(smem_raw + CS_HEADER);      // [npairs][unit 2][2 cout rows][16 B]
  uint4* patch_hi = wbuf + a.w16;                                     // [P positions][16 B]: the cin (<= 8) channels of a position
  uint4* patch_lo = patch_hi + a.P;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int items_total = a.N * a.items_per_sample;
  const int per_cta = (items_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int it_begin = (int)blockIdx.x * per_cta;
  const int it_end = min(items_total, it_begin + per_cta);
  const int tile_cols = 2 * a.cout;

  if (tid == 0) {
    mbar_init(patch_full, CS_CONV_WARPS); mbar_init(patch_empty, 1); mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(tile_full + s, 1); mbar_init(tile_empty + s, CS_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CS_WARP_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();

  if (warp < CS_CONV_WARPS) {
    // =============================== converter warps: fp32 input -> fp16 hi / lo' patch ===============================
    const long long HW = (long long)a.H * a.W;
    const int Prows = (a.RI + a.k - 1) * a.pitch;
    const int ones_c = a.cin;
    for (int it = it_begin; it < it_end; ++it) {
      const int lit = it - it_begin;
      const int n = it / a.items_per_sample, b = it - n * a.items_per_sample;
      const float* xn = a.x + (long long)n * a.x_bs;
      const int y0 = b * a.RI - a.pad;
      // load AND split while the previous item's MMAs still read the patch: after the hand-over only the 2 x CS_KP 16-byte stores are left
      // (splitting after the wait cost ~6 k cycles per item boundary with the tensor pipe idle, scratch/stem_timeline.py)
      uint4 ph[CS_KP], pl[CS_KP];
#pragma unroll
      for (int h = 0; h < CS_KP; h += 5) {                           // batches of 5 positions: 40 raw values in flight per thread
        float xr[5][8];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int p = tid + (h + i) * CS_CONV_THREADS;
#pragma unroll
          for (int c = 0; c < 8; ++c) xr[i][c] = 0.f;
          if (p < Prows) {
            const int r = p / a.pitch, cx = p - r * a.pitch;
            int sy = (y0 + r) % a.H; if (sy < 0) sy += a.H;
            int sx = (cx - a.pad) % a.W; if (sx < 0) sx += a.W;
            const float* src = xn + (long long)sy * a.W + sx;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float v = (c == ones_c) ? 1.f : 0.f;                     // the constant lane that carries the BatchNorm shift
              if (c < a.cin) v = __ldg(src + c * HW);
              xr[i][c] = v;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split2(xr[i][2 * j], xr[i][2 * j + 1], hw + j, lw + j);
          ph[h + i] = make_uint4(hw[0], hw[1], hw[2], hw[3]); pl[h + i] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      if (lit > 0) mbar_wait<64>(patch_empty, (lit - 1) & 1);        // the previous item's MMAs no longer read the patch
#pragma unroll
      for (int i = 0; i < CS_KP; ++i) {
        const int p = tid + i * CS_CONV_THREADS;
        if (p < a.P) { patch_hi[p] = ph[i]; patch_lo[p] = pl[i]; }    // positions past the last patch row: zeros (only invalid outputs read them)
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(patch_full);
    }
  } else if (warp < CS_WARP_MMA) {
    // =============================== epilogue warps ===============================
    const int e = warp - CS_WARP_EPI0;
    const int lanebase = (e & 3) * 32, half = e >> 2;
    const int PWz = (a.W >> 1) + 1, PHz = (a.H >> 1) + 1;
    const long long plane16 = (long long)PHz * PWz, hl_planes = 4 * plane16;
    const int nblk = a.cout >> 4;
    const float lo_scale = 1.f / 2048.f;
    int gt = 0;
    int tlc = 0; (void)tlc;
    for (int it = it_begin; it < it_end; ++it) {
      const int n = it / a.items_per_sample, b = it - n * a.items_per_sample;
      for (int t = 0; t < a.nt; ++t, ++gt) {
        const int slot = gt & 1;
        CS_TL(1, e == 0 && lane == 0);                                // [0] waiting for the tile
        mbar_wait<64>(tile_full + slot, (gt >> 1) & 1);
        tc_fence_after();
        CS_TL(1, e == 0 && lane == 0);                                // [1] tile complete
        const uint32_t trow = tmem + ((uint32_t)lanebase << 16) + (uint32_t)(slot * tile_cols);
        const int q = t * 128 + lanebase + lane;
        const int oyl = q / a.pitch, ox = q - oyl * a.pitch;
        const bool valid = oyl < a.RI && ox < a.W && !(a.dbg & 1);
        const int oy = b * a.RI + oyl;
        const int yy = (oy >> 1) + 1, xx = (ox >> 1) + 1;
        const bool rowmajor = (a.dbg & 4) != 0;                       // dbg 4 (timing experiment only): Z as [c8][PH][hi|lo][plane][PW] -- the 8 rows of a plane-row index adjacent
        const int pl = (oy & 1) * 2 + (ox & 1);
        const long long hl = rowmajor ? 4 * PWz : hl_planes;
        uint4* zp0 = a.z + (long long)((a.dbg & 2) ? (n & 15) : n) * a.z_bs16 + (rowmajor ? (long long)pl * PWz : (long long)pl * plane16);   // dbg 2: all stores into 16 samples of Z (27 MB, L2-resident)
        const bool hx = xx == PWz - 1, hy = yy == PHz - 1;
        // The accumulator slot is handed back as soon as this warp's LAST block has landed in registers: with two slots in TMEM the time a
        // slot is held stalls the MMAs (scratch/stem_timeline.py).  (A two-buffer software pipeline of the TMEM loads was measured slower:
        // 64 accumulator registers + spills, drain time 2.3-3.0 k -> 3.0-4.8 k cycles per tile.)
        // store pointers advance by constants (one channel octet = 8 planes of hi / lo): the epilogue is issue-bound, and 64-bit address
        // arithmetic per store was a quarter of its instructions
        const long long cstride = (a.dbg & 8) ? 8 * PWz : 8 * plane16;   // 16-byte units between channel octets: [c8][hi|lo][plane 4][plane16]
        const bool rowmajor2 = (a.dbg & 8) != 0;                      // dbg 8: Z as [PH][c8][hi|lo][plane][PW] (with dbg 4): all channels of a plane-row index adjacent
        const int rowpitch = rowmajor ? (rowmajor2 ? (a.cout >> 3) * 8 * PWz : 8 * PWz) : PWz;
        const int off_main = yy * rowpitch + xx;
        const long long d_hx = -(long long)xx, d_hy = -(long long)yy * rowpitch, d_c = -(long long)off_main;   // halo column / row / corner relative to the pixel
        uint4* pm = zp0 + (long long)(2 * half) * cstride + off_main;
        const bool halo = hx || hy;
        bool released = false;
        for (int blk = half; blk < nblk; blk += 2, pm += 4 * cstride) {
          const int c0 = blk * 16;
          float v1[16], v2[16];
          tmem_ld16(trow + c0, v1); tmem_ld16(trow + a.cout + c0, v2);
          tmem_ld_wait();
          if (blk + 2 >= nblk) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tile_empty + slot);
            released = true;
          }
          if (valid) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int i0 = 8 * u + 2 * j;
                // BatchNorm scale folded into the prepared weights, shift into the centre tap of the constant input lane: only ReLU is left
                split2(fmaxf(fmaf(v2[i0], lo_scale, v1[i0]), 0.f), fmaxf(fmaf(v2[i0 + 1], lo_scale, v1[i0 + 1]), 0.f), hw + j, lw + j);
              }
              const uint4 hv = make_uint4(hw[0], hw[1], hw[2], hw[3]), lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              uint4* q = pm + u * cstride;
              uint4* ql = q + hl;
              *q = hv; *ql = lv;
              if (halo) {                                                              // circular halo column / row / corner of the plane
                if (hx) { q[d_hx] = hv; ql[d_hx] = lv; }
                if (hy) { q[d_hy] = hv; ql[d_hy] = lv; }
                if (hx && hy) { q[d_c] = hv; ql[d_c] = lv; }
              }
            }
          }
        }
        if (!released) {                                             // a warp without blocks (cout = 16 and the second half) still arrives
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tile_empty + slot);
        }
        CS_TL(1, e == 0 && lane == 0);                                // [2] drained
      }
    }
  } else if (warp == CS_WARP_MMA) {
    // =============================== MMA issuing warp ===============================
    const bool leader = elect_one();
    const uint32_t idesc1 = umma_idesc_f16(128, 2 * a.cout), idesc2 = umma_idesc_f16(128, a.cout);
    const uint64_t hi64 = (uint64_t)((128u >> 4) | (1u << 14)) << 32;          // SBO = 128 B, descriptor version 1
    const uint32_t a_hi = smem_u32(patch_hi) >> 4, a_lo = smem_u32(patch_lo) >> 4;   // + LBO per pair: where the second K unit sits
    const uint32_t b_base = (smem_u32(wbuf) >> 4) | ((uint32_t)(2 * a.cout) << 16);
    const uint32_t pair16 = 4u * (uint32_t)a.cout, cout = (uint32_t)a.cout;
    const int nt = a.nt, k = a.k, pitch = a.pitch;
    int g_loaded = -1, wgen = 0, gt = 0;
    int tlc = 0; (void)tlc;
    for (int it = it_begin; it < it_end; ++it) {
      const int lit = it - it_begin;
      const int n = it / a.items_per_sample, g = n / a.n_per_group;
      if (g != g_loaded) { mbar_wait<64>(w_full, wgen & 1); ++wgen; g_loaded = g; }
      CS_TL(2, leader);                                               // [0] waiting for the patch
      mbar_wait<32>(patch_full, lit & 1);
      tc_fence_after();
      CS_TL(2, leader);                                               // [1] patch in place
      for (int t = 0; t < nt; ++t, ++gt) {
        const int slot = gt & 1;
        if (gt >= 2) mbar_wait<32>(tile_empty + slot, ((gt >> 1) - 1) & 1);
        tc_fence_after();
        CS_TL(2, leader);                                             // [2 + 2t] accumulator slot free
        if (leader) {
          const uint32_t d = tmem + (uint32_t)slot * (2u * cout);
          uint32_t acc = 0u, wb = b_base;
          auto issue = [&](uint32_t off, uint32_t lbo) {
            const uint64_t db = hi64 | wb;
            umma_f16(d, hi64 | (uint64_t)((a_hi + off) | lbo), db, idesc1, acc);          // hi * [hi ; lo']
            umma_f16(d + cout, hi64 | (uint64_t)((a_lo + off) | lbo), db, idesc2, 1u);     // lo' * hi -> correction half
            acc = 1u; wb += pair16;
          };
          // pair order of stem_pair_taps(): vertical pairs (second K unit one row down), then the last filter row in horizontal pairs
          for (int ky = 0; ky + 1 < k; ky += 2) {
            const uint32_t off = (uint32_t)(t * 128 + ky * pitch);
            for (int kx = 0; kx < k; ++kx) issue(off + (uint32_t)kx, (uint32_t)pitch << 16);
          }
          for (int kx = 0; kx < k; kx += 2) issue((uint32_t)(t * 128 + (k - 1) * pitch + kx), 1u << 16);
        }
        __syncwarp();
        if (leader) umma_commit(tile_full + slot);
        CS_TL(2, leader);                                             // [3 + 2t] tile issued
      }
      if (leader) umma_commit(patch_empty);
    }
    __syncwarp();
  } else {
    // =============================== weight copy warp ===============================
    if (lane == 0) {
      int g_loaded = -1;
      for (int it = it_begin; it < it_end; ++it) {
        const int lit = it - it_begin;
        const int n = it / a.items_per_sample, g = n / a.n_per_group;
        // follow every hand-off in order (an mbarrier parity wait is only meaningful one phase behind): at a group change all MMAs that
        // read the old weight image are then known to be complete
        if (lit > 0) mbar_wait<256>(patch_empty, (lit - 1) & 1);
        if (g != g_loaded) {
          const uint32_t bytes = (uint32_t)a.w16 * 16u;
          mbar_arrive_expect_tx(w_full, bytes);
          const uint4* src = a.wz + (long long)g * a.wz_gs16;
          for (uint32_t o = 0; o < bytes; o += 32768u) {
            const uint32_t nb = min(32768u, bytes - o);
            bulk_g2s(reinterpret_cast<unsigned char*>(wbuf) + o, reinterpret_cast<const unsigned char*>(src) + o, nb, w_full);
          }
          g_loaded = g;
        }
      }
    }
    __syncwarp();
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == CS_WARP_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// w [groups][cout][cin][k][k] -> out [groups][pair][unit 2][row 2 cout][8 halfs]: unit u of pair p holds the tap stem_pair_taps() names
// (zeros where the pair has no second tap), rows [0, cout) the fp16 hi part of output channel `row`, rows [cout, 2 cout)
// fp16((w' - hi) * 2^11) of the folded weight w'
__global__ void conv_stem_wprep_kernel(const float* __restrict__ w, long long w_gs, const float* __restrict__ scale, const float* __restrict__ shift,
                                       long long ss_gs, __half* __restrict__ out, long long out_gs, int cin, int cout, int k, int npairs, int nvert,
                                       int groups) {
  const long long per_group = (long long)npairs * 2 * (2 * cout) * 8;
  const int pad = (k - 1) / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_group * groups; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i / per_group);
    long long r = i - (long long)g * per_group;
    const int ci = (int)(r & 7); r >>= 3;
    const int row = (int)(r % (2 * cout)); r /= 2 * cout;
    const int u = (int)(r & 1); r >>= 1;
    int ky, kx, vertical;
    stem_pair_taps((int)r, k, nvert, &ky, &kx, &vertical);
    if (u) { if (vertical) ++ky; else ++kx; }
    const int co = row < cout ? row : row - cout;
    // the consumer's eval-mode BatchNorm folded into the producer: scale * conv(x, w) + shift = conv([x, 1], [scale * w, shift at the centre tap])
    float v = 0.f;
    if (ci < cin && kx < k && ky < k) v = __ldg(scale + (long long)g * ss_gs + co) * w[(long long)g * w_gs + (((long long)co * cin + ci) * k + ky) * k + kx];
    else if (ci == cin && ky == pad && kx == pad) v = __ldg(shift + (long long)g * ss_gs + co);
    const __half hi = __float2half_rn(v);
    out[(long long)g * out_gs + (i - (long long)g * per_group)] = row < cout ? hi : __float2half_rn((v - __half2float(hi)) * 2048.f);
  }
}

static int stem_nvert(int k) { return ((k - 1) / 2) * k; }
static int stem_npairs(int k) { return stem_nvert(k) + (k + 1) / 2; }

bool cs_plan(const ConvDesc& d, ConvStemArgs& a, size_t* smem_out) {
  if (d.kh != d.kw || !(d.kh & 1) || d.kh > 7 || d.stride != 1 || d.up || d.dilate || d.pad_override >= 0 || d.act != 0 || d.relu_in) return false;
  if (d.cin > 7 || d.cout % 16 || d.cout < 16 || d.cout > 128 || (d.H & 1) || (d.W & 1) || d.H < 4 || d.W < 4 || d.W < d.kh) return false;
  memset(&a, 0, sizeof(a));
  a.cin = d.cin; a.H = d.H; a.W = d.W; a.cout = d.cout; a.k = d.kh; a.pad = (d.kh - 1) / 2; a.nvert = stem_nvert(d.kh); a.npairs = stem_npairs(d.kh);
  a.pitch = d.W + d.kh - 1;
  a.w16 = a.npairs * 2 * 2 * d.cout;
  const size_t budget = 226 * 1024;
  // Rows per work item: the largest block that fits (least halo conversion, no padded tile at 64 columns) unless the batch is so small
  // that whole rounds of the persistent grid go half empty -- then the block whose rounds x tiles-per-item is smallest (240 samples on
  // 148 SMs: 32-row items = 4 rounds x 17 tiles, 16-row items = 7 rounds x 9 tiles).  Blocks below 8 rows are only taken when nothing
  // larger fits.
  const int sms = arpde_num_sms();
  long long best_cost = -1;
  for (int RI = d.H; RI >= 1; --RI) {
    if (d.H % RI) continue;
    if (RI < 8 && best_cost >= 0) break;
    const int nt = (RI * a.pitch + 127) / 128;
    int P = nt * 128 + (a.k - 1) * a.pitch + a.k + 1;
    const int Prows = (RI + a.k - 1) * a.pitch;
    if (P < Prows) P = Prows;
    P = (P + 7) & ~7;
    if (P > CS_KP * CS_CONV_THREADS || P > 16000) continue;
    const size_t smem = CS_HEADER + (size_t)a.w16 * 16 + (size_t)2 * P * 16;
    if (smem > budget) continue;
    const long long items = (long long)(d.N > 0 ? d.N : 1) * (d.H / RI);
    const long long grid = items < sms ? items : sms;
    const long long cost = ((items + grid - 1) / grid) * nt;
    if (best_cost >= 0 && cost >= best_cost) continue;
    best_cost = cost;
    a.RI = RI; a.nt = nt; a.P = P; a.items_per_sample = d.H / RI;
    *smem_out = smem;
  }
  return best_cost >= 0;
}

}  // namespace

bool arpde_conv_stem_wanted(const ConvDesc& d) {
  if (!arpde_tc_enabled() || arpde_dev_env("ARPDE_NO_STEM")) return false;
  ConvStemArgs a; size_t smem;
  return cs_plan(d, a, &smem);
}

// prepared fp16 weights per group in fp32 floats of scratch (<= arpde_conv_tc_wprep_floats of the same layer)
long long arpde_conv_stem_wprep_floats(int cout, int k) { return (long long)stem_npairs(k) * 2 * (2 * cout) * 8 / 2; }

int arpde_conv_stem_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs, const ZOut& zo,
                                cudaStream_t stream) {
  ARPDE_CHECK_ARG(zo.scale && zo.shift && (groups == 1 || zo.ss_gs > 0), "conv_stem: the consumer's BatchNorm scale / shift are folded into the prepared weights");
  const long long n = (long long)stem_npairs(d.kh) * 2 * (2 * d.cout) * 8 * groups;
  long long blocks = (n + 255) / 256; if (blocks > 4096) blocks = 4096;
  conv_stem_wprep_kernel<<<(unsigned)blocks, 256, 0, stream>>>(w, w_gs, zo.scale, zo.shift, groups > 1 ? zo.ss_gs : 0, reinterpret_cast<__half*>(wprep), 2 * wprep_gs,
                                                               d.cin, d.cout, d.kh, stem_npairs(d.kh), stem_nvert(d.kh), groups);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

int arpde_conv_stem_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const ZOut& zo, cudaStream_t stream) {
  ARPDE_CHECK_ARG(x && wprep && zo.z, "conv_stem: null pointer");
  if (d.N == 0) return ARPDE_OK;
  ConvStemArgs a; size_t smem;
  if (!cs_plan(d, a, &smem)) { arpde_set_error("conv_stem: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  ARPDE_CHECK_ARG((((uintptr_t)zo.z) & 15) == 0 && (((uintptr_t)wprep) & 15) == 0 && (wprep_gs & 3) == 0, "conv_stem: operands must be 16-byte aligned");
  a.x = x; a.x_bs = x_bs; a.wz = reinterpret_cast<const uint4*>(wprep); a.wz_gs16 = wprep_gs / 4;
  a.z = reinterpret_cast<uint4*>(zo.z); a.z_bs16 = arpde_conv_z_sample16(d.cout, d.H, d.W);
  a.N = d.N; a.n_per_group = d.n_per_group; a.tl = arpde_tc_debug_buffer();
  a.dbg = arpde_dev_env("ARPDE_STEM_DBG") ? atoi(arpde_dev_env("ARPDE_STEM_DBG")) : 0;
  static std::mutex attr_mutex;
  static bool attr_done[64] = {false};
  {
    int dev = 0, optin = 0;
    ARPDE_CUDA(cudaGetDevice(&dev));
    ARPDE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    ARPDE_CHECK_ARG(smem <= (size_t)optin, "conv_stem: plan needs %zu bytes of shared memory, device allows %d", smem, optin);
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      ARPDE_CUDA(cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  const long long items = (long long)d.N * a.items_per_sample;
  ARPDE_CHECK_ARG(items < (1ll << 31), "conv_stem: too many work items");
  long long blocks = (long long)arpde_num_sms();
  if (blocks > items) blocks = items;
  ARPDE_CUDA(arpde_launch1(conv_stem_kernel, (unsigned)blocks, CS_THREADS, smem, stream, d.pdl != 0, a));
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
