// conv_z.cu -- eval-mode "stem pair": the stride-2 3x3 convolution that follows a wide convolution + BatchNorm + ReLU
// (In_conv1 -> In_conv_norm1 -> In_conv_relu1 -> In_conv3 of 2D-Burgers-SWAG/nn/denseEDcirc2d.py:247-255), with the intermediate
// activation handed over in the consumer's tensor-core operand format instead of fp32 NCHW.
//
// Why: on the generic kernel (conv_tc.cu) In_conv3 spends its time moving bytes through shared memory -- raw fp32 rows in (TMA), a
// shared->shared pass that applies BatchNorm + ReLU, splits into TF32 hi/lo and re-arranges into the UMMA layout, then K = 8 MMAs that
// read 4 KB of A per 128x8 MACs -- and the raw ring + patch + streamed weights fill the 227 KB so that conversion and MMAs cannot
// overlap (one patch slot).  Here the PRODUCER (conv_tc's epilogue, `zout`) applies the eval-mode BatchNorm affine + ReLU once and
// writes the activation already split into two fp16 terms, channel-octet interleaved and parity-split:
//
//     Z[n][c8 = C/8][hl = hi|lo][plane = (y&1)*2 + (x&1)][PH = H/2 + 1][PW = W/2 + 1][8 halfs]      (16 bytes per position)
//
// with the circular halo row / column of every parity plane stored in place (row 0 = plane row H/2-1, column 0 = plane column W/2-1),
// so that ANY run of plane rows is one contiguous piece of memory which the consumer brings in with a single 1D bulk copy
// (cp.async.bulk) straight into the K-major no-swizzle UMMA operand layout [c8][position][16 B] -- no conversion warps, no raw ring.
// A tap of the 3x3 stride-2 filter is the same buffer with the descriptor start advanced by (row*PW + col)*16 B (the shifted-window
// formulation of conv_tc.cu on a padded image of pitch PW = Wo + 1).
//
// fp32-grade accuracy from fp16 operands: x = x_hi + 2^-11 x_lo' with x_hi = fp16(x), x_lo' = fp16((x - x_hi) * 2^11) (22 mantissa
// bits; scaling the residual keeps it a normal fp16 number down to |x| ~ 6e-5 and bounds the absolute error by 2^-36 below), the same
// for the weights; per K = 16 step (tcgen05.mma.kind::f16, fp32 accumulators in TMEM)
//     D[:, 0:2N] += A_hi * [W_hi ; W_lo']^T        D[:, N:2N] += A_lo' * W_hi^T          y = D[:, 0:N] + 2^-11 D[:, N:2N]
// Per MAC this moves half the shared-memory bytes and issues half the MMAs of the 3xTF32 scheme.  Range: |x| < 65504 after BN + ReLU
// (fp16 overflow saturates to the largest finite value; the reference's activations are O(1)).
//
// Kernel (persistent, 6 warps, one CTA per SM): work item = (sample, `ntile` M tiles of 128 padded positions).  A pipeline stage is one
// (16-channel chunk, parity plane): 4 bulk copies of the plane rows (2 octets x hi/lo) + 1 of the weights of the plane's taps; a ring
// of 3-6 stages connects the loader lane to the MMA-issuing warp (mbarriers, tcgen05.commit frees a stage).  Accumulators: a ring of
// 512/(2N) tile slots in TMEM, freed tile by tile by the 4 epilogue warps, so the MMAs of the next item start while the previous one
// drains.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"
#include <cuda_fp16.h>
#include <mutex>
#include "tc_ptx.cuh"

#define CZ_THREADS 192
#define CZ_MAX_STAGES 6
#define CZ_MAX_TSLOTS 8
#define CZ_HEADER 256

namespace {

struct ConvZArgs {
  const uint4* z; long long z_bs16;         // activation in operand format, per-sample stride in 16-byte units
  const uint4* wz; long long wz_gs16;       // prepared weights, per-group stride in 16-byte units
  float* y; int y_ctot, y_coff;
  int N, n_per_group;
  int ncc, PH, PW, Ho, Wo, cout, coutP;
  int ntile, items_per_sample, total_tiles, nts, tile_cols, tmem_cols;
  int SPP, nstages, stage16;                // positions per octet in a stage; stage size in 16-byte units
};

__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {      // D = f32, A = B = f16 (format 0), K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ void item_rows(const ConvZArgs& a, int b, int* q0, int* nt, int* r0, int* R) {
  *q0 = b * a.ntile * 128;
  *nt = min(a.ntile, a.total_tiles - b * a.ntile);
  *r0 = *q0 / a.PW;
  const int r1 = min(a.PH - 1, (*q0 + *nt * 128 + a.PW) / a.PW);
  *R = r1 - *r0 + 1;
}

__global__ void __launch_bounds__(CZ_THREADS, 1) conv_z_kernel(const __grid_constant__ ConvZArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);            // [CZ_MAX_STAGES] stage landed (bulk-copy tx bytes)
  uint64_t* empty = full + CZ_MAX_STAGES;                             // [CZ_MAX_STAGES] stage consumed (tcgen05.commit)
  uint64_t* item_full = empty + CZ_MAX_STAGES;                        // [2] all accumulators of an item complete
  uint64_t* tmem_empty = item_full + 2;                               // [CZ_MAX_TSLOTS] tile slot drained (one arrival per epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + 240);
  unsigned char* stages = smem_raw + CZ_HEADER;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int items_total = a.N * a.items_per_sample;
  // Items are dealt round-robin: the items of one sample (which share halo rows of Z) then run on neighbouring SMs at the same time and
  // the second read of a halo row hits L2 -- with a contiguous block per CTA they were ~20 us (= 100 MB of traffic) apart and the launch
  // read 3.70 GB for 3.21 GB of Z.  The weights ride in the stage ring anyway, so a weight-group change costs nothing here.
  const int it_begin = (int)blockIdx.x, it_step = (int)gridDim.x, it_end = items_total;
  const int tap16 = 4 * a.coutP;                       // 16-byte units of one tap's weights: 2 octets x 2*coutP rows
  const uint32_t stage_bytes = (uint32_t)a.stage16 * 16u;

  if (tid == 0) {
    for (int s = 0; s < CZ_MAX_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) mbar_init(item_full + s, 1);
    for (int s = 0; s < CZ_MAX_TSLOTS; ++s) mbar_init(tmem_empty + s, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  // taps of the 3x3 stride-2 filter by parity plane: (count, first index in the prepared weight order, offsets in padded positions)
  const int PW = a.PW;

  if (warp == 0) {
    // =============================== loader ===============================
    if (lane == 0) {
      const long long plane16 = (long long)a.PH * PW;
      int sseq = 0;
      for (int it = it_begin; it < it_end; it += it_step) {
        const int n = it / a.items_per_sample, b = it - n * a.items_per_sample, g = n / a.n_per_group;
        int q0, nt, r0, R;
        item_rows(a, b, &q0, &nt, &r0, &R);
        const uint32_t bytesA = (uint32_t)(R * PW) * 16u;
        const uint4* zn = a.z + (long long)n * a.z_bs16 + (long long)r0 * PW;
        const uint4* wg = a.wz + (long long)g * a.wz_gs16;
        for (int cc = 0; cc < a.ncc; ++cc)
          for (int p = 0; p < 4; ++p, ++sseq) {
            const int s = sseq % a.nstages, k = sseq / a.nstages;
            if (k > 0) mbar_wait<64>(empty + s, (k - 1) & 1);
            const int ntap = p == 0 ? 1 : (p == 3 ? 4 : 2), tap0 = p == 0 ? 0 : (p == 1 ? 1 : (p == 2 ? 3 : 5));
            const uint32_t bytesW = (uint32_t)(ntap * tap16) * 16u;
            mbar_arrive_expect_tx(full + s, 4u * bytesA + bytesW);
            unsigned char* st = stages + (size_t)s * stage_bytes;
#pragma unroll
            for (int hl = 0; hl < 2; ++hl)
#pragma unroll
              for (int u = 0; u < 2; ++u)
                bulk_g2s(st + (size_t)(hl * 2 + u) * a.SPP * 16, zn + ((long long)((2 * cc + u) * 2 + hl) * 4 + p) * plane16, bytesA, full + s);
            bulk_g2s(st + (size_t)4 * a.SPP * 16, wg + (long long)(cc * 9 + tap0) * tap16, bytesW, full + s);
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== MMA issuing warp ===============================
    const bool leader = elect_one();
    const uint32_t idesc1 = umma_idesc_f16(128, 2 * a.coutP), idesc2 = umma_idesc_f16(128, a.coutP);
    const uint64_t hi64 = (uint64_t)((128u >> 4) | (1u << 14)) << 32;          // SBO = 128 B, descriptor version 1
    const uint32_t SPP = (uint32_t)a.SPP, coutP = (uint32_t)a.coutP, tile_cols = (uint32_t)a.tile_cols;
    const uint32_t a_lbo = SPP << 16, b_lbo = (2u * coutP) << 16;
    const int nstages = a.nstages, nts = a.nts, ncc = a.ncc;
    int sseq = 0, gt = 0;
    for (int it = it_begin, lit = 0; it < it_end; it += it_step, ++lit) {
      const int n = it / a.items_per_sample, b = it - n * a.items_per_sample;
      int q0, nt, r0, R;
      item_rows(a, b, &q0, &nt, &r0, &R);
      const uint32_t qlocal0 = (uint32_t)(q0 - r0 * PW);
      for (int t = 0; t < nt; ++t) {                      // the tile slots this item accumulates into must have been drained
        const int gg = gt + t, ts = gg % nts, kk = gg / nts;
        if (kk > 0) mbar_wait<64>(tmem_empty + ts, (kk - 1) & 1);
      }
      tc_fence_after();
      for (int cc = 0; cc < ncc; ++cc)
        for (int p = 0; p < 4; ++p, ++sseq) {
          const int s = sseq % nstages, k = sseq / nstages;
          mbar_wait<32>(full + s, k & 1);
          tc_fence_after();
          if (leader) {
            const uint32_t st16 = smem_u32(stages + (size_t)s * stage_bytes) >> 4;
            const uint32_t a_base = st16 + qlocal0, b_base = st16 + 4u * SPP;
            const int ntap = p == 0 ? 1 : (p == 3 ? 4 : 2);
            uint32_t off[4];
            off[0] = p == 0 ? PW + 1 : (p == 1 ? PW : (p == 2 ? 1 : 0));
            off[1] = p == 3 ? 1 : PW + 1;
            off[2] = PW; off[3] = PW + 1;
            for (int t = 0; t < nt; ++t) {
              const uint32_t d = tmem + (uint32_t)((gt + t) % nts) * tile_cols;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (j < ntap) {
                  const uint64_t da_h = hi64 | (uint64_t)((a_base + (uint32_t)(t * 128) + off[j]) | a_lbo);
                  const uint64_t da_l = da_h + 2u * SPP;
                  const uint64_t db = hi64 | (uint64_t)((b_base + (uint32_t)(j * tap16)) | b_lbo);
                  umma_f16(d, da_h, db, idesc1, (cc | p | j) ? 1u : 0u);      // hi * [hi ; lo']
                  umma_f16(d + coutP, da_l, db, idesc2, 1u);                  // lo' * hi -> correction half
                }
              }
            }
          }
          __syncwarp();
          if (leader) umma_commit(empty + s);
        }
      if (leader) umma_commit(item_full + (lit & 1));
      gt += nt;
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps ===============================
    const int lanebase = (warp & 3) * 32;                 // the TMEM lane quarter this warp may read
    const long long HoWo = (long long)a.Ho * a.Wo;
    const int nts = a.nts;
    const float lo_scale = 1.f / 2048.f;
    int gt = 0;
    for (int it = it_begin, lit = 0; it < it_end; it += it_step, ++lit) {
      const int n = it / a.items_per_sample, b = it - n * a.items_per_sample;
      int q0, nt, r0, R;
      item_rows(a, b, &q0, &nt, &r0, &R);
      mbar_wait<128>(item_full + (lit & 1), (lit >> 1) & 1);
      tc_fence_after();
      for (int t = 0; t < nt; ++t) {
        const int ts = (gt + t) % nts;
        const uint32_t trow = tmem + ((uint32_t)lanebase << 16) + (uint32_t)(ts * a.tile_cols);
        const int q = q0 + t * 128 + lanebase + lane;
        const int oy = q / PW, ox = q - oy * PW;
        const bool valid = oy < a.Ho && ox < a.Wo;
        float* yp = a.y + ((long long)n * a.y_ctot + a.y_coff) * HoWo + (long long)oy * a.Wo + ox;
        for (int c0 = 0; c0 < a.cout; c0 += 16) {
          float v1[16], v2[16];
          tmem_ld16(trow + c0, v1); tmem_ld16(trow + a.coutP + c0, v2);
          tmem_ld_wait();
          if (valid) {
            float* pc = yp + (long long)c0 * HoWo;
            if (c0 + 16 <= a.cout) {
#pragma unroll
              for (int j = 0; j < 16; ++j) { *pc = fmaf(v2[j], lo_scale, v1[j]); pc += HoWo; }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) { if (c0 + j < a.cout) *pc = fmaf(v2[j], lo_scale, v1[j]); pc += HoWo; }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + ts);
      }
      gt += nt;
    }
  }
  pdl_trigger();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(a.tmem_cols));
}

// w [groups][cout][cin][3][3] -> out [groups][cc = cin/16][tap (plane order)][octet 2][row 2*coutP][8 halfs]: rows [0, coutP) = fp16 hi part
// of output channel `row`, rows [coutP, 2 coutP) = fp16((w - hi) * 2^11); exactly the shared-memory image of a stage's weight piece.
__global__ void conv_z_wprep_kernel(const float* __restrict__ w, long long w_gs, __half* __restrict__ out, long long out_gs, int cin, int cout,
                                    int coutP, int groups) {
  const int kyx[9][2] = {{1, 1}, {1, 0}, {1, 2}, {0, 1}, {2, 1}, {0, 0}, {0, 2}, {2, 0}, {2, 2}};
  const long long per_group = (long long)(cin / 16) * 9 * 2 * (2 * coutP) * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_group * groups; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i / per_group);
    long long r = i - (long long)g * per_group;
    const int j = (int)(r & 7); r >>= 3;
    const int row = (int)(r % (2 * coutP)); r /= 2 * coutP;
    const int u = (int)(r & 1); r >>= 1;
    const int tap = (int)(r % 9), cc = (int)(r / 9);
    const int ci = (2 * cc + u) * 8 + j, co = row < coutP ? row : row - coutP;
    float v = 0.f;
    if (co < cout) v = w[(long long)g * w_gs + ((long long)co * cin + ci) * 9 + kyx[tap][0] * 3 + kyx[tap][1]];
    const __half hi = __float2half_rn(v);
    out[(long long)g * out_gs + (i - (long long)g * per_group)] = row < coutP ? hi : __float2half_rn((v - __half2float(hi)) * 2048.f);
  }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

bool cz_plan(const ConvDesc& d, ConvZArgs& a, size_t* smem_out) {
  if (d.kh != 3 || d.kw != 3 || d.stride != 2 || d.up || d.dilate || d.pad_override >= 0 || d.act != 0) return false;
  if (d.cin % 16 || d.cin > 512 || d.cout < 1 || d.cout > 64 || (d.H & 1) || (d.W & 1) || d.H < 4 || d.W < 4) return false;
  memset(&a, 0, sizeof(a));
  a.Ho = d.H / 2; a.Wo = d.W / 2; a.PH = a.Ho + 1; a.PW = a.Wo + 1;
  a.cout = d.cout; a.coutP = round_up(d.cout, 16); a.ncc = d.cin / 16;
  a.tile_cols = 2 * a.coutP;
  a.nts = 512 / a.tile_cols; if (a.nts > CZ_MAX_TSLOTS) a.nts = CZ_MAX_TSLOTS;
  a.total_tiles = (a.Ho * a.PW + 127) / 128;
  int p2 = 32; while (p2 < a.nts * a.tile_cols) p2 <<= 1;
  a.tmem_cols = p2;
  const size_t budget = 226 * 1024;
  for (int ntile = 3; ntile >= 1; --ntile) {
    if (ntile > a.nts) continue;
    const int nt = ntile < a.total_tiles ? ntile : a.total_tiles;
    const int SPP = round_up(nt * 128 + 3 * a.PW, 8);
    if (SPP > 16383) continue;
    const int stage16 = 4 * SPP + 4 * 4 * a.coutP;
    int nstages = (int)((budget - CZ_HEADER) / ((size_t)stage16 * 16));
    if (nstages > CZ_MAX_STAGES) nstages = CZ_MAX_STAGES;
    if (nstages < 3) continue;
    a.ntile = nt; a.SPP = SPP; a.stage16 = stage16; a.nstages = nstages;
    a.items_per_sample = (a.total_tiles + nt - 1) / nt;
    *smem_out = CZ_HEADER + (size_t)nstages * stage16 * 16;
    return true;
  }
  return false;
}

}  // namespace

// 16-byte units per sample of the operand-format activation of a [c, H, W] map
long long arpde_conv_z_sample16(int c, int H, int W) { return (long long)(c / 8) * 2 * 4 * (H / 2 + 1) * (W / 2 + 1); }

// fp32 floats of work-buffer room per sample that a [c, H, W] activation needs so that it can also be held in operand format (0: the
// shape has no such form).  Python's workspace layout calls this for single-producer buffers.
extern "C" int64_t arpde_conv_z_floats(int c, int H, int W) {
  if (c <= 0 || c % 16 || H < 4 || W < 4 || (H & 1) || (W & 1)) return 0;
  return 4 * arpde_conv_z_sample16(c, H, W);
}

// the consumer side of the pair: is `d` (the stride-2 layer, described by its INPUT dims) covered?
bool arpde_conv_z_wanted(const ConvDesc& d) {
  if (!arpde_tc_enabled() || arpde_dev_env("ARPDE_NO_Z")) return false;
  ConvZArgs a; size_t smem;
  return d.relu_in && cz_plan(d, a, &smem);
}

// prepared fp16 weights per group, in fp32 floats of scratch (always <= arpde_conv_tc_wprep_floats of the same layer)
long long arpde_conv_z_wprep_floats(int cin, int cout) { return (long long)(cin / 16) * 9 * 2 * (2 * round_up(cout, 16)) * 8 / 2; }

int arpde_conv_z_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs, cudaStream_t stream) {
  const int coutP = round_up(d.cout, 16);
  const long long n = (long long)(d.cin / 16) * 9 * 2 * (2 * coutP) * 8 * groups;
  long long blocks = (n + 255) / 256; if (blocks > 4096) blocks = 4096;
  conv_z_wprep_kernel<<<(unsigned)blocks, 256, 0, stream>>>(w, w_gs, reinterpret_cast<__half*>(wprep), 2 * wprep_gs, d.cin, d.cout, coutP, groups);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

int arpde_conv_z_launch(const ConvDesc& d, const float* z, const float* wprep, long long wprep_gs, float* y, cudaStream_t stream) {
  ARPDE_CHECK_ARG(z && wprep && y, "conv_z: null pointer");
  if (d.N == 0) return ARPDE_OK;
  ConvZArgs a; size_t smem;
  if (!cz_plan(d, a, &smem)) { arpde_set_error("conv_z: layer not supported"); return ARPDE_ERR_UNSUPPORTED; }
  ARPDE_CHECK_ARG((((uintptr_t)z) & 15) == 0 && (((uintptr_t)wprep) & 15) == 0 && (wprep_gs & 3) == 0, "conv_z: operands must be 16-byte aligned");
  a.z = reinterpret_cast<const uint4*>(z); a.z_bs16 = arpde_conv_z_sample16(d.cin, d.H, d.W);
  a.wz = reinterpret_cast<const uint4*>(wprep); a.wz_gs16 = wprep_gs / 4;
  a.y = y; a.y_ctot = d.y_ctot; a.y_coff = d.y_coff; a.N = d.N; a.n_per_group = d.n_per_group;
  static std::mutex attr_mutex;
  static bool attr_done[64] = {false};
  {
    int dev = 0, optin = 0;
    ARPDE_CUDA(cudaGetDevice(&dev));
    ARPDE_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    ARPDE_CHECK_ARG(smem <= (size_t)optin, "conv_z: plan needs %zu bytes of shared memory, device allows %d", smem, optin);
    std::lock_guard<std::mutex> lock(attr_mutex);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
      ARPDE_CUDA(cudaFuncSetAttribute(conv_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
      if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
  }
  const long long items = (long long)d.N * a.items_per_sample;
  ARPDE_CHECK_ARG(items < (1ll << 31), "conv_z: too many work items");
  long long blocks = (long long)arpde_num_sms();
  if (blocks > items) blocks = items;
  ARPDE_CUDA(arpde_launch1(conv_z_kernel, (unsigned)blocks, CZ_THREADS, smem, stream, d.pdl != 0, a));
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// C ABI: the pair on its own -- y = conv3x3_s2(relu(scale * conv_k1(x) + shift)), both convolutions circular, the first with
// stride 1 on the tensor-core kernel of conv_tc.cu writing operand format, the second by conv_z_kernel.  z: N * arpde_conv_z_floats(c1, H, W)
// floats; scratch: arpde_conv_fwd_tc_scratch_floats(cin, c1, k1, k1, groups) + arpde_conv_fwd_tc_scratch_floats(c1, c3, 3, 3, groups) floats.
extern "C" int arpde_stem_pair_fwd(const float* x, int64_t x_bstride, const float* w1, int64_t w1_gstride, const float* scale, const float* shift,
                                   int64_t ss_gstride, const float* w3, int64_t w3_gstride, float* y, int N, int cin, int H, int W, int c1, int k1,
                                   int c3, int y_ctot, int y_coff, int n_per_group, float* z, int64_t z_floats, float* scratch,
                                   int64_t scratch_floats, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(x && w1 && scale && shift && w3 && y && z && scratch, "stem_pair_fwd: null pointer");
  ARPDE_CHECK_ARG(N >= 0 && cin > 0 && c1 > 0 && c3 > 0 && H > 0 && W > 0, "stem_pair_fwd: bad dims");
  ARPDE_CHECK_ARG(y_coff >= 0 && y_coff + c3 <= y_ctot, "stem_pair_fwd: channel window [%d,%d) outside y_ctot=%d", y_coff, y_coff + c3, y_ctot);
  if (N == 0) return ARPDE_OK;
  ConvDesc d1;
  d1.N = N; d1.cin = cin; d1.H = H; d1.W = W; d1.cout = c1; d1.kh = k1; d1.kw = k1; d1.stride = 1; d1.up = 0; d1.relu_in = 0; d1.act = 0;
  d1.y_ctot = c1; d1.y_coff = 0; d1.n_per_group = n_per_group > 0 ? n_per_group : N;
  ARPDE_CHECK_ARG(N % d1.n_per_group == 0, "stem_pair_fwd: N=%d not a multiple of n_per_group=%d", N, d1.n_per_group);
  ConvDesc d3 = d1;
  d3.cin = c1; d3.cout = c3; d3.kh = 3; d3.kw = 3; d3.stride = 2; d3.relu_in = 1; d3.y_ctot = y_ctot; d3.y_coff = y_coff;
  const int groups = N / d1.n_per_group;
  const long long zneed = arpde_conv_z_floats(c1, H, W) * N;
  ARPDE_CHECK_ARG(arpde_conv_z_floats(c1, H, W) > 0 && arpde_conv_z_wanted(d3), "stem_pair_fwd: shape not covered (needs 16 | c1, even H, W >= 4, c3 <= 64)");
  ARPDE_CHECK_ARG(z_floats >= zneed, "stem_pair_fwd: z scratch too small (%lld < %lld floats)", (long long)z_floats, zneed);
  const long long p1 = arpde_conv_tc_wprep_floats(cin, c1, k1, k1), p3 = arpde_conv_tc_wprep_floats(c1, c3, 3, 3);
  ARPDE_CHECK_ARG(scratch_floats >= (p1 + p3) * groups, "stem_pair_fwd: scratch too small (%lld < %lld floats)", (long long)scratch_floats, (p1 + p3) * groups);
  const bool stem = arpde_conv_stem_wanted(d1);            // few input channels: the dedicated fp16 producer (conv_stem.cu)
  ZOut zo; zo.z = z; zo.scale = scale; zo.shift = shift; zo.ss_gs = ss_gstride;
  int rc = stem ? arpde_conv_stem_prep_launch(d1, w1, w1_gstride, groups, scratch, p1, zo, stream)
                : arpde_conv_tc_prep_launch(d1, w1, w1_gstride, groups, scratch, p1, stream);
  if (rc) return rc;
  float* s3 = scratch + p1 * groups;
  rc = arpde_conv_z_prep_launch(d3, w3, w3_gstride, groups, s3, p3, stream); if (rc) return rc;
  rc = stem ? arpde_conv_stem_launch(d1, x, x_bstride, scratch, p1, zo, stream)
            : arpde_conv_tc_launch(d1, x, x_bstride, scratch, p1, nullptr, nullptr, 0, nullptr, nullptr, stream, &zo);
  if (rc) return rc;
  return arpde_conv_z_launch(d3, z, s3, p3, y, stream);
}

// The two halves of the pair on their own (per-kernel timing in bench.py, tests): producer x -> z (weights prepared into `scratch`,
// arpde_conv_fwd_tc_scratch_floats(cin, c1, k1, k1, groups) floats), consumer z -> y (arpde_conv_fwd_tc_scratch_floats(c1, c3, 3, 3, groups)).
extern "C" int arpde_stem_pair_producer(const float* x, int64_t x_bstride, const float* w1, int64_t w1_gstride, const float* scale, const float* shift,
                                        int64_t ss_gstride, int N, int cin, int H, int W, int c1, int k1, int n_per_group, float* z, int64_t z_floats,
                                        float* scratch, int64_t scratch_floats, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(x && w1 && scale && shift && z && scratch, "stem_pair_producer: null pointer");
  ARPDE_CHECK_ARG(N >= 0 && cin > 0 && c1 > 0 && H > 0 && W > 0, "stem_pair_producer: bad dims");
  if (N == 0) return ARPDE_OK;
  ConvDesc d1;
  d1.N = N; d1.cin = cin; d1.H = H; d1.W = W; d1.cout = c1; d1.kh = k1; d1.kw = k1; d1.stride = 1; d1.up = 0; d1.relu_in = 0; d1.act = 0;
  d1.y_ctot = c1; d1.y_coff = 0; d1.n_per_group = n_per_group > 0 ? n_per_group : N;
  ARPDE_CHECK_ARG(N % d1.n_per_group == 0, "stem_pair_producer: N=%d not a multiple of n_per_group=%d", N, d1.n_per_group);
  const int groups = N / d1.n_per_group;
  ARPDE_CHECK_ARG(arpde_conv_z_floats(c1, H, W) > 0 && z_floats >= arpde_conv_z_floats(c1, H, W) * N, "stem_pair_producer: z scratch too small or shape not covered");
  const long long p1 = arpde_conv_tc_wprep_floats(cin, c1, k1, k1);
  ARPDE_CHECK_ARG(scratch_floats >= p1 * groups, "stem_pair_producer: scratch too small (%lld < %lld floats)", (long long)scratch_floats, p1 * groups);
  const bool stem = arpde_conv_stem_wanted(d1);
  ZOut zo; zo.z = z; zo.scale = scale; zo.shift = shift; zo.ss_gs = ss_gstride;
  int rc = stem ? arpde_conv_stem_prep_launch(d1, w1, w1_gstride, groups, scratch, p1, zo, stream)
                : arpde_conv_tc_prep_launch(d1, w1, w1_gstride, groups, scratch, p1, stream);
  if (rc) return rc;
  return stem ? arpde_conv_stem_launch(d1, x, x_bstride, scratch, p1, zo, stream)
              : arpde_conv_tc_launch(d1, x, x_bstride, scratch, p1, nullptr, nullptr, 0, nullptr, nullptr, stream, &zo);
}

extern "C" int arpde_stem_pair_consumer(const float* z, const float* w3, int64_t w3_gstride, float* y, int N, int H, int W, int c1, int c3, int y_ctot,
                                        int y_coff, int n_per_group, float* scratch, int64_t scratch_floats, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(z && w3 && y && scratch, "stem_pair_consumer: null pointer");
  ARPDE_CHECK_ARG(N >= 0 && c1 > 0 && c3 > 0 && H > 0 && W > 0, "stem_pair_consumer: bad dims");
  ARPDE_CHECK_ARG(y_coff >= 0 && y_coff + c3 <= y_ctot, "stem_pair_consumer: channel window [%d,%d) outside y_ctot=%d", y_coff, y_coff + c3, y_ctot);
  if (N == 0) return ARPDE_OK;
  ConvDesc d3;
  d3.N = N; d3.cin = c1; d3.H = H; d3.W = W; d3.cout = c3; d3.kh = 3; d3.kw = 3; d3.stride = 2; d3.up = 0; d3.relu_in = 1; d3.act = 0;
  d3.y_ctot = y_ctot; d3.y_coff = y_coff; d3.n_per_group = n_per_group > 0 ? n_per_group : N;
  ARPDE_CHECK_ARG(N % d3.n_per_group == 0, "stem_pair_consumer: N=%d not a multiple of n_per_group=%d", N, d3.n_per_group);
  ARPDE_CHECK_ARG(arpde_conv_z_wanted(d3), "stem_pair_consumer: shape not covered (needs 16 | c1, even H, W >= 4, c3 <= 64)");
  const int groups = N / d3.n_per_group;
  const long long p3 = arpde_conv_tc_wprep_floats(c1, c3, 3, 3);
  ARPDE_CHECK_ARG(scratch_floats >= p3 * groups, "stem_pair_consumer: scratch too small (%lld < %lld floats)", (long long)scratch_floats, p3 * groups);
  int rc = arpde_conv_z_prep_launch(d3, w3, w3_gstride, groups, scratch, p3, stream); if (rc) return rc;
  return arpde_conv_z_launch(d3, z, scratch, p3, y, stream);
}
