// dense_block.cu -- a whole dense block of the 2D DenseED in ONE launch (eval mode: rollout / inference).
//
// Replaces _DenseBlock.forward (2D-Burgers-SWAG/nn/denseEDcirc2d.py:34-79): num_layers x [BatchNorm, ReLU, circular conv3x3 -> growth = 4
// channels, torch.cat].  Layer by layer the block costs one launch per dense layer and each launch re-reads all 48..60 input channels
// from HBM (1.86 ms of the 6.8 ms AR step at 1920 samples for 8 % of the FLOPs, ~16 TFLOP/s).  Here a CTA owns 16 rows of one sample:
//   * the 48 raw input channels of its rows plus `layers` halo rows on either side (circular in y) are loaded from HBM ONCE into shared
//     memory (147 KB), every layer appends its 4 channels there, and only the 16 owned rows of the new channels are written back; the
//     halo rows of layer j are recomputed (rows shrink by one per layer and side: 22, 20, 18, 16 -- 19 % extra FLOPs instead of a
//     cluster-wide exchange);
//   * a row is 32 columns = 32 lanes, so the x-neighbours of a pixel come from the neighbouring lanes by shuffle and the circular wrap in
//     x is the shuffle's modulo; BN + ReLU is applied once per loaded value (216 FMAs per 8 shared loads + 16 shuffles + 9 broadcast weight loads);
//   * a warp computes 6 rows x 4 output channels per lane; the 3-4 row groups of a layer are each shared by 4-5 warps that split the
//     input channels (partial sums meet in shared memory in a fixed two-round tree: bit-reproducible), which keeps 15-16 warps busy.
// Bound: fp32 FMA pipe (4-channel layers cannot feed a tensor-core tile: an MMA costs >= 32 cycles whatever N is).
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"

namespace {

constexpr int DB_W = 32, DB_OWN = 16, DB_R = 6, DB_WARPS = 16, DB_MAXL = 4, DB_MAXC = 64, DB_KSMAX = 5, DB_RED = 8 * DB_R * 4 * DB_W;

struct DenseBlockArgs {
  float* buf; long long bs;                     // dense-block buffer [N][ctot][H][32]
  int c0, nl, H, N, n_per_group;
  const float* w[DB_MAXL]; long long w_gs[DB_MAXL];   // layer weights [4][cin][3][3], group stride
  const float* scale; const float* shift; long long ss_gs; int bn_off[DB_MAXL];   // folded eval BatchNorm of layer j: scale[g*ss_gs + bn_off[j] + ci]
};

__global__ void __launch_bounds__(512, 1) dense_block_kernel(DenseBlockArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int rows = DB_OWN + 2 * a.nl;
  const int ctot = a.c0 + 4 * a.nl;
  float* X = sm;                                          // [ctot][rows][32]
  float* Wt = X + (size_t)ctot * rows * DB_W;             // [cin][9][4]
  float* SS = Wt + (size_t)(ctot - 4) * 36;               // [2][cin]
  float* RED = SS + 2 * DB_MAXC;                          // [4 items][2 slots][DB_R * 4][32]: partial sums of the channel slices
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int tiles = a.H / DB_OWN;
  const long long HW = (long long)a.H * DB_W;
  const int lm = (lane + 31) & 31, lp = (lane + 1) & 31;
  for (int work = blockIdx.x; work < a.N * tiles; work += gridDim.x) {
    const int n = work / tiles, y0 = (work - n * tiles) * DB_OWN, g = n / a.n_per_group;
    float* bn_ = a.buf + (long long)n * a.bs;
    // ---- raw input channels of rows y0 - nl .. y0 + 16 + nl - 1 (circular): one float4 per lane covers 4 rows per warp step
    {
      const int nrow = a.c0 * rows;                       // (channel, local row) pairs, 128 B each
      for (int q = warp * 4 + (lane >> 3); q < nrow; q += DB_WARPS * 4) {
        const int c = q / rows, r = q - c * rows;
        int y = y0 - a.nl + r; y = y < 0 ? y + a.H : (y >= a.H ? y - a.H : y);
        // 16-byte cp.async: all of a thread's ~18 row pieces are in flight at once (a load / store loop kept one: 28 k cycles per item)
        const float* src = bn_ + c * HW + (long long)y * DB_W + 4 * (lane & 7);
        float* dst = X + ((size_t)c * rows + r) * DB_W + 4 * (lane & 7);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int j = 0; j < a.nl; ++j) {
      const int cin = a.c0 + 4 * j;
      {  // stage the layer's weights as [ci][tap][co] and its folded BatchNorm
        const float* w = a.w[j] + (long long)g * a.w_gs[j];
        const int per_co = 9 * cin;
        for (int i = tid; i < 4 * per_co; i += 512) { const int co = i / per_co, r = i - co * per_co; Wt[r * 4 + co] = __ldg(w + i); }
        const float* sc = a.scale + (long long)g * a.ss_gs + a.bn_off[j];
        const float* sh = a.shift + (long long)g * a.ss_gs + a.bn_off[j];
        for (int i = tid; i < cin; i += 512) { SS[i] = __ldg(sc + i); SS[DB_MAXC + i] = __ldg(sh + i); }
      }
      if (j == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");      // the input tile has landed (weights were staged meanwhile)
      __syncthreads();
      // Work split: `nitems` groups of DB_R output rows, each shared by KS warps that take a slice of the input channels.  Per input
      // channel a warp issues DB_R + 2 shared loads, 2 (DB_R + 2) shuffles and 9 broadcast weight loads for 36 DB_R FMAs per lane: with
      // DB_R = 6 the FMA pipe (not the load/shuffle path) is the busier one (measured with DB_R = 3: FMA pipe 47 %, both co-limited).
      const int r_lo = j + 1, r_hi = rows - j - 1;        // local output rows [r_lo, r_hi)
      const int nitems = (r_hi - r_lo + DB_R - 1) / DB_R;
      const int KS = DB_WARPS / nitems > DB_KSMAX ? DB_KSMAX : DB_WARPS / nitems;
      const int item = warp / KS, kg = warp - item * KS;
      const bool on = item < nitems;
      float acc[DB_R][4];
#pragma unroll
      for (int r = 0; r < DB_R; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
      const int rb = r_lo + item * DB_R;
      if (on) {
        const int c_lo = kg * cin / KS, c_hi = (kg + 1) * cin / KS;
        int rix[DB_R + 2];
#pragma unroll
        for (int r = 0; r < DB_R + 2; ++r) { int q = rb - 1 + r; rix[r] = (q > rows - 1 ? rows - 1 : q) * DB_W + lane; }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
          const float s = SS[c], t = SS[DB_MAXC + c];
          const float* xc = X + (size_t)c * rows * DB_W;
          float v[DB_R + 2], vl[DB_R + 2], vr[DB_R + 2];
#pragma unroll
          for (int r = 0; r < DB_R + 2; ++r) v[r] = fmaxf(fmaf(xc[rix[r]], s, t), 0.f);
#pragma unroll
          for (int r = 0; r < DB_R + 2; ++r) { vl[r] = __shfl_sync(0xffffffffu, v[r], lm); vr[r] = __shfl_sync(0xffffffffu, v[r], lp); }
          const float4* wp = reinterpret_cast<const float4*>(Wt + (size_t)c * 36);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float4 w0 = wp[ky * 3], w1 = wp[ky * 3 + 1], w2 = wp[ky * 3 + 2];
#pragma unroll
            for (int r = 0; r < DB_R; ++r) {
              const float xl = vl[r + ky], xm = v[r + ky], xr = vr[r + ky];
              acc[r][0] = fmaf(w0.x, xl, fmaf(w1.x, xm, fmaf(w2.x, xr, acc[r][0])));
              acc[r][1] = fmaf(w0.y, xl, fmaf(w1.y, xm, fmaf(w2.y, xr, acc[r][1])));
              acc[r][2] = fmaf(w0.z, xl, fmaf(w1.z, xm, fmaf(w2.z, xr, acc[r][2])));
              acc[r][3] = fmaf(w0.w, xl, fmaf(w1.w, xm, fmaf(w2.w, xr, acc[r][3])));
            }
          }
        }
      }
      // deterministic two-round tree over the channel slices: odd slices -> even slices, then slices 2 and 4 -> slice 0
      float* slot1 = RED + ((size_t)item * 2 + (kg >> 1)) * (DB_R * 4 * DB_W) + lane;
      if (on && (kg & 1)) {
#pragma unroll
        for (int r = 0; r < DB_R; ++r)
#pragma unroll
          for (int co = 0; co < 4; ++co) slot1[(r * 4 + co) * DB_W] = acc[r][co];
      }
      __syncthreads();
      if (on && !(kg & 1) && kg + 1 < KS) {
#pragma unroll
        for (int r = 0; r < DB_R; ++r)
#pragma unroll
          for (int co = 0; co < 4; ++co) acc[r][co] += slot1[(r * 4 + co) * DB_W];
      }
      __syncthreads();
      float* slot2 = RED + ((size_t)item * 2 + (kg >> 2)) * (DB_R * 4 * DB_W) + lane;        // kg = 2 -> slot 0, kg = 4 -> slot 1
      if (on && (kg == 2 || kg == 4)) {
#pragma unroll
        for (int r = 0; r < DB_R; ++r)
#pragma unroll
          for (int co = 0; co < 4; ++co) slot2[(r * 4 + co) * DB_W] = acc[r][co];
      }
      __syncthreads();
      if (on && kg == 0) {
        const float* s2 = RED + ((size_t)item * 2) * (DB_R * 4 * DB_W) + lane;
#pragma unroll
        for (int r = 0; r < DB_R; ++r) {
          if (rb + r < r_hi) {
#pragma unroll
            for (int co = 0; co < 4; ++co) {
              float v = acc[r][co];
              if (KS > 2) v += s2[(r * 4 + co) * DB_W];
              if (KS > 4) v += s2[(DB_R * 4 + r * 4 + co) * DB_W];
              X[((size_t)(cin + co) * rows + rb + r) * DB_W + lane] = v;
            }
          }
        }
      }
      __syncthreads();
    }
    // ---- the owned rows of the new channels go back to the buffer
    {
      const int nrow = 4 * a.nl * DB_OWN;
      for (int q = warp * 4 + (lane >> 3); q < nrow; q += DB_WARPS * 4) {
        const int c = a.c0 + q / DB_OWN, r = q % DB_OWN;
        const float4 v = *reinterpret_cast<const float4*>(X + ((size_t)c * rows + a.nl + r) * DB_W + 4 * (lane & 7));
        *(reinterpret_cast<float4*>(bn_ + c * HW + (long long)(y0 + r) * DB_W) + (lane & 7)) = v;
      }
    }
    __syncthreads();
  }
}


// ------------------------------------------------------------------------------------------------------------------------------------
// Streaming formulation for the 32x32 map (the default 2D net): one CTA (8 warps x 4 rows) owns a WHOLE sample, so the circular wrap
// in y stays inside the CTA and nothing is recomputed.  The input channels are not kept resident: they stream through a double-buffered
// ring of 4-channel chunks (16 KB, cp.async), and while a raw channel is in flight its contribution to ALL layers is accumulated (each
// layer applies its own BN + ReLU to the raw value: 16 accumulators per pixel in registers, 64 per lane).  The outputs of layers 0..2
// are then written to shared memory (and HBM) and consumed as further chunks by the later layers: chunk 12 (layer 0) feeds layers 1-3,
// chunk 13 layers 2-3, chunk 14 layer 3.  The layer weights of a chunk (144 floats per channel) ride in a second ring, gathered with
// 4-byte cp.async from the reference layout.  87 KB of shared memory and 256 threads: two CTAs per SM overlap each other's barriers.
// Per channel and warp: 6 shared loads + 12 shuffles of the raw rows, then per layer 2 + 9 broadcast loads, 36 BN/ReLU ops, 144 FMAs.
constexpr int DS_H = 32, DS_CH = 4, DS_CHUNK = DS_CH * DS_H * DB_W, DS_WCH = DS_CH * DB_MAXL * 36, DS_THREADS = 256;

template <int L0>
__device__ __forceinline__ void ds_chunk(const float* __restrict__ data, const float* __restrict__ wst, const float* __restrict__ SS, int cbase,
                                         float2 (&acc)[DB_MAXL][4][2], int r0, int lane, int lm, int lp) {
#pragma unroll 1
  for (int e = 0; e < DS_CH; ++e) {
    const float* xe = data + e * DS_H * DB_W + lane;
    float x[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = xe[((r0 - 1 + i) & (DS_H - 1)) * DB_W];
#pragma unroll
    for (int l = L0; l < DB_MAXL; ++l) {
      const float s = SS[l * DB_MAXC + cbase + e], t = SS[(DB_MAXL + l) * DB_MAXC + cbase + e];
      // the layer's BN + ReLU on the lane's own column, the x-neighbours' ACTIVATED values by shuffle (12 ALU + 12 SHFL per layer and channel
      // instead of 36 ALU on raw neighbour values shuffled once per channel: the kernel is issue-bound)
      float a[6], al[6], ar[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] = fmaxf(fmaf(x[i], s, t), 0.f);
#pragma unroll
      for (int i = 0; i < 6; ++i) { al[i] = __shfl_sync(0xffffffffu, a[i], lm); ar[i] = __shfl_sync(0xffffffffu, a[i], lp); }
      const float4* wp = reinterpret_cast<const float4*>(wst + (e * DB_MAXL + l) * 36);
#pragma unroll
      // packed fp32 FMAs (fma.rn.f32x2, sm_100): one issue slot for the two output channels of a pair -- the scalar form was issue-bound
      // (80 % of the slots, FMA pipe 61 %); each lane of the pair is an IEEE fma, so the results are bit-identical
      float2 a2[6], al2[6], ar2[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) { a2[i] = make_float2(a[i], a[i]); al2[i] = make_float2(al[i], al[i]); ar2[i] = make_float2(ar[i], ar[i]); }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const float4 w0 = wp[ky * 3], w1 = wp[ky * 3 + 1], w2 = wp[ky * 3 + 2];
        const float2 w0a = make_float2(w0.x, w0.y), w0b = make_float2(w0.z, w0.w), w1a = make_float2(w1.x, w1.y), w1b = make_float2(w1.z, w1.w),
                     w2a = make_float2(w2.x, w2.y), w2b = make_float2(w2.z, w2.w);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          acc[l][r][0] = __ffma2_rn(w0a, al2[r + ky], __ffma2_rn(w1a, a2[r + ky], __ffma2_rn(w2a, ar2[r + ky], acc[l][r][0])));
          acc[l][r][1] = __ffma2_rn(w0b, al2[r + ky], __ffma2_rn(w1b, a2[r + ky], __ffma2_rn(w2b, ar2[r + ky], acc[l][r][1])));
        }
      }
    }
  }
}

__global__ void __launch_bounds__(DS_THREADS, 2) dense_block_stream_kernel(DenseBlockArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* BUF = sm;                                        // [2][4][32][32] raw-channel ring
  float* NEW = BUF + 2 * DS_CHUNK;                        // [3][4][32][32] outputs of layers 0..2 (chunks 12..14)
  float* WST = NEW + 3 * DS_CHUNK;                        // [2][4 e][4 l][9][4] weight ring
  float* SS = WST + 2 * DS_WCH;                           // [2][4 l][64] folded BatchNorm of the four layers
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int lm = (lane + 31) & 31, lp = (lane + 1) & 31;
  const int r0 = warp * 4;
  const long long HW = (long long)DS_H * DB_W;
  const int nraw = a.c0 / DS_CH;                          // raw chunks (c0 is a multiple of 4)
  pdl_wait();
  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    const int g = n / a.n_per_group;
    float* bn_ = a.buf + (long long)n * a.bs;
    auto prefetch = [&](int k) {                          // chunk k < nraw: data + weights; chunk >= nraw: weights only (data is in NEW)
      const int cbase = k * DS_CH;
      if (k < nraw) {
        float* dst = BUF + (k & 1) * DS_CHUNK;
        const float* src = bn_ + (long long)cbase * HW;
        for (int q = tid; q < DS_CHUNK / 4; q += DS_THREADS)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 4 * q)), "l"(src + 4 * q) : "memory");
      }
      float* wd = WST + (k & 1) * DS_WCH;
      for (int q = tid; q < DS_WCH; q += DS_THREADS) {
        const int co = q & 3, tap = (q >> 2) % 9, l = ((q >> 2) / 9) % DB_MAXL, e = (q >> 2) / (9 * DB_MAXL);
        const int cin = a.c0 + 4 * l, c = cbase + e;
        if (l < a.nl && c < cin) {
          const float* src = a.w[l] + (long long)g * a.w_gs[l] + ((long long)co * cin + c) * 9 + tap;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(wd + q)), "l"(src) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch(0);
    for (int i = tid; i < DB_MAXL * DB_MAXC; i += DS_THREADS) {
      const int l = i / DB_MAXC, c = i - l * DB_MAXC;
      const bool ok = l < a.nl && c < a.c0 + 4 * l;
      SS[i] = ok ? __ldg(a.scale + (long long)g * a.ss_gs + a.bn_off[l] + c) : 0.f;
      SS[DB_MAXL * DB_MAXC + i] = ok ? __ldg(a.shift + (long long)g * a.ss_gs + a.bn_off[l] + c) : 0.f;
    }
    float2 acc[DB_MAXL][4][2];                             // [layer][row][output-channel pair]
#pragma unroll
    for (int l = 0; l < DB_MAXL; ++l)
#pragma unroll
      for (int r = 0; r < 4; ++r) { acc[l][r][0] = make_float2(0.f, 0.f); acc[l][r][1] = make_float2(0.f, 0.f); }
    const int nchunks = nraw + a.nl - 1;                  // raw chunks, then the outputs of layers 0 .. nl-2
    // finalize layer l: its accumulators are complete -> shared memory (input of the later layers) and the dense-block buffer
    auto finalize = [&](int l, const float2 (&v2)[4][2]) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int co = 0; co < 4; ++co) {
          const float v = (co & 1) ? v2[r][co >> 1].y : v2[r][co >> 1].x;
          if (l < DB_MAXL - 1) NEW[((size_t)(l * DS_CH + co) * DS_H + r0 + r) * DB_W + lane] = v;
          bn_[(long long)(a.c0 + 4 * l + co) * HW + (r0 + r) * DB_W + lane] = v;
        }
    };
    for (int k = 0; k < nchunks; ++k) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();                                    // chunk k (data + weights) visible; everyone is done with chunk k - 1
      if (k + 1 < nchunks) prefetch(k + 1);
      const float* wst = WST + (k & 1) * DS_WCH;
      if (k < nraw) {
        ds_chunk<0>(BUF + (k & 1) * DS_CHUNK, wst, SS, k * DS_CH, acc, r0, lane, lm, lp);
        if (k == nraw - 1) finalize(0, acc[0]);
      } else if (k == nraw) {
        ds_chunk<1>(NEW, wst, SS, k * DS_CH, acc, r0, lane, lm, lp);
        finalize(1, acc[1]);
      } else if (k == nraw + 1) {
        ds_chunk<2>(NEW + DS_CHUNK, wst, SS, k * DS_CH, acc, r0, lane, lm, lp);
        finalize(2, acc[2]);
      } else {
        ds_chunk<3>(NEW + 2 * DS_CHUNK, wst, SS, k * DS_CH, acc, r0, lane, lm, lp);
      }
    }
    finalize(DB_MAXL - 1, acc[DB_MAXL - 1]);               // (the launcher only takes this kernel for four-layer blocks)
    __syncthreads();                                      // NEW / rings are reused by the next sample
  }
}

}  // namespace

// number of consecutive ops starting at `first` that form a fusable dense block (0 = none)
int arpde_dense_block_match(const arpde_op_t* ops, int n_ops, int first, const arpde_buf_t* bufs, int N) {
  const arpde_op_t& o0 = ops[first];
  if (o0.in_buf < 0 || o0.out_buf != o0.in_buf) return 0;
  const arpde_buf_t& b = bufs[o0.in_buf];
  if (b.w != DB_W || b.h < DB_OWN || b.h % DB_OWN) return 0;
  int nl = 0;
  for (int i = first; i < n_ops && nl < DB_MAXL; ++i, ++nl) {
    const arpde_op_t& o = ops[i];
    if (o.in_buf != o0.in_buf || o.out_buf != o0.in_buf || o.cout != 4 || o.kh != 3 || o.kw != 3 || o.stride != 1 || o.up || o.bn_param < 0 || o.act ||
        o.cin != o0.cin + 4 * nl || o.out_coff != o.cin) break;
  }
  if (nl < 2 || (o0.cin & 1) || o0.cin + 4 * nl > DB_MAXC || o0.cin + 4 * nl > b.ctot) return 0;
  // one CTA per (sample, 16-row tile): worth it once the grid fills the machine (small batches stay on the per-layer kernels)
  if ((long long)N * (b.h / DB_OWN) < arpde_num_sms()) return 0;
  if (arpde_dev_env("ARPDE_DENSE_FUSED_OFF")) return 0;
  return nl;
}

int arpde_dense_block_launch(const arpde_op_t* ops, int first, int nl, const arpde_buf_t* bufs, void* const* params, const int64_t* gstride, float* work,
                             const float* scale, const float* shift, long long ss_gs, int N, int n_per_group, cudaStream_t stream, bool pdl) {
  const arpde_buf_t& b = bufs[ops[first].in_buf];
  DenseBlockArgs a; memset(&a, 0, sizeof(a));
  a.buf = work + b.offset; a.bs = (long long)b.ctot * b.h * b.w;
  a.c0 = ops[first].cin; a.nl = nl; a.H = b.h; a.N = N; a.n_per_group = n_per_group;
  for (int j = 0; j < nl; ++j) {
    a.w[j] = (const float*)params[ops[first + j].w_param]; a.w_gs[j] = gstride[ops[first + j].w_param];
    a.bn_off[j] = ops[first + j].bn_off;
  }
  a.scale = scale; a.shift = shift; a.ss_gs = ss_gs;
  ARPDE_CHECK_ARG((((uintptr_t)a.buf) & 15) == 0 && (a.bs & 3) == 0, "dense_block: buffer must be 16-byte aligned");
  if (b.h == DS_H && nl == DB_MAXL && (a.c0 % DS_CH) == 0 && !arpde_dev_env("ARPDE_DENSE_STREAM_OFF")) {
    // whole-sample streaming kernel (no halo recompute, two CTAs per SM)
    const size_t smem_s = ((size_t)5 * DS_CHUNK + 2 * DS_WCH + 2 * DB_MAXL * DB_MAXC) * sizeof(float);
    ARPDE_CUDA(cudaFuncSetAttribute(dense_block_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    long long grid_s = 2ll * arpde_num_sms(); if (grid_s > N) grid_s = N;
    ARPDE_CUDA(arpde_launch1(dense_block_stream_kernel, (unsigned)grid_s, DS_THREADS, smem_s, stream, pdl, a));
    ARPDE_LAUNCH_CHECK();
    return ARPDE_OK;
  }
  const int rows = DB_OWN + 2 * nl, ctot = a.c0 + 4 * nl;
  const size_t smem = ((size_t)ctot * rows * DB_W + (size_t)(ctot - 4) * 36 + 2 * DB_MAXC + DB_RED) * sizeof(float);
  ARPDE_CHECK_ARG(smem <= 227 * 1024, "dense_block: %zu bytes of shared memory needed", smem);
  ARPDE_CUDA(cudaFuncSetAttribute(dense_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)N * (b.h / DB_OWN);
  long long grid = arpde_num_sms(); if (grid > items) grid = items;
  dense_block_kernel<<<(unsigned)grid, 512, smem, stream>>>(a);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
