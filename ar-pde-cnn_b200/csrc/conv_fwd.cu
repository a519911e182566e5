// conv_fwd.cu -- circular (periodic) convolution, forward, fp32 FFMA register-tiled direct form.
//
// Replaces, per call (reference 2D-Burgers-SWAG/nn/denseEDcirc2d.py, 1D-*/nn/denseEDcirc.py):
//   F.pad(mode='circular') + aten::convolution of every nn.Conv{1,2}d(padding_mode='circular'),
//   the BatchNorm+ReLU that precedes it (_DenseLayer.norm1/relu1 :58-62, _Transition :96-97,:106-107,
//   last_decoding :169-187, In_conv_norm1/relu1 :250-251), UpsamplingNearest{1,2}d (:15-21),
//   torch.cat of the dense layer (:66-68, written in place at a channel offset) and, in train()
//   mode, the batch-statistics reduction of the BatchNorm that consumes the output.
//
// Semantics: torch<=1.2 circular padding == symmetric wrap pad p=(k-1)/2 (k=4: p=1):
//   y[n,co,oy,ox] = sum_{ci,ky,kx} w[g(n),co,ci,ky,kx] * a[n,ci,(oy*s+ky-p) mod Hc,(ox*s+kx-p) mod Wc]
//   a = relu(scale[ci]*x + shift[ci]) (optional), nearest-upsampled x2 when up=1 (Hc = 2H).
//
// Mapping: one CTA = one sample x one tile of TH x TW output pixels x TCO output channels; threads
// (32*NPW px-threads) x (NCG channel groups); each thread keeps COT x PXT accumulators; the input
// patch (BN+ReLU applied while staging) and the weight slab of a CI chunk live in shared memory.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"

template <int KH, int KW, int S, int COT, int PXT>
__global__ void __launch_bounds__(256) conv_fwd_kernel(ConvFwdArgs a) {
  constexpr int KK = KH * KW;
  constexpr int SEG = ((PXT - 1) * S + KW + 3) & ~3;      // input floats a thread reads per (ci,ky), as float4s
  extern __shared__ __align__(16) float smem[];
  // 1D signals (rows_are_samples): a tile row is a sample, rows are independent (KH == 1) and never strided
  const int SV = a.rows_are_samples ? 1 : S;
  const int PH = (a.TH - 1) * SV + KH, PW = (a.TW - 1) * S + KW;
  const int PITCH = a.pitch;
  const int TCO = blockDim.y * COT;
  int* rowmap = reinterpret_cast<int*>(smem);
  int* colmap = rowmap + PH;
  float* patch = smem + ((PH + PW + 3) & ~3);
  float* wsm = patch + a.ci_chunk * PH * PITCH;

  // ---- tile decode
  int t = blockIdx.x;
  const int tx_tile = t % a.tiles_x; t /= a.tiles_x;
  const int ty_tile = t % a.tiles_y; t /= a.tiles_y;
  const int co_tile = t % a.co_tiles; t /= a.co_tiles;
  const int n = a.rows_are_samples ? ty_tile * a.TH : t;          // (first) sample of the tile
  const int g = n / a.n_per_group;
  const int oy0 = a.rows_are_samples ? 0 : ty_tile * a.TH, ox0 = tx_tile * a.TW, co0 = co_tile * TCO;
  const int Hc = a.H << a.up_h, Wc = a.W << a.up;
  const int pad_y = a.pad_y, pad_x = a.pad_x;

  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthreads = blockDim.x * blockDim.y;
  for (int i = tid; i < PH; i += nthreads) {
    if (a.rows_are_samples) { rowmap[i] = n + i < a.N ? i : -1; continue; }      // row i = sample n + i
    int iy = (oy0 * S - pad_y + i) % Hc; if (iy < 0) iy += Hc;
    rowmap[i] = (a.dilate && a.up_h && (iy & 1)) ? -1 : (iy >> a.up_h);     // zero-inserted rows read as 0
  }
  for (int i = tid; i < PW; i += nthreads) {
    int ix = (ox0 * S - pad_x + i) % Wc; if (ix < 0) ix += Wc;
    colmap[i] = (a.dilate && (ix & 1)) ? -1 : (ix >> a.up);
  }

  const int nxt = a.TW / PXT;                              // px-threads per tile row
  const int prow = threadIdx.x / nxt, pcol = (threadIdx.x - prow * nxt) * PXT;
  const bool px_valid = prow < a.TH;
  const int tc = threadIdx.y;

  float acc[COT][PXT];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int p = 0; p < PXT; ++p) acc[c][p] = 0.f;

  const float* xn = a.x + (long long)n * a.x_bs;
  const float* wg = a.w + (long long)g * a.w_gs;
  const float* scale = a.scale ? a.scale + (long long)g * a.ss_gs : nullptr;
  const float* shift = a.scale ? a.shift + (long long)g * a.ss_gs : nullptr;
  const long long HW = (long long)a.H * a.W;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;

  for (int ci0 = 0; ci0 < a.cin; ci0 += a.ci_chunk) {
    const int cic = min(a.ci_chunk, a.cin - ci0);
    __syncthreads();
    // ---- stage the input patch: one warp per pair of (ci,row) lines; BN scale/shift + ReLU applied here.  All global loads of a
    //      batch (2 lines x 4 x 32 columns) are issued before the first one is used: ncu showed 42% of this kernel's stall samples
    //      on the FFMA that consumed a just-issued load when every element was loaded and converted one at a time.
    {
      constexpr int RU = 2, PU = 4;
      const int nlines = cic * PH;
      for (int r0 = warp * RU; r0 < nlines; r0 += nwarps * RU) {
        for (int px0 = 0; px0 < PITCH; px0 += 32 * PU) {
          float v[RU][PU];
          float sc[RU], sh[RU];
          unsigned inside = 0;                              // bit u*PU+p: the element comes from the input (else it is a zero)
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int r = r0 + u;
            const bool rok = r < nlines;
            const int ci = rok ? r / PH : 0, py = rok ? r - ci * PH : 0;
            const int rm = rok ? rowmap[py] : -1;
            const float* src = a.rows_are_samples ? xn + (long long)(rm < 0 ? 0 : rm) * a.x_bs + (ci0 + ci) * HW
                                                  : xn + (ci0 + ci) * HW + (long long)(rm < 0 ? 0 : rm) * a.W;
            sc[u] = 1.f; sh[u] = 0.f;
            if (scale && rok) { sc[u] = scale[ci0 + ci]; sh[u] = shift[ci0 + ci]; }
#pragma unroll
            for (int p = 0; p < PU; ++p) {
              const int px = px0 + p * 32 + lane;
              const int cm = (px < PW) ? colmap[px] : -1;
              const bool in = rm >= 0 && cm >= 0;
              v[u][p] = in ? __ldg(src + cm) : 0.f;
              if (in) inside |= 1u << (u * PU + p);
            }
          }
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int r = r0 + u;
            if (r < nlines) {
              float* dst = patch + r * PITCH;
#pragma unroll
              for (int p = 0; p < PU; ++p) {
                const int px = px0 + p * 32 + lane;
                if (px < PITCH) {
                  float t = 0.f;                            // outside the patch / zero-inserted position
                  if ((inside >> (u * PU + p)) & 1u) { t = fmaf(v[u][p], sc[u], sh[u]); if (a.relu_in) t = fmaxf(t, 0.f); }
                  dst[px] = t;
                }
              }
            }
          }
        }
      }
    }
    // ---- stage weights: global [co][ci][kk] -> smem [ci][kk][co] (zero for co >= cout)
    const int per_co = cic * KK;
    for (int i = tid; i < TCO * per_co; i += nthreads) {
      const int co = i / per_co, j = i - co * per_co;
      float v = 0.f;
      if (co0 + co < a.cout) v = __ldg(wg + ((long long)(co0 + co) * a.cin + ci0) * KK + j);
      wsm[j * TCO + co] = v;
    }
    __syncthreads();
    if (px_valid) {
      for (int ci = 0; ci < cic; ++ci) {
#pragma unroll
        for (int ky = 0; ky < KH; ++ky) {
          const float* rowp = patch + (ci * PH + prow * SV + ky) * PITCH + pcol * S;
          float in[SEG];
#pragma unroll
          for (int q = 0; q < SEG / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(rowp + 4 * q);
            in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
          }
          const float* wp = wsm + (ci * KK + ky * KW) * TCO + tc * COT;
#pragma unroll
          for (int kx = 0; kx < KW; ++kx) {
            float w[COT];
            if (COT % 4 == 0) {
#pragma unroll
              for (int q = 0; q < COT / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4*>(wp + kx * TCO + 4 * q);
                w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
              }
            } else {
#pragma unroll
              for (int c = 0; c < COT; ++c) w[c] = wp[kx * TCO + c];
            }
#pragma unroll
            for (int c = 0; c < COT; ++c)
#pragma unroll
              for (int p = 0; p < PXT; ++p) acc[c][p] = fmaf(w[c], in[p * S + kx], acc[c][p]);
          }
        }
      }
    }
  }

  // ---- epilogue: activation, masked store at the channel offset, optional batch-stat sums
  const int oy = a.rows_are_samples ? 0 : oy0 + prow;
  const int ns = a.rows_are_samples ? n + prow : n;           // sample of this thread's output row
  const bool row_ok = px_valid && oy < a.Ho && ns < a.N;
  const long long HoWo = (long long)a.Ho * a.Wo;
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    const int co = co0 + tc * COT + c;
    float s1 = 0.f, s2 = 0.f;
    if (row_ok && co < a.cout) {
      float* yp = a.y + ((long long)ns * a.y_ctot + a.y_coff + co) * HoWo + (long long)oy * a.Wo + ox0 + pcol;
      float v[PXT];
#pragma unroll
      for (int p = 0; p < PXT; ++p) v[p] = apply_act(acc[c][p], a.act);
      if (PXT % 4 == 0 && (a.Wo & 3) == 0 && ox0 + pcol + PXT <= a.Wo) {
#pragma unroll
        for (int q = 0; q < PXT / 4; ++q)
          *reinterpret_cast<float4*>(yp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
#pragma unroll
        for (int p = 0; p < PXT; ++p) { s1 += v[p]; s2 = fmaf(v[p], v[p], s2); }
      } else {
#pragma unroll
        for (int p = 0; p < PXT; ++p)
          if (ox0 + pcol + p < a.Wo) { yp[p] = v[p]; s1 += v[p]; s2 = fmaf(v[p], v[p], s2); }
      }
    }
    if (a.stats) {   // uniform branch: every lane of the warp shares tc (blockDim.x multiple of 32)
      s1 = warp_sum(s1); s2 = warp_sum(s2);
      if ((threadIdx.x & 31) == 0 && co < a.cout) {
        double* st = a.stats + 2 * (a.y_coff + co);
        atomicAdd(st, (double)s1); atomicAdd(st + 1, (double)s2);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row-tiled variant for the narrow output layers (cout <= 4, stride 1, 2D, no upsampling; e.g. LastTransUp's 5x5 16 -> 2).
// ncu on the generic kernel for that layer (N = 1920): FMA pipe 23 % active, 52 % issue-active -- a thread owned 2 x 8 outputs,
// re-read its weights from shared memory for every (ci, ky) and the 8-row tiles staged a 50 % halo.  Here a thread owns
// COT x RY x 8 outputs (RY rows), keeps the KH*KW*COT weights of the current input channel in registers and walks the
// KH + RY - 1 patch rows once: 43 shared loads per 800 FMAs (5x5, COT = 2, RY = 2) instead of 65 per 400; tiles are up to 64 rows.
template <int KH, int KW, int COT, int RY, int PXT>
__global__ void __launch_bounds__(256) conv_fwd_rt_kernel(ConvFwdArgs a) {
  constexpr int KK = KH * KW;
  constexpr int SEG = (PXT + KW - 1 + 3) & ~3;
  extern __shared__ __align__(16) float smem[];
  const int PH = a.TH + KH - 1, PW = a.TW + KW - 1;
  const int PITCH = a.pitch;
  int* rowmap = reinterpret_cast<int*>(smem);
  int* colmap = rowmap + PH;
  float* patch = smem + ((PH + PW + 3) & ~3);
  float* wsm = patch + a.ci_chunk * PH * PITCH;            // [ci][kk][COT]

  int t = blockIdx.x;
  const int tx_tile = t % a.tiles_x; t /= a.tiles_x;
  const int ty_tile = t % a.tiles_y; t /= a.tiles_y;
  const int n = t;
  const int g = n / a.n_per_group;
  const int oy0 = ty_tile * a.TH, ox0 = tx_tile * a.TW;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  for (int i = tid; i < PH; i += nthreads) { int iy = (oy0 - a.pad_y + i) % a.H; if (iy < 0) iy += a.H; rowmap[i] = iy; }
  for (int i = tid; i < PW; i += nthreads) { int ix = (ox0 - a.pad_x + i) % a.W; if (ix < 0) ix += a.W; colmap[i] = ix; }

  const int nxt = a.TW / PXT;
  const int trow = tid / nxt, pcol = (tid - trow * nxt) * PXT;
  const int prow = trow * RY;
  const bool px_valid = prow < a.TH;

  float acc[COT][RY][PXT];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int r = 0; r < RY; ++r)
#pragma unroll
      for (int p = 0; p < PXT; ++p) acc[c][r][p] = 0.f;

  const float* xn = a.x + (long long)n * a.x_bs;
  const float* wg = a.w + (long long)g * a.w_gs;
  const float* scale = a.scale ? a.scale + (long long)g * a.ss_gs : nullptr;
  const float* shift = a.scale ? a.shift + (long long)g * a.ss_gs : nullptr;
  const long long HW = (long long)a.H * a.W;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;

  for (int ci0 = 0; ci0 < a.cin; ci0 += a.ci_chunk) {
    const int cic = min(a.ci_chunk, a.cin - ci0);
    __syncthreads();
    // stage the patch: the tile spans the full width (TW == W, checked by the launcher), see stage_full_rows
    // 8 row groups in flight per lane: with 4 warps per CTA the staging phase is otherwise a chain of exposed global-load latencies
    // (0.44 -> 0.39 ms on a dense layer; a cp.async double buffer with the prologue applied at compute time and an aligned patch
    //  layout with float4 stores were both tried and were no faster: 0.40 / 0.44 ms)
    stage_full_rows<8>(patch, PITCH, xn, (int)HW, ci0, cic * PH, PH, rowmap, a.W, a.pad_x, PW - a.W, scale, shift, a.relu_in, tid, nthreads);
    for (int i = tid; i < cic * KK * COT; i += nthreads) {          // weights: global [co][ci][kk] -> smem [ci][kk][COT]
      const int c = i % COT, j = i / COT;                          // j = ci * KK + kk
      wsm[i] = c < a.cout ? __ldg(wg + ((long long)c * a.cin + ci0) * KK + j) : 0.f;
    }
    __syncthreads();
    if (px_valid) {
      for (int ci = 0; ci < cic; ++ci) {
        float w[KK][COT];
        const float* wp = wsm + ci * KK * COT;
#pragma unroll
        for (int k = 0; k < KK; ++k)
#pragma unroll
          for (int c = 0; c < COT; ++c) w[k][c] = wp[k * COT + c];
        const float* base = patch + (ci * PH + prow) * PITCH + pcol;
#pragma unroll
        for (int pr = 0; pr < KH + RY - 1; ++pr) {
          float in[SEG];
#pragma unroll
          for (int q = 0; q < SEG / 4; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(base + pr * PITCH + 4 * q);
            in[4 * q] = v4.x; in[4 * q + 1] = v4.y; in[4 * q + 2] = v4.z; in[4 * q + 3] = v4.w;
          }
#pragma unroll
          for (int r = 0; r < RY; ++r) {
            const int ky = pr - r;
            if (ky >= 0 && ky < KH) {
#pragma unroll
              for (int kx = 0; kx < KW; ++kx)
#pragma unroll
                for (int c = 0; c < COT; ++c)
#pragma unroll
                  for (int p = 0; p < PXT; ++p) acc[c][r][p] = fmaf(w[ky * KW + kx][c], in[p + kx], acc[c][r][p]);
            }
          }
        }
      }
    }
  }

  const long long HoWo = (long long)a.Ho * a.Wo;
#pragma unroll
  for (int c = 0; c < COT; ++c) {
    if (px_valid && c < a.cout) {
#pragma unroll
      for (int r = 0; r < RY; ++r) {
        const int oy = oy0 + prow + r;
        if (oy < a.Ho && prow + r < a.TH) {
          float* yp = a.y + ((long long)n * a.y_ctot + a.y_coff + c) * HoWo + (long long)oy * a.Wo + ox0 + pcol;
          float v[PXT];
#pragma unroll
          for (int p = 0; p < PXT; ++p) v[p] = apply_act(acc[c][r][p], a.act);
#pragma unroll
          for (int q = 0; q < PXT / 4; ++q) *reinterpret_cast<float4*>(yp + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
typedef void (*conv_fwd_fn)(ConvFwdArgs);

template <int KH, int KW, int S>
static conv_fwd_fn pick_cot(int cot, int* pxt) {
  constexpr int P = (S == 1) ? 8 : 4;
  *pxt = P;
  switch (cot) {
    case 8: return conv_fwd_kernel<KH, KW, S, 8, P>;
    case 4: return conv_fwd_kernel<KH, KW, S, 4, P>;
    default: return conv_fwd_kernel<KH, KW, S, 2, P>;
  }
}

static conv_fwd_fn pick_kernel(int kh, int kw, int s, int cot, int* pxt) {
#define PK(KH_, KW_, S_) if (kh == KH_ && kw == KW_ && s == S_) return pick_cot<KH_, KW_, S_>(cot, pxt)
  PK(1, 1, 1); PK(3, 3, 1); PK(3, 3, 2); PK(5, 5, 1); PK(4, 4, 2); PK(7, 7, 2);
  PK(1, 3, 1); PK(1, 3, 2); PK(1, 5, 1); PK(1, 4, 2); PK(1, 7, 2); PK(1, 7, 1); PK(7, 7, 1);
  PK(4, 4, 1); PK(1, 4, 1);      // data gradient of the k4/s2 TransDown convolution: flipped 4-tap kernel, stride 1, zero-inserted input
#undef PK
  return nullptr;
}

// shape part of the row-tiled kernel's eligibility (pointer alignment is checked at launch): the narrow layers it is faster on than
// the tensor-core kernel (measured at N = 1920: dense 52->4 3x3 at 32x32 0.44 vs 0.52 ms, 16->2 5x5 at 64x64 0.56 vs 1.1 ms)
bool arpde_conv_fwd_rt_shape_ok(const ConvDesc& d) {
  if (arpde_dev_env("ARPDE_CONV_RT_OFF")) return false;
  if (d.stride != 1 || d.up || d.dilate || d.pad_override >= 0 || d.H <= 1 || d.cout > 4) return false;
  if (!((d.kh == 3 && d.kw == 3) || (d.kh == 5 && d.kw == 5))) return false;
  if ((d.W & 7) || d.W > 64 || (d.W & 3)) return false;
  const int pxt = d.W >= 64 ? 8 : 4, nxt = d.W / pxt;
  int TH = (256 / nxt) * 2; if (TH > d.H) TH = (d.H + 1) / 2 * 2;
  const int threads = ((nxt * ((TH + 1) / 2) + 31) / 32) * 32;
  if (threads < 128 || threads > 256) return false;
  // one CTA per (sample, tile): small batches leave SMs idle, there the persistent tensor-core kernel is the better choice
  // (training mini-batch of 128: forward of the four dense layers 0.16 ms slower on this kernel)
  return (long long)d.N * ((d.H + TH - 1) / TH) >= 2ll * arpde_num_sms();
}

int arpde_conv_fwd_launch(const ConvDesc& d, const float* x, long long x_bs, const float* w, long long w_gs,
                          const float* scale, const float* shift, long long ss_gs, float* y, double* stats,
                          cudaStream_t stream) {
  ARPDE_CHECK_ARG(x && w && y, "conv_fwd: null pointer");
  ARPDE_CHECK_ARG(d.N >= 0 && d.cin > 0 && d.cout > 0 && d.H > 0 && d.W > 0, "conv_fwd: bad dims");
  ARPDE_CHECK_ARG(d.stride == 1 || d.stride == 2, "conv_fwd: stride %d unsupported", d.stride);
  ARPDE_CHECK_ARG(d.n_per_group > 0, "conv_fwd: n_per_group must be > 0");
  if (d.N == 0) return ARPDE_OK;
  ARPDE_CHECK_ARG(!(d.dilate && d.up), "conv_fwd: dilate and up are exclusive");
  const int up_w = (d.up || d.dilate) ? 1 : 0;
  const int up_h = (d.H == 1 && d.kh == 1) ? 0 : up_w;      // 1D signals: H stays 1
  const int Hc = d.H << up_h, Wc = d.W << up_w;
  const int sh = (d.H == 1 && d.kh == 1) ? 1 : d.stride;
  ARPDE_CHECK_ARG(Hc % sh == 0 && Wc % d.stride == 0, "conv_fwd: spatial size %dx%d not divisible by stride %d", Hc, Wc, d.stride);
  const int Ho = Hc / sh, Wo = Wc / d.stride;
  // the 2-channel 5x5 output convolution on 64-column maps: lane-per-two-columns kernel (conv_out.cu)
  if (arpde_conv_out_shape_ok(d, x, x_bs, y, stats) && (scale == nullptr) == (shift == nullptr))
    return arpde_conv_out_launch(d, x, x_bs, w, w_gs, scale, shift, ss_gs, y, stream);
  // narrow output layers: row-tiled kernel (see conv_fwd_rt_kernel)
  if (!arpde_dev_env("ARPDE_CONV_RT_OFF") && d.stride == 1 && !d.up && !d.dilate && d.pad_override < 0 && d.H > 1 && d.cout <= 4 && stats == nullptr &&
      (Wo & 7) == 0 && (((uintptr_t)y) & 15) == 0) {
    conv_fwd_fn rt = nullptr;
    const int cot = d.cout <= 2 ? 2 : 4;
    // 8 columns per thread on wide rows, 4 on narrow ones (a 32x32 image then still fills 128 threads)
    const int PXT = Wo >= 64 ? 8 : 4;
    if (d.kh == 5 && d.kw == 5) rt = PXT == 8 ? (cot == 2 ? conv_fwd_rt_kernel<5, 5, 2, 2, 8> : conv_fwd_rt_kernel<5, 5, 4, 2, 8>)
                                              : (cot == 2 ? conv_fwd_rt_kernel<5, 5, 2, 2, 4> : conv_fwd_rt_kernel<5, 5, 4, 2, 4>);
    else if (d.kh == 3 && d.kw == 3) rt = PXT == 8 ? (cot == 2 ? conv_fwd_rt_kernel<3, 3, 2, 2, 8> : conv_fwd_rt_kernel<3, 3, 4, 2, 8>)
                                                   : (cot == 2 ? conv_fwd_rt_kernel<3, 3, 2, 2, 4> : conv_fwd_rt_kernel<3, 3, 4, 2, 4>);
    if (rt) {
      constexpr int RY = 2;
      ConvFwdArgs a;
      memset(&a, 0, sizeof(a));
      a.x = x; a.x_bs = x_bs; a.cin = d.cin; a.H = d.H; a.W = d.W;
      a.pad_y = (d.kh - 1) / 2; a.pad_x = (d.kw - 1) / 2;
      a.w = w; a.w_gs = w_gs; a.scale = scale; a.shift = shift; a.ss_gs = ss_gs; a.relu_in = d.relu_in;
      a.y = y; a.y_ctot = d.y_ctot; a.y_coff = d.y_coff; a.cout = d.cout; a.Ho = Ho; a.Wo = Wo; a.act = d.act;
      a.n_per_group = d.n_per_group; a.N = d.N;
      int TW = Wo < 64 ? Wo : 64;
      const int nxt = TW / PXT;
      int TH = (256 / nxt) * RY; if (TH > Ho) TH = (Ho + RY - 1) / RY * RY;
      a.TH = TH; a.TW = TW; a.tiles_x = (Wo + TW - 1) / TW; a.tiles_y = (Ho + TH - 1) / TH; a.co_tiles = 1;
      const int PH = TH + d.kh - 1, PW = TW + d.kw - 1;
      a.pitch = (TW + ((d.kw - 1 + 3) & ~3) + 3) & ~3;
      const int KK = d.kh * d.kw;
      const size_t fixed = (size_t)(((PH + PW + 3) & ~3)) * 4;
      const size_t per_ci = (size_t)PH * a.pitch * 4 + (size_t)KK * cot * 4;
      const size_t budget = (size_t)(arpde_dev_env("ARPDE_RT_SMEM_KB") ? atoi(arpde_dev_env("ARPDE_RT_SMEM_KB")) : (Wo >= 64 ? 100 : 56)) * 1024;   // 2 CTAs per SM on 64-wide rows, 4 of the smaller ones
      int cic = (int)((budget - fixed) / per_ci); if (cic < 1) cic = 1; if (cic > d.cin) cic = d.cin;
      cic = (d.cin + ((d.cin + cic - 1) / cic) - 1) / ((d.cin + cic - 1) / cic);
      a.ci_chunk = cic;
      const size_t smem = fixed + per_ci * cic;
      const int threads = ((nxt * ((TH + RY - 1) / RY) + 31) / 32) * 32;
      if (smem <= 200 * 1024 && threads <= 256 && threads >= 128 && TW == Wo && (d.W & 3) == 0 && (x_bs & 3) == 0 && (((uintptr_t)x) & 15) == 0) {
        if (smem > 48 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(rt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const long long blocks = (long long)d.N * a.tiles_y * a.tiles_x;
        ARPDE_CHECK_ARG(blocks < (1ll << 31), "conv_fwd: grid too large");
        rt<<<(unsigned)blocks, threads, smem, stream>>>(a);
        ARPDE_LAUNCH_CHECK();
        return ARPDE_OK;
      }
    }
  }
  const int cot = d.cout > 4 ? 8 : (d.cout > 2 ? 4 : 2);
  int pxt = 8;
  conv_fwd_fn fn = pick_kernel(d.kh, d.kw, d.stride, cot, &pxt);
  if (!fn) { arpde_set_error("conv_fwd: kernel %dx%d stride %d not instantiated", d.kh, d.kw, d.stride); return ARPDE_ERR_UNSUPPORTED; }

  ConvFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.x_bs = x_bs; a.cin = d.cin; a.H = d.H; a.W = d.W; a.up = up_w; a.up_h = up_h; a.dilate = d.dilate;
  a.pad_y = d.pad_override >= 0 ? (d.kh == 1 ? 0 : d.pad_override) : (d.kh == 4 ? 1 : (d.kh - 1) / 2);
  a.pad_x = d.pad_override >= 0 ? d.pad_override : (d.kw == 4 ? 1 : (d.kw - 1) / 2);
  a.w = w; a.w_gs = w_gs; a.scale = scale; a.shift = shift; a.ss_gs = ss_gs; a.relu_in = d.relu_in;
  a.y = y; a.y_ctot = d.y_ctot; a.y_coff = d.y_coff; a.cout = d.cout; a.Ho = Ho; a.Wo = Wo; a.act = d.act;
  a.stats = stats; a.n_per_group = d.n_per_group; a.N = d.N;

  // thread shape: NCG channel groups x 32*NPW px-threads, <= 256 threads
  int ncg = (d.cout + cot - 1) / cot; if (ncg > 6) ncg = 6;
  const int npw_max = 8 / ncg > 0 ? 8 / ncg : 1;
  // tile: TW covers the row in multiples of PXT; 1D signals may use every px-thread along the row
  int TW = ((Wo + pxt - 1) / pxt) * pxt;
  const int tw_cap = (Ho == 1 ? 32 * npw_max : 16) * pxt;
  if (TW > tw_cap) TW = tw_cap;
  const int nxt = TW / pxt;
  int TH = (32 * npw_max) / nxt; if (TH < 1) TH = 1;
  // 1D signals: fill the CTA with consecutive samples as tile rows (they must share a weight group).  A 96-point KS layer used 6 or 12
  // of its 32+ px-threads with one sample per CTA (C1 rollout: 433 us per AR step of 1920 samples)
  const bool one_d_rows = Ho == 1 && d.kh == 1 && !arpde_dev_env("ARPDE_CONV_1D_ROWS_OFF");
  if (one_d_rows) {
    // ... but keep at least ~3 CTAs per SM in the grid (1920 samples in tiles of 32 would be 60 CTAs on 148 SMs)
    const long long per_row_ctas = (long long)((d.cout + ncg * cot - 1) / (ncg * cot)) * ((Wo + TW - 1) / TW);
    while (TH > 1 && (d.n_per_group % TH != 0 || TH > d.N || ((d.N + TH - 1) / TH) * per_row_ctas < 3ll * arpde_num_sms())) --TH;
  }
  else if (TH > Ho) TH = Ho;
  const int npw = (nxt * TH + 31) / 32;
  ARPDE_CHECK_ARG(npw >= 1 && npw <= npw_max, "conv_fwd: internal tiling error (npw=%d)", npw);
  a.TH = TH; a.TW = TW; a.rows_are_samples = one_d_rows ? 1 : 0;
  a.tiles_x = (Wo + TW - 1) / TW; a.tiles_y = one_d_rows ? (d.N + TH - 1) / TH : (Ho + TH - 1) / TH;
  const int TCO = ncg * cot;
  a.co_tiles = (d.cout + TCO - 1) / TCO;
  const int PH = one_d_rows ? TH : (TH - 1) * d.stride + d.kh, PW = (TW - 1) * d.stride + d.kw;
  const int seg = (((pxt - 1) * d.stride + d.kw) + 3) & ~3;
  // a thread reads SEG floats from column pcol*S; the last thread starts at (TW-PXT)*S
  int pitch = (TW - pxt) * d.stride + seg; if (pitch < PW) pitch = PW; pitch = (pitch + 3) & ~3;
  a.pitch = pitch;
  const int KK = d.kh * d.kw;
  const size_t fixed = (size_t)(((PH + PW + 3) & ~3)) * 4;
  const size_t per_ci = (size_t)PH * pitch * 4 + (size_t)KK * TCO * 4;
  const size_t budget = 72 * 1024;
  int cic = (int)((budget - fixed) / per_ci); if (cic < 1) cic = 1; if (cic > d.cin) cic = d.cin;
  cic = (d.cin + ((d.cin + cic - 1) / cic) - 1) / ((d.cin + cic - 1) / cic);      // even chunks
  a.ci_chunk = cic;
  const size_t smem = fixed + per_ci * cic;
  ARPDE_CHECK_ARG(smem <= 200 * 1024, "conv_fwd: tile needs %zu bytes of shared memory", smem);
  if (smem > 48 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (long long)(one_d_rows ? 1 : d.N) * a.co_tiles * a.tiles_y * a.tiles_x;
  ARPDE_CHECK_ARG(blocks < (1ll << 31), "conv_fwd: grid too large");
  fn<<<(unsigned)blocks, dim3(32 * npw, ncg), smem, stream>>>(a);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

extern "C" int arpde_conv_fwd(const float* x, int64_t x_bstride, const float* w, int64_t w_gstride, const float* scale,
                              const float* shift, int64_t ss_gstride, float* y, double* stats, int N, int cin, int H, int W,
                              int cout, int kh, int kw, int stride, int up, int relu_in, int act, int y_ctot, int y_coff,
                              int n_per_group, arpde_stream_t stream) {
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1;
  d.N = N; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride; d.up = up;
  d.relu_in = relu_in; d.act = act; d.y_ctot = y_ctot; d.y_coff = y_coff; d.n_per_group = n_per_group > 0 ? n_per_group : (N > 0 ? N : 1);
  ARPDE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_fwd: scale/shift must both be given or both be null");
  ARPDE_CHECK_ARG(y_coff >= 0 && y_coff + cout <= y_ctot, "conv_fwd: channel window [%d,%d) outside y_ctot=%d", y_coff, y_coff + cout, y_ctot);
  return arpde_conv_fwd_launch(d, x, x_bstride, w, w_gstride, scale, shift, ss_gstride, y, stats, (cudaStream_t)stream);
}
