// conv_common.cuh -- descriptors shared by the convolution kernels and the DenseED program runner.
#pragma once
#include "common.cuh"

struct ConvDesc {
  int N, cin, H, W;        // input dims (H = 1 for 1D), before the optional nearest x2
  int cout, kh, kw, stride, up;
  int dilate = 0;          // input is zero-inserted x2 before the conv (transposed conv of a stride-2 layer)
  int pad_override = -1;   // wrap padding per side; -1 = reference default ((k-1)/2, k=4 -> 1)
  int relu_in;             // ReLU after the per-channel affine of the prologue
  int act;                 // output activation (ARPDE_ACT_*)
  int y_ctot, y_coff;      // output buffer channel count / channel offset (concat-free dense block)
  int n_per_group;         // samples sharing one weight set (SWAG: ICs per weight sample)
  int pdl = 0;             // rollouts: programmatic dependent launch for the kernels that support it (common.cuh)
};

struct ConvFwdArgs {
  const float* x; long long x_bs;
  int cin, H, W, up, up_h, dilate, pad_y, pad_x;
  const float* w; long long w_gs;
  const float* scale; const float* shift; long long ss_gs;
  int relu_in;
  float* y; int y_ctot, y_coff, cout, Ho, Wo, act;
  double* stats;
  int n_per_group, N;
  int tiles_x, tiles_y, co_tiles, TH, TW, pitch, ci_chunk;
  int rows_are_samples;    // 1D signals: the TH rows of a tile are TH consecutive samples (no vertical coupling), see conv_fwd.cu
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case 1: return tanhf(v);
    case 2: return fmaxf(v, 0.f);
    case 3: return v > 0.f ? v : 0.01f * v;
    case 4: return 1.f / (1.f + expf(-v));
    case 5: return 4.f * v > 20.f ? v : 0.25f * log1pf(expf(4.f * v));
    default: return v;
  }
}


// Stage `nlines` = nch x PH full-width patch lines (stride-1, full-width tiles: a line is one input row plus the circular halo) into
// shared memory with the per-channel affine + ReLU prologue applied.  Lanes are packed as (row, float4 column): a 32-wide row uses 8
// lanes, so one warp instruction moves 4 rows (the first version moved one row per instruction: for W = 32 the staging of a dense layer
// issued twice as many instructions as its FFMAs).  rowmap[py] = source row of patch row py; the caller synchronises afterwards.
template <int RU = 2>
__device__ __forceinline__ void stage_full_rows(float* __restrict__ patch, int pitch, const float* __restrict__ xn, int HWi, int c0, int nlines,
                                                int PH, const int* __restrict__ rowmap, int W, int pad_x, int nhalo,
                                                const float* __restrict__ scale, const float* __restrict__ shift, int relu, int tid,
                                                int nthreads) {
  const int W4 = W >> 2;
  const int G = W4 >= 32 ? 1 : 32 / W4;                // rows per warp pass
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int g = lane / W4, x4 = lane - g * W4;
  const bool lane_on = g < G;
  for (int base = warp * G * RU; base < nlines; base += nwarps * G * RU) {
    float4 v[RU]; float sc[RU], sh[RU]; int rr[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      const int r = base + u * G + g;
      const bool ok = lane_on && r < nlines;
      rr[u] = ok ? r : -1;
      const int ci = ok ? r / PH : 0, py = ok ? r - ci * PH : 0;
      sc[u] = 1.f; sh[u] = 0.f;
      if (scale && ok) { sc[u] = scale[c0 + ci]; sh[u] = shift[c0 + ci]; }
      v[u] = ok ? __ldg(reinterpret_cast<const float4*>(xn + (c0 + ci) * HWi + rowmap[py] * W) + x4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      if (rr[u] >= 0) {
        float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { e[k] = fmaf(e[k], sc[u], sh[u]); if (relu) e[k] = fmaxf(e[k], 0.f); }
        float* d4 = patch + rr[u] * pitch + pad_x + 4 * x4;
        if ((pad_x & 1) == 0) { *reinterpret_cast<float2*>(d4) = make_float2(e[0], e[1]); *reinterpret_cast<float2*>(d4 + 2) = make_float2(e[2], e[3]); }
        else { d4[0] = e[0]; d4[1] = e[1]; d4[2] = e[2]; d4[3] = e[3]; }
      }
    }
  }
  // halo columns: patch column h < pad_x <- input column W - pad_x + h, patch column W + h (h >= pad_x) <- input column h - pad_x
  for (int i = tid; i < nlines * nhalo; i += nthreads) {
    const int r = i / nhalo, h = i - r * nhalo;
    const int ci = r / PH, py = r - ci * PH;
    float tv = __ldg(xn + (c0 + ci) * HWi + rowmap[py] * W + (h < pad_x ? W - pad_x + h : h - pad_x));
    if (scale) tv = fmaf(tv, scale[c0 + ci], shift[c0 + ci]);
    if (relu) tv = fmaxf(tv, 0.f);
    patch[r * pitch + (h < pad_x ? h : W + h)] = tv;
  }
}

bool arpde_conv_fwd_rt_shape_ok(const ConvDesc& d);      // the row-tiled FFMA kernel covers (and wins on) this layer shape
int arpde_conv_fwd_launch(const ConvDesc& d, const float* x, long long x_bs, const float* w, long long w_gs,
                          const float* scale, const float* shift, long long ss_gs, float* y, double* stats,
                          cudaStream_t stream);

int arpde_weight_flip_transpose_launch(const float* w, float* wt, int cout, int cin, int kh, int kw, cudaStream_t stream);
// data-parallel hook of the program runner (arpde_densed_forward_dp / _backward_dp): in-place SUM over the ranks of n device doubles,
// enqueued by the callback on the calling stream (or ordered against it); world = number of ranks, every rank holds the same N
struct DpHook { int (*allreduce)(void* user, double* buf, long long n); void* user; int world; };
int arpde_bn_relu_bwd_launch(const float* da, const float* x, long long x_bs, const float* scale, const float* shift, const float* mean,
                             const float* invstd, float* dz, double* sums, float* gx, long long gx_bs, float* dgamma, float* dbeta, int N,
                             int C, int H, int W, int up_h, int up_w, int training, int accumulate, cudaStream_t stream,
                             const DpHook* dp = nullptr);
int arpde_act_bwd_launch(const float* gy, const float* y, float* out, long long n, int act, cudaStream_t stream);
int arpde_conv_wgrad_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                            long long dO_bs, double* dW64, float* dW, cudaStream_t stream, bool generic_only = false);

int arpde_conv_wgrad_best_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                                 long long dO_bs, double* dW64, float* dW, float* tc_scratch, long long tc_scratch_floats, cudaStream_t stream);
int arpde_d2f_launch(const double* src, float* dst, long long n, cudaStream_t stream);
// tensor-core weight gradient, conv_wgrad_tc.cu (ARPDE_WGRAD_TC=0 disables it)
bool arpde_conv_wgrad_tc_wanted(const ConvDesc& d);
long long arpde_conv_wgrad_tc_scratch_floats(const ConvDesc& d);  // fp32 partial-sum slabs (0 if the layer is not covered)
int arpde_conv_wgrad_tc_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                               long long dO_bs, double* dW64, float* dW, float* scratch, long long scratch_floats, cudaStream_t stream);

// persistent 1D rollout (rollout1d.cu): the reference's KS DenseED, one warp per trajectory, all T steps in one launch
struct arpde_op; struct arpde_buf;
bool arpde_ks_program_matches(const arpde_op* ops, int n_ops, const arpde_buf* bufs, int n_bufs, int H, int W, int c_in, int c_out);
int arpde_ks_rollout_launch(const arpde_op* ops, void* const* params, const int64_t* gstride, float* traj, long long bs, int c_hist, int T, int N,
                            int n_per_group, float eps, cudaStream_t stream);

// output convolution of the 2D net (conv_out.cu): 5x5, cout = 2, 64-column maps
bool arpde_conv_out_shape_ok(const ConvDesc& d, const float* x, long long x_bs, const float* y, const double* stats);
int arpde_conv_out_launch(const ConvDesc& d, const float* x, long long x_bs, const float* w, long long w_gs, const float* scale, const float* shift,
                          long long ss_gs, float* y, cudaStream_t stream);

// fused dense block (dense_block.cu): all 4-channel dense layers of a block in one launch, eval mode
int arpde_dense_block_match(const arpde_op* ops, int n_ops, int first, const arpde_buf* bufs, int N);
int arpde_dense_block_launch(const arpde_op* ops, int first, int nl, const arpde_buf* bufs, void* const* params, const int64_t* gstride, float* work,
                             const float* scale, const float* shift, long long ss_gs, int N, int n_per_group, cudaStream_t stream, bool pdl = false);

// tensor-core (tcgen05) implicit-GEMM path, conv_tc.cu
long long* arpde_tc_debug_buffer();                              // arpde_conv_tc_set_debug's buffer (nullptr = off)
int arpde_tc_enabled();                                          // ARPDE_TC=0 in the environment disables it
bool arpde_conv_tc_wanted(const ConvDesc& d);                    // layer heuristics + plannability
long long arpde_conv_tc_wprep_floats(int cin, int cout, int kh, int kw);   // prepared-weight floats per weight group
int arpde_conv_tc_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs,
                              cudaStream_t stream);
// operand-format hand-over to the stride-2 3x3 convolution that follows (conv_z.cu): the epilogue applies the consumer's eval-mode
// BatchNorm affine + ReLU and writes fp16 hi / lo octets into `z` instead of fp32 NCHW into y
struct ZOut { float* z; const float* scale; const float* shift; long long ss_gs; };
int arpde_conv_tc_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const float* scale,
                         const float* shift, long long ss_gs, float* y, double* stats, cudaStream_t stream, const ZOut* zo = nullptr);
// producer of the pair for networks' first layer (cin <= 8, odd k, stride 1, no prologue): conv_stem.cu
bool arpde_conv_stem_wanted(const ConvDesc& d);
long long arpde_conv_stem_wprep_floats(int cout, int k);
int arpde_conv_stem_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs, const ZOut& zo,
                                cudaStream_t stream);      // folds the consumer's BatchNorm scale / shift (zo) into the weights
int arpde_conv_stem_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const ZOut& zo, cudaStream_t stream);
long long arpde_conv_z_sample16(int c, int H, int W);            // 16-byte units per sample of the operand-format activation
bool arpde_conv_z_wanted(const ConvDesc& d);                     // d: the stride-2 consumer, described by its input dims
long long arpde_conv_z_wprep_floats(int cin, int cout);
int arpde_conv_z_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs, cudaStream_t stream);
int arpde_conv_z_launch(const ConvDesc& d, const float* z, const float* wprep, long long wprep_gs, float* y, cudaStream_t stream);
// eval-mode stride-1 1x1 / 3x3 (+ folded nearest x2) convolutions with fp16 two-term operands: conv_h.cu
bool arpde_conv_h_wanted(const ConvDesc& d);
long long arpde_conv_h_wprep_floats(int cin, int cout, int k, int up);
int arpde_conv_h_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs, cudaStream_t stream);
int arpde_conv_h_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const float* scale,
                        const float* shift, long long ss_gs, float* y, cudaStream_t stream);
