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

int arpde_conv_fwd_launch(const ConvDesc& d, const float* x, long long x_bs, const float* w, long long w_gs,
                          const float* scale, const float* shift, long long ss_gs, float* y, double* stats,
                          cudaStream_t stream);

int arpde_weight_flip_transpose_launch(const float* w, float* wt, int cout, int cin, int kh, int kw, cudaStream_t stream);
int arpde_bn_relu_bwd_launch(const float* da, const float* x, long long x_bs, const float* scale, const float* shift, const float* mean,
                             const float* invstd, float* dz, double* sums, float* gx, long long gx_bs, float* dgamma, float* dbeta, int N,
                             int C, int H, int W, int up_h, int up_w, int training, int accumulate, cudaStream_t stream);
int arpde_act_bwd_launch(const float* gy, const float* y, float* out, long long n, int act, cudaStream_t stream);
int arpde_conv_wgrad_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                            long long dO_bs, double* dW64, float* dW, cudaStream_t stream);

int arpde_conv_wgrad_best_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                                 long long dO_bs, double* dW64, float* dW, float* tc_scratch, long long tc_scratch_floats, cudaStream_t stream);
int arpde_d2f_launch(const double* src, float* dst, long long n, cudaStream_t stream);
// tensor-core weight gradient, conv_wgrad_tc.cu (ARPDE_WGRAD_TC=0 disables it)
bool arpde_conv_wgrad_tc_wanted(const ConvDesc& d);
long long arpde_conv_wgrad_tc_scratch_floats(const ConvDesc& d);  // fp32 partial-sum slabs (0 if the layer is not covered)
int arpde_conv_wgrad_tc_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                               long long dO_bs, double* dW64, float* dW, float* scratch, long long scratch_floats, cudaStream_t stream);

// tensor-core (tcgen05) implicit-GEMM path, conv_tc.cu
long long* arpde_tc_debug_buffer();                              // arpde_conv_tc_set_debug's buffer (nullptr = off)
int arpde_tc_enabled();                                          // ARPDE_TC=0 in the environment disables it
bool arpde_conv_tc_wanted(const ConvDesc& d);                    // layer heuristics + plannability
long long arpde_conv_tc_wprep_floats(int cin, int cout, int kh, int kw);   // prepared-weight floats per weight group
int arpde_conv_tc_prep_launch(const ConvDesc& d, const float* w, long long w_gs, int groups, float* wprep, long long wprep_gs,
                              cudaStream_t stream);
int arpde_conv_tc_launch(const ConvDesc& d, const float* x, long long x_bs, const float* wprep, long long wprep_gs, const float* scale,
                         const float* shift, long long ss_gs, float* y, double* stats, cudaStream_t stream);
