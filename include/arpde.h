/*
 * arpde.h -- C ABI of libarpde.so: the B200 (sm_100a) implementation of ar-pde-cnn's hot path.
 *
 * The reference (cics-nd/ar-pde-cnn) has no native layer: its hot path is PyTorch calls made from
 * nn/*.py.  Each entry point below names the reference Python interface it replaces (paths are
 * relative to the reference root).  The Python host side (ar-pde-cnn_b200/nn/...) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; fp32 data, NCHW / NCL;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing
 *     and keeps no state beyond two thread-local diagnostics (last error string, launch counter); scratch memory is passed in
 *     by the caller;
 *   - return value: 0 = ok, otherwise an ARPDE_ERR_* code; arpde_last_error() gives the message of
 *     the calling thread's last failure.  Nothing throws or aborts.
 *   - "bstride" = distance in elements between consecutive batch items (inputs may be channel-slice
 *     views such as input[:, -6:] of a longer history tensor); channel stride is always H*W (or L).
 */
#ifndef ARPDE_H
#define ARPDE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARPDE_VERSION 100          /* major*100 + minor */

/* error codes */
#define ARPDE_OK 0
#define ARPDE_ERR_ARG 1
#define ARPDE_ERR_CUDA 2
#define ARPDE_ERR_UNSUPPORTED 3

/* time integrators */
#define ARPDE_CN 0                 /* crankNicolson */
#define ARPDE_BE 1                 /* backwardEuler */

/* 1D systems */
#define ARPDE_SYS_BURGERS1D 1
#define ARPDE_SYS_KS1D 2

typedef void* arpde_stream_t;      /* cudaStream_t */

int arpde_version(void);
const char* arpde_last_error(void);
/* Diagnostic: number of kernels this process has launched through this library since the last reset (process-wide, atomic: autograd
 * calls the backward entry points from its own thread; reset != 0 clears it after reading).  bench.py reports it as "gpu_launches". */
int64_t arpde_launch_count(int reset);

/* ---------------------------------------------------------------------------------------------
 * Physics integrators (the loss target), forward.
 *   2D: replaces Burger2DIntegrate.crankNicolson / .backwardEuler
 *       (2D-Burgers-SWAG/nn/burger2DFiniteDifference.py:175-211 / :148-173), grad_kernels=[3,3].
 *   1D: replaces BurgerIntegrate.crankNicolson/.backwardEuler
 *       (1D-Burger-SWAG/nn/burgerFiniteDifference.py:284-302 / :264-282) for system=BURGERS1D and
 *       KSIntegrate.crankNicolson/.backwardEuler (1D-KS-SWAG/nn/ksFiniteDifference.py:155-173 /
 *       :139-153) for system=KS1D; k1,k2,k4 = grad_kernels (k4 ignored for Burgers).
 * uPred, uPred0: [N,2,H,W] (2D) / [N,1,L] (1D) with their own batch strides.
 * ustar (nullable): compact output.  sse_out (nullable): device double, receives
 *   sum((ustar - uPred)^2), the data term of BayesNN.calc_neg_log_joint (nn/bayesNN.py:113-115),
 *   without materialising ustar when ustar == NULL.  At least one of the two must be given.
 * ------------------------------------------------------------------------------------------- */
int arpde_cn_burgers2d_fwd(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                           float* ustar, double* sse_out, int N, int H, int W, double dx, double nu, double dt,
                           int mode, arpde_stream_t stream);
/* same arithmetic through the one-thread-per-point kernel (any H, W); parity test hook */
int arpde_cn_burgers2d_fwd_generic(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                                   float* ustar, double* sse_out, int N, int H, int W, double dx, double nu,
                                   double dt, int mode, arpde_stream_t stream);
int arpde_cn_1d_fwd(int system, const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                    float* ustar, double* sse_out, int N, int L, double dx, double nu, double dt, int k1, int k2,
                    int k4, int mode, arpde_stream_t stream);

/* Adjoint of the above (what torch autograd builds through the reference integrators; the loss
 * target is NOT detached, 2D-Burgers-SWAG/main.py:71-73): given grad_ustar (compact, same shape
 * as ustar) writes compact grad_uPred and grad_uPred0. */
int arpde_cn_burgers2d_bwd(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                           const float* grad_ustar, float* grad_uPred, float* grad_uPred0, int N, int H, int W,
                           double dx, double nu, double dt, int mode, arpde_stream_t stream);
/* same adjoint through the one-thread-per-point kernel (any H, W); parity test hook */
int arpde_cn_burgers2d_bwd_generic(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                                   const float* grad_ustar, float* grad_uPred, float* grad_uPred0, int N, int H, int W,
                                   double dx, double nu, double dt, int mode, arpde_stream_t stream);
/* Host-only test hook (no device access): the tiling the bulk-copy ring kernels behind arpde_cn_burgers2d_fwd / _bwd use for a batch of N
 * fields of H x W.  Returns 1 and fills rows_per_item, stages, smem_bytes, items when the shape takes that path (W/4 a power of two <= 32
 * or a multiple of 32, a double-buffered tile within shared memory), 0 when it falls back to the other kernels. */
int arpde_stencil_bulk_plan(int N, int H, int W, int backward, int mode, int* rows_per_item, int* stages, int64_t* smem_bytes,
                            int64_t* items);
int arpde_cn_1d_bwd(int system, const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                    const float* grad_ustar, float* grad_uPred, float* grad_uPred0, int N, int L, double dx,
                    double nu, double dt, int k1, int k2, int k4, int mode, arpde_stream_t stream);

/* Stand-alone periodic filters (forward only): the callable .grad1/.grad2/.grad4 objects of the reference
 * integrators (GradFilter1D / LaplaceFilter1D / FourthOrderFilter1D, ksFiniteDifference.py:6-119;
 * SobelFilter2d.grad_h/.grad_v and LapLaceFilter2d, burger2DFiniteDifference.py:6-129).  taps7_host: centred
 * 7-tap table, w9_host: row-major 3x3 weights, both HOST float arrays already scaled by 1/dx^n. */
int arpde_filter1d(const float* x, float* y, int64_t rows, int L, const float* taps7_host, arpde_stream_t stream);
int arpde_filter2d(const float* x, float* y, int64_t imgs, int H, int W, const float* w9_host, arpde_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Loss.  Replaces BayesNN.calc_neg_log_joint (nn/bayesNN.py:98-129; constants :29-38).
 *   params_host / numels_host: HOST arrays of device pointers / element counts of every tensor of
 *   model.features.parameters().  scratch: 2 device doubles (SSE, prior sum) kept for backward.
 *   output == target == NULL: scratch[0] already holds the SSE (written by arpde_cn_*_fwd).
 *   loss_out: device float[1] (the reference returns a shape-(1,) tensor).
 * bwd: grad_loss device float[1]; any of the grad outputs may be NULL; grad_params_flat is the
 *   concatenation of the per-tensor gradients in params_host order.
 * ------------------------------------------------------------------------------------------- */
int arpde_neg_log_joint_fwd(const float* output, const float* target, int64_t numel, int B, const float* log_beta,
                            const float* const* params_host, const int64_t* numels_host, int nparams, double ntrain,
                            double w_shape, double w_rate, double beta_shape, double beta_rate, float* loss_out,
                            double* scratch, arpde_stream_t stream);
int arpde_neg_log_joint_bwd(const float* output, const float* target, int64_t numel, int B, const float* log_beta,
                            const float* const* params_host, const int64_t* numels_host, int nparams, double ntrain,
                            double w_shape, double w_rate, double beta_shape, double beta_rate, const float* grad_loss,
                            const double* scratch, float* grad_output, float* grad_target, float* grad_log_beta,
                            float* grad_params_flat, arpde_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Circular convolution family.  One call replaces F.pad(circular)+conv of an
 * nn.Conv{1,2}d(padding_mode='circular') of nn/denseEDcirc2d.py / nn/denseEDcirc.py together with
 * the BatchNorm+ReLU in front of it (scale/shift prologue), the nearest x2 upsample (up=1), the
 * torch.cat of _DenseLayer.forward (:66-68; output written at channel y_coff of a y_ctot-channel
 * tensor) and the out_activation.  1D signals use H = 1, kh = 1.
 *   w: [groups][cout][cin][kh][kw]; sample n uses weight set n / n_per_group (w_gstride elements apart;
 *      SWAG: one weight sample for n_per_group initial conditions); scale/shift likewise (ss_gstride).
 *   stats (nullable): device double [y_ctot][2], receives per-channel (sum, sum of squares) of the
 *      written channels (train-mode BatchNorm of the consumer); must be zeroed by the caller.
 *   Effective wrap padding: (k-1)/2, and 1 for k=4 (torch<=1.2 semantics, reference README.md:40).
 * ------------------------------------------------------------------------------------------- */
#define ARPDE_ACT_NONE 0
#define ARPDE_ACT_TANH 1
#define ARPDE_ACT_RELU 2
#define ARPDE_ACT_LRELU 3
#define ARPDE_ACT_SIGMOID 4
#define ARPDE_ACT_SOFTPLUS4 5
int arpde_conv_fwd(const float* x, int64_t x_bstride, const float* w, int64_t w_gstride, const float* scale,
                   const float* shift, int64_t ss_gstride, float* y, double* stats, int N, int cin, int H, int W, int cout,
                   int kh, int kw, int stride, int up, int relu_in, int act, int y_ctot, int y_coff, int n_per_group,
                   arpde_stream_t stream);

/* The same convolution on the 5th-generation tensor cores (tcgen05, accumulators in TMEM) with an error-compensated
 * 3xTF32 split, i.e. fp32-grade results (conv_tc.cu).  scratch: arpde_conv_fwd_tc_scratch_floats(...) device floats for the
 * split weights.  Returns ARPDE_ERR_UNSUPPORTED for shapes the kernel does not cover (cout > 128, > 49 taps). */
int64_t arpde_conv_fwd_tc_scratch_floats(int cin, int cout, int kh, int kw, int groups);
int arpde_conv_fwd_tc(const float* x, int64_t x_bstride, const float* w, int64_t w_gstride, const float* scale,
                      const float* shift, int64_t ss_gstride, float* y, double* stats, int N, int cin, int H, int W, int cout,
                      int kh, int kw, int stride, int up, int relu_in, int act, int y_ctot, int y_coff, int n_per_group,
                      float* scratch, int64_t scratch_floats, arpde_stream_t stream);

/* Diagnostic, host only: the tiling the tensor-core kernel uses for a layer (see conv_tc.cu for the field list; out holds
 * at least 130 int32).  Used by the CPU tests to check the shifted-window index arithmetic without a GPU. */
int arpde_conv_tc_plan(int cin, int H, int W, int cout, int kh, int kw, int stride, int up, int32_t* out);

/* Weight gradient of one circular convolution on the tensor cores (conv_wgrad_tc.cu): the dW that torch autograd derives for
 * nn.Conv2d(padding_mode='circular') under loss.backward() (2D-Burgers-SWAG/main.py:81), with the same fused prologue as the
 * forward op: a = up2?(relu?(scale*x + shift)), dW[co][ci][ky][kx] = sum_{n,oy,ox} dO[n][co][oy][ox] * a_wrapped[n][ci][oy*s+ky-p][ox*s+kx-p].
 * dW64: scratch of cout*cin*kh*kw doubles (fp64 accumulation across CTAs), dW: float result [cout][cin][kh][kw];
 * scratch: arpde_conv_wgrad_tc_scratch(...) floats of device memory (per-CTA fp32 partial sums; 0 = layer not covered).
 * Returns ARPDE_ERR_UNSUPPORTED for layers the kernel does not cover (1D, cout < 16 or > 128, too few taps); the fused backward
 * (arpde_densed_backward) then uses the fp32 FFMA kernel.  arpde_conv_wgrad_tc_plan: host-only diagnostic of the tiling
 * (field list in conv_wgrad_tc.cu; out holds at least 400 int32). */
int64_t arpde_conv_wgrad_tc_scratch(int N, int cin, int H, int W, int cout, int kh, int kw, int stride, int up);
int arpde_conv_wgrad_tc(const float* x, int64_t x_bstride, const float* scale, const float* shift, const float* dO,
                        int64_t dO_bstride, double* dW64, float* dW, int N, int cin, int H, int W, int cout, int kh, int kw,
                        int stride, int up, int relu_in, float* scratch, int64_t scratch_floats, arpde_stream_t stream);
/* Same contract, best available kernel: the sliding-window kernel for narrow layers (cout <= 4, stride 1), the tensor core when
 * `scratch` is given and covers the layer, else the generic fp32 kernel.  path = 0 auto, 1 generic kernel only (cross-check). */
int arpde_conv_wgrad(const float* x, int64_t x_bstride, const float* scale, const float* shift, const float* dO, int64_t dO_bstride,
                     double* dW64, float* dW, int N, int cin, int H, int W, int cout, int kh, int kw, int stride, int up, int relu_in,
                     float* scratch, int64_t scratch_floats, int path, arpde_stream_t stream);
int arpde_conv_wgrad_tc_plan(int cin, int H, int W, int cout, int kh, int kw, int stride, int up, int32_t* out);

/* Profiling hooks of the tensor-core kernel.  They are process-global state, so they exist ONLY in a development build of the
 * library (`ARPDE_DEV=1 bash ar-pde-cnn_b200/csrc/build.sh`, which also enables the ARPDE_* environment switches); the shipped
 * library neither exports them nor reads the environment.  set_debug: per-role cycle counters of CTA 0 of every following
 * arpde_conv_fwd_tc / program launch are written to dev_buf (16 device int64; NULL = off).  set_debug_flags: ablation switches
 * (1 = skip patch staging, 2 = skip the epilogue, 4 = skip the MMAs; results are garbage). */
#ifdef ARPDE_DEV
int arpde_conv_tc_set_debug(long long* dev_buf);
int arpde_conv_tc_set_debug_flags(int flags);
#endif

/* Random periodic initial conditions of the 2D Burgers data set on the device (InitPeriodicCond2d.__getitem__,
 * 2D-Burgers-SWAG/utils/burgerLoader2D.py:99-112): u[b][comp] = 2 f / max|f| + c[b][comp] with
 * f = sum_m lam[b][0][comp][m] cos(2 pi k_m.x) + lam[b][1][comp][m] sin(2 pi k_m.x) over the (2*order+1)^2 modes, evaluated in fp64.
 * lam [B][2][2][(2*order+1)^2] and c [B][2] are DEVICE doubles (the host draws them from numpy's per-index stream); out [B][2][n][n]. */
int arpde_fourier_ic2d(const double* lam, const double* c, float* out, int B, int ncells, int order, arpde_stream_t stream);

/* Error metrics of a (sample-mean) rollout against a target trajectory, on the device (testMSE / testBayesianMSE,
 * 2D-Burgers-SWAG/post/plotMSE.py:78-85,133-137): for every case n < N and saved step t < Tp over the F values of one state
 *   mse[n*Tp+t] = mean_f (pred - target)^2,  ese[n*Tp+t] = (sum_f pred^2 / (2*points) - sum_f target^2 / (2*points))^2   (fp64).
 * Row (n, t) of pred / target starts at n*nstride + t*tstride floats (tstride = test_every * state size selects every k-th step). */
int arpde_error_metrics(const float* pred, int64_t pred_nstride, int64_t pred_tstride, const float* target, int64_t tgt_nstride,
                        int64_t tgt_tstride, int N, int Tp, int64_t F, int64_t points, double* mse, double* ese, arpde_stream_t stream);

/* One-launch Adam step over n fp32 tensors (torch.optim.Adam.step of 2D-Burgers-SWAG/main.py:221-222,84; non-amsgrad):
 * host arrays of DEVICE pointers (param, grad, exp_avg, exp_avg_sq) and element counts; per tensor step_size = lr / (1 - beta1^t),
 * bias_correction2_sqrt = sqrt(1 - beta2^t) and L2 weight decay (NULL = 0), computed by the caller in double as torch does.
 * The state tensors are torch.optim.Adam's own, so optimizer checkpoints interchange. */
int arpde_adam_step(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq, const int64_t* numel,
                    const double* step_size, const double* bias_correction2_sqrt, const double* weight_decay, int n, double beta1,
                    double beta2, double eps, arpde_stream_t stream);

/* BatchNorm bookkeeping (nn.BatchNorm{1,2}d, eps 1e-5, momentum 0.1): per-channel sums of x[N,C(+),HW],
 * train-mode finalize (scale/shift, saved mean/invstd, running-stat update, num_batches_tracked += 1),
 * eval-mode fold of the running statistics. */
int arpde_bn_stats(const float* x, int64_t x_bstride, int N, int C, int64_t HW, double* stats, arpde_stream_t stream);
int arpde_bn_finalize_train(const double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, int64_t* num_batches_tracked, double momentum, double eps, float* scale,
                            float* shift, float* save_mean, float* save_invstd, int C, arpde_stream_t stream);
int arpde_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       double eps, float* scale, float* shift, int C, arpde_stream_t stream);
/* y = act(scale[c]*x + shift[c]) (scale/shift nullable), nearest x2 along H / W when up_h / up_w = 1: the unfused forward of a lone
 * BatchNorm / ReLU / UpsamplingNearest{1,2}d module (DenseED.forward_test, denseEDcirc2d.py:316-323). */
int arpde_affine_act(const float* x, int64_t x_bstride, const float* scale, const float* shift, float* y, int N, int C, int H,
                     int W, int up_h, int up_w, int act, arpde_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-network program.  Replaces DenseED.forward (denseEDcirc2d.py:313-314; 1D :306-307 / :294-295).
 * The host side flattens model.features into conv ops (every BN+ReLU[+upsample] is the prologue of
 * the conv that follows it) over a few activation buffers carved out of one workspace.
 * ------------------------------------------------------------------------------------------- */
typedef struct arpde_op {
  int32_t in_buf;      /* activation buffer read (channels [0,cin)); -1 = network input */
  int32_t out_buf;     /* buffer written; -2 = network output */
  int32_t cin, cout, kh, kw, stride, up;
  int32_t out_coff;    /* channel offset inside out_buf (dense-block concat) */
  int32_t act;         /* ARPDE_ACT_* applied to the output */
  int32_t w_param;     /* index of the conv weight in the parameter table */
  int32_t bn_param;    /* index of BN gamma; beta, running_mean, running_var, num_batches_tracked follow; -1 = none */
  int32_t bn_off;      /* offset of this BN's channels inside the scale/shift scratch */
} arpde_op_t;

typedef struct arpde_buf {
  int64_t offset;      /* floats from the workspace base */
  int64_t stats_off;   /* doubles from the stats base ([ctot][2]) */
  int32_t ctot, h, w;
  int32_t pad_;
} arpde_buf_t;

/* params_host: HOST array of device pointers (n_params entries); param_gstride_host: per-entry distance in
 * elements between weight sets (0 = shared).  scale_shift: 2*groups*bn_total floats.  Training (groups == 1):
 * stats = stats_len doubles (zeroed here), saved_mean_invstd (nullable) = 2*bn_total floats for backward;
 * BatchNorm running statistics in the parameter table are updated in place. */
int arpde_densed_forward(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                         const int64_t* param_gstride_host, int n_params, const float* x, int64_t x_bstride, int N,
                         int n_per_group, int H, int W, float* y, int y_ctot, int y_coff, float* workspace,
                         float* scale_shift, int bn_total, double* stats, int64_t stats_len, float* saved_mean_invstd,
                         int training, double momentum, double eps, float* tc_scratch, int64_t tc_scratch_floats,
                         arpde_stream_t stream);
/* tc_scratch (nullable): arpde_densed_tc_scratch_floats(ops, n_ops, groups) floats of device memory in which the weights of
 * the layers that run on the tensor cores are split into TF32 hi/lo parts; NULL keeps every layer on the fp32 FFMA kernel. */
int64_t arpde_densed_tc_scratch_floats(const arpde_op_t* ops, int n_ops, int groups);

/* Backward of arpde_densed_forward: the autograd pass loss.backward() runs through DenseED in the reference
 * (2D-Burgers-SWAG/main.py:81).  gy: compact gradient of the output; y: forward output (only read when the last op
 * has an activation); workspace / scale_shift / saved_mean_invstd: as left by the forward call (eval mode: the caller
 * fills saved_mean_invstd with running_mean and 1/sqrt(running_var+eps)); gwork: gradient workspace, same size as
 * workspace; gx (nullable): gradient of the input [N,cin,H,W] compact; grads_host: HOST array (n_params entries) of
 * device float* receiving each parameter's gradient (NULL = not needed; running-stat slots are ignored).
 * The stats scratch of the forward: when an op applies BatchNorm directly to the network input (in_buf == -1; a stand-alone
 * _Transition / LastTransUp), stats_len includes 2*cin extra doubles at the tail for the input's batch statistics.
 * Scratch: T1 >= max N*cin*Hc*Wc floats over ops (conv-input resolution), T2 >= max(N*cin*H*W, N*cout_last*Ho*Wo) floats,
 * wT >= max weight numel floats, d64 >= max(weight numel, 2*max cin) doubles.  tc_scratch (nullable): arpde_densed_tc_scratch_floats(ops,
 * n_ops, 1) floats; with it the data-gradient convolutions of the stride-1 layers run on the tensor-core kernel. */
int arpde_densed_backward(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                          int n_params, const float* x, int64_t x_bstride, int N, int H, int W, const float* gy, const float* y,
                          const float* workspace, const float* scale_shift, int bn_total, const float* saved_mean_invstd,
                          float* gwork, float* gx, void* const* grads_host, float* T1, float* T2, float* wT, double* d64,
                          int training, float* tc_scratch, int64_t tc_scratch_floats, arpde_stream_t stream);

/* Data-parallel training (one process per GPU; the mini-batch of initial conditions of 2D-Burgers-SWAG/main.py:47-61 split over
 * `dp_world` ranks holding N samples each): same contracts as arpde_densed_forward / _backward, but train-mode BatchNorm uses the
 * statistics of the GLOBAL mini-batch, which is what the reference computes on one device.  The library holds no communicator:
 * between the kernels that accumulate a layer's sums -- (sum x, sum x^2) forward, (sum dz, sum dz*xhat) backward -- and the kernels
 * that consume them it calls `allreduce(user, buf, n)` from the calling host thread; the callback must sum the n device doubles at
 * buf in place over all ranks, ordered on `stream` (e.g. ncclAllReduce on that stream, or torch.distributed.all_reduce on the
 * current stream), and return 0.  Weight / gamma / beta gradients stay per-rank sums for the caller's gradient all-reduce. */
typedef int (*arpde_allreduce_fn)(void* user, double* buf, int64_t n);
int arpde_densed_forward_dp(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                            const int64_t* param_gstride_host, int n_params, const float* x, int64_t x_bstride, int N,
                            int n_per_group, int H, int W, float* y, int y_ctot, int y_coff, float* workspace,
                            float* scale_shift, int bn_total, double* stats, int64_t stats_len, float* saved_mean_invstd,
                            int training, double momentum, double eps, float* tc_scratch, int64_t tc_scratch_floats,
                            arpde_allreduce_fn allreduce, void* user, int dp_world, arpde_stream_t stream);
int arpde_densed_backward_dp(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                             int n_params, const float* x, int64_t x_bstride, int N, int H, int W, const float* gy, const float* y,
                             const float* workspace, const float* scale_shift, int bn_total, const float* saved_mean_invstd,
                             float* gwork, float* gx, void* const* grads_host, float* T1, float* T2, float* wT, double* d64,
                             int training, float* tc_scratch, int64_t tc_scratch_floats, arpde_allreduce_fn allreduce, void* user,
                             int dp_world, arpde_stream_t stream);

/* Eval-mode stem pair (conv_z.cu): y = conv3x3_stride2(relu(scale * conv_k1xk1(x) + shift)), both circular -- In_conv1 -> In_conv_norm1 ->
 * In_conv_relu1 -> In_conv3 of 2D-Burgers-SWAG/nn/denseEDcirc2d.py:247-255 -- with the intermediate activation handed over in the second
 * convolution's fp16 two-term tensor-core operand format (scratch `z`, arpde_conv_z_floats(c1, H, W) floats per sample; 0 = shape not
 * covered) instead of fp32 NCHW.  arpde_rollout uses the same pair automatically when the producer's work buffer has that much room.
 * scratch: arpde_conv_fwd_tc_scratch_floats(cin, c1, k1, k1, groups) + arpde_conv_fwd_tc_scratch_floats(c1, c3, 3, 3, groups) floats. */
int64_t arpde_conv_z_floats(int c, int H, int W);
int arpde_stem_pair_fwd(const float* x, int64_t x_bstride, const float* w1, int64_t w1_gstride, const float* scale, const float* shift,
                        int64_t ss_gstride, const float* w3, int64_t w3_gstride, float* y, int N, int cin, int H, int W, int c1, int k1,
                        int c3, int y_ctot, int y_coff, int n_per_group, float* z, int64_t z_floats, float* scratch, int64_t scratch_floats,
                        arpde_stream_t stream);
/* The two halves of the pair on their own (per-kernel timing, tests): producer x -> z with scratch of arpde_conv_fwd_tc_scratch_floats(cin, c1,
 * k1, k1, groups) floats; consumer z -> y with arpde_conv_fwd_tc_scratch_floats(c1, c3, 3, 3, groups). */
int arpde_stem_pair_producer(const float* x, int64_t x_bstride, const float* w1, int64_t w1_gstride, const float* scale, const float* shift,
                             int64_t ss_gstride, int N, int cin, int H, int W, int c1, int k1, int n_per_group, float* z, int64_t z_floats,
                             float* scratch, int64_t scratch_floats, arpde_stream_t stream);
int arpde_stem_pair_consumer(const float* z, const float* w3, int64_t w3_gstride, float* y, int N, int H, int W, int c1, int c3, int y_ctot,
                             int y_coff, int n_per_group, float* scratch, int64_t scratch_floats, arpde_stream_t stream);

/* Eval-mode stride-1 convolutions of the decoder on the tensor cores with fp16 two-term operands (conv_h.cu):
 *   y = conv_kxk(up2?(relu?(scale * x + shift))), circular, stride 1, k = 1 or 3 (up = 1 needs k = 3: UpsamplingNearest2d + conv3x3 of
 * LastTransUp / _Transition, 2D-Burgers-SWAG/nn/denseEDcirc2d.py:103-108,128-141,170-183, folded to one 3x3 convolution on the low-resolution
 * grid).  Same argument meaning as arpde_conv_fwd; scale / shift nullable (both).  arpde_rollout uses it for the layers it covers.
 * scratch: arpde_conv_fwd_h_scratch_floats(cin, cout, k, up, groups) floats.  Returns ARPDE_ERR_UNSUPPORTED for shapes the kernel does not
 * cover (needs 16 | cin <= 256, H >= 2; cout <= 128, folded: cout <= 32). */
int64_t arpde_conv_fwd_h_scratch_floats(int cin, int cout, int k, int up, int groups);
int arpde_conv_fwd_h(const float* x, int64_t x_bstride, const float* w, int64_t w_gstride, const float* scale, const float* shift,
                     int64_t ss_gstride, float* y, int N, int cin, int H, int W, int cout, int k, int up, int relu_in, int y_ctot, int y_coff,
                     int n_per_group, float* scratch, int64_t scratch_floats, arpde_stream_t stream);

/* A ready-made arpde_allreduce_fn over NVLink peer memory (one process per GPU on one node; peer.cu): every rank copies its n doubles
 * into its own exchange slot, publishes a sequence number in every peer's flag array, waits for all peers, then reads all slots over
 * NVLink and adds them in rank order (bit-identical on every rank) -- one small kernel on `stream`, no host round trip.  The CALLER owns
 * the context: slots[r] / flags[r] are rank r's exchange buffer ([nslots][slot_doubles] doubles) and flag array ([nslots][world] uint32,
 * zeroed before first use) as mapped into THIS process (e.g. torch.distributed._symmetric_memory buffer_ptrs), seq the number of
 * all-reduces issued so far (advanced by the call; must stay equal on all ranks), error_flag a device int the kernel raises when a
 * peer does not answer within ~2 s.  Never use it with several ranks on one GPU (the kernel waits for its peers' kernels). */
#define ARPDE_PEER_MAX 8
typedef struct arpde_peer_ctx {
  double* slots[ARPDE_PEER_MAX];
  uint32_t* flags[ARPDE_PEER_MAX];
  int32_t world, rank, nslots, slot_doubles;
  uint64_t seq;
  int32_t* error_flag;
  arpde_stream_t stream;
} arpde_peer_ctx_t;
int arpde_peer_allreduce_f64(void* ctx, double* buf, int64_t n);

/* Auto-regressive rollout in eval mode: the AR loops of main.py test()/testSample()
 * (2D-Burgers-SWAG/main.py:112-135,:160-179; 1D main.py :106-121,:148-163) without Python in the loop.
 * traj [N, traj_ctot, H, W], traj_ctot >= c_hist + T*c_out; channels [0,c_hist) hold the repeated initial
 * condition; step t reads channels [c_hist-c_in+t*c_out, +c_in) and writes [c_hist+t*c_out, +c_out). */
int arpde_rollout(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                  const int64_t* param_gstride_host, int n_params, float* traj, int traj_ctot, int c_in, int c_hist, int c_out,
                  int T, int N, int n_per_group, int H, int W, float* workspace, float* scale_shift, int bn_total, double eps,
                  float* tc_scratch, int64_t tc_scratch_floats, arpde_stream_t stream);
/* arpde_rollout runs the reference's 1D Kuramoto-Sivashinsky network (1D-KS-SWAG/main.py defaults: conv7/s2 2->32, four 4-channel
 * dense layers, 48->24->12->1 decoder, 96-point grid) as ONE persistent launch for all T steps (rollout1d.cu: one warp per
 * trajectory, weights and activations in shared memory, 96 floats per step to HBM).  arpde_rollout_layerwise has the same contract
 * and always takes the layer-by-layer path (one launch per convolution and step): the cross-check of the persistent kernel. */
int arpde_rollout_layerwise(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                            const int64_t* param_gstride_host, int n_params, float* traj, int traj_ctot, int c_in, int c_hist,
                            int c_out, int T, int N, int n_per_group, int H, int W, float* workspace, float* scale_shift,
                            int bn_total, double eps, float* tc_scratch, int64_t tc_scratch_floats, arpde_stream_t stream);

/* Ground-truth finite-difference solver of the 1D viscous Burgers equation, fp64, a batch of trajectories in one launch:
 * burgers1d() of 1D-Burger-SWAG/solver/fd_burger1D.py:32-98 (RK4 :45-52; upwind gradient of order grad_order in {1,2} chosen by the sign
 * of u :36-42,58-68; central second derivative of order div_order in {2,4} :70-79; periodic).  u0: [N][L] device doubles; out: [N][nsteps /
 * save_every][L] device floats, row r = the state after (r+1)*save_every steps rounded to fp32 as the reference stores it (save_every = 1
 * reproduces its [nsteps, L] array). */
int arpde_fd_burgers1d(const double* u0, float* out, int N, int L, double dx, double nu, double dt, int nsteps, int save_every,
                       int grad_order, int div_order, arpde_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SWAG.  arpde_swag_sample replaces SwagNN.sample (nn/swag.py:80-150) for S samples at once:
 *   mean/sq_mean/cov_host: HOST arrays of per-tensor device pointers in named_parameters() order
 *   (cov: [K, numel] rows), z1 [S,P] and z2 [S,n_tensors,K] standard-normal noise (the caller draws
 *   it, so results can be reproduced against the reference's RNG stream), out [S,P] flat weights.
 * arpde_predictive_moments replaces the sample mean/variance of post/plotBARContour.py:85-104: traj holds
 *   S*n_ic rows of `inner` floats, sample-major, `bstride` floats apart; var = noise + E[y^2] - E[y]^2. */
int arpde_swag_sample(const float* const* mean_host, const float* const* sq_mean_host, const float* const* cov_host,
                      const int64_t* numels_host, int n_tensors, const float* z1, const float* z2, float* out, int S, int K,
                      int max_models, double scale, double eps, int diag, arpde_stream_t stream);
int arpde_predictive_moments(const float* traj, int64_t bstride, int64_t inner, int n_ic, int S, double noise, float* mean,
                             float* var, arpde_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ARPDE_H */
