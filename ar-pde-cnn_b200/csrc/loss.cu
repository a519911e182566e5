// loss.cu -- negative log joint of the Bayesian surrogate (likelihood + Student-t weight prior +
// Gamma prior on the noise precision), forward and backward, in 2-3 launches instead of ~145.
//
// Replaces BayesNN.calc_neg_log_joint (reference */nn/bayesNN.py:98-129; constants :29-38):
//   loss = -(M/B)(-1/2 e^{lb} SSE + 1/2 N lb) + (a_w+1/2) sum_p sum log1p(w^2 * 0.5/b_w)
//          - ((a_b - 1) lb - e^{lb} r_b),       SSE = sum (target-output)^2, N = target.numel()
// The prior sum runs over every tensor of model.features.parameters() (conv weights, BN gamma/beta).
#include "common.cuh"
#include "../../include/arpde.h"

#define SEG_MAX 64
struct SegTable {
  const float* ptr[SEG_MAX];
  long long off[SEG_MAX + 1];   // prefix sums of numel
  int n;
};

__device__ __forceinline__ int seg_find(const SegTable& t, long long i) {
  int lo = 0, hi = t.n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t.off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// scratch[0] += SSE (if output != null), scratch[1] += sum log1p(cw * w^2)
__global__ void __launch_bounds__(256) nlj_reduce(const float* __restrict__ output, const float* __restrict__ target,
                                                  long long numel, SegTable segs, float cw, double* __restrict__ scratch) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  float sse = 0.f;
  if (output) {
    for (long long i = tid; i < numel; i += nth) { const float d = target[i] - output[i]; sse = fmaf(d, d, sse); }
  }
  double s1 = block_sum_d((double)sse);
  float pr = 0.f;
  const long long P = segs.n ? segs.off[segs.n] : 0;
  for (long long i = tid; i < P; i += nth) {
    const int s = seg_find(segs, i);
    const float w = segs.ptr[s][i - segs.off[s]];
    pr += log1pf(cw * w * w);
  }
  double s2 = block_sum_d((double)pr);
  if (threadIdx.x == 0) {
    if (output && s1 != 0.0) atomicAdd(scratch + 0, s1);
    if (s2 != 0.0) atomicAdd(scratch + 1, s2);
  }
}

struct NljConst { double mb, halfN, w_coef, bshape_m1, brate; };

__global__ void nlj_finalize(const double* __restrict__ scratch, const float* __restrict__ log_beta, NljConst c,
                             float* __restrict__ loss) {
  const double lb = (double)log_beta[0], elb = exp(lb);
  const double loglik = c.mb * (-0.5 * elb * scratch[0] + c.halfN * lb);
  const double prior_w = -c.w_coef * scratch[1];
  const double prior_b = c.bshape_m1 * lb - elb * c.brate;
  loss[0] = (float)(-loglik - prior_w - prior_b);
}

// grads: output/target (elementwise), params (flat, elementwise), log_beta (thread 0 of block 0)
__global__ void __launch_bounds__(256) nlj_bwd(const float* __restrict__ output, const float* __restrict__ target, long long numel,
                                               SegTable segs, float cw, const double* __restrict__ scratch,
                                               const float* __restrict__ log_beta, NljConst c, const float* __restrict__ gloss,
                                               float* __restrict__ g_out, float* __restrict__ g_tgt, float* __restrict__ g_lb,
                                               float* __restrict__ g_params) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const float g = gloss[0];
  const double lb = (double)log_beta[0], elb = exp(lb);
  if (g_out || g_tgt) {
    const float s = (float)(c.mb * elb) * g;            // d loss / d target = (M/B) e^lb (target-output)
    for (long long i = tid; i < numel; i += nth) {
      const float d = s * (target[i] - output[i]);
      if (g_tgt) g_tgt[i] = d;
      if (g_out) g_out[i] = -d;
    }
  }
  if (g_params) {
    const long long P = segs.n ? segs.off[segs.n] : 0;
    const float a = (float)(c.w_coef * 2.0) * cw * g;   // (a_w+1/2) * 2 cw w / (1 + cw w^2)
    for (long long i = tid; i < P; i += nth) {
      const int sg = seg_find(segs, i);
      const float w = segs.ptr[sg][i - segs.off[sg]];
      g_params[i] = a * w / fmaf(cw * w, w, 1.f);
    }
  }
  if (g_lb && tid == 0) {
    const double d = c.mb * (0.5 * elb * scratch[0] - c.halfN) - c.bshape_m1 + elb * c.brate;
    g_lb[0] = (float)(d * (double)g);
  }
}

static int fill_segs(SegTable* t, const float* const* params_host, const int64_t* numels_host, int first, int count) {
  t->n = count; t->off[0] = 0;
  for (int i = 0; i < count; ++i) {
    ARPDE_CHECK_ARG(params_host[first + i] != nullptr && numels_host[first + i] >= 0, "neg_log_joint: bad parameter %d", first + i);
    t->ptr[i] = params_host[first + i];
    t->off[i + 1] = t->off[i] + numels_host[first + i];
  }
  return ARPDE_OK;
}

static NljConst make_const(double ntrain, int B, int64_t numel, double w_shape, double beta_shape, double beta_rate) {
  NljConst c;
  c.mb = ntrain / (double)B; c.halfN = 0.5 * (double)numel; c.w_coef = w_shape + 0.5;
  c.bshape_m1 = beta_shape - 1.0; c.brate = beta_rate;
  return c;
}

extern "C" int arpde_neg_log_joint_fwd(const float* output, const float* target, int64_t numel, int B, const float* log_beta,
                                       const float* const* params_host, const int64_t* numels_host, int nparams,
                                       double ntrain, double w_shape, double w_rate, double beta_shape, double beta_rate,
                                       float* loss_out, double* scratch, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(log_beta && loss_out && scratch, "neg_log_joint_fwd: null pointer");
  ARPDE_CHECK_ARG((output == nullptr) == (target == nullptr), "neg_log_joint_fwd: output/target must both be given or both be null");
  ARPDE_CHECK_ARG(B > 0 && numel >= 0 && nparams >= 0 && w_rate > 0, "neg_log_joint_fwd: bad sizes");
  const float cw = (float)(0.5 / w_rate);
  // output == NULL: scratch[0] already holds the SSE written by arpde_cn_*_fwd(sse_out=scratch)
  if (output) ARPDE_CUDA(cudaMemsetAsync(scratch, 0, 2 * sizeof(double), stream));
  else ARPDE_CUDA(cudaMemsetAsync(scratch + 1, 0, sizeof(double), stream));
  int first = 0;
  bool did_sse = (output == nullptr);
  do {
    const int count = nparams - first < SEG_MAX ? nparams - first : SEG_MAX;
    SegTable t; int rc = fill_segs(&t, params_host, numels_host, first, count); if (rc) return rc;
    const long long work = (did_sse ? 0 : numel) > t.off[count] ? (did_sse ? 0 : numel) : t.off[count];
    int64_t blocks = ceil_div64(work > 0 ? work : 1, 256 * 4); const int64_t cap = (int64_t)arpde_num_sms() * 8;
    if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
    nlj_reduce<<<(unsigned)blocks, 256, 0, stream>>>(did_sse ? nullptr : output, target, numel, t, cw, scratch);
    ARPDE_LAUNCH_CHECK();
    did_sse = true; first += count;
  } while (first < nparams);
  nlj_finalize<<<1, 1, 0, stream>>>(scratch, log_beta, make_const(ntrain, B, numel, w_shape, beta_shape, beta_rate), loss_out);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

extern "C" int arpde_neg_log_joint_bwd(const float* output, const float* target, int64_t numel, int B, const float* log_beta,
                                       const float* const* params_host, const int64_t* numels_host, int nparams,
                                       double ntrain, double w_shape, double w_rate, double beta_shape, double beta_rate,
                                       const float* grad_loss, const double* scratch, float* grad_output, float* grad_target,
                                       float* grad_log_beta, float* grad_params_flat, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(log_beta && grad_loss && scratch, "neg_log_joint_bwd: null pointer");
  ARPDE_CHECK_ARG(!(grad_output || grad_target) || (output && target), "neg_log_joint_bwd: output/target needed for their grads");
  ARPDE_CHECK_ARG(B > 0 && numel >= 0 && nparams >= 0 && w_rate > 0, "neg_log_joint_bwd: bad sizes");
  const float cw = (float)(0.5 / w_rate);
  const NljConst c = make_const(ntrain, B, numel, w_shape, beta_shape, beta_rate);
  int first = 0; int64_t poff = 0; bool first_pass = true;
  do {
    const int count = nparams - first < SEG_MAX ? nparams - first : SEG_MAX;
    SegTable t; int rc = fill_segs(&t, params_host, numels_host, first, count); if (rc) return rc;
    const long long work = (first_pass ? numel : 0) > t.off[count] ? (first_pass ? numel : 0) : t.off[count];
    int64_t blocks = ceil_div64(work > 0 ? work : 1, 256 * 4); const int64_t cap = (int64_t)arpde_num_sms() * 8;
    if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
    nlj_bwd<<<(unsigned)blocks, 256, 0, stream>>>(output, target, numel, t, cw, scratch, log_beta, c, grad_loss,
                                                   first_pass ? grad_output : nullptr, first_pass ? grad_target : nullptr,
                                                   first_pass ? grad_log_beta : nullptr,
                                                   grad_params_flat ? grad_params_flat + poff : nullptr);
    ARPDE_LAUNCH_CHECK();
    poff += t.off[count]; first += count; first_pass = false;
  } while (first < nparams);
  return ARPDE_OK;
}
