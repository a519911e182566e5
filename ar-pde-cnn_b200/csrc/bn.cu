// bn.cu -- BatchNorm bookkeeping around the fused conv prologue.
//
// The reference applies nn.BatchNorm{1,2}d (eps 1e-5, momentum 0.1, affine) + ReLU in front of every
// convolution except the stem (denseEDcirc2d.py:58-62, :96-107, :169-187, :250-251).  Here the
// normalisation is folded into one per-channel (scale, shift) pair consumed by conv_fwd while it
// stages its input:   relu(bn(x)) == relu(scale*x + shift),
//   eval : scale = gamma / sqrt(running_var + eps),  shift = beta - running_mean * scale
//   train: batch mean / biased variance from the (sum, sum of squares) the producing conv left in
//          `stats`; running stats updated with the unbiased variance, num_batches_tracked += 1.
#include "common.cuh"
#include "bn.cuh"
#include "../../include/arpde.h"

__global__ void bn_fold_eval_kernel(BnTable t, float* __restrict__ scale, float* __restrict__ shift, long long ss_gs, float eps) {
  const BnDesc d = t.d[blockIdx.x];
  const int g = blockIdx.y;
  for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
    const float ga = d.gamma[(long long)g * d.g_gs + c], be = d.beta[(long long)g * d.b_gs + c];
    const float inv = 1.0f / sqrtf(d.rvar[c] + eps);
    const float sc = ga * inv;
    scale[(long long)g * ss_gs + d.off + c] = sc;
    shift[(long long)g * ss_gs + d.off + c] = fmaf(-d.rmean[c], sc, be);
  }
}

int arpde_bn_fold_eval_launch(const BnTable& t, int groups, float* scale, float* shift, long long ss_gs, float eps, cudaStream_t stream) {
  if (t.n == 0) return ARPDE_OK;
  bn_fold_eval_kernel<<<dim3(t.n, groups), 128, 0, stream>>>(t, scale, shift, ss_gs, eps);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

__global__ void bn_finalize_train_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                         long long* __restrict__ nbt, float momentum, float eps, float* __restrict__ scale,
                                         float* __restrict__ shift, float* __restrict__ save_mean, float* __restrict__ save_invstd, int C) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const double mean = stats[2 * c] / count;
    double var = stats[2 * c + 1] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = fmaf(-(float)mean, sc, beta[c]);
    if (save_mean) { save_mean[c] = (float)mean; save_invstd[c] = invstd; }
    if (rmean) {
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
    }
  }
  if (nbt && blockIdx.x == 0 && threadIdx.x == 0) nbt[0] += 1;
}

int arpde_bn_finalize_train_launch(const double* stats, double count, const float* gamma, const float* beta, float* rmean,
                                   float* rvar, long long* nbt, float momentum, float eps, float* scale, float* shift,
                                   float* save_mean, float* save_invstd, int C, cudaStream_t stream) {
  bn_finalize_train_kernel<<<(C + 127) / 128, 128, 0, stream>>>(stats, count, gamma, beta, rmean, rvar, nbt, momentum, eps,
                                                              scale, shift, save_mean, save_invstd, C);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// per-channel (sum, sum of squares) of channels [0,C) of x [N, ctot, HW] -- used for standalone BN modules
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long long x_bs, int N, int C, long long HW,
                                                       double* __restrict__ stats, int chunks) {
  const int c = blockIdx.x / chunks, ch = blockIdx.x - c * chunks;
  const long long total = (long long)N * HW, per = (total + chunks - 1) / chunks;
  const long long lo = ch * per, hi = lo + per < total ? lo + per : total;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const long long n = i / HW, r = i - n * HW;
    const float v = x[n * x_bs + (long long)c * HW + r];
    s1 += v; s2 = fmaf(v, v, s2);
  }
  const double d1 = block_sum_d((double)s1);
  const double d2 = block_sum_d((double)s2);
  if (threadIdx.x == 0) { atomicAdd(stats + 2 * c, d1); atomicAdd(stats + 2 * c + 1, d2); }
}

// vectorised variant (16-byte aligned rows): CTA (chunk, channel) sums `per4` consecutive float4 of the channel's N x HW/4 elements;
// (sample, offset) advance incrementally.  The scalar kernel above took 39 us for a 2 MB dense-layer output (8 CTAs per channel, a
// 64-bit division per element); 16 such launches per training step were 8% of it.
__global__ void __launch_bounds__(256) bn_stats_v_kernel(const float* __restrict__ x, long long x_bs, int N, int HW4, double* __restrict__ stats,
                                                         int per4) {
  const int c = blockIdx.y;
  const long long total4 = (long long)N * HW4;
  const long long lo = (long long)blockIdx.x * per4, hi = lo + per4 < total4 ? lo + per4 : total4;
  long long i4 = lo + threadIdx.x;
  int n = (int)(i4 / HW4), r4 = (int)(i4 - (long long)n * HW4);
  float s1 = 0.f, s2 = 0.f;
  const float* xc = x + (long long)c * HW4 * 4;
  for (; i4 < hi; i4 += 256) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(xc + n * x_bs) + r4);
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
    r4 += 256;
    while (r4 >= HW4) { r4 -= HW4; ++n; }
  }
  const double d1 = block_sum_d((double)s1);
  const double d2 = block_sum_d((double)s2);
  if (threadIdx.x == 0) { atomicAdd(stats + 2 * c, d1); atomicAdd(stats + 2 * c + 1, d2); }
}

extern "C" int arpde_bn_stats(const float* x, int64_t x_bstride, int N, int C, int64_t HW, double* stats, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(x && stats && N > 0 && C > 0 && HW > 0, "bn_stats: bad args");
  ARPDE_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * C, stream));
  const long long total = (long long)N * HW;
  if ((HW & 3) == 0 && (x_bstride & 3) == 0 && (((uintptr_t)x) & 15) == 0 && HW / 4 < (1ll << 30)) {
    const long long total4 = total / 4;
    const int per4 = 1024;                             // 4 float4 per thread: short fp32 partial sums, then fp64
    const long long chunks = (total4 + per4 - 1) / per4;
    ARPDE_CHECK_ARG(chunks < (1ll << 31) && C <= 65535, "bn_stats: tensor too large");
    bn_stats_v_kernel<<<dim3((unsigned)chunks, (unsigned)C), 256, 0, stream>>>(x, x_bstride, N, (int)(HW / 4), stats, per4);
    ARPDE_LAUNCH_CHECK();
    return ARPDE_OK;
  }
  int chunks = (int)ceil_div64(total, 256 * 64); if (chunks < 1) chunks = 1; if (chunks > 64) chunks = 64;
  bn_stats_kernel<<<C * chunks, 256, 0, stream>>>(x, x_bstride, N, C, HW, stats, chunks);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

extern "C" int arpde_bn_finalize_train(const double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                                       float* running_var, int64_t* num_batches_tracked, double momentum, double eps, float* scale,
                                       float* shift, float* save_mean, float* save_invstd, int C, arpde_stream_t stream) {
  ARPDE_CHECK_ARG(stats && gamma && beta && scale && shift && C > 0 && count > 0, "bn_finalize_train: bad args");
  ARPDE_CHECK_ARG((save_mean == nullptr) == (save_invstd == nullptr), "bn_finalize_train: save_mean/save_invstd go together");
  return arpde_bn_finalize_train_launch(stats, count, gamma, beta, running_mean, running_var, (long long*)num_batches_tracked,
                                        (float)momentum, (float)eps, scale, shift, save_mean, save_invstd, C, (cudaStream_t)stream);
}

extern "C" int arpde_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                  double eps, float* scale, float* shift, int C, arpde_stream_t stream) {
  ARPDE_CHECK_ARG(gamma && beta && running_mean && running_var && scale && shift && C > 0, "bn_fold_eval: bad args");
  BnTable t; t.n = 1;
  t.d[0].gamma = gamma; t.d[0].beta = beta; t.d[0].rmean = running_mean; t.d[0].rvar = running_var;
  t.d[0].g_gs = 0; t.d[0].b_gs = 0; t.d[0].C = C; t.d[0].off = 0;
  return arpde_bn_fold_eval_launch(t, 1, scale, shift, 0, (float)eps, (cudaStream_t)stream);
}

// standalone elementwise: y = act(scale[c]*x + shift[c]) with optional nearest x2 -- the unfused forward of
// a lone BatchNorm / ReLU / UpsamplingNearest module (DenseED.forward_test walks modules one by one).
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, float* __restrict__ y, int N, int C, int H,
                                                         int W, int up_h, int up_w, int act) {
  const int Ho = H << up_h, Wo = W << up_w;
  const long long total = (long long)N * C * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo); long long t = i / Wo;
    const int oy = (int)(t % Ho); t /= Ho;
    const int c = (int)(t % C); const long long n = t / C;
    float v = x[n * x_bs + ((long long)c * H + (oy >> up_h)) * W + (ox >> up_w)];
    if (scale) v = fmaf(v, scale[c], shift[c]);
    switch (act) {
      case 1: v = tanhf(v); break;
      case 2: v = fmaxf(v, 0.f); break;
      case 3: v = v > 0.f ? v : 0.01f * v; break;
      case 4: v = 1.f / (1.f + expf(-v)); break;
      case 5: v = 4.f * v > 20.f ? v : 0.25f * log1pf(expf(4.f * v)); break;
      default: break;
    }
    y[i] = v;
  }
}

extern "C" int arpde_affine_act(const float* x, int64_t x_bstride, const float* scale, const float* shift, float* y, int N, int C,
                                int H, int W, int up_h, int up_w, int act, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(x && y && N > 0 && C > 0 && H > 0 && W > 0, "affine_act: bad args");
  ARPDE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "affine_act: scale/shift go together");
  const long long total = (long long)N * C * (H << up_h) * (W << up_w);
  int64_t blocks = ceil_div64(total, 256); const int64_t cap = (int64_t)arpde_num_sms() * 16; if (blocks > cap) blocks = cap;
  affine_act_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, x_bstride, scale, shift, y, N, C, H, W, up_h, up_w, act);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
