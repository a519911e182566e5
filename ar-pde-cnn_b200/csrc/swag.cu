// swag.cu -- batched SWAG posterior sampling and predictive moments.
//
// Replaces SwagNN.sample (reference */nn/swag.py:80-150; ~766 ATen ops per call, one Python loop
// iteration per parameter tensor) by one launch over the flat parameter vector x S samples, and the
// numpy post-processing of post/plotBARContour.py:85-104 (2D) / 1D :123,136-139 by one reduction.
//
// Per tensor t and element i (flat index p), sample s:
//   var = max(sq_mean - mean^2, eps)
//   diag: w = mean + scale*sqrt(var)*z1[s,p]
//   full: D[k] = (C[k,i]==0 ? 0 : C[k,i]-mean)      (zero rows = empty slots, left uncentred: swag.py:137)
//         w = mean + sqrt(1/2) * ( scale*sqrt(var)*z1[s,p] + scale/sqrt(K-1) * sum_k D[k]*z2[s,t,k] )
// z2 is drawn once per parameter TENSOR (swag.py:139), K = max_models (constructor value, not n_models).
#include "common.cuh"
#include "../../include/arpde.h"

#define SWAG_SEG_MAX 80
struct SwagTable {
  const float* mean[SWAG_SEG_MAX];
  const float* sq[SWAG_SEG_MAX];
  const float* cov[SWAG_SEG_MAX];
  long long off[SWAG_SEG_MAX + 1];
  int n;
};

__global__ void __launch_bounds__(256) swag_sample_kernel(SwagTable t, const float* __restrict__ z1, const float* __restrict__ z2,
                                                          float* __restrict__ out, int K, float scale, float cov_scale,
                                                          float eps, int diag) {
  const long long P = t.off[t.n];
  const int s = blockIdx.y;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = t.n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (t.off[mid] <= p) lo = mid; else hi = mid - 1; }
    const long long i = p - t.off[lo], numel = t.off[lo + 1] - t.off[lo];
    const float m = t.mean[lo][i];
    // two roundings, like torch's `sq_mean - mean ** 2` (an FMA here changes the result: the difference cancels)
    const float var = fmaxf(__fsub_rn(t.sq[lo][i], __fmul_rn(m, m)), eps);
    const float sds = scale * sqrtf(var) * z1[(long long)s * P + p];
    float w;
    if (diag) {
      w = m + sds;
    } else {
      const float* c = t.cov[lo] + i;
      const float* z = z2 + ((long long)s * t.n + lo) * K;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const float ck = c[(long long)k * numel];
        acc = fmaf(ck == 0.f ? 0.f : ck - m, z[k], acc);
      }
      w = m + 0.70710678118654752440f * (sds + cov_scale * acc);
    }
    out[(long long)s * P + p] = w;
  }
}

extern "C" int arpde_swag_sample(const float* const* mean_host, const float* const* sq_mean_host, const float* const* cov_host,
                                 const int64_t* numels_host, int n_tensors, const float* z1, const float* z2, float* out, int S,
                                 int K, int max_models, double scale, double eps, int diag, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ARPDE_CHECK_ARG(mean_host && sq_mean_host && numels_host && z1 && out, "swag_sample: null pointer");
  ARPDE_CHECK_ARG(diag || (cov_host && z2 && K > 0 && max_models > 1), "swag_sample: full covariance needs cov, z2, K>0, max_models>1");
  ARPDE_CHECK_ARG(n_tensors > 0 && n_tensors <= SWAG_SEG_MAX, "swag_sample: %d parameter tensors (max %d)", n_tensors, SWAG_SEG_MAX);
  ARPDE_CHECK_ARG(S > 0, "swag_sample: S must be > 0");
  SwagTable t; t.n = n_tensors; t.off[0] = 0;
  for (int i = 0; i < n_tensors; ++i) {
    t.mean[i] = mean_host[i]; t.sq[i] = sq_mean_host[i]; t.cov[i] = diag ? nullptr : cov_host[i];
    ARPDE_CHECK_ARG(t.mean[i] && t.sq[i] && (diag || t.cov[i]) && numels_host[i] > 0, "swag_sample: bad tensor %d", i);
    t.off[i + 1] = t.off[i] + numels_host[i];
  }
  const long long P = t.off[n_tensors];
  int64_t bx = ceil_div64(P, 256); if (bx > 4096) bx = 4096;
  const float cov_scale = diag ? 0.f : (float)(scale / sqrt((double)max_models - 1.0));
  swag_sample_kernel<<<dim3((unsigned)bx, S), 256, 0, stream>>>(t, z1, z2, out, K, (float)scale, cov_scale, (float)eps, diag);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// predictive moments over S weight samples: traj [S*n_ic, inner] sample-major, rows bstride apart; mean/var [n_ic, inner]
//   var = noise + E_s[y^2] - (E_s[y])^2, noise = 1/mean(beta) (2D scripts) or mean(1/beta) (1D scripts)
__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ traj, long long bstride, long long inner, int n_ic,
                                                      int S, float noise, float* __restrict__ mean, float* __restrict__ var) {
  const long long total = (long long)n_ic * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / inner, j = i - n * inner;
    float m = 0.f, m2 = 0.f;                 // Welford: E[y^2]-E[y]^2 without the cancellation
    for (int s = 0; s < S; ++s) {
      const float v = traj[((long long)s * n_ic + n) * bstride + j];
      const float d = v - m;
      m += d / (float)(s + 1);
      m2 = fmaf(d, v - m, m2);
    }
    mean[i] = m;
    if (var) var[i] = noise + m2 / (float)S;
  }
}

// vector form: a thread owns four consecutive values, the loads of six samples are issued before the first is used (the scalar kernel had
// one 4-byte load in flight per thread: 2 TB/s over the 6.5 GB trajectory tensor of the C4 job).  Same Welford sequence per element.
__global__ void __launch_bounds__(256) moments_vec4_kernel(const float* __restrict__ traj, long long bstride, long long inner4, int n_ic,
                                                           int S, float noise, float4* __restrict__ mean, float4* __restrict__ var) {
  constexpr int U = 6;
  const long long total = (long long)n_ic * inner4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / inner4, j = i - n * inner4;
    float m[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s0 = 0; s0 < S; s0 += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (s0 + u < S) v[u] = __ldcs(reinterpret_cast<const float4*>(traj + ((long long)(s0 + u) * n_ic + n) * bstride) + j);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (s0 + u < S) {
          const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          const float cnt = (float)(s0 + u + 1);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float d = x[c] - m[c];
            m[c] += d / cnt;
            m2[c] = fmaf(d, x[c] - m[c], m2[c]);
          }
        }
    }
    mean[i] = make_float4(m[0], m[1], m[2], m[3]);
    if (var) var[i] = make_float4(noise + m2[0] / (float)S, noise + m2[1] / (float)S, noise + m2[2] / (float)S, noise + m2[3] / (float)S);
  }
}

extern "C" int arpde_predictive_moments(const float* traj, int64_t bstride, int64_t inner, int n_ic, int S, double noise, float* mean,
                                        float* var, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(traj && mean && inner > 0 && bstride >= inner && n_ic > 0 && S > 0, "predictive_moments: bad args");
  const int64_t cap = (int64_t)arpde_num_sms() * 16;
  const bool vec = inner % 4 == 0 && bstride % 4 == 0 && ((reinterpret_cast<uintptr_t>(traj) | reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(var)) & 15) == 0;
  if (vec) {
    int64_t blocks = ceil_div64((int64_t)n_ic * (inner / 4), 256); if (blocks > cap) blocks = cap;
    moments_vec4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(traj, bstride, inner / 4, n_ic, S, (float)noise, reinterpret_cast<float4*>(mean),
                                                                              reinterpret_cast<float4*>(var));
    ARPDE_LAUNCH_CHECK();
    return ARPDE_OK;
  }
  int64_t blocks = ceil_div64((int64_t)n_ic * inner, 256); if (blocks > cap) blocks = cap;
  moments_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(traj, bstride, inner, n_ic, S, (float)noise, mean, var);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// ------------------------------------------------------------------ error metrics of a rollout against a target trajectory
// testMSE / testBayesianMSE of 2D-Burgers-SWAG/post/plotMSE.py:78-85,133-137 (and the 1D twins): per (case n, saved time t)
//   mse = mean_f (pred - target)^2,   ese = ( sum_f pred^2 / 2 / points  -  sum_f target^2 / 2 / points )^2
// over the F = components x grid values of one state; pred is the (sample-mean) prediction.  The reference moves the whole
// [N, S, T+1, 2, H, W] prediction tensor to the host first; here the rows are reduced where the rollout left them (row strides
// select every test_every-th step), fp64 accumulation, one CTA per (n, t).
__global__ void __launch_bounds__(256) error_metrics_kernel(const float* __restrict__ pred, long long pred_ns, long long pred_ts,
                                                            const float* __restrict__ target, long long tgt_ns, long long tgt_ts, int Tp,
                                                            long long F, double inv_points, double* __restrict__ mse, double* __restrict__ ese) {
  const int n = blockIdx.x / Tp, t = blockIdx.x - n * Tp;
  const float* p = pred + n * pred_ns + t * pred_ts;
  const float* q = target + n * tgt_ns + t * tgt_ts;
  double s_d = 0.0, s_p = 0.0, s_q = 0.0;
  for (long long i = threadIdx.x; i < F; i += blockDim.x) {
    const double a = (double)p[i], b = (double)q[i];
    s_d += (a - b) * (a - b); s_p += a * a; s_q += b * b;
  }
  s_d = block_sum_d(s_d);
  s_p = block_sum_d(s_p);
  s_q = block_sum_d(s_q);
  if (threadIdx.x == 0) {
    mse[blockIdx.x] = s_d / (double)F;
    const double e = 0.5 * s_p * inv_points - 0.5 * s_q * inv_points;
    ese[blockIdx.x] = e * e;
  }
}

extern "C" int arpde_error_metrics(const float* pred, int64_t pred_nstride, int64_t pred_tstride, const float* target, int64_t tgt_nstride,
                                   int64_t tgt_tstride, int N, int Tp, int64_t F, int64_t points, double* mse, double* ese,
                                   arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(pred && target && mse && ese && N >= 0 && Tp > 0 && F > 0 && points > 0, "error_metrics: bad args");
  if (N == 0) return ARPDE_OK;
  ARPDE_CHECK_ARG((long long)N * Tp < (1ll << 31), "error_metrics: too many rows");
  error_metrics_kernel<<<(unsigned)(N * Tp), 256, 0, (cudaStream_t)stream_>>>(pred, pred_nstride, pred_tstride, target, tgt_nstride, tgt_tstride, Tp, F,
                                                                              1.0 / (double)points, mse, ese);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
