// adam.cu -- one-launch Adam step over a list of parameter tensors.
//
// Replaces the optimizer.step() of the reference's training loops (torch.optim.Adam, 2D-Burgers-SWAG/main.py:221-222,84 and the 1D
// twins): for the default 2D net that is 25 tensors / 72 k parameters, which torch's multi-tensor path updates with ~14 launches.
// Same arithmetic as torch's (non-amsgrad, non-maximize) update, in fp32 with the bias corrections computed on the host in double:
//   g' = g + wd * p;  m = m + (1 - b1) (g' - m);  v = b2 v + (1 - b2) g' g';  p = p - step_size * m / (sqrt(v) / sqrt(bc2) + eps)
// Optimizer state stays torch.optim.Adam's (exp_avg, exp_avg_sq, step), so checkpoints interchange (nn/swag.py:207-208 saves it).
#include "common.cuh"
#include "../../include/arpde.h"

#define ADAM_MAX_TENSORS 64

struct AdamTable {
  float* p[ADAM_MAX_TENSORS]; const float* g[ADAM_MAX_TENSORS]; float* m[ADAM_MAX_TENSORS]; float* v[ADAM_MAX_TENSORS];
  int numel[ADAM_MAX_TENSORS];
  int first_block[ADAM_MAX_TENSORS + 1];        // block range of tensor t: [first_block[t], first_block[t+1])
  float step_size[ADAM_MAX_TENSORS], bc2_sqrt[ADAM_MAX_TENSORS], wd[ADAM_MAX_TENSORS];
  int n;
};

#define ADAM_CHUNK 1024

__global__ void __launch_bounds__(256) adam_step_kernel(const __grid_constant__ AdamTable t, float beta2, float w1, float w2, float eps) {
  int ti = 0;
  while (ti + 1 < t.n && (int)blockIdx.x >= t.first_block[ti + 1]) ++ti;
  const int base = ((int)blockIdx.x - t.first_block[ti]) * ADAM_CHUNK;
  float* __restrict__ p = t.p[ti]; const float* __restrict__ g = t.g[ti]; float* __restrict__ m = t.m[ti]; float* __restrict__ v = t.v[ti];
  const float ss = t.step_size[ti], bc2 = t.bc2_sqrt[ti], wd = t.wd[ti];
  for (int i = base + threadIdx.x; i < min(base + ADAM_CHUNK, t.numel[ti]); i += 256) {
    float gi = g[i];
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = m[i] + w1 * (gi - m[i]);                 // Tensor.lerp_(grad, 1 - beta1), weight < 0.5
    const float vi = v[i] * beta2 + (w2 * gi) * gi;           // mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2 + eps;
    p[i] = pi - ss * (mi / denom);                            // addcdiv_(exp_avg, denom, value = -step_size)
  }
}

extern "C" int arpde_adam_step(void* const* params, void* const* grads, void* const* exp_avg, void* const* exp_avg_sq, const int64_t* numel,
                               const double* step_size, const double* bias_correction2_sqrt, const double* weight_decay, int n,
                               double beta1, double beta2, double eps, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(n >= 0 && (n == 0 || (params && grads && exp_avg && exp_avg_sq && numel && step_size && bias_correction2_sqrt)), "adam_step: null pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  for (int t0 = 0; t0 < n; t0 += ADAM_MAX_TENSORS) {
    AdamTable t; memset(&t, 0, sizeof(t));
    t.n = n - t0 < ADAM_MAX_TENSORS ? n - t0 : ADAM_MAX_TENSORS;
    int blocks = 0;
    for (int i = 0; i < t.n; ++i) {
      const int k = t0 + i;
      ARPDE_CHECK_ARG(params[k] && grads[k] && exp_avg[k] && exp_avg_sq[k] && numel[k] >= 0 && numel[k] < (1ll << 31), "adam_step: bad tensor %d", k);
      t.p[i] = (float*)params[k]; t.g[i] = (const float*)grads[k]; t.m[i] = (float*)exp_avg[k]; t.v[i] = (float*)exp_avg_sq[k];
      t.numel[i] = (int)numel[k]; t.first_block[i] = blocks;
      blocks += (int)((numel[k] + ADAM_CHUNK - 1) / ADAM_CHUNK);
      t.step_size[i] = (float)step_size[k]; t.bc2_sqrt[i] = (float)bias_correction2_sqrt[k]; t.wd[i] = weight_decay ? (float)weight_decay[k] : 0.f;
    }
    t.first_block[t.n] = blocks;
    if (blocks == 0) continue;
    // 1 - beta in double, rounded once (torch passes the Python double 1 - beta; 1.f - 0.999f would be 4.7e-5 off)
    adam_step_kernel<<<blocks, 256, 0, stream>>>(t, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps);
    ARPDE_LAUNCH_CHECK();
  }
  return ARPDE_OK;
}
