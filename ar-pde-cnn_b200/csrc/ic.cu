// ic.cu -- on-device synthesis of the periodic random initial conditions of the 2D Burgers data set.
//
// Replaces InitPeriodicCond2d.__getitem__ (2D-Burgers-SWAG/utils/burgerLoader2D.py:99-112): per sample a truncated Fourier series
//   f[comp][p] = sum_m lam[0][comp][m] cos(2 pi (kx_m x_p + ky_m y_p)) + lam[1][comp][m] sin(...),   u = 2 f / max_p |f| + c[comp]
// with (2*order+1)^2 modes, kx = m % (2*order+1) - order, ky = m / (2*order+1) - order, x_p = (p % n) / n, y_p = (p / n) / n.
// The reference evaluates 2 x M x n^2 cosines and sines in numpy per sample on the host (6.4 ms per 64x64 sample, 252 ms at 256x256:
// SURVEY.md 8f) -- three orders of magnitude slower than the training step consumes them.  Here the random coefficients still come
// from the host (numpy's per-index seeded stream is part of the data-set definition); everything else runs in one launch, in fp64 like
// the reference, rounded to fp32 once at the end.  The phase (kx*j + ky*i) mod n is an integer, so cos/sin come from an n-entry table
// (cospi/sinpi of exact rationals) instead of range-reducing 2 pi (kx x + ky y).
#include "common.cuh"
#include "../../include/arpde.h"

#define IC_MAX_MODES 169            // order <= 6
#define IC_MAX_N 1024

// one CTA per (sample, component): pass 1 max |f|, pass 2 recompute and write
__global__ void __launch_bounds__(256) fourier_ic2d_kernel(const double* __restrict__ lam, const double* __restrict__ cc, float* __restrict__ out,
                                                           int n, int order) {
  extern __shared__ double sm[];
  const int K = 2 * order + 1, M = K * K;
  double* tcos = sm;                  // [n]
  double* tsin = sm + n;              // [n]
  double* la = sm + 2 * n;            // [M] cos coefficients
  double* lb = la + M;                // [M] sin coefficients
  const int b = blockIdx.x >> 1, comp = blockIdx.x & 1;
  for (int t = threadIdx.x; t < n; t += blockDim.x) { tcos[t] = cospi(2.0 * t / n); tsin[t] = sinpi(2.0 * t / n); }
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    la[m] = lam[((long long)b * 4 + comp) * M + m];            // lam [B][2 (cos | sin)][2 (comp)][M]
    lb[m] = lam[((long long)b * 4 + 2 + comp) * M + m];
  }
  __syncthreads();
  const int P = n * n;
  double vmax = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    __shared__ double s_red[8];
    double scale = 0.0, shift = 0.0;
    if (pass == 1) {
      // block max of |f|
      double v = vmax;
      for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
      __syncthreads();
      v = s_red[0];
      for (int w = 1; w < 8; ++w) v = fmax(v, s_red[w]);
      scale = 2.0 / v; shift = cc[b * 2 + comp];
    }
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      const int i = p / n, j = p - i * n;
      double f = 0.0;
      // phase index advances by j per kx step and by i per ky step (mod n)
      int ty = ((-order * i) % n + n) % n;
      for (int ky = 0; ky < K; ++ky) {
        int t = (ty + ((-order * j) % n + n)) % n;
        for (int kx = 0; kx < K; ++kx) {
          const int m = ky * K + kx;
          f = fma(la[m], tcos[t], f);
          f = fma(lb[m], tsin[t], f);
          t += j; if (t >= n) t -= n;
        }
        ty += i; if (ty >= n) ty -= n;
      }
      if (pass == 0) vmax = fmax(vmax, fabs(f));
      else out[((long long)b * 2 + comp) * P + p] = (float)(f * scale + shift);
    }
  }
}

extern "C" int arpde_fourier_ic2d(const double* lam, const double* c, float* out, int B, int ncells, int order, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(lam && c && out && B >= 0, "fourier_ic2d: null pointer");
  ARPDE_CHECK_ARG(ncells > 0 && ncells <= IC_MAX_N && order >= 0 && (2 * order + 1) * (2 * order + 1) <= IC_MAX_MODES, "fourier_ic2d: bad ncells/order");
  if (B == 0) return ARPDE_OK;
  const int M = (2 * order + 1) * (2 * order + 1);
  const size_t smem = sizeof(double) * (2 * (size_t)ncells + 2 * (size_t)M);
  fourier_ic2d_kernel<<<2 * B, 256, smem, (cudaStream_t)stream_>>>(lam, c, out, ncells, order);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
