// conv_bwd.cu -- backward of the fused circular-conv op (what torch autograd derives through
// F.pad(circular)+conv, BatchNorm(train or eval)+ReLU and nearest upsampling in the reference,
// triggered by loss.backward() at 2D-Burgers-SWAG/main.py:81 and the 1D twins).
//
//   forward op:  a = up2?( relu( scale*x + shift ) ),  O = conv_circ(a, w, stride)
//   given dO:    dW   = sum_{n,pix} dO (x) a                         (wgrad_kernel, split over samples/tiles,
//                                                                      fp32 per-CTA partials -> fp64 atomics)
//                da   = conv_circ( zero_insert?(dO), flip(w)^T )     (conv_fwd_kernel, stride-1, pad k-1-p)
//                dz   = pool2x2?(da) * [z > 0]                       (bn_relu_bwd_reduce: also S1=sum dz, S2=sum dz*xhat)
//                dx  += scale * (dz - S1/M - xhat*S2/M)   (train)    (bn_bwd_apply; eval: scale*dz)
//                dgamma = S2, dbeta = S1
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"

// ------------------------------------------------------------------ weight flip + transpose
__global__ void weight_flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int cout, int cin, int kh, int kw) {
  const int KK = kh * kw, total = cout * cin * KK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i % KK, ci = (i / KK) % cin, co = i / (KK * cin);
    const int ky = kk / kw, kx = kk - ky * kw;
    wt[((long long)ci * cout + co) * KK + (kh - 1 - ky) * kw + (kw - 1 - kx)] = w[i];
  }
}

int arpde_weight_flip_transpose_launch(const float* w, float* wt, int cout, int cin, int kh, int kw, cudaStream_t stream) {
  const int total = cout * cin * kh * kw;
  weight_flip_transpose_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, wt, cout, cin, kh, kw);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// ------------------------------------------------------------------ BN + ReLU (+ nearest x2) backward, pass 1
// da: [N, C, H<<up_h, W<<up_w] compact; x: forward input (pre-BN) with batch stride; dz out: [N, C, H, W] compact.
// has_bn = 0: dz = pooled da (no mask), sums untouched.
__global__ void __launch_bounds__(256) bn_relu_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ x, long long x_bs,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 float* __restrict__ dz, double* __restrict__ sums, int N, int C, int H,
                                                                 int W, int up_h, int up_w, int chunks) {
  const int c = blockIdx.x / chunks, ch = blockIdx.x - c * chunks;
  const long long HW = (long long)H * W, total = (long long)N * HW, per = (total + chunks - 1) / chunks;
  const long long lo = ch * per, hi = lo + per < total ? lo + per : total;
  const int Hc = H << up_h, Wc = W << up_w;
  const float sc = scale ? scale[c] : 1.f, sh = scale ? shift[c] : 0.f;
  const float mu = mean ? mean[c] : 0.f, is = invstd ? invstd[c] : 1.f;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const long long n = i / HW; const int r = (int)(i - n * HW), y = r / W, xx = r - y * W;
    const float* p = da + ((n * C + c) * Hc + ((long long)y << up_h)) * Wc + (xx << up_w);
    float g = p[0];
    if (up_w) g += p[1];
    if (up_h) { g += p[Wc]; if (up_w) g += p[Wc + 1]; }
    if (scale) {
      const float xv = x[n * x_bs + (long long)c * HW + r];
      if (fmaf(xv, sc, sh) <= 0.f) g = 0.f;
      s1 += g; s2 = fmaf(g, (xv - mu) * is, s2);
    }
    dz[(n * C + c) * HW + r] = g;
  }
  if (scale) {
    const double d1 = block_sum_d((double)s1);
    const double d2 = block_sum_d((double)s2);
    if (threadIdx.x == 0) { atomicAdd(sums + 2 * c, d1); atomicAdd(sums + 2 * c + 1, d2); }
  }
}

// pass 2: gx[n, c] (+)= scale*(dz - S1/M - xhat*S2/M) (train) | scale*dz (eval) ; dgamma = S2, dbeta = S1
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ x, long long x_bs,
                                                           const float* __restrict__ scale, const float* __restrict__ mean,
                                                           const float* __restrict__ invstd, const double* __restrict__ sums,
                                                           float* __restrict__ gx, long long gx_bs, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int N, int C, long long HW, int training,
                                                           int accumulate, int dp_world) {
  const long long total = (long long)N * C * HW;
  const double invM = 1.0 / ((double)N * (double)HW * (double)dp_world);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i % HW; const int c = (int)((i / HW) % C); const long long n = i / (HW * C);
    float g = dz[i];
    float v;
    if (scale) {
      if (training) {
        const float xh = (x[n * x_bs + c * HW + r] - mean[c]) * invstd[c];
        g = g - (float)(sums[2 * c] * invM) - xh * (float)(sums[2 * c + 1] * invM);
      }
      v = scale[c] * g;
    } else {
      v = g;
    }
    float* o = gx + n * gx_bs + c * HW + r;
    *o = accumulate ? *o + v : v;
  }
  if (scale && dgamma && blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) { dgamma[c] = (float)sums[2 * c + 1]; dbeta[c] = (float)sums[2 * c]; }
}

// ---- vectorised variants (16-byte aligned tensors, W % 4 == 0): one CTA per (channel, group of samples), float4 along the row,
// no integer division per element, and no dz round trip -- the apply pass recomputes the masked (pooled) gradient from da and x.
// Traffic per element: reduce reads da + x; apply reads da + x and writes gx (was: +1 write and +1 read of dz).  At the training batch
// (N = 128) the old pair ran at ~2 TB/s on the 96-channel 64x64 tensor; see DESIGN.md 4.5.
template <int UP>
__device__ __forceinline__ float4 bnb_load_g(const float* __restrict__ da_nc, int r, int W) {
  if (UP == 0) return __ldg(reinterpret_cast<const float4*>(da_nc + r));
  const int y = r / W, xx = r - y * W;                 // W % 4 == 0: the four elements share a row
  const int Wc = 2 * W;
  const float* p = da_nc + (size_t)(2 * y) * Wc + 2 * xx;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(p)), a1 = __ldg(reinterpret_cast<const float4*>(p + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p + Wc)), b1 = __ldg(reinterpret_cast<const float4*>(p + Wc + 4));
  // same summation order as the scalar kernel: p[0] + p[1] + p[Wc] + p[Wc+1]
  return make_float4(((a0.x + a0.y) + b0.x) + b0.y, ((a0.z + a0.w) + b0.z) + b0.w, ((a1.x + a1.y) + b1.x) + b1.y, ((a1.z + a1.w) + b1.z) + b1.w);
}

template <int UP>
__global__ void __launch_bounds__(256) bn_relu_bwd_reduce_v_kernel(const float* __restrict__ da, const float* __restrict__ x, long long x_bs,
                                                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                   double* __restrict__ sums, int N, int C, int H, int W, int nper) {
  const int c = blockIdx.y;
  const int n_lo = blockIdx.x * nper, n_hi = min(N, n_lo + nper);
  const int HW = H * W, HW4 = HW >> 2;
  const long long dHW = (long long)HW << (2 * UP);
  const float sc = scale[c], sh = shift[c];
  const float mu = mean ? mean[c] : 0.f, is = invstd ? invstd[c] : 1.f;
  float s1 = 0.f, s2 = 0.f;
  for (int n = n_lo; n < n_hi; ++n) {
    const float* xn = x + n * x_bs + (long long)c * HW;
    const float* dn = da + ((long long)n * C + c) * dHW;
    for (int r4 = threadIdx.x; r4 < HW4; r4 += 256) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(xn + 4 * r4));
      const float4 g = bnb_load_g<UP>(dn, 4 * r4, W);
      const float xe[4] = {xv.x, xv.y, xv.z, xv.w}, ge[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gg = fmaf(xe[e], sc, sh) <= 0.f ? 0.f : ge[e];
        s1 += gg; s2 = fmaf(gg, (xe[e] - mu) * is, s2);
      }
    }
  }
  const double d1 = block_sum_d((double)s1);
  const double d2 = block_sum_d((double)s2);
  if (threadIdx.x == 0) { atomicAdd(sums + 2 * c, d1); atomicAdd(sums + 2 * c + 1, d2); }
}

template <int UP>
__global__ void __launch_bounds__(256) bn_bwd_apply_v_kernel(const float* __restrict__ da, const float* __restrict__ x, long long x_bs,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             const float* __restrict__ mean, const float* __restrict__ invstd,
                                                             const double* __restrict__ sums, float* __restrict__ gx, long long gx_bs,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, int N, int C, int H,
                                                             int W, int nper, int training, int accumulate, int dp_world) {
  const int c = blockIdx.y;
  const int n_lo = blockIdx.x * nper, n_hi = min(N, n_lo + nper);
  const int HW = H * W, HW4 = HW >> 2;
  const long long dHW = (long long)HW << (2 * UP);
  const bool bn = scale != nullptr;
  const float sc = bn ? scale[c] : 1.f, sh = bn ? shift[c] : 0.f;
  const float mu = (bn && mean) ? mean[c] : 0.f, is = (bn && invstd) ? invstd[c] : 1.f;
  const double invM = 1.0 / ((double)N * (double)HW * (double)dp_world);
  const float m1 = (bn && training) ? (float)(sums[2 * c] * invM) : 0.f, m2 = (bn && training) ? (float)(sums[2 * c + 1] * invM) : 0.f;
  for (int n = n_lo; n < n_hi; ++n) {
    const float* xn = x + n * x_bs + (long long)c * HW;
    const float* dn = da + ((long long)n * C + c) * dHW;
    float* on = gx + n * gx_bs + (long long)c * HW;
    for (int r4 = threadIdx.x; r4 < HW4; r4 += 256) {
      const float4 g = bnb_load_g<UP>(dn, 4 * r4, W);
      float ge[4] = {g.x, g.y, g.z, g.w};
      if (bn) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(xn + 4 * r4));
        const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float gg = fmaf(xe[e], sc, sh) <= 0.f ? 0.f : ge[e];
          if (training) { const float xh = (xe[e] - mu) * is; gg = gg - m1 - xh * m2; }
          ge[e] = sc * gg;
        }
      }
      float4* o = reinterpret_cast<float4*>(on + 4 * r4);
      if (accumulate) { const float4 p = *o; ge[0] = p.x + ge[0]; ge[1] = p.y + ge[1]; ge[2] = p.z + ge[2]; ge[3] = p.w + ge[3]; }
      *o = make_float4(ge[0], ge[1], ge[2], ge[3]);
    }
  }
  if (bn && dgamma && blockIdx.x == 0 && threadIdx.x == 0) { dgamma[c] = (float)sums[2 * c + 1]; dbeta[c] = (float)sums[2 * c]; }
}

// dgamma = sum dz*xhat, dbeta = sum dz of THIS rank's samples (data-parallel BatchNorm: written before the sums are all-reduced, the
// gradient all-reduce adds the ranks up)
__global__ void bn_dparam_kernel(const double* __restrict__ sums, float* __restrict__ dgamma, float* __restrict__ dbeta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { dgamma[c] = (float)sums[2 * c + 1]; dbeta[c] = (float)sums[2 * c]; }
}

// dp: optional data-parallel hook -- between the reduction and the apply pass the (sum dz, sum dz*xhat) pairs are all-reduced over the
// ranks (global-batch BatchNorm backward), the element count is N*HW*dp->world
int arpde_bn_relu_bwd_launch(const float* da, const float* x, long long x_bs, const float* scale, const float* shift, const float* mean,
                             const float* invstd, float* dz, double* sums, float* gx, long long gx_bs, float* dgamma, float* dbeta, int N,
                             int C, int H, int W, int up_h, int up_w, int training, int accumulate, cudaStream_t stream, const DpHook* dp) {
  const long long HW = (long long)H * W, total = (long long)N * HW;
  const bool sync = dp && dp->allreduce && dp->world > 1 && scale && training;
  const int dpw = sync ? dp->world : 1;
  auto sync_sums = [&]() -> int {
    if (!sync) return ARPDE_OK;
    if (dgamma) { bn_dparam_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, dgamma, dbeta, C); ARPDE_LAUNCH_CHECK(); }
    if (dp->allreduce(dp->user, sums, 2ll * C)) { arpde_set_error("bn backward: the all-reduce callback failed"); return ARPDE_ERR_CUDA; }
    return ARPDE_OK;
  };
  float* dgamma_k = sync ? nullptr : dgamma;      // the apply kernels write dgamma / dbeta from the (then global) sums otherwise
  if (scale) ARPDE_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream));
  // vector path: full 2D pooling (or none), rows of whole float4s, every base and stride 16-byte aligned; dz is not used
  const bool vec = up_h == up_w && (W & 3) == 0 && (x_bs & 3) == 0 && (gx_bs & 3) == 0 && (((uintptr_t)da | (uintptr_t)x | (uintptr_t)gx) & 15) == 0 &&
                   !arpde_dev_env("ARPDE_BN_BWD_SCALAR");
  if (vec) {
    // samples per CTA: enough CTAs to fill the machine a few times over, at least ~2k float4 per CTA
    int nper = 1;
    while ((long long)C * ((N + nper - 1) / nper) > 8ll * arpde_num_sms() && nper < N) ++nper;
    while (nper < N && (long long)nper * (HW >> 2) < 1024) ++nper;
    const dim3 grid((unsigned)((N + nper - 1) / nper), (unsigned)C);
    if (scale) {
      if (up_w) bn_relu_bwd_reduce_v_kernel<1><<<grid, 256, 0, stream>>>(da, x, x_bs, scale, shift, mean, invstd, sums, N, C, H, W, nper);
      else bn_relu_bwd_reduce_v_kernel<0><<<grid, 256, 0, stream>>>(da, x, x_bs, scale, shift, mean, invstd, sums, N, C, H, W, nper);
      ARPDE_LAUNCH_CHECK();
      if (int rc = sync_sums()) return rc;
    }
    if (up_w) bn_bwd_apply_v_kernel<1><<<grid, 256, 0, stream>>>(da, x, x_bs, scale, shift, mean, invstd, sums, gx, gx_bs, dgamma_k, dbeta, N, C, H, W, nper, training, accumulate, dpw);
    else bn_bwd_apply_v_kernel<0><<<grid, 256, 0, stream>>>(da, x, x_bs, scale, shift, mean, invstd, sums, gx, gx_bs, dgamma_k, dbeta, N, C, H, W, nper, training, accumulate, dpw);
    ARPDE_LAUNCH_CHECK();
    return ARPDE_OK;
  }
  int chunks = (int)ceil_div64(total, 256 * 32); if (chunks < 1) chunks = 1; if (chunks > 128) chunks = 128;
  bn_relu_bwd_reduce_kernel<<<C * chunks, 256, 0, stream>>>(da, x, x_bs, scale, shift, mean, invstd, dz, sums, N, C, H, W, up_h, up_w, chunks);
  ARPDE_LAUNCH_CHECK();
  if (int rc = sync_sums()) return rc;
  int64_t blocks = ceil_div64(total * C, 256 * 4); const int64_t cap = (int64_t)arpde_num_sms() * 16; if (blocks > cap) blocks = cap; if (blocks < 1) blocks = 1;
  bn_bwd_apply_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dz, x, x_bs, scale, mean, invstd, sums, gx, gx_bs, dgamma_k, dbeta, N, C, HW, training, accumulate, dpw);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// ------------------------------------------------------------------ output-activation backward (rare: out_activation != None)
__global__ void act_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, float* __restrict__ out, long long n, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i];
    float d = 1.f;
    switch (act) {
      case 1: d = 1.f - v * v; break;
      case 2: d = v > 0.f ? 1.f : 0.f; break;
      case 3: d = v > 0.f ? 1.f : 0.01f; break;
      case 4: d = v * (1.f - v); break;
      case 5: d = 1.f - expf(-4.f * v); break;      // y = log1p(exp(4x))/4  =>  dy/dx = sigmoid(4x) = 1 - exp(-4y)
      default: break;
    }
    out[i] = gy[i] * d;
  }
}

int arpde_act_bwd_launch(const float* gy, const float* y, float* out, long long n, int act, cudaStream_t stream) {
  int64_t blocks = ceil_div64(n, 256); if (blocks > 4096) blocks = 4096;
  act_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(gy, y, out, n, act);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// ------------------------------------------------------------------ weight gradient
struct WgradArgs {
  const float* x; long long x_bs; int cin, H, W, up, up_h, pad_y, pad_x;
  const float* scale; const float* shift; int relu_in;
  const float* dO; long long dO_bs; int cout, Ho, Wo;
  double* dW;
  int N, tiles_x, tiles_y, co_tiles, ci_chunks, TH, TW, pitch, plane, cic, ncg, dpitch, rgroups;
};

template <int KH, int KW, int S, int COT>
__global__ void __launch_bounds__(256) wgrad_kernel(WgradArgs a) {
  constexpr int KK = KH * KW, PXT = 4;
  constexpr int SEG = ((PXT - 1) * S + KW + 3) & ~3;
  extern __shared__ __align__(16) float smem[];
  const int PH = (a.TH - 1) * S + KH, PW = (a.TW - 1) * S + KW;
  const int TCO = a.ncg * COT;
  int* rowmap = reinterpret_cast<int*>(smem);
  int* colmap = rowmap + PH;
  float* patch = smem + ((PH + PW + 3) & ~3);
  float* dsm = patch + a.cic * a.plane;                  // [TCO][TH][dpitch]
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int ci_l = tid % a.cic, tcr = tid / a.cic;       // ci lane, (co group, row group)
  const int tc = tcr % a.ncg, rg = tcr / a.ncg;          // row group rg handles tile rows rg, rg + RG, ...
  const bool active = rg < a.rgroups;
  const int bt = blockIdx.x;
  const int co_tile = bt % a.co_tiles, ci_chunk = bt / a.co_tiles;
  const int co0 = co_tile * TCO, ci0 = ci_chunk * a.cic;
  const int cic = min(a.cic, a.cin - ci0);
  const int Hc = a.H << a.up_h, Wc = a.W << a.up;
  const long long HW = (long long)a.H * a.W, HoWo = (long long)a.Ho * a.Wo;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  const int tiles = a.tiles_x * a.tiles_y;

  float acc[COT][KK];
#pragma unroll
  for (int c = 0; c < COT; ++c)
#pragma unroll
    for (int k = 0; k < KK; ++k) acc[c][k] = 0.f;

  for (long long item = blockIdx.y; item < (long long)a.N * tiles; item += gridDim.y) {
    const int n = (int)(item / tiles), tl = (int)(item - (long long)n * tiles);
    const int ty_tile = tl / a.tiles_x, tx_tile = tl - ty_tile * a.tiles_x;
    const int oy0 = ty_tile * a.TH, ox0 = tx_tile * a.TW;
    __syncthreads();
    for (int i = tid; i < PH; i += nthreads) { int iy = (oy0 * S - a.pad_y + i) % Hc; if (iy < 0) iy += Hc; rowmap[i] = iy >> a.up_h; }
    for (int i = tid; i < PW; i += nthreads) { int ix = (ox0 * S - a.pad_x + i) % Wc; if (ix < 0) ix += Wc; colmap[i] = ix >> a.up; }
    __syncthreads();
    const float* xn = a.x + (long long)n * a.x_bs;
    // batched staging (all loads of 2 lines x 4 x 32 columns in flight before the first use; see conv_fwd.cu)
    {
      constexpr int RU = 2, PU = 4;
      const int nlines = cic * PH;
      for (int r0 = warp * RU; r0 < nlines; r0 += nwarps * RU) {
        for (int px0 = 0; px0 < a.pitch; px0 += 32 * PU) {
          float v[RU][PU], sc[RU], sh[RU];
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int r = r0 + u;
            const bool rok = r < nlines;
            const int ci = rok ? r / PH : 0, py = rok ? r - ci * PH : 0;
            const float* src = xn + (ci0 + ci) * HW + (long long)rowmap[py] * a.W;
            sc[u] = 1.f; sh[u] = 0.f;
            if (a.scale && rok) { sc[u] = a.scale[ci0 + ci]; sh[u] = a.shift[ci0 + ci]; }
#pragma unroll
            for (int p = 0; p < PU; ++p) {
              const int px = px0 + p * 32 + lane;
              v[u][p] = (rok && px < PW) ? __ldg(src + colmap[px]) : 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int r = r0 + u;
            if (r < nlines) {
              const int ci = r / PH, py = r - ci * PH;
              float* dst = patch + ci * a.plane + py * a.pitch;
#pragma unroll
              for (int p = 0; p < PU; ++p) {
                const int px = px0 + p * 32 + lane;
                if (px < a.pitch) {
                  float t = 0.f;
                  if (px < PW) { t = fmaf(v[u][p], sc[u], sh[u]); if (a.relu_in) t = fmaxf(t, 0.f); }
                  dst[px] = t;
                }
              }
            }
          }
        }
      }
    }
    const float* dn = a.dO + (long long)n * a.dO_bs;
    {
      constexpr int RU = 4;
      const int nlines = TCO * a.TH;
      for (int r0 = warp * RU; r0 < nlines; r0 += nwarps * RU) {
        for (int px0 = 0; px0 < a.dpitch; px0 += 32) {
          float v[RU];
          const int px = px0 + lane;
#pragma unroll
          for (int u = 0; u < RU; ++u) {
            const int r = r0 + u;
            const int co = r / a.TH, py = r - co * a.TH;
            const bool ok = r < nlines && (co0 + co < a.cout) && (oy0 + py < a.Ho) && px < a.TW && ox0 + px < a.Wo;
            v[u] = ok ? __ldg(dn + (long long)(co0 + co) * HoWo + (long long)(oy0 + py) * a.Wo + ox0 + px) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < RU; ++u)
            if (r0 + u < nlines && px < a.dpitch) dsm[(r0 + u) * a.dpitch + px] = v[u];
        }
      }
    }
    __syncthreads();
    if (active && ci_l < cic) {
      const float* pbase = patch + ci_l * a.plane;
      for (int py = rg; py < a.TH; py += a.rgroups) {
        for (int xg = 0; xg < a.TW; xg += PXT) {
          float d[COT][PXT];
#pragma unroll
          for (int c = 0; c < COT; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(dsm + ((tc * COT + c) * a.TH + py) * a.dpitch + xg);
            d[c][0] = v.x; d[c][1] = v.y; d[c][2] = v.z; d[c][3] = v.w;
          }
#pragma unroll
          for (int ky = 0; ky < KH; ++ky) {
            const float* rowp = pbase + (py * S + ky) * a.pitch + xg * S;
            float in[SEG];
#pragma unroll
            for (int q = 0; q < SEG / 4; ++q) {
              const float4 v = *reinterpret_cast<const float4*>(rowp + 4 * q);
              in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int kx = 0; kx < KW; ++kx)
#pragma unroll
              for (int c = 0; c < COT; ++c)
#pragma unroll
                for (int p = 0; p < PXT; ++p) acc[c][ky * KW + kx] = fmaf(d[c][p], in[p * S + kx], acc[c][ky * KW + kx]);
          }
        }
      }
    }
  }
  if (active && ci_l < cic) {
#pragma unroll
    for (int c = 0; c < COT; ++c) {
      const int co = co0 + tc * COT + c;
      if (co < a.cout) {
        double* o = a.dW + ((long long)co * a.cin + ci0 + ci_l) * KK;
#pragma unroll
        for (int k = 0; k < KK; ++k) atomicAdd(o + k, (double)acc[c][k]);
      }
    }
  }
}

__global__ void d2f_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = (float)src[i];
}

int arpde_d2f_launch(const double* src, float* dst, long long n, cudaStream_t stream) {
  d2f_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, stream>>>(src, dst, n);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// ------------------------------------------------------------------ weight gradient of narrow layers (cout <= 4, stride 1, 2D)
// dW has only cout*cin*k^2 entries (800 for the 16 -> 2 5x5 output conv, 2160 for a 60 -> 4 dense layer) while the sum runs over
// N*H*W positions: the work is ~10 us of FMAs, and both other kernels are ~30x off (generic wgrad_kernel: 0.42 ms, 0.2 ms; the
// tensor core pays ~40 cycles per MMA for 4 useful accumulator rows: 0.13 ms).  Here a thread owns one (input channel, ky) pair and
// all KW x CO accumulators of it; it slides along the rows of a tile: per 8 positions three float4 of the input row and the 8 x CO
// output-gradient values (broadcast) feed 8*KW*CO FMAs.  Persistent CTAs keep the accumulators across tiles; one block-level
// reduction over the row groups and cin*KH*KW*CO fp64 atomics per CTA at the end.
struct WgradSmallArgs {
  const float* x; long long x_bs; const float* scale; const float* shift; int relu_in;
  const float* dO; long long dO_bs; double* dW;
  int N, cin, H, W, cout, TH, tiles_y, pitch, gpitch;
};

template <int KH, int KW, int CO>
__global__ void __launch_bounds__(256) wgrad_small_kernel(WgradSmallArgs a) {
  constexpr int PXB = 8, SEG = (PXB + KW - 1 + 3) & ~3;
  extern __shared__ __align__(16) float smem[];
  const int PH = a.TH + KH - 1, PW = a.W + KW - 1;
  const int pad_y = (KH - 1) / 2, pad_x = (KW - 1) / 2;
  float* patch = smem;                                   // [cin][PH][pitch]
  float* gys = patch + (size_t)a.cin * PH * a.pitch;     // [TH][gpitch]: CO values per position, interleaved
  int* rowmap = reinterpret_cast<int*>(gys + (size_t)a.TH * a.gpitch);   // [PH]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int npairs = a.cin * KH;
  const int rgroups = blockDim.x / npairs;               // row groups; threads beyond npairs * rgroups only help staging
  const int pr = tid % npairs, rg = tid / npairs;
  const bool worker = rg < rgroups;
  const int ci = pr / KH, ky = pr - ci * KH;
  const long long HW = (long long)a.H * a.W;
  float acc[KW][CO];
#pragma unroll
  for (int k = 0; k < KW; ++k)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[k][c] = 0.f;

  const int W4 = a.W >> 2;
  for (int t = blockIdx.x; t < a.N * a.tiles_y; t += gridDim.x) {
    const int n = t / a.tiles_y, y0 = (t - n * a.tiles_y) * a.TH;
    const float* xn = a.x + (long long)n * a.x_bs;
    const float* gn = a.dO + (long long)n * a.dO_bs;
    __syncthreads();
    // ---- input patch: whole rows + wrap halo, BN/ReLU applied (stage_full_rows)
    for (int i = tid; i < PH; i += blockDim.x) { int iy = y0 - pad_y + i; if (iy < 0) iy += a.H; else if (iy >= a.H) iy -= a.H; if (iy >= a.H) iy -= a.H; rowmap[i] = iy; }
    __syncthreads();
    stage_full_rows<4>(patch, a.pitch, xn, (int)HW, 0, a.cin * PH, PH, rowmap, a.W, pad_x, PW - a.W, a.scale, a.shift, a.relu_in, tid, blockDim.x);
    // ---- output gradient: [TH][W][CO] interleaved (rows past the image: zeros)
    for (int i = tid; i < a.TH * CO * W4; i += blockDim.x) {
      const int x4 = i % W4, rc = i / W4, c = rc % CO, r = rc / CO;
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < a.cout && y0 + r < a.H) g4 = __ldg(reinterpret_cast<const float4*>(gn + (long long)c * HW + (long long)(y0 + r) * a.W) + x4);
      float* d = gys + (size_t)r * a.gpitch + (size_t)(4 * x4) * CO + c;
      d[0] = g4.x; d[CO] = g4.y; d[2 * CO] = g4.z; d[3 * CO] = g4.w;
    }
    __syncthreads();
    if (worker) {
      for (int r = rg; r < a.TH; r += rgroups) {
        const float* ap = patch + (size_t)(ci * PH + r + ky) * a.pitch;
        const float* gp = gys + (size_t)r * a.gpitch;
        for (int x0 = 0; x0 < a.W; x0 += PXB) {
          float in[SEG], g[PXB][CO];
#pragma unroll
          for (int q = 0; q < SEG / 4; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(ap + x0 + 4 * q);
            in[4 * q] = v4.x; in[4 * q + 1] = v4.y; in[4 * q + 2] = v4.z; in[4 * q + 3] = v4.w;
          }
#pragma unroll
          for (int q = 0; q < PXB * CO / 4; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(gp + (size_t)x0 * CO + 4 * q);
            (&g[0][0])[4 * q] = v4.x; (&g[0][0])[4 * q + 1] = v4.y; (&g[0][0])[4 * q + 2] = v4.z; (&g[0][0])[4 * q + 3] = v4.w;
          }
#pragma unroll
          for (int p = 0; p < PXB; ++p)
#pragma unroll
            for (int k = 0; k < KW; ++k)
#pragma unroll
              for (int c = 0; c < CO; ++c) acc[k][c] = fmaf(g[p][c], in[p + k], acc[k][c]);
        }
      }
    }
  }
  // ---- reduce the row groups in shared memory, then fp64 atomics (one per weight and CTA)
  __syncthreads();
  float* red = smem;                                     // [npairs][KW * CO]
  for (int g = 0; g < rgroups; ++g) {
    if (worker && rg == g) {
#pragma unroll
      for (int k = 0; k < KW; ++k)
#pragma unroll
        for (int c = 0; c < CO; ++c) {
          float* d = red + (size_t)pr * (KW * CO) + k * CO + c;
          *d = g == 0 ? acc[k][c] : *d + acc[k][c];
        }
    }
    __syncthreads();
  }
  for (int i = tid; i < npairs * KW * CO; i += blockDim.x) {
    const int c = i % CO, k = (i / CO) % KW, p2 = i / (CO * KW);
    const int ci2 = p2 / KH, ky2 = p2 - ci2 * KH;
    if (c < a.cout) atomicAdd(a.dW + (((long long)c * a.cin + ci2) * KH + ky2) * KW + k, (double)red[i]);
  }
}

typedef void (*wgrad_small_fn)(WgradSmallArgs);

// returns ARPDE_OK if the narrow-layer kernel took the job, ARPDE_ERR_UNSUPPORTED if the layer is not its kind (caller falls back)
static int wgrad_small_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                              long long dO_bs, double* dW64, cudaStream_t stream) {
  if (d.stride != 1 || d.up || d.dilate || d.pad_override >= 0 || d.H < 2 || d.cout > 4 || d.kh != d.kw) return ARPDE_ERR_UNSUPPORTED;
  if ((d.W & 7) || d.W > 128 || (x_bs & 3) || (dO_bs & 3) || (((uintptr_t)x | (uintptr_t)dO) & 15)) return ARPDE_ERR_UNSUPPORTED;
  const int CO = d.cout <= 2 ? 2 : 4;
  wgrad_small_fn fn = nullptr;
  if (d.kh == 3) fn = CO == 2 ? wgrad_small_kernel<3, 3, 2> : wgrad_small_kernel<3, 3, 4>;
  else if (d.kh == 5) fn = CO == 2 ? wgrad_small_kernel<5, 5, 2> : wgrad_small_kernel<5, 5, 4>;
  if (!fn) return ARPDE_ERR_UNSUPPORTED;
  const int npairs = d.cin * d.kh;
  if (npairs > 256) return ARPDE_ERR_UNSUPPORTED;
  WgradSmallArgs a; memset(&a, 0, sizeof(a));
  a.x = x; a.x_bs = x_bs; a.scale = scale; a.shift = shift; a.relu_in = d.relu_in; a.dO = dO; a.dO_bs = dO_bs; a.dW = dW64;
  a.N = d.N; a.cin = d.cin; a.H = d.H; a.W = d.W; a.cout = d.cout;
  a.pitch = ((d.W + d.kw - 1 + 3) & ~3) + 4;                 // room for the 12-float window of the last 8 columns
  a.gpitch = d.W * CO;
  // tile height: the tallest that keeps the patch + gradient tile under ~100 KB (two CTAs per SM), whole tiles preferred
  int TH = d.H < 16 ? d.H : 16;
  size_t smem = 0;
  for (;; --TH) {
    smem = ((size_t)d.cin * (TH + d.kh - 1) * a.pitch + (size_t)TH * a.gpitch + (TH + d.kh - 1)) * 4;
    if (smem <= 100 * 1024 || TH == 1) break;
  }
  if (smem > 200 * 1024) return ARPDE_ERR_UNSUPPORTED;
  const size_t red_bytes = (size_t)npairs * d.kw * CO * 4;
  if (red_bytes > smem) smem = red_bytes;
  a.TH = TH; a.tiles_y = (d.H + TH - 1) / TH;
  if (smem > 48 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (long long)d.N * a.tiles_y; const long long cap = 2ll * arpde_num_sms(); if (blocks > cap) blocks = cap;
  fn<<<(unsigned)blocks, 256, smem, stream>>>(a);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

typedef void (*wgrad_fn)(WgradArgs);
template <int KH, int KW, int S>
static wgrad_fn pick_w(int cout, int* cot) {
  constexpr int KK = KH * KW;
  constexpr int C = KK <= 9 ? 8 : (KK <= 25 ? 4 : 2);
  if (C == 8 && cout <= 4) { *cot = 4; return wgrad_kernel<KH, KW, S, (C == 8 ? 4 : C)>; }   // dense layers (growth 4): no zero channels
  *cot = C;
  return wgrad_kernel<KH, KW, S, C>;
}
static wgrad_fn pick_wgrad(int kh, int kw, int s, int cout, int* cot) {
#define PK(KH_, KW_, S_) if (kh == KH_ && kw == KW_ && s == S_) return pick_w<KH_, KW_, S_>(cout, cot)
  PK(1, 1, 1); PK(3, 3, 1); PK(3, 3, 2); PK(5, 5, 1); PK(4, 4, 2); PK(7, 7, 2);
  PK(1, 3, 1); PK(1, 3, 2); PK(1, 5, 1); PK(1, 4, 2); PK(1, 7, 2);
#undef PK
  return nullptr;
}

// dW (float, [cout][cin][kh][kw]) = sum over batch; dW64 scratch: cout*cin*kh*kw doubles
int arpde_conv_wgrad_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                            long long dO_bs, double* dW64, float* dW, cudaStream_t stream, bool generic_only) {
  ARPDE_CHECK_ARG(x && dO && dW64 && dW, "conv_wgrad: null pointer");
  int cot = 8;
  wgrad_fn fn = pick_wgrad(d.kh, d.kw, d.stride, d.cout, &cot);
  if (!fn) { arpde_set_error("conv_wgrad: kernel %dx%d stride %d not instantiated", d.kh, d.kw, d.stride); return ARPDE_ERR_UNSUPPORTED; }
  const int KK = d.kh * d.kw;
  const long long wn = (long long)d.cout * d.cin * KK;
  ARPDE_CUDA(cudaMemsetAsync(dW64, 0, sizeof(double) * wn, stream));
  if (d.N > 0 && !generic_only && wgrad_small_launch(d, x, x_bs, scale, shift, dO, dO_bs, dW64, stream) == ARPDE_OK)
    return arpde_d2f_launch(dW64, dW, wn, stream);
  const int one_d = (d.H == 1 && d.kh == 1);
  const int up_h = one_d ? 0 : d.up;
  const int Hc = d.H << up_h, Wc = d.W << d.up;
  const int sh = one_d ? 1 : d.stride;
  const int Ho = Hc / sh, Wo = Wc / d.stride;
  WgradArgs a; memset(&a, 0, sizeof(a));
  a.x = x; a.x_bs = x_bs; a.cin = d.cin; a.H = d.H; a.W = d.W; a.up = d.up; a.up_h = up_h;
  a.pad_y = d.kh == 4 ? 1 : (d.kh - 1) / 2; a.pad_x = d.kw == 4 ? 1 : (d.kw - 1) / 2;
  a.scale = scale; a.shift = shift; a.relu_in = d.relu_in;
  a.dO = dO; a.dO_bs = dO_bs; a.cout = d.cout; a.Ho = Ho; a.Wo = Wo; a.dW = dW64; a.N = d.N;
  int ncg = (d.cout + cot - 1) / cot; if (ncg > 24) ncg = 24;
  int cic = 256 / ncg; if (cic > d.cin) cic = d.cin; if (cic < 1) cic = 1;
  a.ncg = ncg; a.cic = cic;
  // few (co group, ci) pairs (small layers): several row groups share a tile so that the CTA still has ~256 busy threads
  int rgroups = 256 / (ncg * cic); if (rgroups < 1) rgroups = 1; if (rgroups > 8) rgroups = 8;
  a.rgroups = rgroups;
  const int TCO = ncg * cot;
  a.co_tiles = (d.cout + TCO - 1) / TCO; a.ci_chunks = (d.cin + cic - 1) / cic;
  int TW = ((Wo + 3) / 4) * 4; if (TW > 64) TW = 64;
  a.TW = TW; a.tiles_x = (Wo + TW - 1) / TW;
  const int seg = (((4 - 1) * d.stride + d.kw) + 3) & ~3;
  const int PW = (TW - 1) * d.stride + d.kw;
  int pitch = (TW - 4) * d.stride + seg; if (pitch < PW) pitch = PW; pitch = (pitch + 3) & ~3;
  a.pitch = pitch; a.dpitch = TW;
  // rows per tile: keep shared memory under ~96 KB
  int TH = Ho < 8 ? Ho : 8;
  size_t smem = 0;
  for (;; --TH) {
    const int PH = (TH - 1) * sh + d.kh;
    int plane = PH * pitch; if (((plane >> 2) & 1) == 0) plane += 4;      // odd multiple of 4 floats: conflict-free float4 reads across ci
    a.plane = plane;
    smem = ((size_t)((PH + PW + 3) & ~3) + (size_t)cic * plane + (size_t)TCO * TH * TW) * 4;
    if (smem <= 96 * 1024 || TH == 1) break;
  }
  ARPDE_CHECK_ARG(smem <= 200 * 1024, "conv_wgrad: tile needs %zu bytes of shared memory", smem);
  a.TH = TH; a.tiles_y = (Ho + TH - 1) / TH;
  if (smem > 48 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long items = (long long)d.N * a.tiles_x * a.tiles_y;
  const int gx = a.co_tiles * a.ci_chunks;
  long long split = (4ll * arpde_num_sms() + gx - 1) / gx; if (split > items) split = items; if (split < 1) split = 1; if (split > 65535) split = 65535;
  int threads = ((ncg * cic * rgroups + 31) / 32) * 32;
  fn<<<dim3(gx, (unsigned)split), threads, smem, stream>>>(a);
  ARPDE_LAUNCH_CHECK();
  d2f_kernel<<<(unsigned)ceil_div64(wn, 256), 256, 0, stream>>>(dW64, dW, wn);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// narrow layers first (sliding-window kernel, inside arpde_conv_wgrad_launch), then the tensor core, then the generic fp32 kernel
static bool wgrad_small_eligible(const ConvDesc& d, const float* x, long long x_bs, const float* dO, long long dO_bs) {
  if (d.stride != 1 || d.up || d.dilate || d.pad_override >= 0 || d.H < 2 || d.cout > 4 || d.kh != d.kw || (d.kh != 3 && d.kh != 5)) return false;
  if ((d.W & 7) || d.W > 128 || (x_bs & 3) || (dO_bs & 3) || (((uintptr_t)x | (uintptr_t)dO) & 15) || d.cin * d.kh > 256) return false;
  return true;
}

int arpde_conv_wgrad_best_launch(const ConvDesc& d, const float* x, long long x_bs, const float* scale, const float* shift, const float* dO,
                                 long long dO_bs, double* dW64, float* dW, float* tc_scratch, long long tc_scratch_floats, cudaStream_t stream) {
  if (!wgrad_small_eligible(d, x, x_bs, dO, dO_bs) && tc_scratch) {
    const long long need = arpde_conv_wgrad_tc_scratch_floats(d);
    if (need > 0 && need <= tc_scratch_floats && arpde_conv_wgrad_tc_wanted(d) && (((uintptr_t)dO) & 15) == 0 && (dO_bs & 3) == 0)
      return arpde_conv_wgrad_tc_launch(d, x, x_bs, scale, shift, dO, dO_bs, dW64, dW, tc_scratch, tc_scratch_floats, stream);
  }
  return arpde_conv_wgrad_launch(d, x, x_bs, scale, shift, dO, dO_bs, dW64, dW, stream);
}

// C ABI: weight gradient of one circular convolution, best available kernel (narrow-layer sliding window, tensor core when a scratch of
// arpde_conv_wgrad_tc_scratch(...) floats is given and the layer is covered, else the generic fp32 kernel).  Same contract as
// arpde_conv_wgrad_tc; path = 0 auto, 1 generic FFMA only (cross-check).
extern "C" int arpde_conv_wgrad(const float* x, int64_t x_bstride, const float* scale, const float* shift, const float* dO,
                                int64_t dO_bstride, double* dW64, float* dW, int N, int cin, int H, int W, int cout, int kh, int kw,
                                int stride, int up, int relu_in, float* scratch, int64_t scratch_floats, int path, arpde_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ConvDesc d;
  d.dilate = 0; d.pad_override = -1; d.N = N; d.cin = cin; d.H = H; d.W = W; d.cout = cout; d.kh = kh; d.kw = kw; d.stride = stride;
  d.up = up; d.relu_in = relu_in; d.act = 0; d.y_ctot = cout; d.y_coff = 0; d.n_per_group = N > 0 ? N : 1;
  ARPDE_CHECK_ARG((scale == nullptr) == (shift == nullptr), "conv_wgrad: scale/shift must both be given or both be null");
  ARPDE_CHECK_ARG(x && dO && dW64 && dW, "conv_wgrad: null pointer");
  if (path == 1) {
    // generic kernel only
    return arpde_conv_wgrad_launch(d, x, x_bstride, scale, shift, dO, dO_bstride, dW64, dW, stream, true);
  }
  return arpde_conv_wgrad_best_launch(d, x, x_bstride, scale, shift, dO, dO_bstride, dW64, dW, scratch, scratch_floats, stream);
}
