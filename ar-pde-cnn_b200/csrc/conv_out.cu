// conv_out.cu -- the output convolution of the 2D DenseED (LastTransUp.conv3: BN, ReLU, circular 5x5, 16 -> 2 channels on the 64x64 grid).
//
// Replaces the last module of last_decoding (2D-Burgers-SWAG/nn/denseEDcirc2d.py:186-189) in the rollout / inference path; the generic
// row-tiled FFMA kernel (conv_fwd.cu) spends 0.55 ms per AR step of 1920 samples on it (23 TFLOP/s).  Same formulation as the fused
// dense block and the KS rollout kernel:
//   * a CTA owns 16 output rows of one sample; its 20 input rows (2 halo rows on either side, circular in y) of all input channels
//     arrive once by 16-byte cp.async (80 KB for 16 channels: two CTAs per SM, one computes while the other loads);
//   * a row is 64 columns = 32 lanes x 2 adjacent columns: a lane reads one float2 per input row, applies BN + ReLU once, and gets the
//     two halo columns on either side from its neighbour lanes by shuffle (circular wrap in x = the shuffle's modulo);
//   * a warp computes 8 output rows x 2 columns x 2 output channels per lane for a quarter of the input channels, walking the 12 input
//     rows once per channel with the channel's 50 weights in registers (800 FMAs per 12 shared loads + 48 shuffles + 13 weight loads);
//     the four channel quarters meet in shared memory in a fixed order (bit-reproducible).
// Bound: fp32 FMA pipe.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"

namespace {

constexpr int CO_W = 64, CO_TR = 16, CO_R = 8, CO_KS = 4, CO_WARPS = (CO_TR / CO_R) * CO_KS, CO_ROWS = CO_TR + 4, CO_WSTRIDE = 52, CO_MAXC = 16;

struct ConvOutArgs {
  const float* x; long long x_bs;
  const float* w; long long w_gs;
  const float* scale; const float* shift; long long ss_gs;
  float* y; long long y_bs; int y_coff;
  int cin, H, N, n_per_group, relu_in, act;
};

__global__ void __launch_bounds__(32 * CO_WARPS, 2) conv_out5x5_kernel(ConvOutArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm;                                          // [cin][20][64]
  float* Wt = X + (size_t)a.cin * CO_ROWS * CO_W;         // [cin][52]: (ky, kx, co) of one input channel, padded to 13 float4
  float* SS = Wt + CO_MAXC * CO_WSTRIDE;                  // [2][16]
  float* RED = SS + 2 * CO_MAXC;                          // [2 row groups][3 channel quarters][32][32]
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int tiles = a.H / CO_TR;
  const int n = blockIdx.x / tiles, y0 = (blockIdx.x - n * tiles) * CO_TR, g = n / a.n_per_group;
  const long long HW = (long long)a.H * CO_W;
  const float* xn = a.x + (long long)n * a.x_bs;
  pdl_wait();
  {  // input rows y0 - 2 .. y0 + 17 of every channel
    const int npiece = a.cin * CO_ROWS * (CO_W / 4);
    for (int q = tid; q < npiece; q += 32 * CO_WARPS) {
      const int c4 = q & 15, rr = q >> 4;                 // float4 within the row, (channel, local row)
      const int c = rr / CO_ROWS, r = rr - c * CO_ROWS;
      int y = y0 - 2 + r; y = y < 0 ? y + a.H : (y >= a.H ? y - a.H : y);
      const float* src = xn + c * HW + (long long)y * CO_W + 4 * c4;
      float* dst = X + ((size_t)c * CO_ROWS + r) * CO_W + 4 * c4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  {  // weights [2][cin][5][5] -> [ci][(ky, kx, co)], folded BatchNorm
    const float* w = a.w + (long long)g * a.w_gs;
    for (int i = tid; i < 2 * a.cin * 25; i += 32 * CO_WARPS) {
      const int co = i / (a.cin * 25), r = i - co * a.cin * 25, ci = r / 25, t = r - ci * 25;
      Wt[ci * CO_WSTRIDE + t * 2 + co] = __ldg(w + i);
    }
    for (int i = tid; i < a.cin; i += 32 * CO_WARPS) {
      Wt[i * CO_WSTRIDE + 50] = 0.f; Wt[i * CO_WSTRIDE + 51] = 0.f;
      SS[i] = a.scale ? __ldg(a.scale + (long long)g * a.ss_gs + i) : 1.f;
      SS[CO_MAXC + i] = a.scale ? __ldg(a.shift + (long long)g * a.ss_gs + i) : 0.f;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int rg = warp / CO_KS, kg = warp - rg * CO_KS;
  const int rb = rg * CO_R;                               // first output row of the warp (local); its input rows are rb .. rb + 11
  // acc2[o][px] = (channel 0, channel 1) of output row o, column px: packed fp32 FMAs (fma.rn.f32x2) over the output-channel pair, whose
  // weights are adjacent in the staged table -- half the FMA issue slots of the scalar form, bit-identical results
  float2 acc2[CO_R][2];
#pragma unroll
  for (int o = 0; o < CO_R; ++o) { acc2[o][0] = make_float2(0.f, 0.f); acc2[o][1] = make_float2(0.f, 0.f); }
  const int c_lo = kg * a.cin / CO_KS, c_hi = (kg + 1) * a.cin / CO_KS;
  const int lm = (lane + 31) & 31, lp = (lane + 1) & 31;
#pragma unroll 1
  for (int c = c_lo; c < c_hi; ++c) {
    float wr[CO_WSTRIDE];
#pragma unroll
    for (int q = 0; q < CO_WSTRIDE / 4; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(Wt + c * CO_WSTRIDE + 4 * q);
      wr[4 * q] = t.x; wr[4 * q + 1] = t.y; wr[4 * q + 2] = t.z; wr[4 * q + 3] = t.w;
    }
    const float s = SS[c], t0 = SS[CO_MAXC + c];
    const float* xc = X + ((size_t)c * CO_ROWS + rb) * CO_W + 2 * lane;
#pragma unroll
    for (int ir = 0; ir < CO_R + 4; ++ir) {
      const float2 v = *reinterpret_cast<const float2*>(xc + ir * CO_W);
      float a0 = fmaf(v.x, s, t0), a1 = fmaf(v.y, s, t0);
      if (a.relu_in) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
      float win[6];
      win[0] = __shfl_sync(0xffffffffu, a0, lm); win[1] = __shfl_sync(0xffffffffu, a1, lm);
      win[2] = a0; win[3] = a1;
      win[4] = __shfl_sync(0xffffffffu, a0, lp); win[5] = __shfl_sync(0xffffffffu, a1, lp);
      float2 win2[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) win2[q] = make_float2(win[q], win[q]);
#pragma unroll
      for (int ky = 0; ky < 5; ++ky) {
        const int o = ir - ky;                            // output row fed by (input row ir, tap row ky)
        if (o >= 0 && o < CO_R) {
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const float2 w01 = make_float2(wr[(ky * 5 + kx) * 2], wr[(ky * 5 + kx) * 2 + 1]);
            acc2[o][0] = __ffma2_rn(w01, win2[kx], acc2[o][0]);
            acc2[o][1] = __ffma2_rn(w01, win2[kx + 1], acc2[o][1]);
          }
        }
      }
    }
  }
  // channel quarters 1..3 -> shared memory, quarter 0 adds them in order and writes the output
  if (kg > 0) {
    float* r = RED + ((size_t)(rg * 3 + kg - 1) * 32) * 32 + lane;
#pragma unroll
    for (int o = 0; o < CO_R; ++o)
#pragma unroll
      for (int q = 0; q < 4; ++q) r[(o * 4 + q) * 32] = (q >> 1) ? acc2[o][q & 1].y : acc2[o][q & 1].x;
  }
  __syncthreads();
  if (kg == 0) {
    float* yn = a.y + (long long)n * a.y_bs + (long long)a.y_coff * HW;
#pragma unroll
    for (int o = 0; o < CO_R; ++o) {
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float t = (q >> 1) ? acc2[o][q & 1].y : acc2[o][q & 1].x;
#pragma unroll
        for (int k = 0; k < 3; ++k) t += RED[((size_t)(rg * 3 + k) * 32 + o * 4 + q) * 32 + lane];
        v[q] = apply_act(t, a.act);
      }
      const long long off = (long long)(y0 + rb + o) * CO_W + 2 * lane;
      *reinterpret_cast<float2*>(yn + off) = make_float2(v[0], v[1]);
      *reinterpret_cast<float2*>(yn + HW + off) = make_float2(v[2], v[3]);
    }
  }
}

}  // namespace

bool arpde_conv_out_shape_ok(const ConvDesc& d, const float* x, long long x_bs, const float* y, const double* stats) {
  if (arpde_dev_env("ARPDE_CONV_OUT_OFF")) return false;
  if (d.kh != 5 || d.kw != 5 || d.stride != 1 || d.up || d.dilate || d.pad_override >= 0 || d.cout != 2 || stats) return false;
  if (d.W != CO_W || d.H < CO_TR || d.H % CO_TR || d.cin % CO_KS || d.cin > CO_MAXC || d.cin < CO_KS) return false;
  if ((((uintptr_t)x) & 15) || (x_bs & 3) || (((uintptr_t)y) & 7)) return false;
  return true;
}

int arpde_conv_out_launch(const ConvDesc& d, const float* x, long long x_bs, const float* w, long long w_gs, const float* scale, const float* shift,
                          long long ss_gs, float* y, cudaStream_t stream) {
  ConvOutArgs a; memset(&a, 0, sizeof(a));
  a.x = x; a.x_bs = x_bs; a.w = w; a.w_gs = w_gs; a.scale = scale; a.shift = shift; a.ss_gs = ss_gs;
  a.y = y; a.y_bs = (long long)d.y_ctot * d.H * d.W; a.y_coff = d.y_coff;
  a.cin = d.cin; a.H = d.H; a.N = d.N; a.n_per_group = d.n_per_group; a.relu_in = d.relu_in; a.act = d.act;
  const size_t smem = ((size_t)d.cin * CO_ROWS * CO_W + CO_MAXC * CO_WSTRIDE + 2 * CO_MAXC + 2 * 3 * 32 * 32) * sizeof(float);
  ARPDE_CUDA(cudaFuncSetAttribute(conv_out5x5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (long long)d.N * (d.H / CO_TR);
  ARPDE_CHECK_ARG(blocks < (1ll << 31), "conv_out: too many tiles");
  ARPDE_CUDA(arpde_launch1(conv_out5x5_kernel, (unsigned)blocks, 32 * CO_WARPS, smem, stream, d.pdl != 0, a));
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
