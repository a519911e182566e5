// rollout1d.cu -- persistent auto-regressive rollout of the 1D Kuramoto-Sivashinsky DenseED: ONE launch for all T steps.
//
// Replaces the AR loops of 1D-KS-SWAG/main.py test()/testSample() (:117-131, :148-163) around DenseED.forward
// (1D-KS-SWAG/nn/denseEDcirc.py:243-300: conv7/s2 (2->32), 4 x [BN, ReLU, conv3 (->4), cat], BN, ReLU, conv1 (48->24), BN, ReLU,
// nearest x2, conv3 (24->12), BN, ReLU, conv5 (12->1); all circular, no bias) for the configuration the reference drivers build
// (args.py defaults: blocks [4], growth 4, init_features 32, nic 2, nel 96; eval-mode BatchNorm).
//
// The whole network is 4 820 parameters and <= 9 KB of activations per sample, so:
//   * ONE WARP OWNS ONE TRAJECTORY for the entire rollout: its activations live in a private 10 KB slice of shared memory, the last two
//     states (the model input) never leave the SM, and only the 96 new values per step go to HBM.  Layers are separated by __syncwarp only;
//     there is no CTA barrier and no launch inside the time loop (the generic path needs 9 launches per step: 434 us per step of
//     1920 trajectories, launch-latency bound).
//   * A CTA is a batch of up to 16 consecutive trajectories (one per warp) plus the prepared weights of the (at most two, when a
//     batch is not larger than a SWAG group) weight samples they use: weights are transposed to [ci][tap][co] when they are staged,
//     BatchNorm is folded to scale/shift, the nearest-x2 + conv3 pair is folded into four low-resolution taps (w0 | w1+w2 | w0+w1 | w2).
//   * Inside a warp a half-warp covers the 48 low-resolution positions with 3 consecutive positions per lane (stride-3 shared-memory
//     accesses are conflict-free, the two half-warps sit 16 banks apart); halo values come from the neighbouring lanes by shuffle
//     (the circular wrap is the shuffle's modulo).  The two half-warps split the input channels of the 4-channel dense layers (partial
//     sums combined with shfl_xor 16) and the output channels of the wider layers.
// Bound: fp32 FMA pipe (0.51 MFLOP per sample-step, ~30 B per sample-step of HBM traffic).
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/arpde.h"

namespace {

constexpr int L = 96, LH = 48;
// per-group prepared parameters in shared memory (float offsets)
constexpr int W0 = 0;                         // stem   [ci 2][tap 7][co 32]
constexpr int WD0 = 448, WD1 = 832, WD2 = 1264, WD3 = 1744;     // dense j [ci 32+4j][tap 3][co 4]
constexpr int W1 = 2272;                      // conv1  [ci 48][co 24]
constexpr int W2 = 3424;                      // conv2  [ci 24][half 2][fold 4][8 (6 used)]
constexpr int W3 = 4960;                      // conv3  [ci 12][tap 5] (+4 pad)
constexpr int SC = 5024, NBN = 236, SH = SC + NBN;
constexpr int WFLOATS = 5496;                 // SC + 2*NBN = 5496
// BatchNorm layers d0..d3, norm1, norm2, norm3: offsets {0, 32, 68, 108, 152, 200, 224} and channel counts {32, 36, 40, 44, 48, 24, 12}
// per-trajectory activations: X [2][96] (the two latest states), D [48][48] (dense-block buffer; later E [24][48] at D and F [12][96] at D+1152)
constexpr int XOFF = 0, DOFF = 192, SSTRIDE = 2512;             // 2496 + 16: consecutive trajectories sit 16 banks apart

struct KsNet {
  const float* w[8]; long long w_gs[8];       // stem, dense0..3, conv1, conv2, conv3 (reference layout [co][ci][k]); group strides
  const float* g[7]; const float* b[7]; const float* m[7]; const float* v[7];
  long long g_gs[7], b_gs[7];
  float eps;
};

__device__ __forceinline__ float act(float x, float s, float h) { return fmaxf(fmaf(x, s, h), 0.f); }

// stage the prepared parameters of weight group `grp` into `dst`
__device__ void stage_group(const KsNet& n, int grp, float* __restrict__ dst, int tid, int nthreads) {
  {  // stem [32][2][7] -> [ci][k][co]
    const float* w = n.w[0] + (long long)grp * n.w_gs[0];
    for (int i = tid; i < 448; i += nthreads) { const int co = i / 14, r = i - co * 14; dst[W0 + r * 32 + co] = w[i]; }
  }
  const int wd[4] = {WD0, WD1, WD2, WD3};
#pragma unroll
  for (int j = 0; j < 4; ++j) {  // dense j [4][cin][3] -> [ci][k][co]
    const int cin = 32 + 4 * j;
    const float* w = n.w[1 + j] + (long long)grp * n.w_gs[1 + j];
    for (int i = tid; i < 12 * cin; i += nthreads) { const int co = i / (3 * cin), r = i - co * 3 * cin; dst[wd[j] + r * 4 + co] = w[i]; }
  }
  {  // conv1 [24][48] -> [ci][co]
    const float* w = n.w[5] + (long long)grp * n.w_gs[5];
    for (int i = tid; i < 1152; i += nthreads) { const int co = i / 48, ci = i - co * 48; dst[W1 + ci * 24 + co] = w[i]; }
  }
  {  // conv2 [12][24][3] -> folded [ci][half][4][8]: f0 = w0, f1 = w1 + w2, f2 = w0 + w1, f3 = w2
    const float* w = n.w[6] + (long long)grp * n.w_gs[6];
    for (int i = tid; i < 24 * 2 * 4 * 8; i += nthreads) {
      const int c = i & 7, f = (i >> 3) & 3, h = (i >> 5) & 1, ci = i >> 6;
      float val = 0.f;
      if (c < 6) {
        const float* t = w + ((6 * h + c) * 24 + ci) * 3;
        val = f == 0 ? t[0] : f == 1 ? t[1] + t[2] : f == 2 ? t[0] + t[1] : t[2];
      }
      dst[W2 + i] = val;
    }
  }
  {  // conv3 [1][12][5]
    const float* w = n.w[7] + (long long)grp * n.w_gs[7];
    for (int i = tid; i < 64; i += nthreads) dst[W3 + i] = i < 60 ? w[i] : 0.f;
  }
  const int BN_OFF[7] = {0, 32, 68, 108, 152, 200, 224}, BN_C[7] = {32, 36, 40, 44, 48, 24, 12};
#pragma unroll
  for (int j = 0; j < 7; ++j) {  // eval-mode BatchNorm folded exactly like bn_fold_eval_kernel (bn.cu)
    const float* ga = n.g[j] + (long long)grp * n.g_gs[j];
    const float* be = n.b[j] + (long long)grp * n.b_gs[j];
    for (int c = tid; c < BN_C[j]; c += nthreads) {
      const float inv = 1.0f / sqrtf(n.v[j][c] + n.eps);
      const float sc = ga[c] * inv;
      dst[SC + BN_OFF[j] + c] = sc;
      dst[SH + BN_OFF[j] + c] = fmaf(-n.m[j][c], sc, be[c]);
    }
  }
}

// dense layer j: D[CIN + co][p] = sum_ci sum_k w[co][ci][k] relu(bn(D[ci][p + k - 1])); the half-warps split the input channels
template <int CIN>
__device__ __forceinline__ void dense_layer(float* __restrict__ D, const float* __restrict__ wj, const float* __restrict__ sc,
                                            const float* __restrict__ sh, int l, int h) {
  // packed fp32 FMAs (fma.rn.f32x2) over output-channel pairs: acc2[pair][position]; every lane of a pair is the same IEEE fma chain as
  // the scalar form (bit-identical), at half the FMA issue slots
  float2 acc2[2][3];
#pragma unroll
  for (int c = 0; c < 2; ++c) { acc2[c][0] = make_float2(0.f, 0.f); acc2[c][1] = make_float2(0.f, 0.f); acc2[c][2] = make_float2(0.f, 0.f); }
  constexpr int HC = CIN / 2;
  const int c0 = h * HC;
#pragma unroll 2
  for (int i = 0; i < HC; ++i) {
    const int ci = c0 + i;
    const float s = sc[ci], t = sh[ci];
    const float* row = D + ci * LH + 3 * l;
    const float a0 = act(row[0], s, t), a1 = act(row[1], s, t), a2 = act(row[2], s, t);
    const float aL = __shfl_sync(0xffffffffu, a2, l + 15, 16), aR = __shfl_sync(0xffffffffu, a0, l + 1, 16);
    const float4 w0 = *reinterpret_cast<const float4*>(wj + ci * 12), w1 = *reinterpret_cast<const float4*>(wj + ci * 12 + 4),
                 w2 = *reinterpret_cast<const float4*>(wj + ci * 12 + 8);
    const float2 k0[2] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w)}, k1[2] = {make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)},
                 k2[2] = {make_float2(w2.x, w2.y), make_float2(w2.z, w2.w)};
    const float2 vL = make_float2(aL, aL), v0 = make_float2(a0, a0), v1 = make_float2(a1, a1), v2 = make_float2(a2, a2), vR = make_float2(aR, aR);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      acc2[c][0] = __ffma2_rn(k0[c], vL, __ffma2_rn(k1[c], v0, __ffma2_rn(k2[c], v1, acc2[c][0])));
      acc2[c][1] = __ffma2_rn(k0[c], v0, __ffma2_rn(k1[c], v1, __ffma2_rn(k2[c], v2, acc2[c][1])));
      acc2[c][2] = __ffma2_rn(k0[c], v1, __ffma2_rn(k1[c], v2, __ffma2_rn(k2[c], vR, acc2[c][2])));
    }
  }
  float acc[4][3];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int j = 0; j < 3; ++j) { acc[2 * c][j] = acc2[c][j].x; acc[2 * c + 1][j] = acc2[c][j].y; }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[c][j] += __shfl_xor_sync(0xffffffffu, acc[c][j], 16);
  if (h == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float* o = D + (CIN + c) * LH + 3 * l;
      o[0] = acc[c][0]; o[1] = acc[c][1]; o[2] = acc[c][2];
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(512, 1) ks_rollout_kernel(KsNet net, float* __restrict__ traj, long long bs, int c_hist, int T, int N,
                                                            int npg, int kb, int nbatches, int gcap) {
  extern __shared__ __align__(16) float smem[];
  float* wsm = smem;
  float* asm_ = smem + (size_t)gcap * WFLOATS;
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31, l = lane & 15, h = lane >> 4;
  for (int batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
    const int lo = batch * kb, hi = min(N, lo + kb);
    const int g0 = lo / npg, ng = (hi - 1) / npg - g0 + 1;
    __syncthreads();                                       // previous batch done with the weights
    for (int g = 0; g < ng; ++g) stage_group(net, g0 + g, wsm + g * WFLOATS, tid, blockDim.x);
    __syncthreads();
    const int sample = lo + warp;
    if (sample >= hi) continue;                            // whole warps only: no divergence inside a warp
    const float* P = wsm + (sample / npg - g0) * WFLOATS;
    float* X = asm_ + warp * SSTRIDE + XOFF;
    float* D = asm_ + warp * SSTRIDE + DOFF;
    float* tr = traj + (long long)sample * bs;
    for (int i = lane; i < 2 * L; i += 32) X[i] = tr[(long long)(c_hist - 2) * L + i];
    __syncwarp();
    for (int t = 0; t < T; ++t) {
      // ---- stem: conv7 stride 2 (pad 3) on the two latest states, no BN/ReLU; half-warp h computes channel tiles h and h + 2 (8 channels each)
      {
        float v[2][11];
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const float* x = X + (((t + ci) & 1) * L) + 6 * l;
          float o[6];
#pragma unroll
          for (int j = 0; j < 6; ++j) o[j] = x[j];
#pragma unroll
          for (int j = 0; j < 3; ++j) v[ci][j] = __shfl_sync(0xffffffffu, o[3 + j], l + 15, 16);
#pragma unroll
          for (int j = 0; j < 6; ++j) v[ci][3 + j] = o[j];
#pragma unroll
          for (int j = 0; j < 2; ++j) v[ci][9 + j] = __shfl_sync(0xffffffffu, o[j], l + 1, 16);
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const int ct = h + 2 * pass;
          float2 acc2[4][3];                                // [output-channel pair][position]
#pragma unroll
          for (int c = 0; c < 4; ++c) { acc2[c][0] = make_float2(0.f, 0.f); acc2[c][1] = make_float2(0.f, 0.f); acc2[c][2] = make_float2(0.f, 0.f); }
#pragma unroll
          for (int ci = 0; ci < 2; ++ci)
#pragma unroll
            for (int k = 0; k < 7; ++k) {
              const float4 wa = *reinterpret_cast<const float4*>(P + W0 + (ci * 7 + k) * 32 + ct * 8);
              const float4 wb = *reinterpret_cast<const float4*>(P + W0 + (ci * 7 + k) * 32 + ct * 8 + 4);
              const float2 wv[4] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w)};
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const float2 vv = make_float2(v[ci][2 * j + k], v[ci][2 * j + k]);
#pragma unroll
                for (int c = 0; c < 4; ++c) acc2[c][j] = __ffma2_rn(wv[c], vv, acc2[c][j]);
              }
            }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float* o = D + (ct * 8 + 2 * c) * LH + 3 * l;
            o[0] = acc2[c][0].x; o[1] = acc2[c][1].x; o[2] = acc2[c][2].x;
            o[LH] = acc2[c][0].y; o[LH + 1] = acc2[c][1].y; o[LH + 2] = acc2[c][2].y;
          }
        }
        __syncwarp();
      }
      // ---- dense block: 4 x [BN, ReLU, conv3 -> 4 channels appended]
      dense_layer<32>(D, P + WD0, P + SC + 0, P + SH + 0, l, h);
      dense_layer<36>(D, P + WD1, P + SC + 32, P + SH + 32, l, h);
      dense_layer<40>(D, P + WD2, P + SC + 68, P + SH + 68, l, h);
      dense_layer<44>(D, P + WD3, P + SC + 108, P + SH + 108, l, h);
      // ---- LastTransUp.conv1: BN, ReLU, 1x1 conv 48 -> 24; half-warp h computes channels [12h, 12h + 12); the result replaces D
      {
        float2 acc2[6][3];                                  // [output-channel pair][position]
#pragma unroll
        for (int c = 0; c < 6; ++c) { acc2[c][0] = make_float2(0.f, 0.f); acc2[c][1] = make_float2(0.f, 0.f); acc2[c][2] = make_float2(0.f, 0.f); }
#pragma unroll 2
        for (int ci = 0; ci < 48; ++ci) {
          const float s = P[SC + 152 + ci], tt = P[SH + 152 + ci];
          const float* row = D + ci * LH + 3 * l;
          const float a0 = act(row[0], s, tt), a1 = act(row[1], s, tt), a2 = act(row[2], s, tt);
          const float2 v0 = make_float2(a0, a0), v1 = make_float2(a1, a1), v2 = make_float2(a2, a2);
          const float* wp = P + W1 + ci * 24 + 12 * h;
          const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4), wc = *reinterpret_cast<const float4*>(wp + 8);
          const float2 wv[6] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w),
                                make_float2(wc.x, wc.y), make_float2(wc.z, wc.w)};
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            acc2[c][0] = __ffma2_rn(wv[c], v0, acc2[c][0]); acc2[c][1] = __ffma2_rn(wv[c], v1, acc2[c][1]); acc2[c][2] = __ffma2_rn(wv[c], v2, acc2[c][2]);
          }
        }
        __syncwarp();                                      // every lane has read D
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          float* o = D + (12 * h + 2 * c) * LH + 3 * l;    // E [24][48] over the dead dense-block buffer
          o[0] = acc2[c][0].x; o[1] = acc2[c][1].x; o[2] = acc2[c][2].x;
          o[LH] = acc2[c][0].y; o[LH + 1] = acc2[c][1].y; o[LH + 2] = acc2[c][2].y;
        }
        __syncwarp();
      }
      // ---- LastTransUp.conv2: BN, ReLU, nearest x2, conv3 24 -> 12, folded to the low-resolution grid:
      //      out[2p] = w0 a[p-1] + (w1 + w2) a[p],  out[2p+1] = (w0 + w1) a[p] + w2 a[p+1]; half-warp h computes channels [6h, 6h + 6)
      {
        float* F = D + 1152;                               // F [12][96]
        float2 acc2[3][6];                                  // [output-channel pair][high-resolution position]
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int q = 0; q < 6; ++q) acc2[c][q] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int ci = 0; ci < 24; ++ci) {
          const float s = P[SC + 200 + ci], tt = P[SH + 200 + ci];
          const float* row = D + ci * LH + 3 * l;
          float a[5];
          a[1] = act(row[0], s, tt); a[2] = act(row[1], s, tt); a[3] = act(row[2], s, tt);
          a[0] = __shfl_sync(0xffffffffu, a[3], l + 15, 16);
          a[4] = __shfl_sync(0xffffffffu, a[1], l + 1, 16);
          const float* wp = P + W2 + (ci * 2 + h) * 32;
          float2 f[4][3];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 wa = *reinterpret_cast<const float4*>(wp + q * 8);
            const float2 wb = *reinterpret_cast<const float2*>(wp + q * 8 + 4);
            f[q][0] = make_float2(wa.x, wa.y); f[q][1] = make_float2(wa.z, wa.w); f[q][2] = wb;
          }
          float2 a2[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) a2[j] = make_float2(a[j], a[j]);
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              acc2[c][2 * j] = __ffma2_rn(f[0][c], a2[j], __ffma2_rn(f[1][c], a2[j + 1], acc2[c][2 * j]));
              acc2[c][2 * j + 1] = __ffma2_rn(f[2][c], a2[j + 1], __ffma2_rn(f[3][c], a2[j + 2], acc2[c][2 * j + 1]));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float* o = F + (6 * h + 2 * c) * L + 6 * l;
#pragma unroll
          for (int q = 0; q < 6; q += 2) {
            *reinterpret_cast<float2*>(o + q) = make_float2(acc2[c][q].x, acc2[c][q + 1].x);
            *reinterpret_cast<float2*>(o + L + q) = make_float2(acc2[c][q].y, acc2[c][q + 1].y);
          }
        }
        __syncwarp();
      }
      // ---- LastTransUp.conv3: BN, ReLU, conv5 12 -> 1 on the 96-point grid (pad 2): the whole warp, 3 positions per lane
      {
        const float* F = D + 1152;
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 2
        for (int ci = 0; ci < 12; ++ci) {
          const float s = P[SC + 224 + ci], tt = P[SH + 224 + ci];
          const float* row = F + ci * L + 3 * lane;
          float v[7];
          v[2] = act(row[0], s, tt); v[3] = act(row[1], s, tt); v[4] = act(row[2], s, tt);
          v[0] = __shfl_sync(0xffffffffu, v[3], lane + 31);
          v[1] = __shfl_sync(0xffffffffu, v[4], lane + 31);
          v[5] = __shfl_sync(0xffffffffu, v[2], lane + 1);
          v[6] = __shfl_sync(0xffffffffu, v[3], lane + 1);
          const float* wp = P + W3 + ci * 5;
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const float wk = wp[k];
            o0 = fmaf(wk, v[k], o0); o1 = fmaf(wk, v[k + 1], o1); o2 = fmaf(wk, v[k + 2], o2);
          }
        }
        float* xn = X + (t & 1) * L + 3 * lane;            // the new state replaces the older of the two inputs
        xn[0] = o0; xn[1] = o1; xn[2] = o2;
        float* go = tr + (long long)(c_hist + t) * L + 3 * lane;
        go[0] = o0; go[1] = o1; go[2] = o2;
        __syncwarp();
      }
    }
  }
}

}  // namespace

// true when the program is the reference's KS DenseED (1D-KS-SWAG/main.py defaults) on the 96-point grid
bool arpde_ks_program_matches(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, int H, int W, int c_in, int c_out) {
  if (n_ops != 8 || n_bufs < 3 || H != 1 || W != L || c_in != 2 || c_out != 1) return false;
  const arpde_op_t& s = ops[0];
  if (s.in_buf != -1 || s.cin != 2 || s.cout != 32 || s.kh != 1 || s.kw != 7 || s.stride != 2 || s.up || s.bn_param >= 0 || s.act || s.out_coff != 0) return false;
  const int db = s.out_buf;
  if (db < 0 || bufs[db].ctot != 48) return false;
  for (int j = 0; j < 4; ++j) {
    const arpde_op_t& o = ops[1 + j];
    if (o.in_buf != db || o.out_buf != db || o.cin != 32 + 4 * j || o.cout != 4 || o.kh != 1 || o.kw != 3 || o.stride != 1 || o.up || o.bn_param < 0 ||
        o.act || o.out_coff != o.cin) return false;
  }
  const arpde_op_t &a = ops[5], &b = ops[6], &c = ops[7];
  if (a.in_buf != db || a.cin != 48 || a.cout != 24 || a.kw != 1 || a.kh != 1 || a.stride != 1 || a.up || a.bn_param < 0 || a.act || a.out_coff != 0) return false;
  if (b.in_buf != a.out_buf || b.cin != 24 || b.cout != 12 || b.kw != 3 || b.kh != 1 || b.stride != 1 || !b.up || b.bn_param < 0 || b.act || b.out_coff != 0) return false;
  if (c.in_buf != b.out_buf || c.cin != 12 || c.cout != 1 || c.kw != 5 || c.kh != 1 || c.stride != 1 || c.up || c.bn_param < 0 || c.act || c.out_buf != -2) return false;
  return true;
}

int arpde_ks_rollout_launch(const arpde_op_t* ops, void* const* params, const int64_t* gstride, float* traj, long long bs, int c_hist, int T, int N,
                            int n_per_group, float eps, cudaStream_t stream) {
  KsNet net; memset(&net, 0, sizeof(net));
  int nb = 0;
  for (int i = 0; i < 8; ++i) {
    net.w[i] = (const float*)params[ops[i].w_param]; net.w_gs[i] = gstride[ops[i].w_param];
    if (ops[i].bn_param >= 0) {
      const int p = ops[i].bn_param;
      net.g[nb] = (const float*)params[p]; net.b[nb] = (const float*)params[p + 1];
      net.m[nb] = (const float*)params[p + 2]; net.v[nb] = (const float*)params[p + 3];
      net.g_gs[nb] = gstride[p]; net.b_gs[nb] = gstride[p + 1];
      ARPDE_CHECK_ARG(gstride[p + 2] == 0 && gstride[p + 3] == 0, "ks_rollout: BatchNorm running statistics must be shared by all weight groups");
      ++nb;
    }
  }
  ARPDE_CHECK_ARG(nb == 7, "ks_rollout: expected 7 BatchNorm layers, found %d", nb);
  net.eps = eps;
  if (N <= 0 || T <= 0) return ARPDE_OK;
  const int nsm = arpde_num_sms();
  int kb = (N + nsm - 1) / nsm; if (kb > 16) kb = 16; if (kb < 1) kb = 1;
  int gcap = 1; size_t smem = 0;
  for (;; --kb) {
    gcap = kb == 1 ? 1 : (kb + n_per_group - 2) / n_per_group + 1;             // weight groups a batch of kb consecutive trajectories can span
    smem = ((size_t)gcap * WFLOATS + (size_t)kb * SSTRIDE) * sizeof(float);
    if (smem <= 220 * 1024 || kb == 1) break;
  }
  ARPDE_CHECK_ARG(smem <= 227 * 1024, "ks_rollout: %zu bytes of shared memory needed", smem);
  const int nbatches = (N + kb - 1) / kb;
  const int grid = nbatches < nsm ? nbatches : nsm;
  ARPDE_CUDA(cudaFuncSetAttribute(ks_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ks_rollout_kernel<<<grid, kb * 32, smem, stream>>>(net, traj, bs, c_hist, T, N, n_per_group, kb, nbatches, gcap);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
