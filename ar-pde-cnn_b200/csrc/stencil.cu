// stencil.cu -- fused periodic finite-difference time integrators (Crank-Nicolson / backward
// Euler) for 2D coupled Burgers, 1D Burgers and 1D Kuramoto-Sivashinsky, forward + adjoint,
// with an optional fused sum-of-squared-errors reduction (the physics-constrained loss).
//
// Replaces (reference file:line):
//   2D-Burgers-SWAG/nn/burger2DFiniteDifference.py:148-211  Burger2DIntegrate.{backwardEuler,crankNicolson}
//   1D-Burger-SWAG/nn/burgerFiniteDifference.py:264-302     BurgerIntegrate.{backwardEuler,crankNicolson}
//   1D-KS-SWAG/nn/ksFiniteDifference.py:139-173             KSIntegrate.{backwardEuler,crankNicolson}
// and the autograd graphs torch builds through them (12 / 4 / 6 pad+conv pairs per call).
//
// Layout: fp32, NCHW / NCL.  Inputs may be channel-slice views of a history tensor: each input
// carries its own batch stride (elements); channel stride is H*W (resp. L).  Outputs are compact.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/arpde.h"

// =====================================================================================
// 2D coupled Burgers
//   R_u = -gh(u^2/2) - v gv(u) + nu L(u) ;  R_v = -u gh(v) - gv(v^2/2) + nu L(v)
//   gh = Sobel/8 along W, gv = Sobel/8 along H, L = [[.5,0,.5],[0,-2,0],[.5,0,.5]]
//   separable form used here, per column x with rows a=y-1, b=y, c=y+1 of field f:
//       D_f = c - a,  E_f = a + c,  S_f = a + 2b + c
//   gh(f) = (S_f[x+1]-S_f[x-1])/(8dx); gv(f) = (D_f[x-1]+2D_f[x]+D_f[x+1])/(8dx)
//   L(f)  = (E_f[x-1]+E_f[x+1]-4 f)/(2dx^2)
// =====================================================================================
struct B2Coef {
  float k1;    // 1/(8 dx)
  float k1h;   // 1/(16 dx)     (the 1/2 of u^2/2 folded in)
  float k2;    // nu/(2 dx^2)
  float c;     // dt/2 (CN) or dt (BE)
  int cn;      // 1: Crank-Nicolson (adds R(u0)); 0: backward Euler
};

struct Col6 { float Du, Eu, Squ, Sv, Ev, Dqv; };

__device__ __forceinline__ Col6 colq(float ua, float ub, float uc, float va, float vb, float vc) {
  Col6 q;
  q.Du = uc - ua;
  q.Eu = ua + uc;
  q.Squ = fmaf(2.f * ub, ub, fmaf(ua, ua, uc * uc));
  q.Ev = va + vc;
  q.Sv = fmaf(2.f, vb, q.Ev);
  q.Dqv = fmaf(vc, vc, -va * va);
  return q;
}

__device__ __forceinline__ void rhs_point(const Col6& l, const Col6& m, const Col6& r, float ub, float vb,
                                          const B2Coef& k, float& Ru, float& Rv) {
  const float gh_qu = r.Squ - l.Squ;
  const float gv_u = fmaf(2.f, m.Du, l.Du + r.Du);
  const float lap_u = fmaf(-4.f, ub, l.Eu + r.Eu);
  const float gh_v = r.Sv - l.Sv;
  const float gv_qv = fmaf(2.f, m.Dqv, l.Dqv + r.Dqv);
  const float lap_v = fmaf(-4.f, vb, l.Ev + r.Ev);
  Ru = fmaf(-vb, k.k1 * gv_u, fmaf(-k.k1h, gh_qu, k.k2 * lap_u));
  Rv = fmaf(-ub, k.k1 * gh_v, fmaf(-k.k1h, gv_qv, k.k2 * lap_v));
}

// ---- generic path: one thread per grid point, any H, W >= 1 (L1/L2 serve the 3x3 reuse) ----
__device__ __forceinline__ void rhs_generic(const float* __restrict__ u, const float* __restrict__ v, int H, int W,
                                            int y, int x, const B2Coef& k, float& Ru, float& Rv, float& uc_, float& vc_) {
  const int ym = wrap(y - 1, H), yp = wrap(y + 1, H), xm = wrap(x - 1, W), xp = wrap(x + 1, W);
  const int xs[3] = {xm, x, xp};
  Col6 q[3];
  float ub = 0.f, vb = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float ua = u[ym * W + xs[j]], ubj = u[y * W + xs[j]], uc = u[yp * W + xs[j]];
    const float va = v[ym * W + xs[j]], vbj = v[y * W + xs[j]], vc = v[yp * W + xs[j]];
    q[j] = colq(ua, ubj, uc, va, vbj, vc);
    if (j == 1) { ub = ubj; vb = vbj; }
  }
  rhs_point(q[0], q[1], q[2], ub, vb, k, Ru, Rv);
  uc_ = ub; vc_ = vb;
}

__global__ void __launch_bounds__(256) burgers2d_fwd_generic(const float* __restrict__ up, int64_t up_bs,
                                                             const float* __restrict__ u0, int64_t u0_bs,
                                                             float* __restrict__ ustar, double* __restrict__ sse,
                                                             int N, int H, int W, B2Coef k) {
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * HW;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), rem = (int)(i - (int64_t)n * HW), y = rem / W, x = rem - y * W;
    const float* pu = up + n * up_bs; const float* p0 = u0 + n * u0_bs;
    float Ru, Rv, uu, vv, R0u = 0.f, R0v = 0.f, u0c, v0c;
    rhs_generic(pu, pu + HW, H, W, y, x, k, Ru, Rv, uu, vv);
    if (k.cn) rhs_generic(p0, p0 + HW, H, W, y, x, k, R0u, R0v, u0c, v0c);
    else { u0c = p0[rem]; v0c = p0[HW + rem]; }
    const float su = fmaf(k.c, Ru + R0u, u0c), sv = fmaf(k.c, Rv + R0v, v0c);
    if (ustar) { ustar[(int64_t)n * 2 * HW + rem] = su; ustar[(int64_t)n * 2 * HW + HW + rem] = sv; }
    const float du = su - uu, dv = sv - vv;
    acc = fmaf(du, du, fmaf(dv, dv, acc));
  }
  if (sse) {
    const double t = block_sum_d((double)acc);
    if (threadIdx.x == 0) atomicAdd(sse, t);
  }
}

// ---- fast path: each thread owns 4 consecutive x (one float4 per field per row) and rolls down RY
//      rows keeping rows a,b in registers; halo columns come from the neighbour lanes by shuffle.
//      LX lanes (power of two <= 32) span one row segment of 4*LX points; if the row is wider than
//      one warp (EDGE) lanes 0/31 rebuild the halo column from global memory.
struct Row4 { float u[4], v[4]; };

__device__ __forceinline__ Row4 load_row(const float* __restrict__ pu, const float* __restrict__ pv, int off) {
  Row4 r;
  const float4 a = __ldg(reinterpret_cast<const float4*>(pu + off));
  const float4 b = __ldg(reinterpret_cast<const float4*>(pv + off));
  r.u[0] = a.x; r.u[1] = a.y; r.u[2] = a.z; r.u[3] = a.w;
  r.v[0] = b.x; r.v[1] = b.y; r.v[2] = b.z; r.v[3] = b.w;
  return r;
}

__device__ __forceinline__ Col6 shfl_col(const Col6& q, int src) {
  Col6 o;
  o.Du = __shfl_sync(0xffffffffu, q.Du, src); o.Eu = __shfl_sync(0xffffffffu, q.Eu, src);
  o.Squ = __shfl_sync(0xffffffffu, q.Squ, src); o.Sv = __shfl_sync(0xffffffffu, q.Sv, src);
  o.Ev = __shfl_sync(0xffffffffu, q.Ev, src); o.Dqv = __shfl_sync(0xffffffffu, q.Dqv, src);
  return o;
}

struct Halo2 { float ua, ub, va, vb; };   // rolling rows a,b of one halo column (EDGE lanes only)

template <bool EDGE>
__device__ __forceinline__ void rhs_row4(const Row4& a, const Row4& b, const Row4& c, int lsrc, int rsrc,
                                         bool edgeL, bool edgeR, Halo2& hl, Halo2& hr,
                                         const float* __restrict__ pu, const float* __restrict__ pv, int rowc_off,
                                         int xl, int xr, const B2Coef& k, float* Ru, float* Rv) {
  Col6 q[6];
#pragma unroll
  for (int j = 0; j < 4; ++j) q[j + 1] = colq(a.u[j], b.u[j], c.u[j], a.v[j], b.v[j], c.v[j]);
  q[0] = shfl_col(q[4], lsrc);
  q[5] = shfl_col(q[1], rsrc);
  if (EDGE) {
    if (edgeL) {
      const float uc = __ldg(pu + rowc_off + xl), vc = __ldg(pv + rowc_off + xl);
      q[0] = colq(hl.ua, hl.ub, uc, hl.va, hl.vb, vc);
      hl.ua = hl.ub; hl.ub = uc; hl.va = hl.vb; hl.vb = vc;
    }
    if (edgeR) {
      const float uc = __ldg(pu + rowc_off + xr), vc = __ldg(pv + rowc_off + xr);
      q[5] = colq(hr.ua, hr.ub, uc, hr.va, hr.vb, vc);
      hr.ua = hr.ub; hr.ub = uc; hr.va = hr.vb; hr.vb = vc;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) rhs_point(q[j], q[j + 1], q[j + 2], b.u[j], b.v[j], k, Ru[j], Rv[j]);
}

template <int RY, bool EDGE, int MINB>
__global__ void __launch_bounds__(256, MINB) burgers2d_fwd_vec4(const float* __restrict__ up, int64_t up_bs,
                                                          const float* __restrict__ u0, int64_t u0_bs,
                                                          float* __restrict__ ustar, double* __restrict__ sse,
                                                          int N, int H, int W, int lx_log2, B2Coef k) {
  const int LX = 1 << lx_log2;                  // lanes per row segment
  const int lane = threadIdx.x & 31;
  const int G = 32 >> lx_log2;                  // strip-groups per warp
  const int segs = (W >> 2) >> lx_log2;         // row segments (1 unless EDGE)
  const int nstrips = (H + RY - 1) / RY;
  const int64_t ngroups = (int64_t)N * nstrips * segs;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t gid = warp_id * G + (lane >> lx_log2);
  const bool active = gid < ngroups;
  if (!active) gid = ngroups - 1;               // keep the lane alive for the shuffles
  const int seg = (int)(gid % segs);
  const int strip = (int)((gid / segs) % nstrips);
  const int n = (int)(gid / ((int64_t)segs * nstrips));
  const int lxi = lane & (LX - 1);
  const int x0 = ((seg << lx_log2) + lxi) << 2;
  const int lbase = lane & ~(LX - 1);
  const int lsrc = lbase | ((lxi - 1) & (LX - 1)), rsrc = lbase | ((lxi + 1) & (LX - 1));
  const bool edgeL = EDGE && lxi == 0, edgeR = EDGE && lxi == LX - 1;
  const int xl = wrap(x0 - 1, W), xr = wrap(x0 + 4, W);
  const int HW = H * W;
  const float* pu = up + (int64_t)n * up_bs; const float* pv = pu + HW;
  const float* p0u = u0 + (int64_t)n * u0_bs; const float* p0v = p0u + HW;
  const int y0 = strip * RY;

  Row4 a = load_row(pu, pv, wrap(y0 - 1, H) * W + x0), b = load_row(pu, pv, y0 * W + x0);
  Row4 a0, b0;
  Halo2 hl = {0, 0, 0, 0}, hr = {0, 0, 0, 0}, hl0 = {0, 0, 0, 0}, hr0 = {0, 0, 0, 0};
  if (k.cn) { a0 = load_row(p0u, p0v, wrap(y0 - 1, H) * W + x0); b0 = load_row(p0u, p0v, y0 * W + x0); }
  else { a0 = b0 = a; }
  if (EDGE) {
    const int ra = wrap(y0 - 1, H) * W, rb = y0 * W;
    if (edgeL) { hl = {pu[ra + xl], pu[rb + xl], pv[ra + xl], pv[rb + xl]}; if (k.cn) hl0 = {p0u[ra + xl], p0u[rb + xl], p0v[ra + xl], p0v[rb + xl]}; }
    if (edgeR) { hr = {pu[ra + xr], pu[rb + xr], pv[ra + xr], pv[rb + xr]}; if (k.cn) hr0 = {p0u[ra + xr], p0u[rb + xr], p0v[ra + xr], p0v[rb + xr]}; }
  }
  float acc = 0.f;
  int yc = wrap(y0 + 1, H);
  Row4 c = load_row(pu, pv, yc * W + x0), c0;
  if (k.cn) c0 = load_row(p0u, p0v, yc * W + x0); else c0 = c;
#pragma unroll
  for (int r = 0; r < RY; ++r) {
    const int y = y0 + r;
    // prefetch row y+2 while row y is being computed
    const int yn = wrap(yc + 1, H);
    Row4 nx = load_row(pu, pv, yn * W + x0), nx0;
    if (k.cn) nx0 = load_row(p0u, p0v, yn * W + x0); else nx0 = nx;
    float Ru[4], Rv[4], Su[4], Sv[4];
    rhs_row4<EDGE>(a, b, c, lsrc, rsrc, edgeL, edgeR, hl, hr, pu, pv, yc * W, xl, xr, k, Ru, Rv);
    if (k.cn) {
      float R0u[4], R0v[4];
      rhs_row4<EDGE>(a0, b0, c0, lsrc, rsrc, edgeL, edgeR, hl0, hr0, p0u, p0v, yc * W, xl, xr, k, R0u, R0v);
#pragma unroll
      for (int j = 0; j < 4; ++j) { Su[j] = fmaf(k.c, Ru[j] + R0u[j], b0.u[j]); Sv[j] = fmaf(k.c, Rv[j] + R0v[j], b0.v[j]); }
    } else {
      // backward Euler: the previous state enters only as the additive term
      const Row4 z = load_row(p0u, p0v, wrap(y, H) * W + x0);
#pragma unroll
      for (int j = 0; j < 4; ++j) { Su[j] = fmaf(k.c, Ru[j], z.u[j]); Sv[j] = fmaf(k.c, Rv[j], z.v[j]); }
    }
    if (active && y < H) {
      if (ustar) {
        float* o = ustar + (int64_t)n * 2 * HW + y * W + x0;
        *reinterpret_cast<float4*>(o) = make_float4(Su[0], Su[1], Su[2], Su[3]);
        *reinterpret_cast<float4*>(o + HW) = make_float4(Sv[0], Sv[1], Sv[2], Sv[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float du = Su[j] - b.u[j], dv = Sv[j] - b.v[j];
        acc = fmaf(du, du, fmaf(dv, dv, acc));
      }
    }
    a = b; b = c; c = nx; a0 = b0; b0 = c0; c0 = nx0; yc = yn;
  }
  if (sse) {
    const double t = block_sum_d((double)acc);
    if (threadIdx.x == 0) atomicAdd(sse, t);
  }
}

// ---- shared-memory tiled path (alternative, see the dispatch note in arpde_cn_burgers2d_fwd): persistent CTAs, each walks (sample, strip of TY full-width rows)
//      items.  The (TY+2)-row tiles of u, v, u0, v0 of the NEXT item are prefetched with 16-byte cp.async (coalesced, no register
//      staging, periodic row wrap per chunk) into the second half of a double buffer while the current item is computed, so
//      every CTA keeps a full tile (18 KB at 64x64) in flight all the time; three CTAs per SM hold ~55 KB in flight, above
//      the ~40 KB/SM that 6.5 TB/s x ~0.9 us of HBM latency needs.  (The register-rolling kernel above -- 120 registers, two
//      CTAs per SM, four float4 loads in flight per thread -- stops at 52% of the measured HBM peak; a first shared-memory
//      version that loaded, synchronised and computed each tile in sequence reached only 37%.)  Every thread then computes one
//      float4 of ustar_u and ustar_v: rows y-1, y, y+1 as one float4 plus two wrapped halo scalars per field.  The loss
//      reduction is per-thread accumulation over all items, warp shuffles and one double atomic per CTA.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <bool CN>
__global__ void __launch_bounds__(256) burgers2d_fwd_smem(const float* __restrict__ up, int64_t up_bs, const float* __restrict__ u0,
                                                          int64_t u0_bs, float* __restrict__ ustar, double* __restrict__ sse, int N,
                                                          int H, int W, int TY, B2Coef k) {
  extern __shared__ __align__(16) float tile_all[];      // [2 buffers][4 fields][TY+2][W]
  const int nstrips = (H + TY - 1) / TY;
  const int items = N * nstrips;
  const int rows = TY + 2, w4 = W >> 2;
  const int HW = H * W;
  const int per_field = rows * w4, total = 4 * per_field;
  const int fstride = rows * W, tile_floats = 4 * fstride;

  auto prefetch = [&](int it, float* tile) {
    const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
    const float* up_n = up + (int64_t)n * up_bs;
    const float* u0_n = u0 + (int64_t)n * u0_bs;
    for (int i = threadIdx.x; i < total; i += 256) {
      const int f = i / per_field, r = (i - f * per_field) / w4, c = i - f * per_field - r * w4;
      const int y = wrap(y0 - 1 + r, H);                  // rows y0-1 .. y0+TY of the periodic field
      const float* sp = (f < 2 ? up_n : u0_n) + (f & 1) * HW;
      cp_async16(reinterpret_cast<float4*>(tile) + i, reinterpret_cast<const float4*>(sp + (size_t)y * W) + c);
    }
  };

  float acc = 0.f;
  int buf = 0;
  int it = blockIdx.x;
  if (it < items) prefetch(it, tile_all);
  cp_async_commit();
  for (; it < items; it += gridDim.x) {
    const int nxt = it + gridDim.x;
    if (nxt < items) prefetch(nxt, tile_all + (buf ^ 1) * tile_floats);
    cp_async_commit();
    cp_async_wait<1>();                                   // everything but the newest group (the next item) has landed
    __syncthreads();
    const float* tile = tile_all + buf * tile_floats;
    const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
    for (int i = threadIdx.x; i < TY * w4; i += 256) {
      const int r = i / w4, c = i - r * w4;
      const int y = y0 + r;
      if (y >= H) continue;
      const int x0 = c << 2, xl = (x0 == 0 ? W : x0) - 1, xr = x0 + 4 == W ? 0 : x0 + 4;
      float Ru[4], Rv[4], R0u[4] = {0, 0, 0, 0}, R0v[4] = {0, 0, 0, 0};
      float ub[4], vb[4], u0b[4], v0b[4];
#pragma unroll
      for (int pass = 0; pass < (CN ? 2 : 1); ++pass) {
        const float* tu = tile + (2 * pass) * fstride + r * W;     // rows r, r+1, r+2 of the tile = y-1, y, y+1
        const float* tv = tu + fstride;
        float U[3][6], V[3][6];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          const float4 a = *reinterpret_cast<const float4*>(tu + rr * W + x0);
          const float4 b = *reinterpret_cast<const float4*>(tv + rr * W + x0);
          U[rr][0] = tu[rr * W + xl]; U[rr][1] = a.x; U[rr][2] = a.y; U[rr][3] = a.z; U[rr][4] = a.w; U[rr][5] = tu[rr * W + xr];
          V[rr][0] = tv[rr * W + xl]; V[rr][1] = b.x; V[rr][2] = b.y; V[rr][3] = b.z; V[rr][4] = b.w; V[rr][5] = tv[rr * W + xr];
        }
        Col6 q[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) q[j] = colq(U[0][j], U[1][j], U[2][j], V[0][j], V[1][j], V[2][j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (pass == 0) { rhs_point(q[j], q[j + 1], q[j + 2], U[1][j + 1], V[1][j + 1], k, Ru[j], Rv[j]); ub[j] = U[1][j + 1]; vb[j] = V[1][j + 1]; }
          else { rhs_point(q[j], q[j + 1], q[j + 2], U[1][j + 1], V[1][j + 1], k, R0u[j], R0v[j]); u0b[j] = U[1][j + 1]; v0b[j] = V[1][j + 1]; }
        }
      }
      if (!CN) {
        const float4 a = *reinterpret_cast<const float4*>(tile + 2 * fstride + (r + 1) * W + x0);
        const float4 b = *reinterpret_cast<const float4*>(tile + 3 * fstride + (r + 1) * W + x0);
        u0b[0] = a.x; u0b[1] = a.y; u0b[2] = a.z; u0b[3] = a.w; v0b[0] = b.x; v0b[1] = b.y; v0b[2] = b.z; v0b[3] = b.w;
      }
      float Su[4], Sv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Su[j] = fmaf(k.c, Ru[j] + R0u[j], u0b[j]); Sv[j] = fmaf(k.c, Rv[j] + R0v[j], v0b[j]);
        const float du = Su[j] - ub[j], dv = Sv[j] - vb[j];
        acc = fmaf(du, du, fmaf(dv, dv, acc));
      }
      if (ustar) {
        float* o = ustar + (int64_t)n * 2 * HW + (size_t)y * W + x0;
        *reinterpret_cast<float4*>(o) = make_float4(Su[0], Su[1], Su[2], Su[3]);
        *reinterpret_cast<float4*>(o + HW) = make_float4(Sv[0], Sv[1], Sv[2], Sv[3]);
      }
    }
    __syncthreads();                                      // the buffer is refilled by the prefetch of the next iteration
    buf ^= 1;
  }
  cp_async_wait<0>();
  if (sse) {
    const double t = block_sum_d((double)acc);
    if (threadIdx.x == 0) atomicAdd(sse, t);
  }
}

// ---- bulk-copy ring path (default for the eligible shapes): persistent CTAs of 8 consumer warps + 1 producer warp.  An item is
//      (sample, strip of TY full-width rows); its (TY+2)-row tiles of u, v, u0, v0 are contiguous runs of rows in global memory
//      (split only where the periodic wrap falls), so ONE lane moves a whole stage with <= 12 `cp.async.bulk` copies that complete
//      on the stage's `full` mbarrier -- no per-thread address arithmetic, no registers and no occupancy spent on bytes in flight:
//      NS stages of ~70 KB keep > 100 KB per SM outstanding where the register-rolling kernel had 32 KB (long-scoreboard bound,
//      62 % of the HBM peak) and the cp.async tiled kernel spent 2.2x the instructions on chunk indices.  A consumer thread owns 4
//      consecutive x and rolls down RB rows of the tile: one LDS.128 per field and row, the halo columns as RAW neighbour values by
//      two shuffles per field and row (lanes at a warp boundary of rows wider than 128 points read them from the tile), column
//      sums (colq) of the 6 columns, 4 points out; results go straight from registers to global memory as float4.  Consumers hand a
//      stage back through its `empty` mbarrier (one arrive per warp); nobody waits for anybody else inside the CTA.
#include "tc_ptx.cuh"

struct Row6 { float x[6]; };                       // columns x0-1, x0 .. x0+3, x0+4 of one tile row

// predicated shared-memory load without a branch (the edge lanes of rows wider than one warp): keeps the warp converged
// around the shuffles (with `if (edge) v = trow[x]` ptxas built BSSY/BSYNC regions and slow-path collectives around every row)
__device__ __forceinline__ float lds_if(float v, const float* p, bool pred) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p ld.shared.f32 %0, [%1];\n\t}\n" : "+f"(v) : "r"(smem_u32(p)), "r"((int)pred));
  return v;
}

template <bool EDGE>
__device__ __forceinline__ Row6 lds_row6(const float* __restrict__ trow, int x0, int xl, int xr, int lsrc, int rsrc, bool edgeL, bool edgeR) {
  Row6 r;
  const float4 a = *reinterpret_cast<const float4*>(trow + x0);
  r.x[1] = a.x; r.x[2] = a.y; r.x[3] = a.z; r.x[4] = a.w;
  r.x[0] = __shfl_sync(0xffffffffu, a.w, lsrc);
  r.x[5] = __shfl_sync(0xffffffffu, a.x, rsrc);
  if (EDGE) {
    r.x[0] = lds_if(r.x[0], trow + xl, edgeL);
    r.x[5] = lds_if(r.x[5], trow + xr, edgeR);
  }
  return r;
}

// producer side: rows ystart .. ystart+nrows-1 (periodic) of one field into consecutive tile rows
__device__ __forceinline__ void bulk_rows(float* dst, const float* __restrict__ field, int ystart, int nrows, int H, int W, uint64_t* bar) {
  int y = wrap(ystart, H), left = nrows;
  while (left > 0) {
    const int nr = left < H - y ? left : H - y;
    bulk_g2s(dst, field + (size_t)y * W, (uint32_t)(nr * W) * 4u, bar);
    dst += nr * W; left -= nr; y = 0;
  }
}

struct UnitMap { int x0, xl, xr, lsrc, rsrc, rb; bool edgeL, edgeR, valid; };

template <bool EDGE>
__device__ __forceinline__ UnitMap unit_map(int i, int w4, int nrb, int W, int lx_log2) {
  UnitMap m;
  const int lane = threadIdx.x & 31, LX = 1 << lx_log2;
  const int c = i % w4;
  m.rb = i / w4; m.valid = m.rb < nrb; if (!m.valid) m.rb = nrb - 1;
  m.x0 = c << 2; m.xl = wrap(m.x0 - 1, W); m.xr = wrap(m.x0 + 4, W);
  const int lxi = lane & (LX - 1), lbase = lane & ~(LX - 1);
  m.lsrc = lbase | ((lxi - 1) & (LX - 1)); m.rsrc = lbase | ((lxi + 1) & (LX - 1));
  m.edgeL = EDGE && lane == 0; m.edgeR = EDGE && lane == 31;
  return m;
}

// NCW consumer warps + 1 producer warp
template <int RB, int NCW, bool EDGE, bool CN>
__global__ void __launch_bounds__(32 * (NCW + 1), 1) burgers2d_fwd_bulk(const float* __restrict__ up, int64_t up_bs,
                                                                       const float* __restrict__ u0, int64_t u0_bs,
                                                                       float* __restrict__ ustar, double* __restrict__ sse, int N, int H,
                                                                       int W, int TY, int NS, int lx_log2, B2Coef k) {
  extern __shared__ __align__(128) unsigned char b2_smem[];     // [NS][4 fields][TY+2][W] floats | full[NS] | empty[NS]
  const int rows = TY + 2, fstride = rows * W, stage_floats = 4 * fstride;
  float* tiles = reinterpret_cast<float*>(b2_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(b2_smem + (size_t)NS * stage_floats * sizeof(float));
  uint64_t* empty = full + NS;
  const int nstrips = (H + TY - 1) / TY, items = N * nstrips, HW = H * W;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  if (warp == NCW) {
    if (lane == 0) {
      int j = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++j) {
        const int s = j % NS;
        if (j >= NS) mbar_wait(empty + s, (uint32_t)((j / NS) - 1) & 1u);
        mbar_arrive_expect_tx(full + s, (uint32_t)stage_floats * 4u);
        const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
        const float* pu = up + (int64_t)n * up_bs; const float* p0 = u0 + (int64_t)n * u0_bs;
        float* dst = tiles + (size_t)s * stage_floats;
        bulk_rows(dst, pu, y0 - 1, rows, H, W, full + s);
        bulk_rows(dst + fstride, pu + HW, y0 - 1, rows, H, W, full + s);
        bulk_rows(dst + 2 * fstride, p0, y0 - 1, rows, H, W, full + s);
        bulk_rows(dst + 3 * fstride, p0 + HW, y0 - 1, rows, H, W, full + s);
      }
    }
  } else {
    const int w4 = W >> 2, nrb = TY / RB, nunits = nrb * w4;
    int j = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++j) {
      const int s = j % NS;
      mbar_wait(full + s, (uint32_t)(j / NS) & 1u);
      __syncwarp();
      const float* tile = tiles + (size_t)s * stage_floats;
      const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
      for (int i0 = warp * 32; i0 < nunits; i0 += 32 * NCW) {
        const UnitMap m = unit_map<EDGE>(i0 + lane, w4, nrb, W, lx_log2);
        const int r0 = m.rb * RB;                             // tile row r0 = field row y0 + r0 - 1
        const float* tu = tile + r0 * W; const float* tv = tu + fstride;
        const float* t0u = tu + 2 * fstride; const float* t0v = tu + 3 * fstride;
        // pass 0: R(uPred) of the RB rows (kept in registers), pass 1: R(uPred0), combination, store -- one field pair rolling at a time
        float Su[RB][4], Sv[RB][4];
        {
          Row6 U[3], V[3];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            U[rr + 1] = lds_row6<EDGE>(tu + rr * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
            V[rr + 1] = lds_row6<EDGE>(tv + rr * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
          }
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            U[0] = U[1]; U[1] = U[2]; V[0] = V[1]; V[1] = V[2];
            U[2] = lds_row6<EDGE>(tu + (r + 2) * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
            V[2] = lds_row6<EDGE>(tv + (r + 2) * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
            Col6 q[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) q[c] = colq(U[0].x[c], U[1].x[c], U[2].x[c], V[0].x[c], V[1].x[c], V[2].x[c]);
#pragma unroll
            for (int c = 0; c < 4; ++c) rhs_point(q[c], q[c + 1], q[c + 2], U[1].x[c + 1], V[1].x[c + 1], k, Su[r][c], Sv[r][c]);
          }
        }
        Row6 U0[3], V0[3];
        if (CN) {
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            U0[rr + 1] = lds_row6<EDGE>(t0u + rr * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
            V0[rr + 1] = lds_row6<EDGE>(t0v + rr * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
          }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          float Ou[4], Ov[4];
          if (CN) {
            U0[0] = U0[1]; U0[1] = U0[2]; V0[0] = V0[1]; V0[1] = V0[2];
            U0[2] = lds_row6<EDGE>(t0u + (r + 2) * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
            V0[2] = lds_row6<EDGE>(t0v + (r + 2) * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
            Col6 q[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) q[c] = colq(U0[0].x[c], U0[1].x[c], U0[2].x[c], V0[0].x[c], V0[1].x[c], V0[2].x[c]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float R0u, R0v;
              rhs_point(q[c], q[c + 1], q[c + 2], U0[1].x[c + 1], V0[1].x[c + 1], k, R0u, R0v);
              Ou[c] = fmaf(k.c, Su[r][c] + R0u, U0[1].x[c + 1]); Ov[c] = fmaf(k.c, Sv[r][c] + R0v, V0[1].x[c + 1]);
            }
          } else {
            // backward Euler: the previous state enters only as the additive term
            const float4 a = *reinterpret_cast<const float4*>(t0u + (r + 1) * W + m.x0);
            const float4 b = *reinterpret_cast<const float4*>(t0v + (r + 1) * W + m.x0);
            Ou[0] = fmaf(k.c, Su[r][0], a.x); Ou[1] = fmaf(k.c, Su[r][1], a.y); Ou[2] = fmaf(k.c, Su[r][2], a.z); Ou[3] = fmaf(k.c, Su[r][3], a.w);
            Ov[0] = fmaf(k.c, Sv[r][0], b.x); Ov[1] = fmaf(k.c, Sv[r][1], b.y); Ov[2] = fmaf(k.c, Sv[r][2], b.z); Ov[3] = fmaf(k.c, Sv[r][3], b.w);
          }
          const int y = y0 + r0 + r;
          if (m.valid && y < H) {
            if (ustar) {
              float* o = ustar + (int64_t)n * 2 * HW + (size_t)y * W + m.x0;
              *reinterpret_cast<float4*>(o) = make_float4(Ou[0], Ou[1], Ou[2], Ou[3]);
              *reinterpret_cast<float4*>(o + HW) = make_float4(Ov[0], Ov[1], Ov[2], Ov[3]);
            }
            if (sse) {
              const float4 a = *reinterpret_cast<const float4*>(tu + (r + 1) * W + m.x0);     // uPred at the centre row
              const float4 b = *reinterpret_cast<const float4*>(tv + (r + 1) * W + m.x0);
              const float du[4] = {Ou[0] - a.x, Ou[1] - a.y, Ou[2] - a.z, Ou[3] - a.w}, dv[4] = {Ov[0] - b.x, Ov[1] - b.y, Ov[2] - b.z, Ov[3] - b.w};
#pragma unroll
              for (int c = 0; c < 4; ++c) acc = fmaf(du[c], du[c], fmaf(dv[c], dv[c], acc));
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
  }
  if (sse) {
    const double t = block_sum_d((double)acc);
    if (threadIdx.x == 0) atomicAdd(sse, t);
  }
}

// ---- adjoint, generic: one thread per point.  With G = dL/dustar (gu,gv):
//   d/du = c [ u gh(gu) + gv(v gu) + nu L(gu) - gv_ gh(v) ],  d/dv = c [ -gu gv(u) + gh(u gv_) + v gv(gv_) + nu L(gv_) ]
//   (gh, gv antisymmetric => adjoint = -gh, -gv; L symmetric), plus identity G onto uPred0.
__device__ __forceinline__ void adj_generic(const float* __restrict__ u, const float* __restrict__ v,
                                            const float* __restrict__ gu, const float* __restrict__ gv, int H, int W,
                                            int y, int x, const B2Coef& k, float& du, float& dv) {
  const int ys[3] = {wrap(y - 1, H), y, wrap(y + 1, H)};
  const int xs[3] = {wrap(x - 1, W), x, wrap(x + 1, W)};
  float U[3][3], V[3][3], A[3][3], B[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int o = ys[i] * W + xs[j];
      U[i][j] = u[o]; V[i][j] = v[o]; A[i][j] = gu[o]; B[i][j] = gv[o];
    }
  auto GH = [](float f[3][3]) { return (f[0][2] - f[0][0]) + 2.f * (f[1][2] - f[1][0]) + (f[2][2] - f[2][0]); };
  auto GV = [](float f[3][3]) { return (f[2][0] - f[0][0]) + 2.f * (f[2][1] - f[0][1]) + (f[2][2] - f[0][2]); };
  auto LP = [](float f[3][3]) { return f[0][0] + f[0][2] + f[2][0] + f[2][2] - 4.f * f[1][1]; };
  float VA[3][3], UB[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { VA[i][j] = V[i][j] * A[i][j]; UB[i][j] = U[i][j] * B[i][j]; }
  const float uc = U[1][1], vc = V[1][1], ac = A[1][1], bc = B[1][1];
  du = k.c * (k.k1 * (uc * GH(A) + GV(VA) - bc * GH(V)) + k.k2 * LP(A));
  dv = k.c * (k.k1 * (-ac * GV(U) + GH(UB) + vc * GV(B)) + k.k2 * LP(B));
}

__global__ void __launch_bounds__(256) burgers2d_bwd_generic(const float* __restrict__ up, int64_t up_bs,
                                                             const float* __restrict__ u0, int64_t u0_bs,
                                                             const float* __restrict__ g, float* __restrict__ gup,
                                                             float* __restrict__ gu0, int N, int H, int W, B2Coef k) {
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), rem = (int)(i - (int64_t)n * HW), y = rem / W, x = rem - y * W;
    const float* pu = up + n * up_bs; const float* p0 = u0 + n * u0_bs;
    const float* pg = g + (int64_t)n * 2 * HW;
    float du, dv;
    adj_generic(pu, pu + HW, pg, pg + HW, H, W, y, x, k, du, dv);
    gup[(int64_t)n * 2 * HW + rem] = du; gup[(int64_t)n * 2 * HW + HW + rem] = dv;
    float d0u = 0.f, d0v = 0.f;
    if (k.cn) adj_generic(p0, p0 + HW, pg, pg + HW, H, W, y, x, k, d0u, d0v);
    gu0[(int64_t)n * 2 * HW + rem] = d0u + pg[rem]; gu0[(int64_t)n * 2 * HW + HW + rem] = d0v + pg[HW + rem];
  }
}

// ---- adjoint, shared-memory tiled (same structure as burgers2d_fwd_smem): persistent CTAs walk (sample, strip of TY full-width rows)
//      items; the (TY+2)-row tiles of u, v, gu, gv (and u0, v0 for Crank-Nicolson) of the NEXT item are prefetched with 16-byte cp.async
//      into the second half of a double buffer while the current item is computed.  A thread produces one float4 of each of the four
//      gradient fields from separable column sums (the adjoint of the forward kernel's D/E/S columns):
//        S_f = a + 2b + c, E_f = a + c, D_f = c - a over rows y-1, y, y+1;  GH(f) = S_f[x+1] - S_f[x-1], GV(f) = D_f[x-1] + 2 D_f[x] + D_f[x+1],
//        LP(f) = E_f[x-1] + E_f[x+1] - 4 f.   Algorithmic traffic 40 B / point (read uPred, uPred0, grad_ustar; write two gradients).
struct ColA { float SA, EA, DB, EB; };            // columns of the incoming gradient (shared by the uPred and the uPred0 pass)
struct ColU { float DVA, SV, DU, SUB; };          // columns that involve the state (u, v) of the pass

template <bool CN>
__global__ void __launch_bounds__(256) burgers2d_bwd_smem(const float* __restrict__ up, int64_t up_bs, const float* __restrict__ u0,
                                                          int64_t u0_bs, const float* __restrict__ g, float* __restrict__ gup,
                                                          float* __restrict__ gu0, int N, int H, int W, int TY, B2Coef k) {
  extern __shared__ __align__(16) float tile_all[];      // [2 buffers][NF fields][TY+2][W]; fields: gu, gv, u, v, (u0, v0)
  constexpr int NF = CN ? 6 : 4;
  const int nstrips = (H + TY - 1) / TY;
  const int items = N * nstrips;
  const int rows = TY + 2, w4 = W >> 2;
  const int HW = H * W;
  const int per_field = rows * w4, total = NF * per_field;
  const int fstride = rows * W, tile_floats = NF * fstride;

  auto prefetch = [&](int it, float* tile) {
    const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
    const float* srcs[3] = {g + (int64_t)n * 2 * HW, up + (int64_t)n * up_bs, u0 + (int64_t)n * u0_bs};
    for (int i = threadIdx.x; i < total; i += 256) {
      const int f = i / per_field, r = (i - f * per_field) / w4, c = i - f * per_field - r * w4;
      const int y = wrap(y0 - 1 + r, H);
      const float* sp = srcs[f >> 1] + (f & 1) * HW;
      cp_async16(reinterpret_cast<float4*>(tile) + i, reinterpret_cast<const float4*>(sp + (size_t)y * W) + c);
    }
  };

  int buf = 0;
  int it = blockIdx.x;
  if (it < items) prefetch(it, tile_all);
  cp_async_commit();
  for (; it < items; it += gridDim.x) {
    const int nxt = it + gridDim.x;
    if (nxt < items) prefetch(nxt, tile_all + (buf ^ 1) * tile_floats);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* tile = tile_all + buf * tile_floats;
    const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
    for (int i = threadIdx.x; i < TY * w4; i += 256) {
      const int r = i / w4, c = i - r * w4;
      const int y = y0 + r;
      if (y >= H) continue;
      const int x0 = c << 2, xl = (x0 == 0 ? W : x0) - 1, xr = x0 + 4 == W ? 0 : x0 + 4;
      // six columns x0-1 .. x0+4 of rows y-1, y, y+1 (tile rows r, r+1, r+2)
      auto load6 = [&](int f, float (&F)[3][6]) {
        const float* t = tile + f * fstride + r * W;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          const float4 a = *reinterpret_cast<const float4*>(t + rr * W + x0);
          F[rr][0] = t[rr * W + xl]; F[rr][1] = a.x; F[rr][2] = a.y; F[rr][3] = a.z; F[rr][4] = a.w; F[rr][5] = t[rr * W + xr];
        }
      };
      float A[3][6], B[3][6];
      load6(0, A); load6(1, B);
      ColA qa[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        qa[j].EA = A[0][j] + A[2][j]; qa[j].SA = fmaf(2.f, A[1][j], qa[j].EA);
        qa[j].EB = B[0][j] + B[2][j]; qa[j].DB = B[2][j] - B[0][j];
      }
      float out[4][4];                                     // [gup_u, gup_v, gu0_u, gu0_v][4 columns]
#pragma unroll
      for (int pass = 0; pass < (CN ? 2 : 1); ++pass) {
        float U[3][6], V[3][6];
        load6(2 + 2 * pass, U); load6(3 + 2 * pass, V);
        ColU qu[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          qu[j].DVA = fmaf(V[2][j], A[2][j], -V[0][j] * A[0][j]);
          qu[j].SV = fmaf(2.f, V[1][j], V[0][j] + V[2][j]);
          qu[j].DU = U[2][j] - U[0][j];
          qu[j].SUB = fmaf(U[0][j], B[0][j], fmaf(2.f * U[1][j], B[1][j], U[2][j] * B[2][j]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const ColA &la = qa[j], &ra = qa[j + 2];
          const ColU &lu = qu[j], &mu = qu[j + 1], &ru = qu[j + 2];
          const float uc = U[1][j + 1], vc = V[1][j + 1], ac = A[1][j + 1], bc = B[1][j + 1];
          const float gh_a = ra.SA - la.SA, gv_va = fmaf(2.f, mu.DVA, lu.DVA + ru.DVA), gh_v = ru.SV - lu.SV;
          const float lp_a = fmaf(-4.f, ac, la.EA + ra.EA);
          const float gv_u = fmaf(2.f, mu.DU, lu.DU + ru.DU), gh_ub = ru.SUB - lu.SUB;
          const float gv_b = fmaf(2.f, qa[j + 1].DB, la.DB + ra.DB), lp_b = fmaf(-4.f, bc, la.EB + ra.EB);
          const float du = k.c * fmaf(k.k1, fmaf(uc, gh_a, fmaf(-bc, gh_v, gv_va)), k.k2 * lp_a);
          const float dv = k.c * fmaf(k.k1, fmaf(-ac, gv_u, fmaf(vc, gv_b, gh_ub)), k.k2 * lp_b);
          if (pass == 0) { out[0][j] = du; out[1][j] = dv; }
          else { out[2][j] = du + ac; out[3][j] = dv + bc; }
        }
      }
      if (!CN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { out[2][j] = A[1][j + 1]; out[3][j] = B[1][j + 1]; }
      }
      float* o = gup + (int64_t)n * 2 * HW + (size_t)y * W + x0;
      float* o0 = gu0 + (int64_t)n * 2 * HW + (size_t)y * W + x0;
      *reinterpret_cast<float4*>(o) = make_float4(out[0][0], out[0][1], out[0][2], out[0][3]);
      *reinterpret_cast<float4*>(o + HW) = make_float4(out[1][0], out[1][1], out[1][2], out[1][3]);
      *reinterpret_cast<float4*>(o0) = make_float4(out[2][0], out[2][1], out[2][2], out[2][3]);
      *reinterpret_cast<float4*>(o0 + HW) = make_float4(out[3][0], out[3][1], out[3][2], out[3][3]);
    }
    __syncthreads();
    buf ^= 1;
  }
  cp_async_wait<0>();
}

// ---- adjoint on the bulk-copy ring (same producer / consumer structure as burgers2d_fwd_bulk): tiles of gu, gv, u, v (, u0, v0),
//      a consumer thread rolls down RB rows keeping three rows of every field in registers (the cp.async tiled kernel re-read
//      all three rows plus 36 conflicting halo scalars per output row: its shared-memory pipe was 75 % busy).
template <int RB, int NCW, bool EDGE, bool CN>
__global__ void __launch_bounds__(32 * (NCW + 1), 1) burgers2d_bwd_bulk(const float* __restrict__ up, int64_t up_bs,
                                                                        const float* __restrict__ u0, int64_t u0_bs,
                                                                        const float* __restrict__ g, float* __restrict__ gup,
                                                                        float* __restrict__ gu0, int N, int H, int W, int TY, int NS,
                                                                        int lx_log2, B2Coef k) {
  extern __shared__ __align__(128) unsigned char b2_smem[];     // [NS][NF fields][TY+2][W] floats | full[NS] | empty[NS]
  constexpr int NF = CN ? 6 : 4;
  const int rows = TY + 2, fstride = rows * W, stage_floats = NF * fstride;
  float* tiles = reinterpret_cast<float*>(b2_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(b2_smem + (size_t)NS * stage_floats * sizeof(float));
  uint64_t* empty = full + NS;
  const int nstrips = (H + TY - 1) / TY, items = N * nstrips, HW = H * W;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NCW) {
    if (lane == 0) {
      int j = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x, ++j) {
        const int s = j % NS;
        if (j >= NS) mbar_wait(empty + s, (uint32_t)((j / NS) - 1) & 1u);
        mbar_arrive_expect_tx(full + s, (uint32_t)stage_floats * 4u);
        const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
        const float* srcs[3] = {g + (int64_t)n * 2 * HW, up + (int64_t)n * up_bs, u0 + (int64_t)n * u0_bs};
        float* dst = tiles + (size_t)s * stage_floats;
#pragma unroll
        for (int f = 0; f < NF; ++f) bulk_rows(dst + f * fstride, srcs[f >> 1] + (f & 1) * HW, y0 - 1, rows, H, W, full + s);
      }
    }
    return;
  }
  const int w4 = W >> 2, nrb = TY / RB, nunits = nrb * w4;
  int j = 0;
  for (int it = blockIdx.x; it < items; it += gridDim.x, ++j) {
    const int s = j % NS;
    mbar_wait(full + s, (uint32_t)(j / NS) & 1u);
    __syncwarp();
    const float* tile = tiles + (size_t)s * stage_floats;
    const int n = it / nstrips, y0 = (it - n * nstrips) * TY;
    for (int i0 = warp * 32; i0 < nunits; i0 += 32 * NCW) {
      const UnitMap m = unit_map<EDGE>(i0 + lane, w4, nrb, W, lx_log2);
      const int r0 = m.rb * RB;
      const float* t = tile + r0 * W;
      Row6 F[NF][3];                                          // [gu, gv, u, v, u0, v0][rows y-1, y, y+1]
#pragma unroll
      for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) F[f][rr + 1] = lds_row6<EDGE>(t + f * fstride + rr * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          F[f][0] = F[f][1]; F[f][1] = F[f][2];
          F[f][2] = lds_row6<EDGE>(t + f * fstride + (r + 2) * W, m.x0, m.xl, m.xr, m.lsrc, m.rsrc, m.edgeL, m.edgeR);
        }
        const Row6 *A = F[0], *B = F[1];
        ColA qa[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          qa[c].EA = A[0].x[c] + A[2].x[c]; qa[c].SA = fmaf(2.f, A[1].x[c], qa[c].EA);
          qa[c].EB = B[0].x[c] + B[2].x[c]; qa[c].DB = B[2].x[c] - B[0].x[c];
        }
        float out[4][4];                                      // [gup_u, gup_v, gu0_u, gu0_v][4 columns]
#pragma unroll
        for (int pass = 0; pass < (CN ? 2 : 1); ++pass) {
          const Row6 *U = F[2 + 2 * pass], *V = F[3 + 2 * pass];
          ColU qu[6];
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            qu[c].DVA = fmaf(V[2].x[c], A[2].x[c], -V[0].x[c] * A[0].x[c]);
            qu[c].SV = fmaf(2.f, V[1].x[c], V[0].x[c] + V[2].x[c]);
            qu[c].DU = U[2].x[c] - U[0].x[c];
            qu[c].SUB = fmaf(U[0].x[c], B[0].x[c], fmaf(2.f * U[1].x[c], B[1].x[c], U[2].x[c] * B[2].x[c]));
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const ColA &la = qa[c], &ra = qa[c + 2];
            const ColU &lu = qu[c], &mu = qu[c + 1], &ru = qu[c + 2];
            const float uc = U[1].x[c + 1], vc = V[1].x[c + 1], ac = A[1].x[c + 1], bc = B[1].x[c + 1];
            const float gh_a = ra.SA - la.SA, gv_va = fmaf(2.f, mu.DVA, lu.DVA + ru.DVA), gh_v = ru.SV - lu.SV;
            const float lp_a = fmaf(-4.f, ac, la.EA + ra.EA);
            const float gv_u = fmaf(2.f, mu.DU, lu.DU + ru.DU), gh_ub = ru.SUB - lu.SUB;
            const float gv_b = fmaf(2.f, qa[c + 1].DB, la.DB + ra.DB), lp_b = fmaf(-4.f, bc, la.EB + ra.EB);
            const float du = k.c * fmaf(k.k1, fmaf(uc, gh_a, fmaf(-bc, gh_v, gv_va)), k.k2 * lp_a);
            const float dv = k.c * fmaf(k.k1, fmaf(-ac, gv_u, fmaf(vc, gv_b, gh_ub)), k.k2 * lp_b);
            if (pass == 0) { out[0][c] = du; out[1][c] = dv; }
            else { out[2][c] = du + ac; out[3][c] = dv + bc; }
          }
        }
        if (!CN) {
#pragma unroll
          for (int c = 0; c < 4; ++c) { out[2][c] = A[1].x[c + 1]; out[3][c] = B[1].x[c + 1]; }
        }
        const int y = y0 + r0 + r;
        if (m.valid && y < H) {
          float* o = gup + (int64_t)n * 2 * HW + (size_t)y * W + m.x0;
          float* o0 = gu0 + (int64_t)n * 2 * HW + (size_t)y * W + m.x0;
          *reinterpret_cast<float4*>(o) = make_float4(out[0][0], out[0][1], out[0][2], out[0][3]);
          *reinterpret_cast<float4*>(o + HW) = make_float4(out[1][0], out[1][1], out[1][2], out[1][3]);
          *reinterpret_cast<float4*>(o0) = make_float4(out[2][0], out[2][1], out[2][2], out[2][3]);
          *reinterpret_cast<float4*>(o0 + HW) = make_float4(out[3][0], out[3][1], out[3][2], out[3][3]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
}

// shape rule of the bulk-copy ring kernels: whole float4s, 16-byte aligned bases and strides (checked by the callers), W/4 a power of
// two <= 32 or a multiple of 32 (lane mapping of the halo shuffles).  TY = RB x (consumer threads / (W/4)) rows gives every consumer thread one
// unit per item; it is halved while the launch would have fewer than two items per SM (small batches), never below RB.
struct BulkPlan { int TY, NS, lx_log2; bool edge; size_t smem; int64_t items; };
static bool bulk_plan(int N, int H, int W, int RB, int NCW, int NF, BulkPlan* p) {
  const int w4 = W >> 2;
  if (W % 4 != 0 || w4 < 1 || H < 2) return false;
  const bool pow2 = (w4 & (w4 - 1)) == 0;
  if (!((pow2 && w4 <= 32) || w4 % 32 == 0)) return false;
  const int cthreads = 32 * NCW;
  int TY = RB * (w4 >= cthreads ? 1 : cthreads / w4);
  const int hr = (H + RB - 1) / RB * RB; if (TY > hr) TY = hr;
  const int sms = arpde_num_sms();
  while (TY > RB && (TY / 2) % RB == 0 && (int64_t)N * ((H + TY - 1) / TY) < 2 * sms) TY /= 2;
  size_t stage = (size_t)NF * (TY + 2) * W * sizeof(float);
  while (stage * 2 + 256 > 227 * 1024 && TY > RB && (TY / 2) % RB == 0) { TY /= 2; stage = (size_t)NF * (TY + 2) * W * sizeof(float); }
  if (stage * 2 + 256 > 227 * 1024 || stage >= (1u << 20)) return false;
  int NS = (int)((227 * 1024 - 256) / stage); if (NS > 4) NS = 4;
  if (arpde_dev_env("ARPDE_BULK_TY")) {
    const int t = atoi(arpde_dev_env("ARPDE_BULK_TY"));
    if (t >= RB && t % RB == 0) { TY = t; stage = (size_t)NF * (TY + 2) * W * sizeof(float); NS = (int)((227 * 1024 - 256) / stage); if (NS > 4) NS = 4; if (NS < 2) return false; }
  }
  if (arpde_dev_env("ARPDE_BULK_NS")) { const int t = atoi(arpde_dev_env("ARPDE_BULK_NS")); if (t >= 1 && (size_t)t * stage + 256 <= 227 * 1024) NS = t; }
  p->TY = TY; p->NS = NS; p->smem = stage * NS + 2 * NS * sizeof(uint64_t);
  p->lx_log2 = 0; while ((1 << p->lx_log2) < (w4 < 32 ? w4 : 32)) ++p->lx_log2;
  p->edge = w4 > 32;
  p->items = (int64_t)N * ((H + TY - 1) / TY);
  return p->items < (1ll << 31);
}

extern "C" int arpde_stencil_bulk_plan(int N, int H, int W, int backward, int mode, int* rows_per_item, int* stages, int64_t* smem_bytes,
                                       int64_t* items) {
  if (N < 1 || H < 1 || W < 1 || (mode != ARPDE_CN && mode != ARPDE_BE)) return 0;
  BulkPlan bp;
  const bool ok = backward ? bulk_plan(N, H, W, 2, 8, mode == ARPDE_CN ? 6 : 4, &bp) : bulk_plan(N, H, W, 4, 8, 4, &bp);
  if (!ok) return 0;
  if (rows_per_item) *rows_per_item = bp.TY;
  if (stages) *stages = bp.NS;
  if (smem_bytes) *smem_bytes = (int64_t)bp.smem;
  if (items) *items = bp.items;
  return 1;
}

static int make_b2coef(double dx, double nu, double dt, int mode, B2Coef* k) {
  ARPDE_CHECK_ARG(dx > 0 && (mode == ARPDE_CN || mode == ARPDE_BE), "burgers2d: bad dx=%g or mode=%d", dx, mode);
  k->k1 = (float)(0.125 / dx); k->k1h = (float)(0.0625 / dx); k->k2 = (float)(nu * 0.5 / (dx * dx));
  k->cn = mode == ARPDE_CN; k->c = (float)(k->cn ? 0.5 * dt : dt);
  return ARPDE_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int arpde_cn_burgers2d_fwd(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                                      float* ustar, double* sse_out, int N, int H, int W, double dx, double nu,
                                      double dt, int mode, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(N >= 0 && H >= 1 && W >= 1, "burgers2d_fwd: bad dims N=%d H=%d W=%d", N, H, W);
  B2Coef k; int rc = make_b2coef(dx, nu, dt, mode, &k); if (rc) return rc;
  if (sse_out) ARPDE_CUDA(cudaMemsetAsync(sse_out, 0, sizeof(double), (cudaStream_t)stream_));
  if (N == 0) return ARPDE_OK;                       // empty batch: nothing to do (pointers may be null)
  ARPDE_CHECK_ARG(uPred && uPred0 && (ustar || sse_out), "burgers2d_fwd: null pointer");
  const int w4 = W >> 2;
  const bool pow2 = w4 > 0 && (w4 & (w4 - 1)) == 0;
  const bool vec_ok = (W % 4 == 0) && H >= 2 && aligned16(uPred) && aligned16(uPred0) && (!ustar || aligned16(ustar)) &&
                      up_bstride % 4 == 0 && u0_bstride % 4 == 0 && (pow2 || w4 % 32 == 0);
  const bool base_ok = (W % 4 == 0) && aligned16(uPred) && aligned16(uPred0) && (!ustar || aligned16(ustar)) && up_bstride % 4 == 0 &&
                       u0_bstride % 4 == 0 && (int64_t)H * W % 4 == 0;
  static const int ty_env = arpde_dev_env("ARPDE_STENCIL_TY") ? atoi(arpde_dev_env("ARPDE_STENCIL_TY")) : 0;
  // rows wider than one warp's 128 columns: the rolling kernel needs its EDGE variant (halo scalars from global memory, 152 registers,
  // one CTA per SM: 2.3-2.4 TB/s at 256x256); the tiled kernel with the tallest tile that fits 96 KB reaches 2.9 TB/s there
  // (TY = 4 / 6 / 8 / 10 rows: 2.58 / 2.86 / 2.88 / 2.94 TB/s)
  const bool wide = W > 128;
  int TY = ty_env ? ty_env : (W > 0 ? (wide ? 3072 / W - 2 : 1024 / W) : 0); if (TY > H) TY = H;
  const size_t tile_bytes = (size_t)2 * 4 * (TY + 2) * W * sizeof(float);   // double buffer
  // Dispatch: the register-rolling kernel (RY = 4 rows per thread, 120 registers, no spills) is the measured best at 64x64
  // (3.96 TB/s = 60% of the measured HBM peak; RY = 8: 3.4, RY = 2: 3.6; forcing 3 CTAs/SM spills and loses 6%); the
  // shared-memory tiled kernel is issue-bound (ncu: 70% issue-active, 2.2x the instructions, 48% of its shared-memory wavefronts
  // are bank conflicts of the halo scalars) and reaches 2.75 TB/s -- it serves widths the rolling kernel's lane mapping cannot
  // (W/4 neither a power of two nor a multiple of 32) and is selectable with ARPDE_STENCIL_SMEM=1 for A/B runs.
  static const bool force_smem = arpde_dev_env("ARPDE_STENCIL_SMEM") != nullptr;
  const bool no_bulk = arpde_dev_env("ARPDE_STENCIL_NOBULK") != nullptr;
  BulkPlan bp;
  // 8 consumer warps x 4 rows per unit (measured against 16 warps x 2 rows at 88 registers: 5.04 vs 4.78 TB/s at 64x64, 5.73 vs 5.49 at 256x256)
  if (base_ok && !no_bulk && !force_smem && bulk_plan(N, H, W, 4, 8, 4, &bp)) {
    // default: bulk-copy ring kernel (one persistent CTA per SM)
    int64_t blocks = arpde_num_sms(); if (blocks > bp.items) blocks = bp.items;
#define LAUNCH_BULK(EDGE_, CN_)                                                                                                          \
    do {                                                                                                                                 \
      ARPDE_CUDA(cudaFuncSetAttribute(burgers2d_fwd_bulk<4, 8, EDGE_, CN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.smem)); \
      burgers2d_fwd_bulk<4, 8, EDGE_, CN_><<<(unsigned)blocks, 32 * 9, bp.smem, (cudaStream_t)stream_>>>(                                 \
          uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, bp.TY, bp.NS, bp.lx_log2, k);                                  \
    } while (0)
    if (bp.edge) { if (k.cn) LAUNCH_BULK(true, true); else LAUNCH_BULK(true, false); }
    else { if (k.cn) LAUNCH_BULK(false, true); else LAUNCH_BULK(false, false); }
#undef LAUNCH_BULK
  } else if (base_ok && TY >= 1 && tile_bytes <= 96 * 1024 && (force_smem || !vec_ok || (wide && TY >= 4))) {
    const int nstrips = (H + TY - 1) / TY;
    const int64_t items = (int64_t)N * nstrips;
    ARPDE_CHECK_ARG(items < (1ll << 31), "burgers2d_fwd: too many strips");
    int per_sm = (int)((200 * 1024) / (tile_bytes + 1024)); if (per_sm > 4) per_sm = 4; if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)arpde_num_sms() * per_sm; if (blocks > items) blocks = items;
    if (k.cn) {
      if (tile_bytes > 40 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(burgers2d_fwd_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
      burgers2d_fwd_smem<true><<<(unsigned)blocks, 256, tile_bytes, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, TY, k);
    } else {
      if (tile_bytes > 40 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(burgers2d_fwd_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
      burgers2d_fwd_smem<false><<<(unsigned)blocks, 256, tile_bytes, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, TY, k);
    }
  } else if (vec_ok) {
    static const int ry_env = arpde_dev_env("ARPDE_STENCIL_RY") ? atoi(arpde_dev_env("ARPDE_STENCIL_RY")) : 0;
    const int RY = ry_env ? ry_env : 4;
    static const int minb = arpde_dev_env("ARPDE_STENCIL_MINB") ? atoi(arpde_dev_env("ARPDE_STENCIL_MINB")) : 2;
    int lx_log2 = 0; while ((1 << lx_log2) < (w4 < 32 ? w4 : 32)) ++lx_log2;
    const bool edge = w4 > 32;
    const int G = 32 >> lx_log2, segs = w4 >> lx_log2, nstrips = (H + RY - 1) / RY;
    const int64_t ngroups = (int64_t)N * nstrips * segs, nwarps = ceil_div64(ngroups, G);
    const int64_t blocks = ceil_div64(nwarps, 8);
#define LAUNCH_ROLL(RY_)                                                                                                                  \
    do {                                                                                                                                  \
      if (edge) burgers2d_fwd_vec4<RY_, true, 1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, lx_log2, k); \
      else if (minb == 2) burgers2d_fwd_vec4<RY_, false, 2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, lx_log2, k);    \
      else if (minb == 4) burgers2d_fwd_vec4<RY_, false, 4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, lx_log2, k);    \
      else burgers2d_fwd_vec4<RY_, false, 3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, lx_log2, k);    \
    } while (0)
    if (RY == 2) LAUNCH_ROLL(2); else if (RY == 3) LAUNCH_ROLL(3); else if (RY == 4) LAUNCH_ROLL(4); else if (RY == 5) LAUNCH_ROLL(5); else if (RY == 6) LAUNCH_ROLL(6); else LAUNCH_ROLL(8);
#undef LAUNCH_ROLL
  } else {
    const int64_t total = (int64_t)N * H * W;
    int64_t blocks = ceil_div64(total, 256); const int64_t cap = (int64_t)arpde_num_sms() * 32; if (blocks > cap) blocks = cap;
    burgers2d_fwd_generic<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, k);
  }
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// test hook: force the generic kernel (parity of the two paths is a test)
extern "C" int arpde_cn_burgers2d_fwd_generic(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                                              float* ustar, double* sse_out, int N, int H, int W, double dx, double nu,
                                              double dt, int mode, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(uPred && uPred0 && (ustar || sse_out) && N > 0, "burgers2d_fwd_generic: bad args");
  B2Coef k; int rc = make_b2coef(dx, nu, dt, mode, &k); if (rc) return rc;
  if (sse_out) ARPDE_CUDA(cudaMemsetAsync(sse_out, 0, sizeof(double), (cudaStream_t)stream_));
  const int64_t total = (int64_t)N * H * W;
  int64_t blocks = ceil_div64(total, 256); const int64_t cap = (int64_t)arpde_num_sms() * 32; if (blocks > cap) blocks = cap;
  burgers2d_fwd_generic<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, H, W, k);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

static int burgers2d_bwd_impl(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                              const float* grad_ustar, float* grad_uPred, float* grad_uPred0, int N, int H, int W,
                              double dx, double nu, double dt, int mode, arpde_stream_t stream_, bool force_generic) {
  ARPDE_CHECK_ARG(N >= 0 && H >= 1 && W >= 1, "burgers2d_bwd: bad dims");
  B2Coef k; int rc = make_b2coef(dx, nu, dt, mode, &k); if (rc) return rc;
  if (N == 0) return ARPDE_OK;
  ARPDE_CHECK_ARG(uPred && uPred0 && grad_ustar && grad_uPred && grad_uPred0, "burgers2d_bwd: null pointer");
  // vector path: whole float4s per row, every base and batch stride 16-byte aligned, a double-buffered tile of >= 2 rows within 96 KB
  const bool base_ok = (W % 4 == 0) && ((int64_t)H * W) % 4 == 0 && aligned16(uPred) && aligned16(uPred0) && aligned16(grad_ustar) &&
                       aligned16(grad_uPred) && aligned16(grad_uPred0) && up_bstride % 4 == 0 && u0_bstride % 4 == 0 && H >= 2;
  const int NF = k.cn ? 6 : 4;
  int TY = W > 0 ? (W > 128 ? (96 * 1024) / (2 * NF * W * 4) - 2 : 1024 / W) : 0; if (TY > H) TY = H;
  const size_t tile_bytes = (size_t)2 * NF * (TY + 2) * W * sizeof(float);
  const bool no_bulk = arpde_dev_env("ARPDE_STENCIL_NOBULK") != nullptr;
  BulkPlan bp;
  if (base_ok && !force_generic && !no_bulk && bulk_plan(N, H, W, 2, 8, NF, &bp)) {
    int64_t blocks = arpde_num_sms(); if (blocks > bp.items) blocks = bp.items;
#define LAUNCH_BULK(EDGE_, CN_)                                                                                                          \
    do {                                                                                                                                 \
      ARPDE_CUDA(cudaFuncSetAttribute(burgers2d_bwd_bulk<2, 8, EDGE_, CN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.smem)); \
      burgers2d_bwd_bulk<2, 8, EDGE_, CN_><<<(unsigned)blocks, 32 * 9, bp.smem, (cudaStream_t)stream_>>>(                                 \
          uPred, up_bstride, uPred0, u0_bstride, grad_ustar, grad_uPred, grad_uPred0, N, H, W, bp.TY, bp.NS, bp.lx_log2, k);             \
    } while (0)
    if (bp.edge) { if (k.cn) LAUNCH_BULK(true, true); else LAUNCH_BULK(true, false); }
    else { if (k.cn) LAUNCH_BULK(false, true); else LAUNCH_BULK(false, false); }
#undef LAUNCH_BULK
    ARPDE_LAUNCH_CHECK();
    return ARPDE_OK;
  }
  if (base_ok && TY >= 2 && tile_bytes <= 96 * 1024 && !force_generic) {
    const int nstrips = (H + TY - 1) / TY;
    const int64_t items = (int64_t)N * nstrips;
    ARPDE_CHECK_ARG(items < (1ll << 31), "burgers2d_bwd: too many strips");
    int per_sm = (int)((200 * 1024) / (tile_bytes + 1024)); if (per_sm > 4) per_sm = 4; if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)arpde_num_sms() * per_sm; if (blocks > items) blocks = items;
    if (k.cn) {
      if (tile_bytes > 40 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(burgers2d_bwd_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
      burgers2d_bwd_smem<true><<<(unsigned)blocks, 256, tile_bytes, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, grad_ustar, grad_uPred, grad_uPred0, N, H, W, TY, k);
    } else {
      if (tile_bytes > 40 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(burgers2d_bwd_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
      burgers2d_bwd_smem<false><<<(unsigned)blocks, 256, tile_bytes, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, grad_ustar, grad_uPred, grad_uPred0, N, H, W, TY, k);
    }
    ARPDE_LAUNCH_CHECK();
    return ARPDE_OK;
  }
  const int64_t total = (int64_t)N * H * W;
  int64_t blocks = ceil_div64(total, 256); const int64_t cap = (int64_t)arpde_num_sms() * 32; if (blocks > cap) blocks = cap;
  burgers2d_bwd_generic<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, grad_ustar, grad_uPred, grad_uPred0, N, H, W, k);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

extern "C" int arpde_cn_burgers2d_bwd(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                                      const float* grad_ustar, float* grad_uPred, float* grad_uPred0, int N, int H, int W,
                                      double dx, double nu, double dt, int mode, arpde_stream_t stream) {
  return burgers2d_bwd_impl(uPred, up_bstride, uPred0, u0_bstride, grad_ustar, grad_uPred, grad_uPred0, N, H, W, dx, nu, dt, mode, stream, false);
}

// test hook: the one-thread-per-point adjoint (any H, W); parity of the two kernels is a test
extern "C" int arpde_cn_burgers2d_bwd_generic(const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                                              const float* grad_ustar, float* grad_uPred, float* grad_uPred0, int N, int H, int W,
                                              double dx, double nu, double dt, int mode, arpde_stream_t stream) {
  return burgers2d_bwd_impl(uPred, up_bstride, uPred0, u0_bstride, grad_ustar, grad_uPred, grad_uPred0, N, H, W, dx, nu, dt, mode, stream, true);
}

// =====================================================================================
// 1D systems (periodic, one channel): tap tables up to radius 3.
//   Burgers: ustar = u0 - c[(u g(u) + u0 g(u0)) - nu (L(u)+L(u0))]           (CN)
//            ustar = u0 + dt(-0.25 g(u^2) + nu L(u))                          (BE, literal)
//   KS:      ustar = u0 - c[(g(u^2)/2 + L(u) + nu D4(u)) + (same at u0)]      (CN; BE drops u0 rhs, c=dt)
// =====================================================================================
struct Taps1D {
  float g[7], l[7], d[7];   // centred tap tables (index 3 = centre), pre-divided, times 1/dx^n
  int r;                    // max radius used
  float nu, c;
  int cn, system;
};

__device__ __forceinline__ void load7(const float* __restrict__ p, int i, int L, int r, float* w) {
#pragma unroll
  for (int t = -3; t <= 3; ++t) w[t + 3] = (t >= -r && t <= r) ? p[wrap(i + t, L)] : 0.f;
}

__device__ __forceinline__ float rhs1d(const float* w, const Taps1D& k) {
  float g = 0.f, l = 0.f, d = 0.f;
  if (k.system == ARPDE_SYS_BURGERS1D && k.cn) {
#pragma unroll
    for (int t = 0; t < 7; ++t) { g = fmaf(k.g[t], w[t], g); l = fmaf(k.l[t], w[t], l); }
    return fmaf(w[3], g, -k.nu * l);                       // u g(u) - nu L(u)
  }
#pragma unroll
  for (int t = 0; t < 7; ++t) { g = fmaf(k.g[t], w[t] * w[t], g); l = fmaf(k.l[t], w[t], l); d = fmaf(k.d[t], w[t], d); }
  if (k.system == ARPDE_SYS_BURGERS1D) return fmaf(0.25f, g, -k.nu * l);   // BE: 0.5*0.5 g(u^2) - nu L(u)
  return fmaf(0.5f, g, fmaf(k.nu, d, l));                                  // KS
}

__global__ void __launch_bounds__(256) stencil1d_fwd(const float* __restrict__ up, int64_t up_bs, const float* __restrict__ u0,
                                                     int64_t u0_bs, float* __restrict__ ustar, double* __restrict__ sse,
                                                     int N, int L, Taps1D k) {
  const int64_t total = (int64_t)N * L;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / L), x = (int)(i - (int64_t)n * L);
    float w[7];
    load7(up + n * up_bs, x, L, k.r, w);
    const float uc = w[3];
    float R = rhs1d(w, k), u0c;
    if (k.cn) { load7(u0 + n * u0_bs, x, L, k.r, w); R += rhs1d(w, k); u0c = w[3]; }
    else u0c = u0[n * u0_bs + x];
    const float s = fmaf(-k.c, R, u0c);
    if (ustar) ustar[i] = s;
    const float d = s - uc;
    acc = fmaf(d, d, acc);
  }
  if (sse) {
    const double t = block_sum_d((double)acc);
    if (threadIdx.x == 0) atomicAdd(sse, t);
  }
}

// adjoint of -c*rhs: g taps antisymmetric (g^T = -g), l and d symmetric.
//   Burgers CN : d/du = -c[ G g(u) - g(u G) - nu L(G) ]
//   Burgers BE : d/du = -c[ -0.5 u g(G) - nu L(G) ]
//   KS         : d/du = -c[ -u g(G) + L(G) + nu D4(G) ]
__device__ __forceinline__ float adj1d(const float* wu, const float* wg, const Taps1D& k) {
  float gG = 0.f, lG = 0.f, dG = 0.f, guG = 0.f, gu = 0.f;
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    gG = fmaf(k.g[t], wg[t], gG); lG = fmaf(k.l[t], wg[t], lG); dG = fmaf(k.d[t], wg[t], dG);
    guG = fmaf(k.g[t], wu[t] * wg[t], guG); gu = fmaf(k.g[t], wu[t], gu);
  }
  if (k.system == ARPDE_SYS_BURGERS1D && k.cn) return -k.c * (wg[3] * gu - guG - k.nu * lG);
  if (k.system == ARPDE_SYS_BURGERS1D) return -k.c * (-0.5f * wu[3] * gG - k.nu * lG);
  return -k.c * (-wu[3] * gG + lG + k.nu * dG);
}

__global__ void __launch_bounds__(256) stencil1d_bwd(const float* __restrict__ up, int64_t up_bs, const float* __restrict__ u0,
                                                     int64_t u0_bs, const float* __restrict__ g, float* __restrict__ gup,
                                                     float* __restrict__ gu0, int N, int L, Taps1D k) {
  const int64_t total = (int64_t)N * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / L), x = (int)(i - (int64_t)n * L);
    float wu[7], wg[7];
    load7(g + (int64_t)n * L, x, L, k.r, wg);
    load7(up + n * up_bs, x, L, k.r, wu);
    gup[i] = adj1d(wu, wg, k);
    float d0 = wg[3];
    if (k.cn) { load7(u0 + n * u0_bs, x, L, k.r, wu); d0 += adj1d(wu, wg, k); }
    gu0[i] = d0;
  }
}

static void put_taps(float* dst, const double* taps, int n, double den, double scale) {
  for (int i = 0; i < 7; ++i) dst[i] = 0.f;
  const int r = (n - 1) / 2;
  for (int i = 0; i < n; ++i) {
    const float w = (float)taps[i] / (float)den;            // fp32 division as the reference builds its weights
    dst[3 - r + i] = (float)((double)w / scale);
  }
}

static int make_taps(int system, int k1, int k2, int k4, double dx, double nu, double dt, int mode, Taps1D* t) {
  static const double G3[] = {-1, 0, 1}, G5[] = {1, -8, 0, 8, -1}, G7[] = {-1, 9, -45, 0, 45, -9, 1};
  static const double L3[] = {1, -2, 1}, L5[] = {-1, 16, -30, 16, -1};
  static const double D5[] = {1, -4, 6, -4, 1}, D7[] = {-1, 12, -39, 56, -39, 12, -1};
  ARPDE_CHECK_ARG(system == ARPDE_SYS_BURGERS1D || system == ARPDE_SYS_KS1D, "stencil1d: bad system %d", system);
  ARPDE_CHECK_ARG(mode == ARPDE_CN || mode == ARPDE_BE, "stencil1d: bad mode %d", mode);
  ARPDE_CHECK_ARG(dx > 0, "stencil1d: dx must be > 0");
  memset(t, 0, sizeof(*t));
  if (k1 == 3) put_taps(t->g, G3, 3, 2, dx); else if (k1 == 5) put_taps(t->g, G5, 5, 12, dx);
  else if (k1 == 7 && system == ARPDE_SYS_BURGERS1D) put_taps(t->g, G7, 7, 60, dx);
  else { arpde_set_error("stencil1d: unsupported gradient kernel size %d", k1); return ARPDE_ERR_UNSUPPORTED; }
  if (k2 == 3) put_taps(t->l, L3, 3, 1, dx * dx); else if (k2 == 5) put_taps(t->l, L5, 5, 12, dx * dx);
  else { arpde_set_error("stencil1d: unsupported laplacian kernel size %d", k2); return ARPDE_ERR_UNSUPPORTED; }
  int r = ((k1 > k2 ? k1 : k2) - 1) / 2;
  if (system == ARPDE_SYS_KS1D) {
    if (k4 == 5) put_taps(t->d, D5, 5, 1, dx * dx * dx * dx); else if (k4 == 7) put_taps(t->d, D7, 7, 6, dx * dx * dx * dx);
    else { arpde_set_error("stencil1d: unsupported 4th-derivative kernel size %d", k4); return ARPDE_ERR_UNSUPPORTED; }
    if ((k4 - 1) / 2 > r) r = (k4 - 1) / 2;
  }
  t->r = r; t->nu = (float)nu; t->cn = mode == ARPDE_CN; t->c = (float)(t->cn ? 0.5 * dt : dt); t->system = system;
  return ARPDE_OK;
}

extern "C" int arpde_cn_1d_fwd(int system, const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                               float* ustar, double* sse_out, int N, int L, double dx, double nu, double dt, int k1, int k2,
                               int k4, int mode, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(N >= 0 && L >= 1, "cn_1d_fwd: bad dims N=%d L=%d", N, L);
  ARPDE_CHECK_ARG(N == 0 || (uPred && uPred0 && (ustar || sse_out)), "cn_1d_fwd: null pointer");
  Taps1D t; int rc = make_taps(system, k1, k2, k4, dx, nu, dt, mode, &t); if (rc) return rc;
  ARPDE_CHECK_ARG(L > t.r, "cn_1d_fwd: L=%d must exceed the stencil radius %d", L, t.r);
  if (sse_out) ARPDE_CUDA(cudaMemsetAsync(sse_out, 0, sizeof(double), (cudaStream_t)stream_));
  if (N == 0) return ARPDE_OK;
  int64_t blocks = ceil_div64((int64_t)N * L, 256); const int64_t cap = (int64_t)arpde_num_sms() * 32; if (blocks > cap) blocks = cap;
  stencil1d_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, ustar, sse_out, N, L, t);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

extern "C" int arpde_cn_1d_bwd(int system, const float* uPred, int64_t up_bstride, const float* uPred0, int64_t u0_bstride,
                               const float* grad_ustar, float* grad_uPred, float* grad_uPred0, int N, int L, double dx,
                               double nu, double dt, int k1, int k2, int k4, int mode, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(N >= 0 && L >= 1, "cn_1d_bwd: bad dims");
  ARPDE_CHECK_ARG(N == 0 || (uPred && uPred0 && grad_ustar && grad_uPred && grad_uPred0), "cn_1d_bwd: null pointer");
  Taps1D t; int rc = make_taps(system, k1, k2, k4, dx, nu, dt, mode, &t); if (rc) return rc;
  ARPDE_CHECK_ARG(L > t.r, "cn_1d_bwd: L=%d must exceed the stencil radius %d", L, t.r);
  if (N == 0) return ARPDE_OK;
  int64_t blocks = ceil_div64((int64_t)N * L, 256); const int64_t cap = (int64_t)arpde_num_sms() * 32; if (blocks > cap) blocks = cap;
  stencil1d_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(uPred, up_bstride, uPred0, u0_bstride, grad_ustar, grad_uPred, grad_uPred0, N, L, t);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

// =====================================================================================
// stand-alone periodic filters (the callable filter objects .grad1/.grad2/.grad4 of the
// reference integrators: GradFilter1D, LaplaceFilter1D, FourthOrderFilter1D, SobelFilter2d.grad_h/
// grad_v, LapLaceFilter2d).  Every channel is filtered independently.
// =====================================================================================
struct W7 { float w[7]; };
struct W9 { float w[9]; };

__global__ void __launch_bounds__(256) filter1d_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int L, W7 k) {
  const long long total = rows * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / L; const int c = (int)(i - r * L);
    const float* p = x + r * L;
    float acc = 0.f;
#pragma unroll
    for (int t = -3; t <= 3; ++t)
      if (k.w[t + 3] != 0.f) { int j = (c + t) % L; if (j < 0) j += L; acc = fmaf(k.w[t + 3], p[j], acc); }
    y[i] = acc;
  }
}

__global__ void __launch_bounds__(256) filter2d_kernel(const float* __restrict__ x, float* __restrict__ y, long long imgs, int H, int W, W9 k) {
  const long long HW = (long long)H * W, total = imgs * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / HW; const int rem = (int)(i - n * HW), yy = rem / W, xx = rem - yy * W;
    const float* p = x + n * HW;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        int r = (yy + a - 1) % H; if (r < 0) r += H;
        int c = (xx + b - 1) % W; if (c < 0) c += W;
        acc = fmaf(k.w[a * 3 + b], p[r * W + c], acc);
      }
    y[i] = acc;
  }
}

extern "C" int arpde_filter1d(const float* x, float* y, int64_t rows, int L, const float* taps7_host, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(x && y && taps7_host && rows > 0 && L > 0, "filter1d: bad args");
  W7 k; for (int i = 0; i < 7; ++i) k.w[i] = taps7_host[i];
  int64_t blocks = ceil_div64(rows * L, 256); const int64_t cap = (int64_t)arpde_num_sms() * 32; if (blocks > cap) blocks = cap;
  filter1d_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, y, rows, L, k);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}

extern "C" int arpde_filter2d(const float* x, float* y, int64_t imgs, int H, int W, const float* w9_host, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(x && y && w9_host && imgs > 0 && H > 0 && W > 0, "filter2d: bad args");
  W9 k; for (int i = 0; i < 9; ++i) k.w[i] = w9_host[i];
  int64_t blocks = ceil_div64(imgs * H * W, 256); const int64_t cap = (int64_t)arpde_num_sms() * 32; if (blocks > cap) blocks = cap;
  filter2d_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, y, imgs, H, W, k);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
