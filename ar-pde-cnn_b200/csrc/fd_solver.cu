// fd_solver.cu -- ground-truth finite-difference solver for the 1D viscous Burgers equation on the GPU (fp64).
//
// Replaces burgers1d() of 1D-Burger-SWAG/solver/fd_burger1D.py:32-98 (the script that produces reference / validation data: classical
// RK4 in time :45-52, first- or second-order UPWIND gradient chosen by the sign of u :36-42,58-68, 2nd- or 4th-order central second
// derivative :70-79, periodic domain, double precision, one trajectory at a time through nn.Conv1d).  Here a batch of trajectories is
// integrated in ONE launch: one CTA per trajectory keeps its state in registers for all time steps, the four RK4 stages exchange
// neighbour values through a double-buffered shared-memory line (one barrier per stage), and only the saved steps go to HBM, rounded to
// fp32 exactly as the reference stores them (torch.zeros(nsteps, L) is a float32 tensor, :82).
#include "common.cuh"
#include "../../include/arpde.h"

namespace {

constexpr int FD_PPT = 4;          // points per thread (L <= 1024 * FD_PPT)

struct FdCoef {
  double bw[5], fw[5], lp[5];      // backward / forward one-sided gradient and second-derivative taps, centred (index 2 = the point)
  double nu, dt;
  int r;                           // stencil radius (1 or 2)
};

__device__ __forceinline__ double fd_rhs(const double* __restrict__ s, int i, int L, const FdCoef& k) {
  const double u = s[i];
  double b = 0.0, f = 0.0, l = 0.0;
#pragma unroll
  for (int t = -2; t <= 2; ++t) {
    if (t < -k.r || t > k.r) continue;
    int j = i + t; j = j < 0 ? j + L : (j >= L ? j - L : j);
    const double v = s[j];
    b += k.bw[t + 2] * v; f += k.fw[t + 2] * v; l += k.lp[t + 2] * v;
  }
  const double ux = u > 0.0 ? b : f;                     // upwind: backward difference where the wave moves right
  return -u * ux + k.nu * l;
}

__global__ void __launch_bounds__(1024) fd_burgers1d_kernel(const double* __restrict__ u0, float* __restrict__ out, int L, int nsteps,
                                                            int save_every, FdCoef k) {
  extern __shared__ double line[];                       // [2][L]
  const int n = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int nsave = nsteps / save_every;
  double u[FD_PPT], acc[FD_PPT];
#pragma unroll
  for (int p = 0; p < FD_PPT; ++p) { const int i = tid + p * nt; u[p] = i < L ? u0[(long long)n * L + i] : 0.0; }
  int buf = 0;
  for (int step = 0; step < nsteps; ++step) {
    // RK4 (fd_burger1D.py:45-52): k1 = f(u), k2 = f(u + k1 dt/2), k3 = f(u + k2 dt/2), k4 = f(u + k3 dt); u += (k1 + 2k2 + 2k3 + k4) dt/6
    double kprev[FD_PPT];
#pragma unroll
    for (int stage = 0; stage < 4; ++stage) {
      double* s = line + buf * L;
#pragma unroll
      for (int p = 0; p < FD_PPT; ++p) {
        const int i = tid + p * nt;
        if (i < L) s[i] = stage == 0 ? u[p] : (stage == 3 ? u[p] + kprev[p] * k.dt : u[p] + kprev[p] * k.dt / 2);
      }
      __syncthreads();
#pragma unroll
      for (int p = 0; p < FD_PPT; ++p) {
        const int i = tid + p * nt;
        if (i < L) {
          const double kk = fd_rhs(s, i, L, k);
          kprev[p] = kk;
          acc[p] = stage == 0 ? kk : acc[p] + ((stage == 1 || stage == 2) ? 2 * kk : kk);
        }
      }
      buf ^= 1;                                          // the next stage writes the other line: one barrier per stage suffices
    }
#pragma unroll
    for (int p = 0; p < FD_PPT; ++p) u[p] = u[p] + acc[p] * k.dt / 6;
    if ((step + 1) % save_every == 0) {
      const int row = (step + 1) / save_every - 1;
      if (row < nsave) {
#pragma unroll
        for (int p = 0; p < FD_PPT; ++p) {
          const int i = tid + p * nt;
          if (i < L) out[((long long)n * nsave + row) * L + i] = (float)u[p];
        }
      }
    }
  }
}

}  // namespace

extern "C" int arpde_fd_burgers1d(const double* u0, float* out, int N, int L, double dx, double nu, double dt, int nsteps, int save_every,
                                  int grad_order, int div_order, arpde_stream_t stream_) {
  ARPDE_CHECK_ARG(N >= 0 && L >= 5 && L <= 1024 * FD_PPT, "fd_burgers1d: bad dims N=%d L=%d (5 <= L <= %d)", N, L, 1024 * FD_PPT);
  ARPDE_CHECK_ARG(dx > 0 && dt > 0 && nsteps >= 0 && save_every >= 1, "fd_burgers1d: bad dx/dt/nsteps/save_every");
  ARPDE_CHECK_ARG((grad_order == 1 || grad_order == 2) && (div_order == 2 || div_order == 4), "fd_burgers1d: grad_order in {1,2}, div_order in {2,4}");
  if (N == 0 || nsteps == 0) return ARPDE_OK;
  ARPDE_CHECK_ARG(u0 && out, "fd_burgers1d: null pointer");
  FdCoef k; memset(&k, 0, sizeof(k));
  // tap tables as the reference builds them: integer taps divided by the spacing in double (:60-79)
  if (grad_order == 1) { k.bw[1] = -1.0 / dx; k.bw[2] = 1.0 / dx; k.fw[2] = -1.0 / dx; k.fw[3] = 1.0 / dx; }
  else { k.bw[0] = 1.0 / (dx * 2); k.bw[1] = -4.0 / (dx * 2); k.bw[2] = 3.0 / (dx * 2); k.fw[2] = -3.0 / (dx * 2); k.fw[3] = 4.0 / (dx * 2); k.fw[4] = -1.0 / (dx * 2); }
  if (div_order == 2) { k.lp[1] = 1.0 / (dx * dx); k.lp[2] = -2.0 / (dx * dx); k.lp[3] = 1.0 / (dx * dx); }
  else { k.lp[0] = (-1.0 / 12) / (dx * dx); k.lp[1] = (4.0 / 3) / (dx * dx); k.lp[2] = (-5.0 / 2) / (dx * dx); k.lp[3] = (4.0 / 3) / (dx * dx); k.lp[4] = (-1.0 / 12) / (dx * dx); }
  k.r = (grad_order == 2 || div_order == 4) ? 2 : 1;
  k.nu = nu; k.dt = dt;
  int threads = (L + FD_PPT - 1) / FD_PPT; threads = ((threads + 31) / 32) * 32; if (threads < 64) threads = 64;
  // fewer points per thread when the line is short: one point per thread up to 1024
  if (L <= 1024) threads = ((L + 31) / 32) * 32;
  const size_t smem = (size_t)2 * L * sizeof(double);
  if (smem > 48 * 1024) ARPDE_CUDA(cudaFuncSetAttribute(fd_burgers1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fd_burgers1d_kernel<<<N, threads, smem, (cudaStream_t)stream_>>>(u0, out, L, nsteps, save_every, k);
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
