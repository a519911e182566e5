// common.cuh -- shared helpers for libarpde (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define ARPDE_OK 0
#define ARPDE_ERR_ARG 1
#define ARPDE_ERR_CUDA 2
#define ARPDE_ERR_UNSUPPORTED 3

// thread-local last-error string (the only mutable global state of the library)
void arpde_set_error(const char* fmt, ...);

#define ARPDE_CHECK_ARG(cond, ...)                 \
  do {                                             \
    if (!(cond)) {                                 \
      arpde_set_error(__VA_ARGS__);                \
      return ARPDE_ERR_ARG;                        \
    }                                              \
  } while (0)

#define ARPDE_CUDA(call)                                                             \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      arpde_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return ARPDE_ERR_CUDA;                                                         \
    }                                                                                \
  } while (0)

// every kernel launch of the library is followed by this macro: error check + the calling thread's launch counter
// (arpde_launch_count, a diagnostic like the error string; bench.py reports it as gpu_launches)
void arpde_count_launch();
#define ARPDE_LAUNCH_CHECK()          \
  do {                                \
    ARPDE_CUDA(cudaGetLastError());   \
    arpde_count_launch();             \
  } while (0)

// Development switches (kernel ablations, tile-size overrides).  They read the environment ONLY in a library built with
// -DARPDE_DEV (`ARPDE_DEV=1 bash build.sh`); the shipped library reads no environment variable and keeps no state but the
// thread-local error string, as include/arpde.h promises.
#ifdef ARPDE_DEV
#include <stdlib.h>
static inline const char* arpde_dev_env(const char* key) { return getenv(key); }
#else
static inline const char* arpde_dev_env(const char*) { return nullptr; }
#endif

// Programmatic dependent launch (rollouts): with `pdl` the kernel may be scheduled while its predecessor in the stream is still draining --
// its CTAs take the SMs the predecessor's CTAs leave, run their prologue (barrier init, TMEM allocation) and block in pdl_wait() until the
// predecessor has completed and its writes are visible.  Only kernels that call pdl_wait() before their first access to global memory are
// launched this way; a kernel launched without the attribute is serialised as usual, whatever its predecessor did.
template <typename T>
static inline cudaError_t arpde_launch1(void (*kernel)(T), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, bool pdl, const T& arg) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr;
  memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization; attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, arg);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }   // the next kernel may be scheduled
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }                  // predecessor complete, its writes visible
// Warp index as a value the compiler KNOWS to be warp-uniform.  With `threadIdx.x >> 5` every branch or loop bound derived from the warp index
// counts as potentially divergent, and each __shfl_sync inside such a region is compiled to a WARPSYNC / ENDCOLLECTIVE slow-path sequence
// instead of one SHFL (found in the SASS of the stencil, output-conv and KS-rollout kernels: one such pair per shuffle).
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
#endif

int arpde_num_sms();   // cached cudaDevAttrMultiProcessorCount of the current device

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide double sum; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double s_part[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum_d(v);
  __syncthreads();
  if (lane == 0) s_part[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? s_part[threadIdx.x] : 0.0;
  if (wid == 0) v = warp_sum_d(v);
  return v;
}

__device__ __forceinline__ int wrap(int i, int n) {  // i in [-n, 2n)
  return i < 0 ? i + n : (i >= n ? i - n : i);
}
