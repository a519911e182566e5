// peer.cu -- all-reduce of a few doubles over NVLink peer memory (one process per GPU, one node).
//
// Data-parallel training (arpde_densed_forward_dp / _backward_dp) sums the fp64 BatchNorm statistics of every layer over the ranks:
// 16 all-reduces of <= 192 doubles per AR step, each sitting on the critical path between the kernel that accumulates the sums and the
// kernel that consumes them.  Through NCCL (from a Python callback) one of them costs ~35 us of latency; here it is ONE small kernel:
// every rank copies its values into its own exchange slot (symmetric memory: all ranks map all slots), publishes a sequence number in
// every peer's flag array (st.release.sys), waits for the same number from every peer (ld.acquire.sys), then reads all ranks' slots over
// NVLink and adds them in rank order -- so every rank obtains bit-identical sums.  No host round trip, no communicator in the library:
// the caller owns the context (peer pointers from torch.distributed._symmetric_memory, the running sequence number).
// The kernel waits for kernels of OTHER GPUs: it must only be used with one GPU per rank (two such ranks on one GPU could deadlock);
// the wait gives up after ~2 s and raises the context's error flag instead of hanging the device.
#include "common.cuh"
#include "../../include/arpde.h"

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct PeerArgs {
  double* slots[ARPDE_PEER_MAX];
  uint32_t* flags[ARPDE_PEER_MAX];
  int world, rank, slot_doubles;
  int* error;
};

__global__ void __launch_bounds__(128) peer_allreduce_kernel(PeerArgs c, double* __restrict__ buf, int n, uint32_t seq1, int slot) {
  const int tid = threadIdx.x;
  double* mine = c.slots[c.rank] + (size_t)slot * c.slot_doubles;
  for (int i = tid; i < n; i += blockDim.x) mine[i] = buf[i];
  __threadfence_system();
  __syncthreads();
  if (tid < c.world) st_release_sys(c.flags[tid] + slot * c.world + c.rank, seq1);      // my values for this round are visible
  if (tid < c.world) {
    const uint32_t* f = c.flags[c.rank] + slot * c.world + tid;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - seq1) < 0) {
      __nanosleep(64);
      if (clock64() - t0 > 4000000000ll) { if (c.error) *c.error = 1; break; }           // ~2 s: a peer is gone; do not hang the device
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < c.world; ++r) s += __ldcv(c.slots[r] + (size_t)slot * c.slot_doubles + i);      // rank order: identical on every rank
    buf[i] = s;
  }
}

}  // namespace

extern "C" int arpde_peer_allreduce_f64(void* ctx_, double* buf, int64_t n) {
  arpde_peer_ctx_t* ctx = (arpde_peer_ctx_t*)ctx_;
  ARPDE_CHECK_ARG(ctx && buf && n >= 0, "peer_allreduce: null argument");
  ARPDE_CHECK_ARG(ctx->world >= 1 && ctx->world <= ARPDE_PEER_MAX && ctx->rank >= 0 && ctx->rank < ctx->world, "peer_allreduce: bad rank %d / world %d", ctx->rank, ctx->world);
  ARPDE_CHECK_ARG(ctx->nslots >= 2 && n <= ctx->slot_doubles, "peer_allreduce: %lld doubles do not fit a slot of %d", (long long)n, ctx->slot_doubles);
  if (n == 0 || ctx->world == 1) return ARPDE_OK;
  PeerArgs a;
  for (int r = 0; r < ctx->world; ++r) { a.slots[r] = ctx->slots[r]; a.flags[r] = ctx->flags[r]; ARPDE_CHECK_ARG(a.slots[r] && a.flags[r], "peer_allreduce: rank %d not mapped", r); }
  a.world = ctx->world; a.rank = ctx->rank; a.slot_doubles = ctx->slot_doubles; a.error = ctx->error_flag;
  const uint64_t seq = ctx->seq++;
  peer_allreduce_kernel<<<1, 128, 0, (cudaStream_t)ctx->stream>>>(a, buf, (int)n, (uint32_t)(seq + 1), (int)(seq % (uint64_t)ctx->nslots));
  ARPDE_LAUNCH_CHECK();
  return ARPDE_OK;
}
