// api.cu -- version, error string, device attribute cache.
#include <stdarg.h>
#include "common.cuh"
#include "../../include/arpde.h"
#include <atomic>

static thread_local char g_err[512] = "";

void arpde_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void arpde_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// kernels this PROCESS has launched through the library since the last reset (reset != 0 clears the counter after reading it); process-wide
// because autograd runs the backward calls on its own thread
extern "C" int64_t arpde_launch_count(int reset) {
  return reset ? g_launches.exchange(0, std::memory_order_relaxed) : g_launches.load(std::memory_order_relaxed);
}

int arpde_num_sms() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
    cached_dev = dev;
  }
  return cached;
}

extern "C" int arpde_version(void) { return ARPDE_VERSION; }
extern "C" const char* arpde_last_error(void) { return g_err; }
