// api.cu -- version, error string, device attribute cache.
#include <stdarg.h>
#include "common.cuh"
#include "../../include/arpde.h"

static thread_local char g_err[512] = "";

void arpde_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local long long g_launches = 0;
void arpde_count_launch() { ++g_launches; }
// kernels this thread has launched through the library since the last reset (reset != 0 clears the counter after reading it)
extern "C" int64_t arpde_launch_count(int reset) {
  const long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
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
