// Library-level entry points of libb2t.so: version, error string, device check.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void b2t_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;
void b2t_count_launches(int n) { g_launches += (unsigned long long)n; }

static int g_coop_limit = 0, g_trace_limit = 0;
int b2t_coop_limit() { return g_coop_limit; }
int b2t_trace_limit() { return g_trace_limit; }
// Cap the resident blocks per SM of the cooperative sweeps / of the path-loop kernel (0 = no cap), so that
// a private-arena pipeline on a second stream can run next to the path loop of the main arena.
B2T_EXPORT int b2t_set_launch_limits(int coop_blocks_per_sm, int trace_blocks_per_sm) {
  g_coop_limit = coop_blocks_per_sm < 0 ? 0 : coop_blocks_per_sm;
  g_trace_limit = trace_blocks_per_sm < 0 ? 0 : trace_blocks_per_sm;
  return B2T_OK;
}

B2T_EXPORT int b2t_version(void) { return 100; }

// number of kernels this library has launched since load (or since the last reset)
B2T_EXPORT unsigned long long b2t_launch_count(int reset) {
  const unsigned long long v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

B2T_EXPORT const char* b2t_last_error(void) { return g_err; }

B2T_EXPORT int b2t_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    b2t_set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B2T_ERR_DEVICE;
  }
  int major = 0, minor = 0;   // attribute queries are cheap; cudaGetDeviceProperties is not
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) {
    b2t_set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return B2T_ERR_DEVICE;
  }
  if (major != 10) {
    b2t_set_error("libb2t.so is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return B2T_ERR_DEVICE;
  }
  return B2T_OK;
}
