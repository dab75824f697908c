// Error plumbing + device info for the C ABI (include/aum_b200.h).
#include <stdarg.h>

#include "common.cuh"

namespace aum {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

}  // namespace aum

extern "C" {

int aum_version(void) { return AUM_B200_VERSION; }

const char* aum_last_error(void) { return aum::g_err; }

int aum_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { aum::set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return 2; }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) { aum::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return 2; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

}  // extern "C"
