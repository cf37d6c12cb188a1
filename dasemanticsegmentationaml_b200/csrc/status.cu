#include "status.h"
#include "b200seg.h"

#include <stdarg.h>
#include <stdio.h>

namespace b200 {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(B200_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return B200_OK;
}

}  // namespace b200

extern "C" {

const char* b200_last_error() { return b200::g_err; }

int b200_version() { return 100; }

// 0 when the current device can run the library (compute capability 10.x), else B200_EARCH.
int b200_check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b200::set_error(B200_ECUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return b200::set_error(B200_EARCH, "device %d is sm_%d%d; libb200seg is built for sm_100a only", dev, major, minor);
  return B200_OK;
}

}  // extern "C"
