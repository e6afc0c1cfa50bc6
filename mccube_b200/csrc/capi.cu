// Library-level plumbing of the C ABI: version, thread-local error text, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mccb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// Launch-configuration errors only (cudaPeekAtLastError does not synchronise, so the
// call stays asynchronous and graph-capturable).
int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(MCCB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return MCCB200_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace mccb

extern "C" int mccb200_version(void) { return MCCB200_VERSION; }
extern "C" const char* mccb200_last_error(void) { return mccb::g_err; }
extern "C" int mccb200_sm_count(void) { return mccb::sm_count(); }
