// api.cu — library-level entry points and error plumbing of librsuper_b200.so.
// Contract: 0 / negative return codes, no exceptions across the boundary, no exit,
// thread-local last-error string (forward runs on the Python main thread, backward on the autograd
// engine's worker thread).
#include "rsb_common.cuh"

#include <cstdarg>
#include <cstdio>

#include "../../include/rsuper_b200.h"

namespace rsb {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -3;
  }
  return 0;
}

}  // namespace rsb

extern "C" const char* rsb_version(void) { return "rsuper_b200 0.1.0 (sm_100a)"; }

extern "C" const char* rsb_last_error(void) { return rsb::g_last_error; }

extern "C" int rsb_num_sms(void) {
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached_dev = dev;
    cached_sms = sms;
  }
  return cached_sms;
}
