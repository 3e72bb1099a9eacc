// libdv_b200: version / error plumbing of the C ABI (include/dv_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dv

extern "C" int dv_version(void) { return 1; }
extern "C" const char* dv_last_error_string(void) { return dv::g_err; }
extern "C" int dv_device_sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
    dv::set_error("dv_device_sm_count: cudaDeviceGetAttribute failed");
    (void)cudaGetLastError();
    return DV_ECUDA;
  }
  return n;
}
