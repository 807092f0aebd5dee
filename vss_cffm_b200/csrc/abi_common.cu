// ABI bookkeeping: version, thread-local error text, device check.
#include <stdarg.h>

#include "common.cuh"

namespace cffm {
namespace {
thread_local char g_err[512] = "";
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("CFFM_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
}  // namespace cffm

extern "C" int cffm_abi_version(void) { return CFFM_ABI_VERSION; }

extern "C" const char* cffm_last_error(void) { return cffm::g_err; }

extern "C" int cffm_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cffm::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return -(int)e;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    cffm::set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return -(int)e;
  }
  CFFM_REQUIRE(major == 10, CFFM_E_ARCH, "device %d has compute capability %d.x; this library is built for sm_100a only",
               dev, major);
  return CFFM_OK;
}

extern "C" int cffm_current_device(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cffm::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return -(int)e;
  }
  return dev;
}
