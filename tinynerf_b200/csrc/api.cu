// Library-level entry points of the C ABI (include/tinynerf_b200.h): version, errors, device info.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

namespace tnf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  // per-device cache; benign race (same value written)
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

}  // namespace tnf

extern "C" int tnf_version(void) { return 1000; }

extern "C" const char* tnf_last_error(void) { return tnf::g_err; }

extern "C" int tnf_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  using namespace tnf;
  TNF_REQUIRE(sm_count && cc_major && cc_minor, "null output pointer");
  int dev = 0;
  TNF_CUDA(cudaGetDevice(&dev));
  TNF_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  TNF_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  TNF_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return TNF_OK;
}
