// C-ABI housekeeping: thread-local error string, version, device info.
#include <stdarg.h>
#include <string.h>

#include "pmu_common.cuh"

namespace pmu {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace pmu

extern "C" const char* pmu_last_error(void) { return pmu::g_err; }

extern "C" int pmu_version(void) { return 100; }  // 0.1.0

extern "C" int pmu_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  PMU_CUDA(cudaGetDevice(&dev));
  if (sm_count) PMU_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) PMU_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) PMU_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return PMU_OK;
}

extern "C" int pmu_set_device(int device) {
  PMU_CHECK_ARG(device >= 0, "pmu_set_device: negative device index");
  PMU_CUDA(cudaSetDevice(device));
  return PMU_OK;
}
