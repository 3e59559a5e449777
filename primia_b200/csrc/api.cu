// Error plumbing + misc entry points of the C ABI.
#include "common.cuh"
#include <string.h>

char g_pm_err[512] = "";

extern "C" const char* pm_last_error(void) { return g_pm_err; }
extern "C" int pm_version(void) { return 100; }
extern "C" int pm_device_sm_count(int* out) {
  if (!out) return PM_EINVAL;
  int dev = 0, n = 0;
  PM_CUDA(cudaGetDevice(&dev));
  PM_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *out = n;
  return PM_OK;
}
