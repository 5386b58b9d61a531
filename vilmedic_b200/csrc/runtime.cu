// Runtime glue of libvlmb200.so: thread-local error string, device properties, ABI version.
#include "common.cuh"
#include "vlm_b200.h"
#include <cstdarg>
#include <cstring>

namespace vlm {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(err));
    return -2;
  }
  return 0;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace vlm

extern "C" const char* vlm_last_error(void) { return vlm::g_err; }

extern "C" int vlm_abi_version(void) { return VLM_B200_ABI_VERSION; }

extern "C" int vlm_device_check(void) {
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) {
    vlm::set_error("vlm_device_check: no CUDA device: %s", cudaGetErrorString(err));
    return -1;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    vlm::set_error("vlm_device_check: device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return -1;
  }
  return 0;
}
