// Runtime glue of libvlmb200.so: thread-local error string, device properties, ABI version.
#include "common.cuh"
#include "vlm_b200.h"
#include <cstdarg>
#include <cstring>
#include <cstdlib>

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

// SMs left to other work (the NCCL kernels of the data-parallel exchange): every persistent grid is sized num_sms() = SM count -
// margin.  A persistent kernel with a static tile schedule whose CTAs cannot all be resident runs a second wave — leaving a few
// SMs free costs margin/148 of the compute instead.  Default: env VLM_SM_MARGIN (0); vlm_set_sm_margin() overrides.
static int g_sm_margin = -1;

int sm_margin() {
  if (g_sm_margin < 0) {
    const char* v = getenv("VLM_SM_MARGIN");
    int m = v ? atoi(v) : 0;
    g_sm_margin = (m < 0 || m > 64) ? 0 : m;
  }
  return g_sm_margin;
}

static int num_sms_physical() {
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

int num_sms() { return num_sms_physical() - sm_margin(); }

// Background mode: while set, the HBM-bound helper kernels that the host issues on a SIDE stream (optimizer update of a gradient
// bucket, bias-gradient column sums) are launched as many small CTAs (128 threads, <= 48 registers, next to no shared memory) so
// that one of them fits on an SM NEXT TO the resident persistent tensor CTA (448 threads x <= 128 registers leave 8 K registers)
// and their memory traffic runs under the tensor-bound kernels of the main stream instead of after them.
static int g_background = 0;
int background_mode() { return g_background; }
int num_sms_all() { return num_sms_physical(); }

}  // namespace vlm

extern "C" const char* vlm_last_error(void) { return vlm::g_err; }

extern "C" int vlm_abi_version(void) { return VLM_B200_ABI_VERSION; }

extern "C" int vlm_set_sm_margin(int margin) {
  if (margin < 0 || margin > 64) {
    vlm::set_error("vlm_set_sm_margin: margin must be in [0, 64], got %d", margin);
    return -1;
  }
  vlm::g_sm_margin = margin;
  return 0;
}

extern "C" int vlm_get_sm_margin(void) { return vlm::sm_margin(); }

extern "C" int vlm_set_background(int on) {
  const int prev = vlm::g_background;
  vlm::g_background = on ? 1 : 0;
  return prev;
}

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
