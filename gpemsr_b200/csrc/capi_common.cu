// Error reporting, device gate and launch counter shared by every entry point.
#include "capi_common.h"
#include <cstring>
#include <cstdlib>

namespace gpemsr {

static thread_local char tl_err[512] = "";
std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_err, sizeof(tl_err), fmt, ap);
  va_end(ap);
  return code;
}

static int device_is_sm100(int dev) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return set_error(GPEMSR_ERR_CUDA, "cannot query compute capability of device %d (no CUDA device?)", dev);
  }
  if (major != 10)
    return set_error(GPEMSR_ERR_UNSUPPORTED_ARCH,
                     "device %d is compute capability %d.x; this library is built for sm_100a only "
                     "and has no fallback path", dev, major);
  return GPEMSR_OK;
}

int check_device_current() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return set_error(GPEMSR_ERR_CUDA, "cudaGetDevice failed (no CUDA device?)");
  }
  static thread_local int cached_dev = -1;
  if (cached_dev == dev) return GPEMSR_OK;
  int rc = device_is_sm100(dev);
  if (rc == GPEMSR_OK) cached_dev = dev;
  return rc;
}

bool use_clusters() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPEMSR_CLUSTER"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

int num_sms() {
  static thread_local int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
  cached = n;
  return n;
}

}  // namespace gpemsr

extern "C" {
int gpemsr_version(void) { return 0x000100; }
const char* gpemsr_last_error_string(void) { return gpemsr::tl_err; }
int gpemsr_device_check(int device) { return gpemsr::device_is_sm100(device); }
int64_t gpemsr_kernel_launches(void) { return (int64_t)gpemsr::g_launches.load(); }
}
