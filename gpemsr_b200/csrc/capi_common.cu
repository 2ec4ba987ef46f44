// Error reporting, device gate and launch counter shared by every entry point.
#include "capi_common.h"
#include <cstring>
#include <cstdlib>
#include <cuda.h>            // CUtensorMap types only: the encoder is resolved at run time, libcuda is not linked

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

bool use_pair_mma() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPEMSR_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

bool use_tensor_maps() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GPEMSR_TMA"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

static std::atomic<long long> g_maps_built{0}, g_maps_rejected{0};
void tensor_map_stats(long long* built, long long* rejected) { *built = g_maps_built.load(); *rejected = g_maps_rejected.load(); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

bool encode_u64_map(TensorMap* out, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                     const unsigned* box) {
  static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) >= alignof(CUtensorMap), "TensorMap mirrors CUtensorMap");
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || rank < 2 || rank > 5) { g_maps_rejected.fetch_add(1); return false; }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 1; i < rank; ++i) gstr[i - 1] = strides_bytes[i];
  const CUresult r = fn(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT64, (cuuint32_t)rank, const_cast<void*>(base),
                        gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  (r == CUDA_SUCCESS ? g_maps_built : g_maps_rejected).fetch_add(1);
  return r == CUDA_SUCCESS;
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
void gpemsr_tensor_map_stats(int64_t* built, int64_t* rejected) {
  long long b = 0, r = 0;
  gpemsr::tensor_map_stats(&b, &r);
  if (built) *built = b;
  if (rejected) *rejected = r;
}
}
