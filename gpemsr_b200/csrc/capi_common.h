// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "../../include/gpemsr_b200.h"
#include "sm100.cuh"

namespace gpemsr {

int set_error(int code, const char* fmt, ...);
int check_device_current();                 // GPEMSR_OK iff the current device is sm_100-class
extern std::atomic<long long> g_launches;   // kernels launched by this library
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int num_sms();
bool use_clusters();                      // GPEMSR_CLUSTER=0 disables the cluster-multicast kernels (A/B testing)
bool use_pair_mma();                      // GPEMSR_PAIR=0: the wide streaming GEMMs stay on cta_group::1 + multicast (A/B testing)
bool use_tensor_maps();                   // GPEMSR_TMA=0: activation tiles by plain bulk copies instead of tensor-map TMA (A/B testing)

using TensorMap = sm100::TensorMap;      // a CUtensorMap as an opaque kernel parameter
// Tiled tensor map over K8-blocked activation cells, element type = 8 bytes (half a 16-byte cell: a 16-byte inner dimension
// makes the TMA engine work cell by cell -- measured 1.35x slower than bulk copies -- while a run of seg_len cells exceeds the
// 256-element box limit, so a run is described as 2 x (seg_len 8-byte elements)).  dims[0 .. rank-1] / strides_bytes[1 .. rank-1]
// / box[] innermost first.  No swizzle, out-of-bounds elements read as zero.  false when the driver entry point is missing or
// rejects the shape (callers then use the bulk-copy path; tensor_map_stats() counts both outcomes).
bool encode_u64_map(TensorMap* out, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                    const unsigned* box);
void tensor_map_stats(long long* built, long long* rejected);

#define GPEMSR_CUDA_OK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::gpemsr::set_error(GPEMSR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                                 cudaGetErrorString(_e), __FILE__, __LINE__);             \
  } while (0)

#define GPEMSR_LAUNCH_OK(name)                                                            \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess)                                                                \
      return ::gpemsr::set_error(GPEMSR_ERR_CUDA, "launch of %s failed: %s", name,        \
                                 cudaGetErrorString(_e));                                 \
    ::gpemsr::count_launch();                                                             \
  } while (0)

// launch with a thread-block cluster of `cluster_x` CTAs along x
template <class Kern, class... Args>
inline cudaError_t launch_cluster(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t s, unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

}  // namespace gpemsr
