// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "../../include/gpemsr_b200.h"

namespace gpemsr {

int set_error(int code, const char* fmt, ...);
int check_device_current();                 // GPEMSR_OK iff the current device is sm_100-class
extern std::atomic<long long> g_launches;   // kernels launched by this library
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int num_sms();
bool use_clusters();                      // GPEMSR_CLUSTER=0 disables the cluster-multicast kernels (A/B testing)

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
