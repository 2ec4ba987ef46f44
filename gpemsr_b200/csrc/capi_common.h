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

}  // namespace gpemsr
