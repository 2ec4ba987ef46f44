// Self-test entry points: exercise the tcgen05 GEMM core on caller-supplied matrices so the descriptor /
// TMEM / pipeline plumbing can be validated in isolation (tests/test_gemm_core_gpu.py).
#include "capi_common.h"
#include "gemm_core.cuh"
#include "pack.cuh"

namespace {

struct EpiRowMajor {
  float* d;        // [m_alloc, ld]
  int ld;
  int n;           // valid columns
  long long m;     // valid rows
  template <int BLOCK_N>
  __device__ __forceinline__ void run(uint32_t tmem_acc, long long m_tile, int n_tile, int row) const {
    const long long gr = m_tile * gemm::BLOCK_M + row;
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t r[32];
      sm100::tmem_ld_32x32(tmem_acc + c0, r);
      sm100::tmem_ld_wait();
      if (gr < m) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n_tile * BLOCK_N + c0 + j;
          if (col < n) d[gr * ld + col] = __uint_as_float(r[j]);
        }
      }
    }
  }
};
template <int BLOCK_N>
struct EpiRowMajorN : EpiRowMajor {
  static constexpr int WARPS = 4;
  struct State {};
  __device__ __forceinline__ void tile(State&, uint32_t tmem_acc, long long m_tile, int n_tile, int, int row, int) const {
    this->template run<BLOCK_N>(tmem_acc, m_tile, n_tile, row);
  }
};

template <int BLOCK_N, int BLOCK_K, int SPLIT, int NSTAGE>
int launch(const gemm::Operands& op, const EpiRowMajor& e, cudaStream_t s) {
  using Cfg = gemm::Config<BLOCK_N, BLOCK_K, SPLIT, NSTAGE>;
  EpiRowMajorN<BLOCK_N> epi;
  static_cast<EpiRowMajor&>(epi) = e;
  auto kern = gemm::gemm_kernel<BLOCK_N, BLOCK_K, SPLIT, NSTAGE, EpiRowMajorN<BLOCK_N>>;
  GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  const int grid = (int)std::min<long long>(op.m_tiles, gpemsr::num_sms());
  kern<<<grid, gemm::NUM_THREADS, Cfg::SMEM_BYTES, s>>>(op, epi);
  GPEMSR_LAUNCH_OK("gemm_kernel(selftest)");
  return GPEMSR_OK;
}

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

}  // namespace

extern "C" {

// workspace: blocked hi/lo planes of A and B + error flag
size_t gpemsr_selftest_gemm_workspace_bytes(int64_t m, int n, int k) {
  const long long ma = round_up(m, 128), na = round_up(n, 256), kp = round_up(k, 64);
  return (size_t)(2 * (ma + na) * kp * 2 + 256);
}

// D[m, n] = A[m, k] * B[n, k]^T, fp32 row-major in and out.  split = 1 (one bf16 pass) or 3 (hi/lo, fp32-faithful).
// block_n in {64, 128, 256}.
int gpemsr_selftest_gemm(const float* a, const float* b, int64_t m, int n, int k, int split, int block_n,
                         float* d, void* ws, size_t ws_bytes, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (m <= 0 || n <= 0 || k <= 0) return set_error(GPEMSR_ERR_BAD_SHAPE, "selftest_gemm: empty problem");
  if (ws_bytes < gpemsr_selftest_gemm_workspace_bytes(m, n, k))
    return set_error(GPEMSR_ERR_WORKSPACE, "selftest_gemm: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const long long ma = round_up(m, 128), na = round_up(n, 256), kp = round_up(k, 64);
  __nv_bfloat16* a_hi = (__nv_bfloat16*)ws;
  __nv_bfloat16* a_lo = a_hi + ma * kp;
  __nv_bfloat16* b_hi = a_lo + ma * kp;
  __nv_bfloat16* b_lo = b_hi + na * kp;
  int* err = (int*)(b_lo + na * kp);
  GPEMSR_CUDA_OK(cudaMemsetAsync(err, 0, sizeof(int), s));
  {
    long long cells = ma * (kp / 8);
    pack::pack_rowmajor_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(a, m, k, k, a_hi, a_lo, ma, ma, (int)kp, 0);
    GPEMSR_LAUNCH_OK("pack_rowmajor_kernel(A)");
    cells = na * (kp / 8);
    pack::pack_rowmajor_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(b, n, k, k, b_hi, b_lo, na, na, (int)kp, 0);
    GPEMSR_LAUNCH_OK("pack_rowmajor_kernel(B)");
  }
  gemm::Operands op{};
  op.a_hi = a_hi; op.a_lo = a_lo; op.b_hi = b_hi; op.b_lo = b_lo;
  op.a_rows = ma; op.b_rows = (int)na; op.k = (int)kp; op.taps = 1; op.a_row_off[0] = 0;
  op.m_tiles = ma / 128; op.a_row0 = 0; op.err_flag = err;
  EpiRowMajor e{d, n, n, m};
  op.n_tiles = (int)((n + block_n - 1) / block_n);
  if (split == 1 && block_n == 256) rc = launch<256, 64, 1, 4>(op, e, s);
  else if (split == 1 && block_n == 128) rc = launch<128, 64, 1, 4>(op, e, s);
  else if (split == 1 && block_n == 64) rc = launch<64, 64, 1, 4>(op, e, s);
  else if (split == 3 && block_n == 256) rc = launch<256, 32, 3, 4>(op, e, s);
  else if (split == 3 && block_n == 128) rc = launch<128, 32, 3, 4>(op, e, s);
  else if (split == 3 && block_n == 64) rc = launch<64, 32, 3, 4>(op, e, s);
  else return set_error(GPEMSR_ERR_UNSUPPORTED, "selftest_gemm: split=%d block_n=%d", split, block_n);
  return rc;
}

// reads the pipeline error flag left by the last selftest_gemm on this workspace (synchronises the stream)
int gpemsr_selftest_gemm_status(const void* ws, int64_t m, int n, int k, gpemsr_stream_t stream) {
  using namespace gpemsr;
  const long long ma = round_up(m, 128), na = round_up(n, 256), kp = round_up(k, 64);
  const int* err = (const int*)((const __nv_bfloat16*)ws + 2 * (ma + na) * kp);
  int h = 0;
  GPEMSR_CUDA_OK(cudaMemcpyAsync(&h, err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  GPEMSR_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  if (h != 0) return set_error(GPEMSR_ERR_CUDA, "gemm pipeline timed out at wait site %d", h);
  return GPEMSR_OK;
}

}  // extern "C"
