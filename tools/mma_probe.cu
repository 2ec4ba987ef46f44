// Micro-probe: what does one SS-mode tcgen05.mma (M = 128 per CTA, K = 16, bf16) cost as a function of N and of cta_group?
// No global loads, no epilogue: operands sit in shared memory (zeros), one thread issues REPS back-to-back MMAs into one TMEM
// accumulator, commits to an mbarrier and the CTA measures clock64() from first issue to completion.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gpemsr_b200/csrc -o /tmp/mma_probe tools/mma_probe.cu && /tmp/mma_probe
//
// Prints cycles per MMA for: N = 64 / 128 / 256 with cta_group::1; the paired pattern of the tap-fused kernel (N = 128 then
// N = 64, different A per MMA); and cta_group::2 (M = 256 over a CTA pair, each CTA supplying N / 2 rows of B).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "sm100.cuh"

using namespace sm100;

__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// mode 0: cta_group::1, every MMA has shape N; mode 1: the tap-fused pattern (N2 = 2N then N, A alternates hi / lo, 9 taps);
template <int CTA_GROUP, int mode>
__global__ void __launch_bounds__(128, 1) probe(int n, int reps, long long* out, int var) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (var & 16) {                                  // random bf16 operands in [-2, 2): does the cost depend on the data?
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      auto next = [&]() { h ^= h << 13; h ^= h >> 17; h ^= h << 5; return (h & 0x807F807Fu) | 0x3F803F80u; };
      v = make_uint4(next(), next(), next(), next());
    }
    reinterpret_cast<uint4*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1000000); fence_mbar_init(); }
  if (warp == 0) {
    if constexpr (CTA_GROUP == 1) tmem_alloc<512>(&tmem_base);
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  if constexpr (CTA_GROUP == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t crank = CTA_GROUP == 2 ? cluster_ctarank() : 0;
  long long t0 = 0, t1 = 0;
  // the issuing lane is elected in a converged warp (as the product kernels do): a divergent `threadIdx.x == 32` makes the
  // compiler wrap every UTCHMMA in an ELECT / BRA.U.ANY loop, ~47 cycles per MMA slot, and the probe measures that instead
  if (warp == 1) {
  if (elect_one()) {
  if (crank == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
    const uint32_t idesc_n = idesc_bf16_f32(128 * CTA_GROUP, n), idesc_2n = idesc_bf16_f32(128 * CTA_GROUP, 2 * n);
    // descriptors are built once; per MMA only the 14-bit start-address field moves (one 32-bit add), as in the product kernels
    const uint32_t brows = (mode == 1 ? 2 * n : n) / CTA_GROUP;
    // var bits: 1 = B fixed (same descriptor every MMA), 2 = A without the tap shift, 4 = A stage fixed, 8 = A in aligned 128-row segments
    const uint32_t b_var = (var & 1) ? 0 : 1, a_shift = (var & 2) ? 0 : 1, a_stage = (var & 4) ? 0 : 1;
    const uint64_t da0 = smem_desc_kmajor_noswz(a0, (var & 8) ? 128 * 16 : 130 * 16, 128);
    const uint64_t db0 = smem_desc_kmajor_noswz(b0, brows * 16, 128);
    const uint32_t b_tap16 = (2 * brows * 16 <= 7168 ? 2 * brows * 16 : 7168) >> 4;   // 9 taps stay inside the 64 KB B region
    t0 = clock64();
    for (int g = 0; g < reps / 9; ++g) {
      const uint64_t da = da0 + (uint64_t)((((g & 7) * a_stage) * 12480) >> 4);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        // A: 128 rows x 16 k = 2 cells of 130-row segments (like the tap-fused stage); a different start row per MMA (tap shift)
        const uint64_t da_hi = da + (uint64_t)((((tap / 3) * 4160 + (tap % 3) * 16) >> 4) * a_shift);
        const uint64_t da_lo = da_hi + (6240 >> 4);
        const uint64_t db = db0 + (uint64_t)(tap * b_tap16 * b_var);
        const bool acc = (g | tap) != 0;
        if (mode == 0) {
          if constexpr (CTA_GROUP == 1) umma_bf16(tmem_base, da_hi, db, idesc_n, acc); else umma2_bf16(tmem_base, da_hi, db, idesc_n, acc);
        } else {
          if constexpr (CTA_GROUP == 1) { umma_bf16(tmem_base, da_hi, db, idesc_2n, acc); umma_bf16(tmem_base, da_lo, db, idesc_n, true); }
          else { umma2_bf16(tmem_base, da_hi, db, idesc_2n, acc); umma2_bf16(tmem_base, da_lo, db, idesc_2n, true); }
        }
      }
      // var & 32: a commit after every 9 trips (the product's per-stage release); var & 64: after every 18 trips
      if constexpr (CTA_GROUP == 1) { if ((var & 32) || ((var & 64) && (g & 1))) umma_commit(&bar2); }
    }
    if constexpr (CTA_GROUP == 1) umma_commit(&bar); else umma2_commit_mc(&bar, 3);
  }
  {
    bool ok = mbar_wait(&bar, 0, nullptr, 0);
    t1 = clock64();
    if (crank == 0) out[blockIdx.x] = ok ? t1 - t0 : -1;
  }
  }
  __syncwarp();
  }
  tc_fence_before();
  if constexpr (CTA_GROUP == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (CTA_GROUP == 1) tmem_dealloc<512>(tmem_base);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

template <int G, int MODE>
double run(int n, int reps, int var = 0) {
  long long* d;
  const int grid = 148 / G * G;
  cudaMalloc(&d, grid * sizeof(long long));
  cudaMemset(d, 0, grid * sizeof(long long));
  cudaFuncSetAttribute(probe<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 160 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe<G, MODE>, n, reps, d, var);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); return -1; }
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(d);
  double s = 0; int c = 0;
  for (int i = 0; i < grid; i += G) { if (h[i] > 0) { s += (double)h[i]; ++c; } }
  return c ? s / c / reps : -1;
}

int main() {
  const int reps = 9 * 455;
  printf("cycles per loop trip (all SMs busy, %d trips; mode 0: one MMA per trip, mode 1: the paired pattern = 2 MMAs per trip)\n", reps);
  for (int n : {16, 32, 64, 128, 256}) printf("cta_group::1  M=128 N=%3d          : %7.1f cycles / MMA\n", n, run<1, 0>(n, reps));
  for (int n : {32, 64}) printf("cta_group::1  paired N=%3d+%3d        : %7.1f cycles / (tap, k-slab) pair\n", 2 * n, n, run<1, 1>(n, reps));
  for (int n : {64, 128, 256}) printf("cta_group::2  M=256 N=%3d          : %7.1f cycles / MMA (covers 256 rows)\n", n, run<2, 0>(n, reps));
  for (int n : {32, 64}) printf("cta_group::2  paired N=%3d+%3d (x2 rows): %7.1f cycles / pair (covers 256 rows)\n", 2 * n, 2 * n, run<2, 1>(n, reps));
  for (int var : {0, 32, 64})
    printf("variant %2d (1 = B fixed, 2 = no tap shift, 4 = A stage fixed, 8 = aligned A segments, 16 = random data, 32 = commit / 9 trips, 64 = commit / 18 trips): N=64 %7.1f  N=128 %7.1f  N=256 %7.1f  paired 128+64 %7.1f  cta2 N=256 %7.1f\n",
           var, run<1, 0>(64, reps, var), run<1, 0>(128, reps, var), run<1, 0>(256, reps, var), run<1, 1>(64, reps, var), run<2, 0>(256, reps, var));
  return 0;
}
