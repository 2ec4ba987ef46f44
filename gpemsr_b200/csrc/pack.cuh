// Layout helpers for the K8-blocked bf16 operand format (see gemm_core.cuh).
#pragma once
#include "sm100.cuh"

namespace pack {

// split fp32 -> (hi, lo) bf16 with hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Row-major fp32 X[rows, k] (row stride ld) -> blocked [k_pad/8][rows_alloc][8] hi (and lo if non-null).
// Rows >= rows and k >= k are written as zero up to rows_fill / k_pad.
static __global__ void pack_rowmajor_kernel(const float* __restrict__ x, long long rows, int k, long long ld,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                     long long rows_alloc, long long rows_fill, int k_pad, long long row_dst0) {
  const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte cell per thread
  const long long cells = rows_fill * (k_pad / 8);
  if (cell >= cells) return;
  const long long r = cell % rows_fill;
  const int kc = (int)(cell / rows_fill);
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kk = kc * 8 + j * 2 + e;
      v[e] = (r < rows && kk < k) ? x[r * ld + kk] : 0.0f;
    }
    __nv_bfloat16 h0, l0, h1, l1;
    split_bf16(v[0], h0, l0);
    split_bf16(v[1], h1, l1);
    h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  const long long dst = ((long long)kc * rows_alloc + row_dst0 + r) * 8;
  *reinterpret_cast<uint4*>(hi + dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace pack
