// Shared device / host helpers for the padded K8-blocked activation format (see include/gpemsr_b200.h): geometry, row
// placement, the (hi, lo) bf16 split and argument checks.  Included by every translation unit that touches activations.
#pragma once
#include "capi_common.h"
#include "sm100.cuh"

namespace {

constexpr int kRowTile = 128;      // gemm::BLOCK_M: r_img is a multiple of it

struct Geom {
  int n, h, w, padded;
  long long r_img, m0, rows_alloc;
  __host__ __device__ int wp() const { return w + 2 * padded; }       // `padded` is the zero-ring width (0 = compact)
};
Geom to_geom(const gpemsr_geom_t& g) { return Geom{g.n, g.h, g.w, g.padded, g.r_img, g.m0, g.rows_alloc}; }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == GPEMSR_ACT_RELU) return fmaxf(v, 0.f);
  if (act == GPEMSR_ACT_LRELU) return v > 0.f ? v : v * slope;
  return v;
}

__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  // packed conversions (F2FP.BF16.PACK_AB: two values per instruction on the FMA pipe) instead of one F2F each on the
  // conversion pipe: the epilogues convert 2 x 64 values per output row
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);              // .x (low half) = v[2j]
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    h[j] = *reinterpret_cast<const uint32_t*>(&h2);
    l[j] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// decode a row of geometry g (relative to g.m0) into (image, y, x); false for ring / tail rows
__device__ __forceinline__ bool decode_row(const Geom& g, long long rel, int& img, int& y, int& x) {
  // every epilogue thread decodes its row once per tile: 32-bit divisions whenever the numbers fit (always, for real shapes) --
  // a 64-bit division is ~100 dependent instructions and the narrow-tile epilogues are latency bound
  long long q;
  if ((((unsigned long long)rel | (unsigned long long)g.r_img) >> 31) == 0) {
    const unsigned r32 = (unsigned)rel, ri = (unsigned)g.r_img;
    img = (int)(r32 / ri);
    q = (long long)(r32 - (unsigned)img * ri);
  } else {
    img = (int)(rel / g.r_img);
    q = rel - (long long)img * g.r_img;
  }
  if (g.padded) {
    const unsigned wp = (unsigned)(g.w + 2 * g.padded), q32 = (unsigned)q;        // q < r_img < 2^31 here
    const int yp = (int)(q32 / wp), xp = (int)(q32 - (unsigned)yp * wp);
    y = yp - g.padded; x = xp - g.padded;
    return img < g.n && y >= 0 && y < g.h && x >= 0 && x < g.w;
  }
  if (q >> 31) { y = (int)(q / g.w); x = (int)(q - (long long)y * g.w); }
  else { const unsigned q32 = (unsigned)q, w32 = (unsigned)g.w; y = (int)(q32 / w32); x = (int)(q32 - (unsigned)y * w32); }
  return img < g.n && q < (long long)g.h * g.w;
}
__device__ __forceinline__ long long place_row(const Geom& g, int img, int y, int x) {
  return g.m0 + (long long)img * g.r_img + (long long)(y + g.padded) * (g.w + 2 * g.padded) + (x + g.padded);
}

int check_geom(const gpemsr_geom_t& g, const char* what) {
  using namespace gpemsr;
  if (g.n <= 0 || g.h <= 0 || g.w <= 0 || g.r_img <= 0 || (g.r_img % kRowTile) != 0)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "%s: bad geometry n=%d h=%d w=%d r_img=%lld", what, g.n, g.h, g.w, (long long)g.r_img);
  if (g.padded < 0 || g.padded > 3) return set_error(GPEMSR_ERR_BAD_SHAPE, "%s: ring width %d (0..3 supported)", what, g.padded);
  if (g.r_img >= (1LL << 31)) return set_error(GPEMSR_ERR_BAD_SHAPE, "%s: more than 2^31 rows per image", what);
  const long long need = (long long)(g.h + 2 * g.padded) * (g.w + 2 * g.padded);
  const long long margin = g.padded ? (long long)g.padded * (g.w + 2 * g.padded) + g.padded : 0;     // largest tap shift
  if (g.r_img < need || g.m0 < margin || g.rows_alloc < g.m0 + (long long)g.n * g.r_img + margin)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "%s: geometry does not leave room for the zero ring / tap shifts", what);
  return GPEMSR_OK;
}

}  // namespace
