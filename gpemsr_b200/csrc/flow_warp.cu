// a-5: flow_warp -- bilinear gather driven by a dense pixel-space flow field.
//
// Replaces basicsr.archs.arch_util.flow_warp (third-party; called from SpyNet.process, which the
// reference reaches at model/GPEMSR.py:99-100):
//     grid  = meshgrid + flow ; g = 2*grid/max(size-1,1) - 1 ; out = grid_sample(x, g, bilinear, pad, ac)
//
// HBM-bound: algorithmic bytes = 2*C*H*W*4 (x once in, out once) + H*W*8 (flow).
//
// Mapping.  A CTA owns a TILE_W x TILE_H pixel tile of one image and a slice of its channels.
//   phase 1  every thread turns ONE flow vector (coalesced float2 load) into four clamped tap
//            offsets and four bilinear weights and parks them in shared memory: the flow tile is
//            consumed C times but the coordinate arithmetic (two IEEE divisions) runs once;
//   phase 2  warps sweep channel planes: lane <-> consecutive x, so the four tap loads of a warp
//            fall into one or two 128-byte lines per input row (the flow is smooth) and every
//            store is a full 128-byte line.  Neighbouring rows of a tile reuse each other's input
//            lines out of L1 (2-D tile => ~ (TILE_H+1)/TILE_H re-fetch instead of 2x for a row strip).
//            CH_UNROLL planes are in flight per thread.  The east taps of lane i are normally the west taps of lane i+1, so
//            they are taken by warp shuffle and only loaded when the addresses differ.
//
// Measured (profiles/README.md): 64 x 1250^2 in 0.225 ms = 3.6 TB/s = 55 % of the measured 6.55 TB/s copy peak; DRAM traffic
// equals the algorithmic bytes (413 MB read, ~400 MB written).  Variants tried on the B200 and rejected because they were no
// faster: 64x4 / 128x2 / 32x32 tiles, CH_UNROLL 2 / 8, 32-bit precomputed offsets (-37 % instructions), two pixels per lane,
// four rows per thread with vertical tap reuse, channel-split grids.  All land at 0.22-0.25 ms: the limiter is the L1 / memory
// path of sector-granular gathers over 64 interleaved planes, not issue rate or occupancy.  Also rejected: staging each plane's
// tap box (11 rows x 192 B per 32 x 8 tile) in shared memory with one cp.async.bulk per row, 8 planes in flight per CTA:
// correct, but 0.395 ms -- the bulk-copy engine sustains only ~1 such small copy per ~26 cycles per SM.
//
// Bit-faithful coordinates: each elementwise op of the reference is one separately
// rounded fp32 op here (__fadd_rn/__fmul_rn/__fdiv_rn, no contraction), in the reference order:
//     v = g + f ; t = 2*v ; q = t / max(size-1,1) ; n = q - 1            (BasicSR)
//     u = ((n + 1) / 2) * (size - 1) ; [border: u = min(size-1, max(u, 0))]   (ATen unnormalize/clip)
// then ATen's bilinear weights (x_se - x)(y_se - y)... and the accumulation order nw, ne, sw, se.
#include "capi_common.h"

namespace {


struct Taps {
  int o_nw, o_ne, o_sw, o_se;      // offsets inside one (n, c) plane, clamped into the plane
  float w_nw, w_ne, w_sw, w_se;    // bilinear weights, 0 where the tap is out of bounds
};

__device__ __forceinline__ float unnormalize(float n, int size, bool align_corners) {
  if (align_corners) {
    return __fmul_rn(__fmul_rn(__fadd_rn(n, 1.0f), 0.5f), (float)(size - 1));   // x/2 == x*0.5 exactly
  }
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(n, 1.0f), (float)size), 1.0f), 0.5f);
}

// recip: `tensor / python_scalar` the way ATen's CUDA true-divide kernel evaluates it (BinaryDivTrueKernel.cu: a CPU-scalar
// divisor becomes a multiplication by inv_b = 1.0f / b, one rounding more than the division ATen's CPU kernel performs).
__device__ __forceinline__ Taps make_taps(float fx, float fy, int px, int py, int h, int w,
                                          bool border, bool align_corners, bool recip) {
  const float dw = (float)max(w - 1, 1), dh = (float)max(h - 1, 1);
  const float tx2 = __fmul_rn(2.0f, __fadd_rn((float)px, fx)), ty2 = __fmul_rn(2.0f, __fadd_rn((float)py, fy));
  float nx = __fsub_rn(recip ? __fmul_rn(tx2, __fdiv_rn(1.0f, dw)) : __fdiv_rn(tx2, dw), 1.0f);
  float ny = __fsub_rn(recip ? __fmul_rn(ty2, __fdiv_rn(1.0f, dh)) : __fdiv_rn(ty2, dh), 1.0f);
  float ix = unnormalize(nx, w, align_corners);
  float iy = unnormalize(ny, h, align_corners);
  if (border) {
    ix = fminf((float)(w - 1), fmaxf(ix, 0.0f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.0f));
  }
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float x1f = __fadd_rn(x0f, 1.0f), y1f = __fadd_rn(y0f, 1.0f);
  // huge |flow| would overflow the int conversion: saturate first (those taps are out of bounds anyway)
  const float lim = 1.0e9f;
  const int x0 = (int)fminf(fmaxf(x0f, -lim), lim), y0 = (int)fminf(fmaxf(y0f, -lim), lim);
  const int x1 = x0 + 1, y1 = y0 + 1;
  const bool vx0 = (x0 >= 0) & (x0 < w), vx1 = (x1 >= 0) & (x1 < w);
  const bool vy0 = (y0 >= 0) & (y0 < h), vy1 = (y1 >= 0) & (y1 < h);
  const int cx0 = min(max(x0, 0), w - 1), cx1 = min(max(x1, 0), w - 1);
  const int cy0 = min(max(y0, 0), h - 1), cy1 = min(max(y1, 0), h - 1);
  const float ax1 = __fsub_rn(x1f, ix), ax0 = __fsub_rn(ix, x0f);
  const float ay1 = __fsub_rn(y1f, iy), ay0 = __fsub_rn(iy, y0f);
  Taps t;
  t.o_nw = cy0 * w + cx0; t.o_ne = cy0 * w + cx1; t.o_sw = cy1 * w + cx0; t.o_se = cy1 * w + cx1;
  // (a NaN flow saturates to an out-of-bounds tap: weight 0)
  t.w_nw = (vx0 & vy0) ? __fmul_rn(ax1, ay1) : 0.0f;
  t.w_ne = (vx1 & vy0) ? __fmul_rn(ax0, ay1) : 0.0f;
  t.w_sw = (vx0 & vy1) ? __fmul_rn(ax1, ay0) : 0.0f;
  t.w_se = (vx1 & vy1) ? __fmul_rn(ax0, ay0) : 0.0f;
  return t;
}

__device__ __forceinline__ Taps zero_taps() { Taps t; t.o_nw = t.o_ne = t.o_sw = t.o_se = 0; t.w_nw = t.w_ne = t.w_sw = t.w_se = 0.f; return t; }

template <int TILE_W, int TILE_H, int CH_UNROLL, int MIN_BLOCKS = 5>
__global__ void __launch_bounds__(TILE_W * TILE_H, MIN_BLOCKS)
flow_warp_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                 int c, int h, int w, int c_per_cta, int border, int align_corners, int recip) {
  constexpr int NT = TILE_W * TILE_H;
  __shared__ int4 s_off[NT];
  __shared__ float4 s_wgt[NT];

  static_assert(TILE_W % 32 == 0 && NT <= 1024, "tile");
  const int tx = threadIdx.x & (TILE_W - 1), ty = threadIdx.x / TILE_W;
  const int px = blockIdx.x * TILE_W + tx, py = blockIdx.y * TILE_H + ty;
  const int c_splits = (c + c_per_cta - 1) / c_per_cta;
  const int n = blockIdx.z / c_splits;
  const int c0 = (blockIdx.z % c_splits) * c_per_cta;
  const int c1 = min(c0 + c_per_cta, c);
  const size_t plane = (size_t)h * w;
  const bool inside = (px < w) & (py < h);

  // phase 1: flow tile -> tap table in shared memory
  if (inside) {
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)n * plane + (size_t)py * w + px));
    const Taps t = make_taps(f.x, f.y, px, py, h, w, border != 0, align_corners != 0, recip != 0);
    s_off[threadIdx.x] = make_int4(t.o_nw, t.o_ne, t.o_sw, t.o_se);
    s_wgt[threadIdx.x] = make_float4(t.w_nw, t.w_ne, t.w_sw, t.w_se);
  } else {
    s_off[threadIdx.x] = make_int4(0, 0, 0, 0);
    s_wgt[threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  // (threads outside the image keep running: they take part in the shuffles below, with offset 0 and weight 0)

  // phase 2: sweep the channel planes of this CTA's slice.
  // The east taps of lane i are usually the west taps of lane i+1 (smooth flow): fetch them by shuffle and only issue
  // the load when the addresses differ -- this nearly halves the L1 traffic, which is what bounds the kernel.
  const int4 o = s_off[threadIdx.x];
  const float4 wt = s_wgt[threadIdx.x];
  const unsigned lane = threadIdx.x & 31;
  const int nx_nw = __shfl_down_sync(0xffffffffu, o.x, 1), nx_sw = __shfl_down_sync(0xffffffffu, o.z, 1);
  const bool sh_n = lane < 31 && nx_nw == o.y, sh_s = lane < 31 && nx_sw == o.w;
  const float* xp = x + ((size_t)n * c + c0) * plane;
  float* op = out + ((size_t)n * c + c0) * plane + (size_t)(inside ? py : 0) * w + (inside ? px : 0);
  int ch = c0;
  for (; ch + CH_UNROLL <= c1; ch += CH_UNROLL) {
    float a[CH_UNROLL], b[CH_UNROLL], d[CH_UNROLL], e[CH_UNROLL];
#pragma unroll
    for (int u = 0; u < CH_UNROLL; ++u) {
      const float* p = xp + (size_t)u * plane;
      a[u] = __ldg(p + o.x); d[u] = __ldg(p + o.z);
      if (!sh_n) b[u] = __ldg(p + o.y);
      if (!sh_s) e[u] = __ldg(p + o.w);
    }
#pragma unroll
    for (int u = 0; u < CH_UNROLL; ++u) {
      const float bs = __shfl_down_sync(0xffffffffu, a[u], 1), es = __shfl_down_sync(0xffffffffu, d[u], 1);
      const float bv = sh_n ? bs : b[u], ev = sh_s ? es : e[u];
      float acc = __fmul_rn(a[u], wt.x);
      acc = __fmaf_rn(bv, wt.y, acc);
      acc = __fmaf_rn(d[u], wt.z, acc);
      acc = __fmaf_rn(ev, wt.w, acc);
      if (inside) __stcs(op + (size_t)u * plane, acc);
    }
    xp += (size_t)CH_UNROLL * plane;
    op += (size_t)CH_UNROLL * plane;
  }
  for (; ch < c1; ++ch) {
    const float a0 = __ldg(xp + o.x), d0 = __ldg(xp + o.z);
    const float bs = __shfl_down_sync(0xffffffffu, a0, 1), es = __shfl_down_sync(0xffffffffu, d0, 1);
    const float bv = sh_n ? bs : __ldg(xp + o.y), ev = sh_s ? es : __ldg(xp + o.w);
    float acc = __fmul_rn(a0, wt.x);
    acc = __fmaf_rn(bv, wt.y, acc);
    acc = __fmaf_rn(d0, wt.z, acc);
    acc = __fmaf_rn(ev, wt.w, acc);
    if (inside) __stcs(op, acc);
    xp += plane;
    op += plane;
  }
}


}  // namespace

extern "C" int gpemsr_flow_warp_ex(const float* x, const float* flow, int n, int c, int h, int w,
                                   int padding_mode, int align_corners, int coord_form, float* out, gpemsr_stream_t stream) {
  using namespace gpemsr;
  if (coord_form != GPEMSR_COORD_DIV && coord_form != GPEMSR_COORD_RECIP)
    return set_error(GPEMSR_ERR_UNSUPPORTED, "flow_warp: coord_form %d (0 = true division, 1 = reciprocal multiply)", coord_form);
  if (n < 0 || c < 0 || h < 0 || w < 0)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "flow_warp: negative dimension n=%d c=%d h=%d w=%d", n, c, h, w);
  if (padding_mode != GPEMSR_PAD_ZEROS && padding_mode != GPEMSR_PAD_BORDER)
    return set_error(GPEMSR_ERR_UNSUPPORTED, "flow_warp: padding_mode %d (only zeros=0, border=1; the "
                     "reference never uses reflection)", padding_mode);
  if ((size_t)h * (size_t)w >= (size_t)1 << 31)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "flow_warp: h*w must be < 2^31");
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (n == 0 || c == 0 || h == 0 || w == 0) return GPEMSR_OK;     // empty input: nothing to do
  if (!x || !flow || !out) return set_error(GPEMSR_ERR_BAD_SHAPE, "flow_warp: null pointer");
  if (reinterpret_cast<uintptr_t>(flow) & 7)
    return set_error(GPEMSR_ERR_BAD_ALIGN, "flow_warp: flow must be 8-byte aligned");

  constexpr int TILE_W = 32, TILE_H = 8, CH_UNROLL = 4;
  const int tiles_x = (w + TILE_W - 1) / TILE_W, tiles_y = (h + TILE_H - 1) / TILE_H;
  // split channels across CTAs only when the pixel tiles alone cannot fill the chip (>= ~4 waves)
  const long long tiles = (long long)tiles_x * tiles_y * n;
  const long long want = 4LL * num_sms() * 8;
  int c_splits = 1;
  if (tiles < want) c_splits = (int)min((long long)((c + CH_UNROLL - 1) / CH_UNROLL), (want + tiles - 1) / tiles);
  if (c_splits < 1) c_splits = 1;
  int c_per_cta = (c + c_splits - 1) / c_splits;
  c_per_cta = ((c_per_cta + CH_UNROLL - 1) / CH_UNROLL) * CH_UNROLL;
  c_splits = (c + c_per_cta - 1) / c_per_cta;
  if ((long long)n * c_splits > 65535 || tiles_y > 65535)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "flow_warp: grid too large (n*c_splits=%lld, tiles_y=%d)",
                     (long long)n * c_splits, tiles_y);
  dim3 grid(tiles_x, tiles_y, n * c_splits);
  flow_warp_kernel<TILE_W, TILE_H, CH_UNROLL><<<grid, TILE_W * TILE_H, 0, (cudaStream_t)stream>>>(
      x, flow, out, c, h, w, c_per_cta, padding_mode == GPEMSR_PAD_BORDER, align_corners, coord_form == GPEMSR_COORD_RECIP);
  GPEMSR_LAUNCH_OK("flow_warp_kernel");
  return GPEMSR_OK;
}

extern "C" int gpemsr_flow_warp(const float* x, const float* flow, int n, int c, int h, int w,
                                int padding_mode, int align_corners, float* out, gpemsr_stream_t stream) {
  return gpemsr_flow_warp_ex(x, flow, n, c, h, w, padding_mode, align_corners, GPEMSR_COORD_DEFAULT, out, stream);
}
