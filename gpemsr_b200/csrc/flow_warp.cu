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
//            CH_UNROLL planes are in flight per thread (4*CH_UNROLL independent loads).
//
// Bit-faithful coordinates: each elementwise op of the reference is one separately
// rounded fp32 op here (__fadd_rn/__fmul_rn/__fdiv_rn, no contraction), in the reference order:
//     v = g + f ; t = 2*v ; q = t / max(size-1,1) ; n = q - 1            (BasicSR)
//     u = ((n + 1) / 2) * (size - 1) ; [border: u = min(size-1, max(u, 0))]   (ATen unnormalize/clip)
// then ATen's bilinear weights (x_se - x)(y_se - y)... and the accumulation order nw, ne, sw, se.
#include "capi_common.h"
#include <cstdlib>

namespace {

constexpr int THREADS = 256;

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

__device__ __forceinline__ Taps make_taps(float fx, float fy, int px, int py, int h, int w,
                                          bool border, bool align_corners) {
  const float dw = (float)max(w - 1, 1), dh = (float)max(h - 1, 1);
  float nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)px, fx)), dw), 1.0f);
  float ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn((float)py, fy)), dh), 1.0f);
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
                 int c, int h, int w, int c_per_cta, int border, int align_corners, int diag = 0) {
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
    const Taps t = make_taps(f.x, f.y, px, py, h, w, border != 0, align_corners != 0);
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
      if (diag == 2) { a[u] = wt.x; d[u] = wt.y; b[u] = wt.z; e[u] = wt.w; continue; }     // diagnostic: no loads
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
      if (inside && (diag != 1 || acc == 12345.678f)) __stcs(op + (size_t)u * plane, acc);                 // diag 1: no stores
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


// ---------------------------------------------------------------------------------------------------------------------
// Same mapping as flow_warp_kernel<32, 8, 4>, restructured for instruction count: the first version spent ~43 instructions
// per (pixel, channel), mostly 64-bit address chains, and was issue-bound.  Here every (tap, unrolled channel) offset is a
// 32-bit element offset computed ONCE per thread; inside the loop an address is one IMAD.WIDE off a single running
// pointer that advances CH_UNROLL planes per iteration.
template <int CH_UNROLL>
__global__ void __launch_bounds__(256, 3)
flow_warp_lean_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                      int c, int h, int w, int c_per_cta, int border, int align_corners) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int px = blockIdx.x * 32 + tx, py = blockIdx.y * 8 + ty;
  const int c_splits = (c + c_per_cta - 1) / c_per_cta;
  const int n = blockIdx.z / c_splits;
  const int c0 = (blockIdx.z % c_splits) * c_per_cta;
  const int c1 = min(c0 + c_per_cta, c);
  const size_t plane = (size_t)h * w;
  const bool inside = (px < w) & (py < h);
  Taps t = zero_taps();
  if (inside) {
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)n * plane + (size_t)py * w + px));
    t = make_taps(f.x, f.y, px, py, h, w, border != 0, align_corners != 0);
  }
  const unsigned lane = threadIdx.x & 31;
  const int nx_nw = __shfl_down_sync(0xffffffffu, t.o_nw, 1), nx_sw = __shfl_down_sync(0xffffffffu, t.o_sw, 1);
  const bool sh_n = lane < 31 && nx_nw == t.o_ne, sh_s = lane < 31 && nx_sw == t.o_se;
  const int pl = (int)plane, pix = inside ? py * w + px : 0;
  int e_nw[CH_UNROLL], e_ne[CH_UNROLL], e_sw[CH_UNROLL], e_se[CH_UNROLL], e_o[CH_UNROLL];
#pragma unroll
  for (int u = 0; u < CH_UNROLL; ++u) {
    e_nw[u] = t.o_nw + u * pl; e_ne[u] = t.o_ne + u * pl; e_sw[u] = t.o_sw + u * pl; e_se[u] = t.o_se + u * pl; e_o[u] = pix + u * pl;
  }
  const float* xp = x + ((size_t)n * c + c0) * plane;
  float* op = out + ((size_t)n * c + c0) * plane;
  const size_t step = (size_t)CH_UNROLL * plane;
  int ch = c0;
  for (; ch + CH_UNROLL <= c1; ch += CH_UNROLL) {
    float a[CH_UNROLL], b[CH_UNROLL], d[CH_UNROLL], e[CH_UNROLL];
#pragma unroll
    for (int u = 0; u < CH_UNROLL; ++u) {
      a[u] = __ldg(xp + e_nw[u]); d[u] = __ldg(xp + e_sw[u]);
      if (!sh_n) b[u] = __ldg(xp + e_ne[u]);
      if (!sh_s) e[u] = __ldg(xp + e_se[u]);
    }
#pragma unroll
    for (int u = 0; u < CH_UNROLL; ++u) {
      const float bs = __shfl_down_sync(0xffffffffu, a[u], 1), es = __shfl_down_sync(0xffffffffu, d[u], 1);
      const float bv = sh_n ? bs : b[u], ev = sh_s ? es : e[u];
      float acc = __fmul_rn(a[u], t.w_nw);
      acc = __fmaf_rn(bv, t.w_ne, acc);
      acc = __fmaf_rn(d[u], t.w_sw, acc);
      acc = __fmaf_rn(ev, t.w_se, acc);
      if (inside) __stcs(op + e_o[u], acc);
    }
    xp += step; op += step;
  }
  for (; ch < c1; ++ch) {
    const float a0 = __ldg(xp + e_nw[0]), d0 = __ldg(xp + e_sw[0]);
    const float bs = __shfl_down_sync(0xffffffffu, a0, 1), es = __shfl_down_sync(0xffffffffu, d0, 1);
    const float bv = sh_n ? bs : __ldg(xp + e_ne[0]), ev = sh_s ? es : __ldg(xp + e_se[0]);
    float acc = __fmul_rn(a0, t.w_nw);
    acc = __fmaf_rn(bv, t.w_ne, acc);
    acc = __fmaf_rn(d0, t.w_sw, acc);
    acc = __fmaf_rn(ev, t.w_se, acc);
    if (inside) __stcs(op + e_o[0], acc);
    xp += plane; op += plane;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// ROWS vertically adjacent pixels per thread.  ncu shows the L1 data pipe (one 128-byte wavefront per clock per SM) is what
// bounds this kernel: a warp-wide tap load of 32 neighbouring pixels spans ~3 cache lines, and every input row is fetched
// twice -- as the south taps of output row y and as the north taps of row y+1.  A thread that owns rows y..y+ROWS-1 keeps
// the row values in registers: for a smooth flow the south pair of row r IS the north pair of row r+1, so ROWS+1 row loads
// serve ROWS output rows (instead of 2*ROWS); east taps come from the right-hand lane by shuffle as before.  Rows / lanes
// whose taps do not line up fall back to their own (predicated) loads, so any flow is handled exactly.
template <int ROWS, int CH_UNROLL>
__global__ void __launch_bounds__(256, 2)
flow_warp_rows_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                      int c, int h, int w, int c_per_cta, int border, int align_corners) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int px = blockIdx.x * 32 + lane, py0 = (blockIdx.y * 8 + wrp) * ROWS;
  const int c_splits = (c + c_per_cta - 1) / c_per_cta;
  const int n = blockIdx.z / c_splits;
  const int c0 = (blockIdx.z % c_splits) * c_per_cta;
  const int c1 = min(c0 + c_per_cta, c);
  const size_t plane = (size_t)h * w;

  Taps T[ROWS];
  bool in[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    in[r] = (px < w) & (py0 + r < h);
    T[r] = zero_taps();
    if (in[r]) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)n * plane + (size_t)(py0 + r) * w + px));
      T[r] = make_taps(f.x, f.y, px, py0 + r, h, w, border != 0, align_corners != 0);
    }
  }
  // row loads k = 0..ROWS: k = 0 is the north pair of row 0, k >= 1 the south pair of row k-1
  int W[ROWS + 1], E[ROWS + 1];
  W[0] = T[0].o_nw; E[0] = T[0].o_ne;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) { W[r + 1] = T[r].o_sw; E[r + 1] = T[r].o_se; }
  bool he[ROWS + 1];           // east value of load k comes from the right-hand lane
#pragma unroll
  for (int k = 0; k <= ROWS; ++k) he[k] = (lane < 31) & (__shfl_down_sync(0xffffffffu, W[k], 1) == E[k]);
  bool vs[ROWS];               // north pair of row r is row load r (r = 0: by definition)
  vs[0] = true;
#pragma unroll
  for (int r = 1; r < ROWS; ++r) vs[r] = (T[r].o_nw == W[r]) & (T[r].o_ne == E[r]);

  const float* xp = x + ((size_t)n * c + c0) * plane;
  float* op = out + ((size_t)n * c + c0) * plane + (size_t)py0 * w + px;
  for (int ch = c0; ch < c1; ch += CH_UNROLL) {
    float wv[CH_UNROLL][ROWS + 1], ev[CH_UNROLL][ROWS + 1], fn[CH_UNROLL][ROWS], fe[CH_UNROLL][ROWS];
#pragma unroll
    for (int u = 0; u < CH_UNROLL; ++u) {
      const float* q = xp + (size_t)u * plane;
      const bool live = ch + u < c1;
#pragma unroll
      for (int k = 0; k <= ROWS; ++k) {
        wv[u][k] = live ? __ldg(q + W[k]) : 0.f;
        if (live && !he[k]) ev[u][k] = __ldg(q + E[k]);
      }
#pragma unroll
      for (int r = 1; r < ROWS; ++r)
        if (live && !vs[r]) { fn[u][r] = __ldg(q + T[r].o_nw); fe[u][r] = __ldg(q + T[r].o_ne); }
    }
#pragma unroll
    for (int u = 0; u < CH_UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k <= ROWS; ++k) {
        const float sh = __shfl_down_sync(0xffffffffu, wv[u][k], 1);
        if (he[k]) ev[u][k] = sh;
      }
      if (ch + u < c1) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float nwv = vs[r] ? wv[u][r] : fn[u][r], nev = vs[r] ? ev[u][r] : fe[u][r];
          float acc = __fmul_rn(nwv, T[r].w_nw);
          acc = __fmaf_rn(nev, T[r].w_ne, acc);
          acc = __fmaf_rn(wv[u][r + 1], T[r].w_sw, acc);
          acc = __fmaf_rn(ev[u][r + 1], T[r].w_se, acc);
          if (in[r]) __stcs(op + (size_t)u * plane + (size_t)r * w, acc);
        }
      }
    }
    xp += (size_t)CH_UNROLL * plane; op += (size_t)CH_UNROLL * plane;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two pixels per lane.  The first version spent ~43 instructions per (pixel, channel) -- mostly 64-bit address arithmetic
// for four independent tap pointers -- and was issue-bound at 54 % of HBM bandwidth.  Here a lane owns the pixel pair
// (x, x+1) of a 64-pixel row segment.  For a smooth flow the six taps of the pair are three consecutive floats in each of
// two rows, so per channel a lane issues   n0 = pn[0], n1 = pn[1], s0 = ps[0], s1 = ps[1]   (immediate offsets off TWO
// running pointers), gets the third float of each row from its right-hand neighbour by shuffle, and stores a float2.
// Warps whose taps do not line up this way (large / noisy flow, clamped borders) take the generic 8-load path.

template <int CH_UNROLL>
__global__ void __launch_bounds__(256, 4)
flow_warp_pair_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                      int c, int h, int w, int c_per_cta, int border, int align_corners) {
  const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
  const int px = blockIdx.x * 64 + 2 * lane, py = blockIdx.y * 8 + wrow;
  const int c_splits = (c + c_per_cta - 1) / c_per_cta;
  const int n = blockIdx.z / c_splits;
  const int c0 = (blockIdx.z % c_splits) * c_per_cta;
  const int c1 = min(c0 + c_per_cta, c);
  const size_t plane = (size_t)h * w;
  const bool in_a = (px < w) & (py < h), in_b = (px + 1 < w) & (py < h);
  const bool vec = (w & 1) == 0;                         // pixel pairs are 8-byte aligned in every row

  Taps A = zero_taps(), B = zero_taps();
  if (in_a) {
    const float* fp = flow + ((size_t)n * plane + (size_t)py * w + px) * 2;
    float4 f;
    if (vec && in_b) f = __ldg(reinterpret_cast<const float4*>(fp));
    else { const float2 fa = __ldg(reinterpret_cast<const float2*>(fp)); f.x = fa.x; f.y = fa.y; f.z = f.w = 0.f;
           if (in_b) { const float2 fb = __ldg(reinterpret_cast<const float2*>(fp) + 1); f.z = fb.x; f.w = fb.y; } }
    A = make_taps(f.x, f.y, px, py, h, w, border != 0, align_corners != 0);
    if (in_b) B = make_taps(f.z, f.w, px + 1, py, h, w, border != 0, align_corners != 0);
  }
  // does the pair line up as three consecutive floats per row, and does the right neighbour continue it?
  const bool lined = (A.o_ne == A.o_nw + 1) & (A.o_se == A.o_sw + 1) & (!in_b | ((B.o_nw == A.o_ne) & (B.o_sw == A.o_se)));
  const int nb_n = __shfl_down_sync(0xffffffffu, A.o_nw, 1), nb_s = __shfl_down_sync(0xffffffffu, A.o_sw, 1);
  const bool sh_n = lane < 31 && nb_n == B.o_ne, sh_s = lane < 31 && nb_s == B.o_se;
  const bool fast = __all_sync(0xffffffffu, lined | !in_a);

  const float* xp = x + ((size_t)n * c + c0) * plane;
  float* op = out + ((size_t)n * c + c0) * plane + (in_a ? (size_t)py * w + px : 0);
  if (fast) {
    const float* pn = xp + A.o_nw;
    const float* ps = xp + A.o_sw;
    const int d_n2 = B.o_ne - A.o_nw, d_s2 = B.o_se - A.o_sw;      // third float of each row (only read when not shuffled)
    int ch = c0;
    for (; ch + CH_UNROLL <= c1; ch += CH_UNROLL) {
      float n0[CH_UNROLL], n1[CH_UNROLL], n2[CH_UNROLL], s0[CH_UNROLL], s1[CH_UNROLL], s2[CH_UNROLL];
#pragma unroll
      for (int u = 0; u < CH_UNROLL; ++u) {
        const float* qn = pn + (size_t)u * plane;
        const float* qs = ps + (size_t)u * plane;
        n0[u] = __ldg(qn); n1[u] = __ldg(qn + 1); s0[u] = __ldg(qs); s1[u] = __ldg(qs + 1);
        if (!sh_n) n2[u] = __ldg(qn + d_n2);
        if (!sh_s) s2[u] = __ldg(qs + d_s2);
      }
#pragma unroll
      for (int u = 0; u < CH_UNROLL; ++u) {
        const float tn = __shfl_down_sync(0xffffffffu, n0[u], 1), ts = __shfl_down_sync(0xffffffffu, s0[u], 1);
        const float bne = sh_n ? tn : n2[u], bse = sh_s ? ts : s2[u];
        float ra = __fmul_rn(n0[u], A.w_nw);
        ra = __fmaf_rn(n1[u], A.w_ne, ra); ra = __fmaf_rn(s0[u], A.w_sw, ra); ra = __fmaf_rn(s1[u], A.w_se, ra);
        float rb = __fmul_rn(n1[u], B.w_nw);
        rb = __fmaf_rn(bne, B.w_ne, rb); rb = __fmaf_rn(s1[u], B.w_sw, rb); rb = __fmaf_rn(bse, B.w_se, rb);
        float* q = op + (size_t)u * plane;
        if (vec && in_b) __stcs(reinterpret_cast<float2*>(q), make_float2(ra, rb));
        else { if (in_a) __stcs(q, ra); if (in_b) __stcs(q + 1, rb); }
      }
      pn += (size_t)CH_UNROLL * plane; ps += (size_t)CH_UNROLL * plane; op += (size_t)CH_UNROLL * plane;
    }
    for (; ch < c1; ++ch) {
      const float a0 = __ldg(pn), a1 = __ldg(pn + 1), b0 = __ldg(ps), b1 = __ldg(ps + 1);
      const float tn = __shfl_down_sync(0xffffffffu, a0, 1), ts = __shfl_down_sync(0xffffffffu, b0, 1);
      const float bne = sh_n ? tn : __ldg(pn + d_n2), bse = sh_s ? ts : __ldg(ps + d_s2);
      float ra = __fmul_rn(a0, A.w_nw);
      ra = __fmaf_rn(a1, A.w_ne, ra); ra = __fmaf_rn(b0, A.w_sw, ra); ra = __fmaf_rn(b1, A.w_se, ra);
      float rb = __fmul_rn(a1, B.w_nw);
      rb = __fmaf_rn(bne, B.w_ne, rb); rb = __fmaf_rn(b1, B.w_sw, rb); rb = __fmaf_rn(bse, B.w_se, rb);
      if (vec && in_b) __stcs(reinterpret_cast<float2*>(op), make_float2(ra, rb));
      else { if (in_a) __stcs(op, ra); if (in_b) __stcs(op + 1, rb); }
      pn += plane; ps += plane; op += plane;
    }
  } else {
    // generic path: eight independent taps
    for (int ch = c0; ch < c1; ++ch) {
      float ra = __fmul_rn(__ldg(xp + A.o_nw), A.w_nw);
      ra = __fmaf_rn(__ldg(xp + A.o_ne), A.w_ne, ra); ra = __fmaf_rn(__ldg(xp + A.o_sw), A.w_sw, ra);
      ra = __fmaf_rn(__ldg(xp + A.o_se), A.w_se, ra);
      float rb = __fmul_rn(__ldg(xp + B.o_nw), B.w_nw);
      rb = __fmaf_rn(__ldg(xp + B.o_ne), B.w_ne, rb); rb = __fmaf_rn(__ldg(xp + B.o_sw), B.w_sw, rb);
      rb = __fmaf_rn(__ldg(xp + B.o_se), B.w_se, rb);
      if (in_a) __stcs(op, ra);
      if (in_b) __stcs(op + 1, rb);
      xp += plane; op += plane;
    }
  }
}

}  // namespace

extern "C" int gpemsr_flow_warp(const float* x, const float* flow, int n, int c, int h, int w,
                                int padding_mode, int align_corners, float* out, gpemsr_stream_t stream) {
  using namespace gpemsr;
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

  static int variant = -1, diag = 0;
  if (variant < 0) {
    const char* e = getenv("GPEMSR_FLOW_VARIANT"); variant = e ? atoi(e) : 0;
    const char* dg = getenv("GPEMSR_FLOW_DIAG"); diag = dg ? atoi(dg) : 0;
  }
  const int TILE_W = variant == 2 || variant == 4 || variant == 10 ? 64 : variant == 3 ? 128 : 32;
  const int TILE_H = variant == 9 ? 32 : variant == 10 || variant == 11 ? 16 : THREADS / TILE_W;
  const int CH_UNROLL = variant == 1 || variant == 4 ? 8 : 4;
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
  const int bd = padding_mode == GPEMSR_PAD_BORDER;
  cudaStream_t st = (cudaStream_t)stream;
  if (variant == 16 || variant == 17 || variant == 18) {
    const int rows = variant == 17 ? 2 : 4;
    dim3 g3((w + 31) / 32, (h + 8 * rows - 1) / (8 * rows), n * c_splits);
    if (variant == 16) flow_warp_rows_kernel<4, 2><<<g3, 256, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners);
    else if (variant == 17) flow_warp_rows_kernel<2, 4><<<g3, 256, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners);
    else flow_warp_rows_kernel<4, 1><<<g3, 256, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners);
    GPEMSR_LAUNCH_OK("flow_warp_rows_kernel");
    return GPEMSR_OK;
  }
  if (variant == 14 || variant == 15) {
    if ((size_t)h * w * 8 >= ((size_t)1 << 31)) return set_error(GPEMSR_ERR_BAD_SHAPE, "flow_warp: plane too large for 32-bit offsets");
    if (variant == 14) flow_warp_lean_kernel<4><<<grid, 256, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners);
    else flow_warp_lean_kernel<8><<<grid, 256, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners);
    GPEMSR_LAUNCH_OK("flow_warp_lean_kernel");
    return GPEMSR_OK;
  }
  if (variant == 12 || variant == 13) {
    dim3 g2((w + 63) / 64, (h + 7) / 8, n * c_splits);
    if (variant == 12) flow_warp_pair_kernel<4><<<g2, 256, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners);
    else flow_warp_pair_kernel<2><<<g2, 256, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners);
    GPEMSR_LAUNCH_OK("flow_warp_pair_kernel");
    return GPEMSR_OK;
  }
  switch (variant) {
    case 1: flow_warp_kernel<32, 8, 8><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 2: flow_warp_kernel<64, 4, 4><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 3: flow_warp_kernel<128, 2, 4><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 4: flow_warp_kernel<64, 4, 8><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 5: flow_warp_kernel<32, 8, 4, 8><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 6: flow_warp_kernel<32, 8, 4, 6><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 7: flow_warp_kernel<32, 8, 8, 4><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 9: flow_warp_kernel<32, 32, 4, 1><<<grid, 1024, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 10: flow_warp_kernel<64, 16, 4, 1><<<grid, 1024, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 11: flow_warp_kernel<32, 16, 4, 2><<<grid, 512, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    case 8: flow_warp_kernel<32, 8, 2, 8><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners); break;
    default: flow_warp_kernel<32, 8, 4><<<grid, THREADS, 0, st>>>(x, flow, out, c, h, w, c_per_cta, bd, align_corners, diag); break;
  }
  GPEMSR_LAUNCH_OK("flow_warp_kernel");
  return GPEMSR_OK;
}
