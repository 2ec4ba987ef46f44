// Memory-bound glue of GPEMSR.forward around the implicit-GEMM convolutions (model/GPEMSR.py:64-234, 323-456): the
// reference-feature fusion, the POD alignment pyramid and the ThreeDA fusion are chains of convolutions (gpemsr_igemm)
// connected by the element-wise / resampling steps below.  All of them work on the padded K8-blocked activation format
// (one thread = one 8-channel cell of one pixel: 32-byte fp32 loads, 16-byte bf16 stores, lanes <-> consecutive pixels),
// write straight into a channel slot of the consumer's operand buffer (torch.cat never materialises) and are HBM bound.
#include "act_layout.cuh"

namespace {

struct Cell { float v[8]; };

__device__ __forceinline__ Cell load_cell(const float* __restrict__ p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  return Cell{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}

__device__ __forceinline__ void store_cell(float (&v)[8], size_t cell, float* __restrict__ f32, __nv_bfloat16* __restrict__ hi,
                                           __nv_bfloat16* __restrict__ lo) {
  if (f32) {
    *reinterpret_cast<float4*>(f32 + cell) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(f32 + cell + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (hi) {
    uint4 h, l;
    split8(v, h, l);
    *reinterpret_cast<uint4*>(hi + cell) = h;
    if (lo) *reinterpret_cast<uint4*>(lo + cell) = l;
  }
}

__device__ __forceinline__ float sigmoidf(float v) { return 1.0f / (1.0f + expf(-v)); }

// ATen upsample_bilinear2d source index for align_corners=False with a given scale factor s: max((dst + 0.5) / s - 0.5, 0)
struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_src(int dst, float rscale, int in_size) {
  float s = rscale * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  Lerp r;
  r.i0 = min((int)s, in_size - 1);
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = s - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

// decode the flat thread index (img, cell, pixel) with the pixel fastest
__device__ __forceinline__ bool decode_t(long long t, const Geom& g, int cells, int& img, int& cc, int& y, int& x) {
  const long long hw = (long long)g.h * g.w, total = (long long)g.n * cells * hw;
  if (t >= total) return false;
  if (total < (1LL << 32)) {        // every shape of the model: 32-bit divisions (a 64-bit one is ~4x the instructions, and these
    const unsigned tt = (unsigned)t, uhw = (unsigned)hw, q = tt / uhw, p = tt - q * uhw;      // kernels move 64 bytes per thread)
    img = (int)(q / (unsigned)cells); cc = (int)(q - (unsigned)img * (unsigned)cells);
    y = (int)(p / (unsigned)g.w); x = (int)(p - (unsigned)y * (unsigned)g.w);
    return true;
  }
  const long long p = t % hw;
  cc = (int)((t / hw) % cells); img = (int)(t / (hw * cells));
  y = (int)(p / g.w); x = (int)(p % g.w);
  return true;
}

// F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False) * mul  (model/GPEMSR.py:130,132,136,142,144,148,226,231)
__global__ void cells_upsample2x_kernel(const float* __restrict__ x, Geom gi, int cells, float mul, Geom go, int cell_off,
                                        float* __restrict__ of32, __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  int img, cc, oy, ox;
  if (!decode_t((long long)blockIdx.x * blockDim.x + threadIdx.x, go, cells, img, cc, oy, ox)) return;
  const Lerp ly = lerp_src(oy, 0.5f, gi.h), lx = lerp_src(ox, 0.5f, gi.w);
  const float* base = x + (size_t)cc * gi.rows_alloc * 8;
  const Cell a = load_cell(base + place_row(gi, img, ly.i0, lx.i0) * 8), b = load_cell(base + place_row(gi, img, ly.i0, lx.i1) * 8);
  const Cell c = load_cell(base + place_row(gi, img, ly.i1, lx.i0) * 8), d = load_cell(base + place_row(gi, img, ly.i1, lx.i1) * 8);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    v[j] = (ly.l0 * (lx.l0 * a.v[j] + lx.l1 * b.v[j]) + ly.l1 * (lx.l0 * c.v[j] + lx.l1 * d.v[j])) * mul;
  store_cell(v, ((size_t)(cell_off + cc) * go.rows_alloc + place_row(go, img, oy, ox)) * 8, of32, ohi, olo);
}

// x * F.interpolate(sigmoid?(mask), scale_factor=s, bilinear, align_corners=False)  (model/GPEMSR.py:357-362, 368-376)
__global__ void cells_mul_mask_kernel(const float* __restrict__ x, Geom g, int cells, const float* __restrict__ mask, int hm, int wm,
                                      float rscale, int sig, int cell_off, float* __restrict__ of32,
                                      __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  int img, cc, y, xx;
  if (!decode_t((long long)blockIdx.x * blockDim.x + threadIdx.x, g, cells, img, cc, y, xx)) return;
  const Lerp ly = lerp_src(y, rscale, hm), lx = lerp_src(xx, rscale, wm);
  const float* m = mask + (long long)img * hm * wm;
  float m00 = __ldg(m + (long long)ly.i0 * wm + lx.i0), m01 = __ldg(m + (long long)ly.i0 * wm + lx.i1);
  float m10 = __ldg(m + (long long)ly.i1 * wm + lx.i0), m11 = __ldg(m + (long long)ly.i1 * wm + lx.i1);
  if (sig) { m00 = sigmoidf(m00); m01 = sigmoidf(m01); m10 = sigmoidf(m10); m11 = sigmoidf(m11); }
  const float mv = ly.l0 * (lx.l0 * m00 + lx.l1 * m01) + ly.l1 * (lx.l0 * m10 + lx.l1 * m11);
  const size_t row = place_row(g, img, y, xx);
  const Cell a = load_cell(x + ((size_t)cc * g.rows_alloc + row) * 8);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = a.v[j] * mv;
  store_cell(v, ((size_t)(cell_off + cc) * g.rows_alloc + row) * 8, of32, ohi, olo);
}

// cat([MaxPool2d(3, 2, 1)(x), AvgPool2d(3, 2, 1)(x)], dim=1)  (model/GPEMSR.py:217-218, 222-223; the average counts the
// zero padding: count_include_pad=True divides by 9 everywhere; the maximum ignores it)
__global__ void cells_pool3x3s2_kernel(const float* __restrict__ x, Geom gi, int cells, Geom go, __nv_bfloat16* __restrict__ ohi,
                                       __nv_bfloat16* __restrict__ olo) {
  int img, cc, oy, ox;
  if (!decode_t((long long)blockIdx.x * blockDim.x + threadIdx.x, go, cells, img, cc, oy, ox)) return;
  float mx[8], sm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mx[j] = -INFINITY; sm[j] = 0.f; }
  const float* base = x + (size_t)cc * gi.rows_alloc * 8;
  for (int dy = -1; dy <= 1; ++dy) {
    const int y = 2 * oy + dy;
    if (y < 0 || y >= gi.h) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = 2 * ox + dx;
      if (xx < 0 || xx >= gi.w) continue;
      const Cell a = load_cell(base + place_row(gi, img, y, xx) * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) { mx[j] = fmaxf(mx[j], a.v[j]); sm[j] += a.v[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[j] = sm[j] / 9.0f;
  const size_t row = place_row(go, img, oy, ox);
  store_cell(mx, ((size_t)cc * go.rows_alloc + row) * 8, nullptr, ohi, olo);
  store_cell(sm, ((size_t)(cells + cc) * go.rows_alloc + row) * 8, nullptr, ohi, olo);
}

// copy `cells` channel cells between activation buffers; bcast_t > 0: destination image i takes source image
// (i / bcast_t) * bcast_t + center (the centre frame's features repeated for every neighbour, model/GPEMSR.py:421-431)
__global__ void cells_copy_kernel(const float* __restrict__ sf32, const uint4* __restrict__ shi, const uint4* __restrict__ slo, Geom gs,
                                  int src_cell_off, int cells, int bcast_t, int center, Geom gd, int cell_off,
                                  float* __restrict__ df32, uint4* __restrict__ dhi, uint4* __restrict__ dlo) {
  int img, cc, y, x;
  if (!decode_t((long long)blockIdx.x * blockDim.x + threadIdx.x, gd, cells, img, cc, y, x)) return;
  const int simg = bcast_t > 0 ? (img / bcast_t) * bcast_t + center : img;
  const size_t s = (size_t)(src_cell_off + cc) * gs.rows_alloc + place_row(gs, simg, y, x);
  const size_t d = (size_t)(cell_off + cc) * gd.rows_alloc + place_row(gd, img, y, x);
  if (df32 && sf32) {
    const float4* sp = reinterpret_cast<const float4*>(sf32 + s * 8);
    float4* dp = reinterpret_cast<float4*>(df32 + d * 8);
    dp[0] = __ldg(sp); dp[1] = __ldg(sp + 1);
  }
  if (dhi && shi) dhi[d] = __ldg(shi + s);
  if (dlo && slo) dlo[d] = __ldg(slo + s);
}

// ThreeDA temporal attention (model/GPEMSR.py:181-196): for frame i of a window of t frames
//   corr = sigmoid(sum_c emb[i, c] * emb_ref[c]);  out[b, i * c + ch] = aligned[i, ch] * corr
// emb / aligned have n = b * t images, emb_ref and the output have n = b images; one thread = one (frame, pixel).
__global__ void temporal_attn_scale_kernel(const float* __restrict__ emb, const float* __restrict__ emb_ref,
                                           const float* __restrict__ aligned, Geom g, Geom gr, int cells, int t, Geom go,
                                           __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  int img, one, y, x;
  if (!decode_t((long long)blockIdx.x * blockDim.x + threadIdx.x, g, 1, img, one, y, x)) return;
  const int b = img / t, i = img - b * t;
  const size_t row = place_row(g, img, y, x), rrow = place_row(gr, b, y, x), orow = place_row(go, b, y, x);
  float corr = 0.f;
  for (int cc = 0; cc < cells; ++cc) {
    const Cell e = load_cell(emb + ((size_t)cc * g.rows_alloc + row) * 8), r = load_cell(emb_ref + ((size_t)cc * gr.rows_alloc + rrow) * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) corr = fmaf(e.v[j], r.v[j], corr);
  }
  corr = sigmoidf(corr);
  for (int cc = 0; cc < cells; ++cc) {
    const Cell a = load_cell(aligned + ((size_t)cc * g.rows_alloc + row) * 8);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = a.v[j] * corr;
    store_cell(v, ((size_t)(i * cells + cc) * go.rows_alloc + orow) * 8, nullptr, ohi, olo);
  }
}

// ThreeDA output (model/GPEMSR.py:233): feat * sigmoid(attn) * 2 + attn_add + fea_3d2 + fea_3d3
__global__ void threeda_combine_kernel(const float* __restrict__ feat, const float* __restrict__ attn, const float* __restrict__ attn_add,
                                       const float* __restrict__ f3d2, const float* __restrict__ f3d3, Geom g, int cells,
                                       float* __restrict__ of32, __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  int img, cc, y, x;
  if (!decode_t((long long)blockIdx.x * blockDim.x + threadIdx.x, g, cells, img, cc, y, x)) return;
  const size_t cell = ((size_t)cc * g.rows_alloc + place_row(g, img, y, x)) * 8;
  const Cell f = load_cell(feat + cell), a = load_cell(attn + cell), ad = load_cell(attn_add + cell), c2 = load_cell(f3d2 + cell),
             c3 = load_cell(f3d3 + cell);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = ((f.v[j] * sigmoidf(a.v[j]) * 2.0f + ad.v[j]) + c2.v[j]) + c3.v[j];
  store_cell(v, cell, of32, ohi, olo);
}

// Small strided 3x3 convolution on CUDA cores (POD.flowdsconv*, model/GPEMSR.py:71-76, 101-106: 2 -> 16 channels at stride 4,
// 16 -> 16 at stride 2, padding 1): a few hundred MACs per output on LR-sized maps -- far too small for a tensor-core tile.
__global__ void conv3x3_direct_kernel(const float* __restrict__ x, int n, int cin, int h, int w, const float* __restrict__ wgt,
                                      const float* __restrict__ bias, int cout, int stride, int ho, int wo, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * cout * ho * wo) return;
  const int ox = (int)(t % wo), oy = (int)((t / wo) % ho), co = (int)((t / ((long long)wo * ho)) % cout);
  const int img = (int)(t / ((long long)wo * ho * cout));
  float acc = bias ? __ldg(bias + co) : 0.f;
  for (int ci = 0; ci < cin; ++ci) {
    const float* p = x + ((long long)img * cin + ci) * h * w;
    const float* wk = wgt + ((long long)co * cin + ci) * 9;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int y = oy * stride - 1 + ky;
      if (y < 0 || y >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = ox * stride - 1 + kx;
        if (xx < 0 || xx >= w) continue;
        acc = fmaf(__ldg(p + (long long)y * w + xx), __ldg(wk + ky * 3 + kx), acc);
      }
    }
  }
  out[t] = acc;
}

inline unsigned blocks_for(long long total) { return (unsigned)((total + 255) / 256); }


// 3x3 convolutions with very few output channels (conv_last 64 -> 1, refmaskconv3 64 -> 1, the composed last decoder stage
// 64 -> 4 phases): on the tensor pipe an M = 128 MMA costs the same ~60 cycles for N = 16 as for N = 64, and the nine taps make
// it nine of them per k-step -- 1.4 % of peak.  Instead the nine taps become nine OUTPUT COLUMNS of ONE 1x1 GEMM
//     D[row, tap * n_out + o] = sum_c x[row, c] * w[o, c, tap]          (gpemsr_igemm, fp32 cells; ring rows stay zero)
// and this kernel finishes the convolution with a nine-point shifted sum on CUDA cores:
//     out[o](y, x) = act(bias[o] + sum_tap D[(y + dy_tap, x + dx_tap), tap * n_out + o])   (+ the bilinear base image, :452-455)
// up == 2: column o = phase * co + ch is channel ch of output pixel (2y + phase / 2, 2x + phase % 2) (the composed up-block).
__global__ void tap_gather_sum_kernel(const float* __restrict__ d, Geom g, int n_out, int up, int co, const float* __restrict__ bias,
                                      int act, float slope, const float* __restrict__ base, int bh, int bw, int bscale,
                                      float* __restrict__ out) {
  const long long hw = (long long)g.h * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)g.n * hw) return;
  const int img = (int)(t / hw), y = (int)((t % hw) / g.w), x = (int)(t % g.w);
  const long long row = place_row(g, img, y, x);
  const int wp = g.wp();
  float acc[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = (bias && o < n_out) ? __ldg(bias + o) : 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const long long r = row + (long long)(tap / 3 - 1) * wp + (tap % 3 - 1);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      if (o < n_out) {
        const int col = tap * n_out + o;
        acc[o] += __ldg(d + ((size_t)(col >> 3) * g.rows_alloc + r) * 8 + (col & 7));
      }
    }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) acc[o] = apply_act(acc[o], act, slope);
  if (up == 1) {
    float b = 0.f;
    if (base) {                                  // ATen upsample_bilinear2d(x_center, scale_factor, align_corners=False)
      const float rs = 1.0f / (float)bscale;
      const Lerp ly = lerp_src(y, rs, bh), lx = lerp_src(x, rs, bw);
      const float* p = base + (long long)img * bh * bw;
      b = ly.l0 * (lx.l0 * __ldg(p + (long long)ly.i0 * bw + lx.i0) + lx.l1 * __ldg(p + (long long)ly.i0 * bw + lx.i1)) +
          ly.l1 * (lx.l0 * __ldg(p + (long long)ly.i1 * bw + lx.i0) + lx.l1 * __ldg(p + (long long)ly.i1 * bw + lx.i1));
    }
    for (int o = 0; o < n_out; ++o) out[((long long)img * n_out + o) * hw + (long long)y * g.w + x] = acc[o] + b;
  } else {
    const long long Wo = 2LL * g.w, plane = 4 * hw;
    for (int o = 0; o < n_out; ++o) {
      const int ph = o / co, ch = o - ph * co;
      out[((long long)img * co + ch) * plane + (long long)(2 * y + (ph >> 1)) * Wo + 2 * x + (ph & 1)] = acc[o];
    }
  }
}

}  // namespace

extern "C" {

int gpemsr_cells_upsample2x(const float* x_f32, const gpemsr_geom_t* gi, int c, float mul, const gpemsr_geom_t* go, int c_off,
                            float* out_f32, void* out_hi, void* out_lo, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_f32 || !gi || !go || c <= 0 || c % 8 || c_off % 8 || !(out_f32 || out_hi))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_upsample2x: bad arguments (c and c_off must be multiples of 8)");
  if ((rc = check_geom(*gi, "cells_upsample2x(in)")) != GPEMSR_OK || (rc = check_geom(*go, "cells_upsample2x(out)")) != GPEMSR_OK) return rc;
  if (go->n != gi->n || go->h != 2 * gi->h || go->w != 2 * gi->w)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_upsample2x: the output must be the x2 grid of the input");
  const long long total = (long long)go->n * (c / 8) * go->h * go->w;
  cells_upsample2x_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(x_f32, to_geom(*gi), c / 8, mul, to_geom(*go), c_off / 8,
                                                                              out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  GPEMSR_LAUNCH_OK("cells_upsample2x_kernel");
  return GPEMSR_OK;
}

int gpemsr_cells_mul_mask(const float* x_f32, const gpemsr_geom_t* g, int c, const float* mask, int hm, int wm, int scale,
                          int sigmoid, int c_off, float* out_f32, void* out_hi, void* out_lo, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_f32 || !g || !mask || c <= 0 || c % 8 || c_off % 8 || scale < 1 || !(out_f32 || out_hi))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_mul_mask: bad arguments");
  if ((rc = check_geom(*g, "cells_mul_mask")) != GPEMSR_OK) return rc;
  if (g->h != hm * scale || g->w != wm * scale) return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_mul_mask: mask size x scale != tensor size");
  const long long total = (long long)g->n * (c / 8) * g->h * g->w;
  cells_mul_mask_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(x_f32, to_geom(*g), c / 8, mask, hm, wm, 1.0f / (float)scale,
                                                                            sigmoid, c_off / 8, out_f32, (__nv_bfloat16*)out_hi,
                                                                            (__nv_bfloat16*)out_lo);
  GPEMSR_LAUNCH_OK("cells_mul_mask_kernel");
  return GPEMSR_OK;
}

int gpemsr_cells_pool3x3s2(const float* x_f32, const gpemsr_geom_t* gi, int c, const gpemsr_geom_t* go, void* out_hi, void* out_lo,
                           gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_f32 || !gi || !go || c <= 0 || c % 8 || !out_hi) return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_pool3x3s2: bad arguments");
  if ((rc = check_geom(*gi, "cells_pool3x3s2(in)")) != GPEMSR_OK || (rc = check_geom(*go, "cells_pool3x3s2(out)")) != GPEMSR_OK) return rc;
  if (go->n != gi->n || go->h != (gi->h - 1) / 2 + 1 || go->w != (gi->w - 1) / 2 + 1)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_pool3x3s2: the output must be the (k3, s2, p1) grid of the input");
  const long long total = (long long)go->n * (c / 8) * go->h * go->w;
  cells_pool3x3s2_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(x_f32, to_geom(*gi), c / 8, to_geom(*go),
                                                                             (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  GPEMSR_LAUNCH_OK("cells_pool3x3s2_kernel");
  return GPEMSR_OK;
}

int gpemsr_cells_copy(const float* src_f32, const void* src_hi, const void* src_lo, const gpemsr_geom_t* gs, int src_c_off, int c,
                      int bcast_t, int center, const gpemsr_geom_t* gd, int c_off, float* dst_f32, void* dst_hi, void* dst_lo,
                      gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!gs || !gd || c <= 0 || c % 8 || c_off % 8 || src_c_off % 8 || bcast_t < 0 || (bcast_t > 0 && (center < 0 || center >= bcast_t)))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_copy: bad arguments");
  if ((rc = check_geom(*gs, "cells_copy(src)")) != GPEMSR_OK || (rc = check_geom(*gd, "cells_copy(dst)")) != GPEMSR_OK) return rc;
  if (gs->h != gd->h || gs->w != gd->w || (bcast_t == 0 && gs->n != gd->n) || (bcast_t > 0 && (gd->n % bcast_t || gs->n != gd->n)))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "cells_copy: geometries differ");
  const long long total = (long long)gd->n * (c / 8) * gd->h * gd->w;
  cells_copy_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(src_f32, (const uint4*)src_hi, (const uint4*)src_lo, to_geom(*gs),
                                                                        src_c_off / 8, c / 8, bcast_t, center, to_geom(*gd), c_off / 8,
                                                                        dst_f32, (uint4*)dst_hi, (uint4*)dst_lo);
  GPEMSR_LAUNCH_OK("cells_copy_kernel");
  return GPEMSR_OK;
}

int gpemsr_temporal_attn_scale(const float* emb_f32, const float* emb_ref_f32, const float* aligned_f32, const gpemsr_geom_t* g,
                               const gpemsr_geom_t* g_ref, int c, int t, const gpemsr_geom_t* g_out, void* out_hi, void* out_lo,
                               gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!emb_f32 || !emb_ref_f32 || !aligned_f32 || !g || !g_ref || !g_out || !out_hi || c <= 0 || c % 8 || t <= 0)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "temporal_attn_scale: bad arguments");
  if ((rc = check_geom(*g, "temporal_attn_scale(in)")) != GPEMSR_OK || (rc = check_geom(*g_ref, "temporal_attn_scale(ref)")) != GPEMSR_OK ||
      (rc = check_geom(*g_out, "temporal_attn_scale(out)")) != GPEMSR_OK) return rc;
  if (g->n % t || g_ref->n != g->n / t || g_out->n != g->n / t || g_ref->h != g->h || g_ref->w != g->w || g_out->h != g->h || g_out->w != g->w)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "temporal_attn_scale: geometries do not describe b*t frames / b windows of one size");
  const long long total = (long long)g->n * g->h * g->w;
  temporal_attn_scale_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(emb_f32, emb_ref_f32, aligned_f32, to_geom(*g),
                                                                                 to_geom(*g_ref), c / 8, t, to_geom(*g_out),
                                                                                 (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  GPEMSR_LAUNCH_OK("temporal_attn_scale_kernel");
  return GPEMSR_OK;
}

int gpemsr_threeda_combine(const float* feat, const float* attn, const float* attn_add, const float* fea_3d2, const float* fea_3d3,
                           const gpemsr_geom_t* g, int c, float* out_f32, void* out_hi, void* out_lo, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!feat || !attn || !attn_add || !fea_3d2 || !fea_3d3 || !g || c <= 0 || c % 8 || !(out_f32 || out_hi))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "threeda_combine: bad arguments");
  if ((rc = check_geom(*g, "threeda_combine")) != GPEMSR_OK) return rc;
  const long long total = (long long)g->n * (c / 8) * g->h * g->w;
  threeda_combine_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(feat, attn, attn_add, fea_3d2, fea_3d3, to_geom(*g), c / 8,
                                                                             out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  GPEMSR_LAUNCH_OK("threeda_combine_kernel");
  return GPEMSR_OK;
}

int gpemsr_conv3x3_direct(const float* x, int n, int cin, int h, int w, const float* wgt, const float* bias, int cout, int stride,
                          float* out, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x || !wgt || !out || n <= 0 || cin <= 0 || cout <= 0 || h <= 0 || w <= 0 || stride < 1)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "conv3x3_direct: bad arguments");
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;       // floor((h + 2 - 3) / stride) + 1
  const long long total = (long long)n * cout * ho * wo;
  conv3x3_direct_kernel<<<blocks_for(total), 256, 0, (cudaStream_t)stream>>>(x, n, cin, h, w, wgt, bias, cout, stride, ho, wo, out);
  GPEMSR_LAUNCH_OK("conv3x3_direct_kernel");
  return GPEMSR_OK;
}

int gpemsr_tap_gather_sum(const float* taps_f32, const gpemsr_geom_t* g, int n_out, int up, int co, const float* bias,
                          int act, float slope, const float* base, int base_h, int base_w, int base_scale, float* out_nchw, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!taps_f32 || !g || !out_nchw || n_out < 1 || n_out > 4 || (up != 1 && up != 2) || !g->padded ||
      (up == 2 && (co < 1 || n_out != 4 * co || base)) || (base && (base_h * base_scale != g->h || base_w * base_scale != g->w)))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "tap_gather_sum: n_out in 1..4 on a ringed geometry; up = 2 needs n_out == 4 * co; the base image "
                     "must be (h / scale) x (w / scale)");
  if ((rc = check_geom(*g, "tap_gather_sum")) != GPEMSR_OK) return rc;
  const long long total = (long long)g->n * g->h * g->w;
  tap_gather_sum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(taps_f32, to_geom(*g), n_out, up, co, bias, act, slope, base,
                                                                                         base_h, base_w, base_scale, out_nchw);
  GPEMSR_LAUNCH_OK("tap_gather_sum_kernel");
  return GPEMSR_OK;
}

}  // extern "C"
