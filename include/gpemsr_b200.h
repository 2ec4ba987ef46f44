/*
 * gpemsr_b200 -- C ABI of the B200 (sm_100a) kernels behind GPEMSR's inference hot path.
 *
 * The reference (jtshou/GPEMSR) is pure Python/PyTorch and has no FFI of its own; its
 * "operator interface" for this path is the nn.Module method surface listed in SURVEY.md
 * section 8(b).  Each entry point below names the reference method (file:line, relative to
 * GPEMSR-CREMI/GPEMSR/) whose device work it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (inputs, outputs and workspace);
 *     the library allocates nothing persistent;
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue asynchronously and never
 *     synchronise; they are re-entrant across streams;
 *   - return value: 0 = OK, negative = GPEMSR_ERR_*; gpemsr_last_error_string() returns a
 *     thread-local description of the last failure;
 *   - there is NO CPU fallback: on a device that is not compute capability 10.x every
 *     compute entry point returns GPEMSR_ERR_UNSUPPORTED_ARCH.
 *   - activations on this boundary use the reference's layouts (NCHW / [N,H,W,2] / [B,H,W,K],
 *     fp32, contiguous); indices are int64 like torch.argmin's.
 */
#ifndef GPEMSR_B200_H
#define GPEMSR_B200_H

#if defined(GPEMSR_BUILDING)
#define GPEMSR_API __attribute__((visibility("default")))
#else
#define GPEMSR_API
#endif

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPEMSR_OK                     0
#define GPEMSR_ERR_BAD_SHAPE         -1
#define GPEMSR_ERR_BAD_ALIGN         -2
#define GPEMSR_ERR_UNSUPPORTED_ARCH  -3
#define GPEMSR_ERR_CUDA              -4
#define GPEMSR_ERR_WORKSPACE         -5
#define GPEMSR_ERR_UNSUPPORTED       -6

#define GPEMSR_PAD_ZEROS   0
#define GPEMSR_PAD_BORDER  1
/* how `2 * v / max(size - 1, 1)` of BasicSR's flow_warp is rounded: ATen's CPU kernel divides; ATen's CUDA kernel
 * (BinaryDivTrueKernel.cu, CPU-scalar divisor) multiplies by the fp32 reciprocal 1.0f / b.  The reference runs on CUDA
 * (output_GPEMSR.py:44), so RECIP is what its own path computes; DIV reproduces a CPU run of the same code. */
#define GPEMSR_COORD_DIV     0
#define GPEMSR_COORD_RECIP   1
#define GPEMSR_COORD_DEFAULT GPEMSR_COORD_RECIP

/* activation fused into a convolution epilogue */
#define GPEMSR_ACT_NONE    0
#define GPEMSR_ACT_RELU    1
#define GPEMSR_ACT_LRELU   2   /* slope passed separately */
#define GPEMSR_ACT_EXP     3   /* exp(v - row_max[row]): the softmax numerator, see gpemsr_igemm_desc_t.row_max */

typedef void* gpemsr_stream_t;

/* ---- housekeeping ------------------------------------------------------------------- */
GPEMSR_API int         gpemsr_version(void);                  /* 0xMMmmpp */
GPEMSR_API const char* gpemsr_last_error_string(void);
GPEMSR_API int         gpemsr_device_check(int device);       /* OK iff `device` is sm_100-class */
GPEMSR_API int64_t     gpemsr_kernel_launches(void);          /* kernels launched by this library so far (process-wide) */
/* Tensor maps (TMA descriptors of the activation tiles of the tap-fused / dy-fused GEMMs) encoded so far, and shapes the driver
 * refused (those launches ran the bulk-copy producer instead).  GPEMSR_TMA=0 in the environment disables them (A/B timing). */
GPEMSR_API void        gpemsr_tensor_map_stats(int64_t* built, int64_t* rejected);

/* ---- a-5: flow_warp -------------------------------------------------------------------
 * Replaces basicsr.archs.arch_util.flow_warp (third-party; reached from model/GPEMSR.py:99-100
 * via SpyNet.process): out = grid_sample(x, normalise(meshgrid + flow), bilinear, padding_mode,
 * align_corners).  x [n,c,h,w], flow [n,h,w,2] (dx,dy in pixels), out [n,c,h,w]; fp32.
 * Coordinates are evaluated with the reference's sequence of separately rounded fp32 ops. */
GPEMSR_API int gpemsr_flow_warp(const float* x, const float* flow, int n, int c, int h, int w,
                     int padding_mode, int align_corners, float* out, gpemsr_stream_t stream);
/* same with the coordinate rounding form chosen explicitly (GPEMSR_COORD_*); gpemsr_flow_warp uses GPEMSR_COORD_DEFAULT */
GPEMSR_API int gpemsr_flow_warp_ex(const float* x, const float* flow, int n, int c, int h, int w,
                     int padding_mode, int align_corners, int coord_form, float* out, gpemsr_stream_t stream);

/* ---- a-1: Codebook.forward  (model/codebook.py:15-32) ---------------------------------
 * z [b,d,hw] (NCHW with hw = H*W), emb [k,d]  ->  zq [b,d,hw], idx [b*hw] (row order b,h,w),
 * sq_err_sum: optional device float, receives sum((zq - z)^2) so the caller can form the
 * reference loss (1 + beta) * sum / (b*d*hw) (:26).  Lowest index wins ties (:23).
 * The distance GEMM runs on tcgen05 in bf16; rows whose best/second-best gap is inside the
 * bf16 error bound are re-scored in fp32, so idx is the arg-min of fp32-accumulated distances. */
GPEMSR_API size_t gpemsr_vq_workspace_bytes(int64_t n_rows, int d, int k);
GPEMSR_API int gpemsr_vq_lookup_nchw(const float* z, const float* emb, int b, int d, int64_t hw, int k,
                          float* zq, int64_t* idx, float* sq_err_sum,
                          void* ws, size_t ws_bytes, gpemsr_stream_t stream);

/* ---- a-2: Indexer head + Codebook.inference_lr  (model/indexer.py:47,53 / 96,100 and
 *      model/codebook.py:34-43) ----------------------------------------------------------
 * feat [b,d,hw] (Indexer output_layer result, NCHW), w [k,d], bias [k] (Indexer.embedding),
 * emb [k,dq] (Codebook.embedding) -> zq [b,dq,hw], idx [b*hw] = argmax_k (feat.w_k + bias_k)
 * (= top-1 of the softmax, :38-39).  `logits` optional [b*hw,k] output (nullptr to skip). */
GPEMSR_API int gpemsr_logits_argmax_gather(const float* feat, const float* w, const float* bias,
                                const float* emb, int b, int d, int64_t hw, int k, int dq,
                                float* zq, int64_t* idx, float* logits,
                                void* ws, size_t ws_bytes, gpemsr_stream_t stream);

/* Codebook.inference_lr on materialised logits p [n_rows,k] (model/codebook.py:34-43). */
GPEMSR_API int gpemsr_argmax_gather(const float* p, const float* emb, int b, int64_t hw, int k, int dq,
                         float* zq, int64_t* idx, gpemsr_stream_t stream);

/* ---- a-3 / a-4: implicit-GEMM convolutions, GroupNorm, non-local attention --------------------------
 * Replaces the cuDNN / cuBLAS / ATen calls behind Decoder.forward / multi_scale_feat_calculate
 * (model/decoder.py:37-57), ResidualBlock / UpBlock / NonLocalBlock (model/blocks.py:8-83) and the SR tail of
 * GPEMSR.forward (model/GPEMSR.py:441-455).  gpemsr_b200/decoder.py and gpemsr_b200/sr_tail.py chain these.
 *
 * Internal activation format ("padded K8-blocked"): an image batch [n, c, h, w] is flattened to rows
 *     row(i, y, x) = m0 + i * r_img + (y + 1) * (w + 2) + (x + 1)          (one zero ring around every image)
 * and stored per group of 8 channels as [c_pad/8][rows_alloc][8]: an fp32 master and/or (hi, lo) bf16 planes with
 * hi = bf16(v), lo = bf16(v - hi).  A 3x3 tap is then a ROW SHIFT of the A operand, so a convolution is 9 shifted
 * GEMMs accumulated in TMEM -- no im2col buffer.  Ring / tail rows are never written and must stay zero (buffers
 * are zeroed once by the caller).  "Compact" tensors (attention tokens) use row(i, y, x) = i * r_img + y * w + x. */
typedef struct gpemsr_geom {
  int32_t n, h, w;         /* images, height, width */
  int32_t padded;          /* zero-ring width P: 0 = compact rows, 1 = 3x3 taps, 3 = 7x7 taps (wp = w + 2P) */
  int64_t r_img;           /* rows reserved per image (multiple of 128) */
  int64_t m0;              /* first row of image 0 (>= P * wp + P: the largest negative tap shift) */
  int64_t rows_alloc;      /* rows allocated per 8-channel plane */
} gpemsr_geom_t;

typedef struct gpemsr_igemm_desc {
  /* A operand: activations (rows) ; B operand: weights / keys (columns), both K8-blocked bf16 */
  const void* a_hi; const void* a_lo;       /* [k_pad/8][a_geom.rows_alloc][8] */
  const void* b_hi; const void* b_lo;       /* [taps][k_pad/8][b_rows][8] */
  gpemsr_geom_t a_geom;                     /* rows computed = n * r_img starting at m0 */
  int32_t k_pad;                            /* reduction length per tap (multiple of 32 for split 3, of 64 for split 1) */
  int32_t taps;                             /* 1..49 */
  int32_t tap_dy[49], tap_dx[49];           /* A row shift of tap t = dy * (w + 2) + dx */
  int32_t b_rows;                           /* allocated B rows per tap (multiple of block_n) */
  int32_t b_packed;                         /* 1: b_hi is a gpemsr_pack_weights_tiled() buffer (b_lo unused) */
  int32_t n_cols;                           /* valid output columns (channels) */
  int32_t split;                            /* 3: hi/lo planes, fp32-faithful ; 1: single bf16 pass (a_lo/b_lo unused) */
  /* epilogue: v = act(scale * acc + bias) + residual */
  float scale;
  const float* bias;                        /* [n_cols] or, if bias_per_row, indexed by output row within the image */
  int32_t bias_per_row;
  int32_t act; float slope;                 /* GPEMSR_ACT_* */
  const float* residual;                    /* fp32 master in the OUTPUT geometry / channel blocking, or NULL */
  /* output placement */
  gpemsr_geom_t o_geom;                     /* geometry of the blocked outputs */
  int32_t up;                               /* 1: same resolution ; 2: output pixel (2y + py, 2x + px) */
  int32_t py, px;
  int32_t pixel_shuffle;                    /* 1: column c*4 + dy*2 + dx -> channel c at (2y + dy, 2x + dx) (up must be 2) */
  int32_t phase_cols;                       /* > 0: parity phases merged along the columns: column p*phase_cols + c is channel c
                                               of output pixel (2y + p/2, 2x + p%2) (up must be 2, phase_cols % 32 == 0) */
  int32_t c_off;                            /* first output channel (multiple of 8) inside the output tensors */
  float* out_f32;                           /* fp32 master [co_pad/8][o_geom.rows_alloc][8] or NULL */
  void* out_hi; void* out_lo;               /* bf16 planes or NULL */
  float* out_nchw; int32_t nchw_c;          /* reference layout [n, nchw_c, up*h, up*w] or NULL */
  float* out_rowmajor; int64_t ld;          /* [rows, ld] fp32 (attention scores), rows relative to m0, or NULL */
  double* gn_sums; int32_t gn_cpg;          /* optional fused GroupNorm statistics of the stored values: [n][n_cols/gn_cpg][2]
                                               doubles (sum, sum of squares per image and group of gn_cpg channels), zeroed by
                                               the call; gn_cpg in {1,2,4,8,16,32} */
  const void* patch_other; float* patch_sums; int32_t patch_size;
                                            /* optional fused patch correlation (the VGG relu1_2 similarity mask,
                                               model/GPEMSR.py:345-353): patch_other = fp32 cells of a second tensor in the
                                               output geometry; patch_sums [n][h/ps][w/ps][3] receives, per ps x ps pixel block
                                               and over all n_cols channels, (sum v*o, sum v*v, sum o*o) of the stored values v;
                                               zeroed by the call; needs up == 1 and h, w divisible by patch_size */
  int32_t* err_flag;                        /* device int: set when the pipeline times out (never hangs) */
  /* softmax fused into the attention GEMMs (model/blocks.py:74-79; rows = queries, columns = keys; the rows of a launch are
   * indexed from 0 at a_geom.m0).  Three launches replace bmm -> scale -> softmax -> bmm without a fp32 score matrix:
   *   1. pre-pass (split 1): row_max_out[row] = max over the valid columns of v           (nothing else is stored)
   *   2. scores:  act = GPEMSR_ACT_EXP, row_max = that buffer, row_sum: v = exp(v - row_max[row]) is stored as operand planes
   *      (the unnormalised probabilities, zero beyond n_cols) and summed per row into row_sum[row]
   *   3. P v^T:   row_div = that buffer: v = acc / row_div[row]
   * row_max_out / row_sum are initialised by the call. */
  float* row_max_out;
  const float* row_max;
  float* row_sum;
  const float* row_div;
  int32_t patch_other_bf16;                 /* 1: patch_other points at bf16 cells (the hi operand plane of the second tensor: what a
                                               one-pass bf16 branch keeps of it anyway) instead of fp32 cells */
} gpemsr_igemm_desc_t;

GPEMSR_API int gpemsr_igemm(const gpemsr_igemm_desc_t* desc, gpemsr_stream_t stream);

/* NCHW fp32 <-> internal format.  pack writes whichever of f32 / hi / lo are non-NULL (channels c_off..c_off+c-1). */
GPEMSR_API int gpemsr_act_pack_nchw(const float* x, int c, const gpemsr_geom_t* g, int c_off,
                         float* f32, void* hi, void* lo, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_act_unpack_nchw(const float* f32, int c, const gpemsr_geom_t* g, int c_off, float* x,
                           gpemsr_stream_t stream);
/* ---- DCNv2 (BasicSR DCNv2Pack -> torchvision.ops.deform_conv2d; model/GPEMSR.py:79-94, 123-150) ----
 * Deformable im2col for the modulated 3x3 / stride 1 / padding 1 case with channels == 8 * deformable_groups: x_f32 = fp32
 * master cells of the input in geometry gx; offset_mask = the conv_offset output, NCHW fp32 [n, 3 * groups * 9, h, w]
 * (channels 2(g*9+k), +1 = the (y, x) offset of group g, tap k; channel 2*groups*9 + g*9 + k = its mask logit); the result
 * is the [pixels x 9 * channels] operand (cell k * groups + g) in geometry `go`, ready for a one-tap gpemsr_igemm() with the
 * weight reordered to [co][k * channels + ci].  Samples follow torchvision's border rule (zero outside (-1, H) x (-1, W)). */
GPEMSR_API int gpemsr_deform_im2col(const float* x_f32, const gpemsr_geom_t* gx, int channels, int deform_groups,
                         const float* offset_mask, void* out_hi, void* out_lo, const gpemsr_geom_t* go, gpemsr_stream_t stream);
/* ---- SpyNet helpers (basicsr SpyNet.process / forward; reached from model/GPEMSR.py:99-100) ----
 * gpemsr_resize_bilinear: F.interpolate(mode='bilinear') with ATen's source-index arithmetic (rh / rw = ATen's
 *   area_pixel_compute_scale, computed by the host in fp32), plus the elementwise work SpyNet wraps around it:
 *   out[n, co, y, x] (=|+=) mul[co] * (bilinear(x[n, co % c_in]) - sub[co]) / div[co]  (NULL = identity); rep_h / rep_w > 0
 *   repeat the last natural row / column (F.pad replicate after an upsampling).  NCHW fp32 in; NCHW and / or NHWC (the
 *   layout flow_warp takes its flow in) out.
 * gpemsr_avg_pool2: F.avg_pool2d(x, 2, 2) on `planes` images of even size.
 * gpemsr_pack_concat3: torch.cat of up to three NCHW tensors (<= 8 channels in total) written as one operand cell column. */
GPEMSR_API int gpemsr_resize_bilinear(const float* x, int n, int c_in, int h, int w, int c_out, int ho, int wo, int align_corners,
                           float rh, float rw, int rep_h, int rep_w, const float* sub, const float* div, const float* mul,
                           int accumulate, float* out /*NCHW, nullable*/, float* out_nhwc /*[n, ho, wo, c_out], nullable*/,
                           gpemsr_stream_t stream);
GPEMSR_API int gpemsr_avg_pool2(const float* x, int64_t planes, int h, int w, float* out, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_pack_concat3(const float* a, int ca, const float* b, int cb, const float* c, int cc, const gpemsr_geom_t* g,
                        void* out_hi, void* out_lo, gpemsr_stream_t stream);
/* First VGG19 layer on a one-channel image (model/VGG.py:21 slice1.0 applied to `img.expand(-1, 3, -1, -1)`,
 * model/GPEMSR.py:345,349): y = relu(conv3x3(x, w1) + bias) with w1[co][ky][kx] = sum over the 3 identical input channels
 * of the reference weight, zero padding 1.  x fp32 [n, 1, h, w] (reference layout); output = bf16 (hi, lo) operand planes
 * with `co` channels (multiple of 8, <= 64) in geometry g (padded).  CUDA cores: 9 MACs per output, HBM bound. */
GPEMSR_API int gpemsr_conv3x3_c1_relu(const float* x, const float* w1 /*[co][9]*/, const float* bias /*[co]*/, int co,
                           const gpemsr_geom_t* g, void* out_hi, void* out_lo, gpemsr_stream_t stream);
/* mask[i] = s0 / (max(sqrt(s1), eps) * max(sqrt(s2), eps)) over the n triples of a patch_sums buffer: the cosine similarity of
 * F.normalize()d patch vectors (model/GPEMSR.py:346-351). */
GPEMSR_API int gpemsr_patch_cosine(const float* patch_sums, int64_t n, float eps, float* mask, gpemsr_stream_t stream);
/* Space-to-depth of bf16 operand planes for stride-2 convolutions (DownBlock, model/blocks.py:41-47: Conv2d(k3, s2, p1)):
 * channel block (p*2 + q) of output pixel (y, x) = input pixel (2y + p, 2x + q), zero past an odd edge; c % 8 == 0; the
 * output holds 4*c channels on the ceil(h/2) x ceil(w/2) grid.  The stride-2 conv is then a 2x2-tap (offsets {-1,0}^2)
 * stride-1 gpemsr_igemm() over it. */
GPEMSR_API int gpemsr_space_to_depth(const void* in_hi, const void* in_lo, const gpemsr_geom_t* in_geom, int c,
                          void* out_hi, void* out_lo, const gpemsr_geom_t* out_geom, gpemsr_stream_t stream);
/* weights -> B operand planes.  src element (col n, k, tap t) = w[n * n_stride + k * k_stride + tap_src[t]].
 * Conv2d weight [co, ci, kh, kw]: n_stride = ci*kh*kw, k_stride = kh*kw ; ConvTranspose2d [ci, co, kh, kw]:
 * n_stride = kh*kw, k_stride = co*kh*kw ; Linear [n, k]: n_stride = k, k_stride = 1, one tap. */
GPEMSR_API int gpemsr_pack_weights(const float* w, int n, int k, int64_t n_stride, int64_t k_stride, int taps,
                        const int32_t* tap_src, int b_rows, int k_pad, void* hi, void* lo, gpemsr_stream_t stream);

/* Which kernel variant gpemsr_igemm() will run for this descriptor (only n_cols, k_pad, taps, tap_dy/dx, split and
 * pixel_shuffle are read): *block_n = column tile, *tapfused = 1 when the B-resident tap-fused kernel is used (it needs
 * the plain gpemsr_pack_weights() layout; the streaming kernel is fastest with gpemsr_pack_weights_tiled()). */
GPEMSR_API int gpemsr_igemm_plan(const gpemsr_igemm_desc_t* desc, int32_t* block_n, int32_t* tapfused);
/* weights -> ONE buffer [n_tile][tap][k-chunk][plane][k-cell][block_n][8] so that every pipeline stage of the streaming
 * kernel is a single contiguous bulk copy.  split = 3: planes (hi, lo), k-chunk = 32 ; split = 1: hi only, k-chunk = 64. */
GPEMSR_API size_t gpemsr_pack_weights_tiled_bytes(int n, int k_pad, int taps, int block_n, int split);
GPEMSR_API int gpemsr_pack_weights_tiled(const float* w, int n, int k, int64_t n_stride, int64_t k_stride, int taps,
                              const int32_t* tap_src, int block_n, int k_pad, int split, void* out,
                              gpemsr_stream_t stream);

/* GroupNorm(32 groups, eps) over an fp32 master (model/blocks.py:5-6): stats -> per-(image, channel) scale/shift,
 * then y = act(x * scale + shift) (+ residual) written as fp32 master and/or hi/lo planes, optionally re-rowed
 * into a compact geometry (attention tokens).  chan_sums: [n, c, 2] doubles (zeroed by gn_stats). */
GPEMSR_API int gpemsr_gn_stats(const float* x_f32, int c, const gpemsr_geom_t* g, double* chan_sums,
                    gpemsr_stream_t stream);
GPEMSR_API int gpemsr_gn_scale_shift(const double* sums, int sums_per_group /* 0: [n,c,2] per channel, 1: [n,groups,2] */,
                          const float* gamma, const float* beta, int n, int c, int groups, double count_per_channel,
                          float eps, float* scale_shift /* [n, c, 2] */, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_affine_act(const float* x_f32, int c, const gpemsr_geom_t* g, const float* scale_shift,
                      int act, float slope, const float* residual, const gpemsr_geom_t* og,
                      float* out_f32, void* out_hi, void* out_lo, float* out_nchw /* [n,c,h,w] or NULL */,
                      gpemsr_stream_t stream);

/* softmax over the keys of attention scores (model/blocks.py:76) stored as K8-blocked fp32 cells [t_pad/8][rows_alloc][8] (the
 * igemm out_f32 format with keys as "channels"): coalesced without a transpose; probabilities come out as K8-blocked bf16 A
 * operand planes [t_pad/8][t_pad][8] (hi, lo).  scratch: 2*t*(1+16) floats. */
GPEMSR_API int gpemsr_softmax_cells_blocked(const float* s_cells, int64_t t, int64_t rows_alloc, int64_t t_pad, float* scratch,
                                 void* p_hi, void* p_lo, gpemsr_stream_t stream);

/* Border ring of a 2x-upsampling, 4-phase, 3x3-tap linear map whose weights depend on the output position class (top /
 * interior / bottom) x (left / interior / right): wc [9][4][cout][cin][3][3], bias [9][cout].  Used by the decoder's final
 * stage, where ConvTranspose2d (model/blocks.py:35) and the output conv (model/decoder.py:33) are composed into one
 * GEMM and only the ring, where the conv's zero padding cuts taps off, needs these exact per-class weights. */
GPEMSR_API int gpemsr_border_phase_conv(const float* x_f32, int cin, const gpemsr_geom_t* g, const float* wc,
                             const float* bias, int cout, float* out_nchw, gpemsr_stream_t stream);

/* out[i, 0, y, x] += bilinear_upsample(x_center, scale, align_corners=False)  (model/GPEMSR.py:452-455) */
GPEMSR_API int gpemsr_add_bilinear_base(const float* x_center, int n, int h, int w, int scale, float* out,
                             gpemsr_stream_t stream);

/* ---- glue of GPEMSR.forward between the convolutions (model/GPEMSR.py:64-234 POD / ThreeDA, :323-440 reference-feature
 *      fusion and alignment) -- element-wise / resampling kernels on the padded K8-blocked format.  `c`, `c_off` are multiples
 *      of 8; outputs go to channel slot c_off of the destination buffers (any of f32 / hi / lo may be NULL), so the operands
 *      of the reference's torch.cat(...) -> conv pairs are assembled in place. ----
 * gpemsr_cells_upsample2x: F.interpolate(x, scale_factor=2, 'bilinear', align_corners=False) * mul  (:130,132,136,142,144,148,226,231).
 * gpemsr_cells_mul_mask:   x * F.interpolate(sigmoid?(mask), scale_factor=scale, 'bilinear', align_corners=False); mask is NCHW
 *                          fp32 [n, 1, hm, wm] (:354-362); sigmoid != 0 applies torch.sigmoid to the mask values first (:357).
 * gpemsr_cells_pool3x3s2:  cat([MaxPool2d(3,2,1)(x), AvgPool2d(3,2,1)(x)], 1) -> 2c channels (hi, lo planes) (:217-218, 222-223).
 * gpemsr_cells_copy:       channel-slot copy between buffers of one image size; bcast_t > 0: destination image i reads source
 *                          image (i / bcast_t) * bcast_t + center (the centre frame's features next to every neighbour's, :421-431).
 * gpemsr_temporal_attn_scale: ThreeDA temporal attention (:181-196): out[b, i*c + ch] = aligned[b*t + i, ch] *
 *                          sigmoid(sum_ch emb[b*t + i, ch] * emb_ref[b, ch]); emb / aligned: n = b*t images, emb_ref / out: n = b.
 * gpemsr_threeda_combine:  feat * sigmoid(attn) * 2 + attn_add + fea_3d2 + fea_3d3 (:232-233).
 * gpemsr_conv3x3_direct:   Conv2d(cin, cout, 3, stride, padding 1) on CUDA cores, NCHW fp32 in / out [n, cout, (h-1)/stride+1,
 *                          (w-1)/stride+1] -- POD.flowdsconv* (:71-76, 101-106), a few hundred MACs per output. */
GPEMSR_API int gpemsr_cells_upsample2x(const float* x_f32, const gpemsr_geom_t* gi, int c, float mul, const gpemsr_geom_t* go,
                            int c_off, float* out_f32, void* out_hi, void* out_lo, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_cells_mul_mask(const float* x_f32, const gpemsr_geom_t* g, int c, const float* mask, int hm, int wm, int scale,
                          int sigmoid, int c_off, float* out_f32, void* out_hi, void* out_lo, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_cells_pool3x3s2(const float* x_f32, const gpemsr_geom_t* gi, int c, const gpemsr_geom_t* go, void* out_hi,
                           void* out_lo, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_cells_copy(const float* src_f32, const void* src_hi, const void* src_lo, const gpemsr_geom_t* gs, int src_c_off,
                      int c, int bcast_t, int center, const gpemsr_geom_t* gd, int c_off, float* dst_f32, void* dst_hi,
                      void* dst_lo, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_temporal_attn_scale(const float* emb_f32, const float* emb_ref_f32, const float* aligned_f32,
                               const gpemsr_geom_t* g, const gpemsr_geom_t* g_ref, int c, int t, const gpemsr_geom_t* g_out,
                               void* out_hi, void* out_lo, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_threeda_combine(const float* feat, const float* attn, const float* attn_add, const float* fea_3d2,
                           const float* fea_3d3, const gpemsr_geom_t* g, int c, float* out_f32, void* out_hi, void* out_lo,
                           gpemsr_stream_t stream);
GPEMSR_API int gpemsr_conv3x3_direct(const float* x, int n, int cin, int h, int w, const float* wgt, const float* bias, int cout,
                          int stride, float* out, gpemsr_stream_t stream);

/* 3x3 convolutions with <= 4 output columns (model/GPEMSR.py:450 conv_last, :356 refmaskconv3, the composed last stage of
 * model/decoder.py:31,33) as ONE 1x1 GEMM whose columns are the nine taps (gpemsr_igemm writing fp32 cells, column
 * tap * n_out + o, tap = ky * 3 + kx) followed by this nine-point shifted sum on CUDA cores:
 *   out[o](y, x) = act(bias[o] + sum_tap taps[(y + ky - 1, x + kx - 1), tap * n_out + o])  (+ bilinear x`base_scale` of `base` [n, 1, h/s, w/s],
 *   align_corners=False: the base image of :452-455).  up == 2: column o = phase * co + ch is channel ch of output pixel
 *   (2y + phase / 2, 2x + phase % 2), out_nchw [n, co, 2h, 2w]; up == 1: out_nchw [n, n_out, h, w]. */
GPEMSR_API int gpemsr_tap_gather_sum(const float* taps_f32, const gpemsr_geom_t* g, int n_out, int up, int co, const float* bias,
                          int act, float slope, const float* base, int base_h, int base_w, int base_scale, float* out_nchw, gpemsr_stream_t stream);


/* ---- self-test of the tcgen05 GEMM core (used by tests/, not by the product path) ------
 * D[m,n] = A[m,k] * B[n,k]^T, fp32 row-major; split = 1 (single bf16 pass) or 3 (hi/lo bf16,
 * fp32-faithful); block_n in {64,128,256}.  _status() synchronises and reports pipeline time-outs. */
GPEMSR_API size_t gpemsr_selftest_gemm_workspace_bytes(int64_t m, int n, int k);
GPEMSR_API int gpemsr_selftest_gemm(const float* a, const float* b, int64_t m, int n, int k, int split,
                         int block_n, float* d, void* ws, size_t ws_bytes, gpemsr_stream_t stream);
GPEMSR_API int gpemsr_selftest_gemm_status(const void* ws, int64_t m, int n, int k, gpemsr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GPEMSR_B200_H */
