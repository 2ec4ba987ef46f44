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

/* activation fused into a convolution epilogue */
#define GPEMSR_ACT_NONE    0
#define GPEMSR_ACT_RELU    1
#define GPEMSR_ACT_LRELU   2   /* slope passed separately */

typedef void* gpemsr_stream_t;

/* ---- housekeeping ------------------------------------------------------------------- */
GPEMSR_API int         gpemsr_version(void);                  /* 0xMMmmpp */
GPEMSR_API const char* gpemsr_last_error_string(void);
GPEMSR_API int         gpemsr_device_check(int device);       /* OK iff `device` is sm_100-class */
GPEMSR_API int64_t     gpemsr_kernel_launches(void);          /* kernels launched by this library so far (process-wide) */

/* ---- a-5: flow_warp -------------------------------------------------------------------
 * Replaces basicsr.archs.arch_util.flow_warp (third-party; reached from model/GPEMSR.py:99-100
 * via SpyNet.process): out = grid_sample(x, normalise(meshgrid + flow), bilinear, padding_mode,
 * align_corners).  x [n,c,h,w], flow [n,h,w,2] (dx,dy in pixels), out [n,c,h,w]; fp32.
 * Coordinates are evaluated with the reference's sequence of separately rounded fp32 ops. */
GPEMSR_API int gpemsr_flow_warp(const float* x, const float* flow, int n, int c, int h, int w,
                     int padding_mode, int align_corners, float* out, gpemsr_stream_t stream);

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
