// a-3 / a-4: implicit-GEMM convolution family on the tcgen05 GEMM core.
//
// Replaces the cuDNN / cuBLAS calls of Decoder / ResidualBlock / UpBlock / NonLocalBlock (model/decoder.py:37-57,
// model/blocks.py:8-83) and of the SR tail (model/GPEMSR.py:441-455).  One entry point, gpemsr_igemm(), covers
//   * Conv2d 3x3 / 1x1, stride 1 (9 or 1 row-shifted GEMMs over the zero-ringed, flattened activation),
//   * ConvTranspose2d(k3, s2, p1, op1) as four output-parity phases with 1/2/2/4 taps (no zero-stuffed MACs),
//   * the attention products q^T k and P v^T and the Linear head (plain GEMMs on compact rows),
// with the epilogue fused: scale, bias (per column or per row), ReLU / LeakyReLU, residual add, PixelShuffle(2),
// parity-phase scatter, and stores as fp32 master, (hi, lo) bf16 operand planes for the next layer, NCHW or row-major.
#include "capi_common.h"
#include "gemm_core.cuh"
#include "act_layout.cuh"
#include <algorithm>
#include <cstdlib>

namespace {

// PAIR: the accumulator is split over the column ranges (c, c + BLOCK_N) -- the paired-N MMAs of the tap-fused / dy-fused kernels in
// the fp32-faithful split keep a_hi * w_lo apart from a_hi * w_hi + a_lo * w_hi; the epilogue sums them.  A compile-time property
// of the launch: the kernels that never pair (the streaming kernel, every single-pass launch) carry no second TMEM load.
// LEAN: a compile-time promise made by the host dispatch (lean_mode()) about which special epilogues can be active.  The generic
// epilogue keeps GroupNorm sums, patch correlation, softmax row statistics, exp, PixelShuffle, phase scatter, NCHW / row-major
// stores, per-row bias and partial column tiles behind run-time branches: ~600 issued instructions per thread and tile, and the
// epilogue warps of the narrow kernels are ISSUE bound (profiles/r02_tapfuse_roles.txt).  A specialised instantiation contains the
// plain scale / bias / activation / residual / store path plus only the features its mask names.
template <int BLOCK_N, bool PAIR = false, int LEAN = 0>      // LEAN: 0 = generic; else bit 0 set + the features that stay enabled:
struct EpiConv {                                             //   2 patch correlation (VGG mask), 4 partial column tiles / row-major
  static constexpr bool GEN = LEAN == 0;                     //   stores (tap GEMMs), 8 parity-phase scatter (merged ConvTranspose2d),
  static constexpr bool PATCH = GEN || (LEAN & 2);           //   16 fused softmax (row max / exp / row sum / row division)
  static constexpr bool ROWM = GEN || (LEAN & 4), FULLCOLS = !GEN && !(LEAN & 4);
  static constexpr bool PHASE = GEN || (LEAN & 8);
  static constexpr bool SOFTMAX = GEN || (LEAN & 16);
  Geom ag, og;
  int n_cols;
  float scale;
  const float* bias;
  int bias_per_row, act;
  float slope;
  const float* residual;
  int up, py, px, pixel_shuffle, phase_cols, c_off;
  float* out_f32;
  __nv_bfloat16 *out_hi, *out_lo;
  float* out_nchw;
  int nchw_c;
  float* out_rowmajor;
  long long ld;
  double* gn_sums;
  int gn_cpg;
  static constexpr int pair_off = PAIR ? BLOCK_N : 0;      // column distance of the two partial accumulators (see PAIR above)
  const float* patch_other;   // fp32 cells in the output geometry: the tensor the stored values are correlated with
  int patch_other_bf16;       // 1: patch_other points at bf16 cells (the hi plane of that tensor) instead
  float* patch_sums;          // [n][h / patch][w / patch][3] = (sum v*o, sum v*v, sum o*o) per patch x patch block of pixels
  int patch_size;
  // softmax fused into the attention GEMMs (model/blocks.py:74-79): rows = queries, columns = keys
  float* row_max_out;         // pre-pass: only the per-row maximum of v over the valid columns is produced (atomic max)
  const float* row_max;       // GPEMSR_ACT_EXP: v = exp(v - row_max[row])
  float* row_sum;             // per-row sum of the stored v over the columns (atomicAdd once per CTA column sweep)
  const float* row_div;       // v = acc / row_div[row]  (the deferred softmax normalisation of P v^T)

  // epilogue warps: (WARPS / 4) warps share a TMEM lane quarter and split the tile's columns.  The narrow tiles are EPILOGUE bound
  // (ncu source view, 64 -> 64 conv @ 5 x 640^2 with 8 warps: the epilogue warps are busy 92 % of the time at 10 cycles per
  // instruction, half of it long-scoreboard stalls on bias / residual / constant loads that two warps per scheduler cannot hide,
  // tile period 7900 cycles against 5000 of MMA time), so N = 64 runs 16 epilogue warps and N = 32 runs 8: -7 % / -14 % on the
  // split-3 / single-pass 64-channel convs.  (Measured and dropped: 16 warps on the N >= 128 tiles -- the 112-register cap spills and the
  // wide kernels get 2-4 % slower.)
  static constexpr int WARPS = BLOCK_N == 64 ? 16 : BLOCK_N >= 128 ? 8 : BLOCK_N == 32 ? 8 : 4;
  // In the tap-fused / dy-fused kernels (one column tile per row tile) the epilogue warps form GROUPS groups that own one TMEM
  // accumulator buffer each and take alternate tiles: TWO tile epilogues are in flight.  An epilogue alone needs ~3 400 cycles per
  // tile whatever the warp count (a latency chain: barrier wake-up, bias loads, TMEM loads, stores, fence, arrive), more than the
  // MMAs of a single-plane tile; tools/ablate.py: the VGG layer ran 0.417 ms with and 0.277 ms without its epilogue.
  static constexpr int GROUPS = (BLOCK_N == 64 || BLOCK_N == 32) ? 2 : 1;
  struct State {
    bool init = false, valid = false;
    int img = 0, y = 0, x = 0;
    float row_bias = 0.f;
    long long orow = 0;        // output row of (up*y + py, up*x + px) in the blocked outputs
    long long nchw0 = 0;       // offset of (img, channel 0, Y, X) in out_nchw
    float p_ab = 0.f, p_aa = 0.f, p_bb = 0.f;      // patch-correlation partial sums of this row (patch_sums)
    float r_max = -3.0e38f, r_sum = 0.f, r_sub = 0.f, r_inv = 1.f;     // fused softmax: running max / sum, subtrahend, 1 / divisor
  };

  // one cell = 8 consecutive output channels of one output pixel; (dy, dx) only differ from 0 under PixelShuffle
  __device__ __forceinline__ void store_cell(const State& st, float (&v)[8], int ch0, int dy, int dx) const {
#if GPEMSR_ABLATE & 16
    if (v[0] != 12345.678f) return;          // profiling build: the epilogue computes but does not store
#endif
    if (out_f32 || out_hi || residual) {
      const long long orow = st.orow + (long long)dy * (og.w + 2 * og.padded) + dx;
      const size_t cell = ((size_t)(ch0 >> 3) * og.rows_alloc + orow) * 8;
      if (residual) {
        const float4 r0 = *reinterpret_cast<const float4*>(residual + cell), r1 = *reinterpret_cast<const float4*>(residual + cell + 4);
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      }
      if (out_f32) {
        *reinterpret_cast<float4*>(out_f32 + cell) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(out_f32 + cell + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (out_hi) {
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4*>(out_hi + cell) = hi;
        if (out_lo) *reinterpret_cast<uint4*>(out_lo + cell) = lo;
      }
    }
    if (GEN && out_nchw) {
      const long long Wo = (long long)up * ag.w, plane = (long long)up * ag.h * Wo;
      float* p = out_nchw + st.nchw0 + (long long)ch0 * plane + (long long)dy * Wo + dx;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (ch0 + j < nchw_c) p[(long long)j * plane] = v[j];
    }
  }

  // fused GroupNorm statistics: per group of CPG channels, the sum and sum of squares of this warp's 32 rows
  template <int CHUNK, int CPG>
  __device__ __forceinline__ void group_stats_t(const State& st, const float (&f)[CHUNK], int col0) const {
    const int lane = threadIdx.x & 31;
    constexpr int G = CPG < CHUNK ? CPG : CHUNK;           // a 32-channel group spans the whole 16-column chunk of N = 16 tiles
#pragma unroll
    for (int g0 = 0; g0 < CHUNK; g0 += G) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int j = 0; j < G; ++j) { s += f[g0 + j]; q = fmaf(f[g0 + j], f[g0 + j], q); }
#pragma unroll
      for (int o = 16; o; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      const int col = col0 + g0;
      if (lane == 0 && col < n_cols) {
        double* dst = gn_sums + ((size_t)st.img * (n_cols / gn_cpg) + col / gn_cpg) * 2;
        atomicAdd(dst, (double)s);
        atomicAdd(dst + 1, (double)q);
      }
    }
  }
  template <int CHUNK>
  __device__ __forceinline__ void group_stats(const State& st, const float (&f)[CHUNK], int col0) const {
    switch (gn_cpg) {
      case 1: group_stats_t<CHUNK, 1>(st, f, col0); break;
      case 2: group_stats_t<CHUNK, 2>(st, f, col0); break;
      case 4: group_stats_t<CHUNK, 4>(st, f, col0); break;
      case 8: group_stats_t<CHUNK, 8>(st, f, col0); break;
      case 16: group_stats_t<CHUNK, 16>(st, f, col0); break;
      default: group_stats_t<CHUNK, 32>(st, f, col0); break;
    }
  }

  // the bias of a chunk's columns: loaded BEFORE the chunk's TMEM loads are issued, so the two latencies overlap (the epilogue
  // warps stall on exactly these loads: ncu source view)
  template <int CHUNK>
  __device__ __forceinline__ void load_bias(const State& st, float (&bv)[CHUNK], int col0) const {
    if (GEN && bias_per_row) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) bv[j] = st.row_bias;
    } else if (bias && col0 + CHUNK <= n_cols) {
#pragma unroll
      for (int j = 0; j < CHUNK; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
        bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) bv[j] = (bias && col0 + j < n_cols) ? __ldg(bias + col0 + j) : 0.f;
    }
  }

  template <int CHUNK>
  __device__ __forceinline__ void chunk(State& st, const uint32_t (&r)[CHUNK], const float (&bv)[CHUNK], const uint4 (&po)[CHUNK / 8],
                                        int col0, long long rel) const {
    float f[CHUNK];
    const bool full = FULLCOLS || col0 + CHUNK <= n_cols;
    // v = scale * acc + bias
#pragma unroll
    for (int j = 0; j < CHUNK; ++j) f[j] = fmaf(scale, __uint_as_float(r[j]), bv[j]);
    if (SOFTMAX && row_max_out) {      // softmax pre-pass: nothing is stored
      float m = st.r_max;
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) if (full || col0 + j < n_cols) m = fmaxf(m, f[j]);
      st.r_max = m;
      return;
    }
    if (SOFTMAX && row_div) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) f[j] *= st.r_inv;
    }
    if (act == GPEMSR_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) f[j] = fmaxf(f[j], 0.f);
    } else if (act == GPEMSR_ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * slope;
    } else if (SOFTMAX && act == GPEMSR_ACT_EXP) {
      // exp(v - max) = 2^(v * log2(e) - max * log2(e)): one FFMA + MUFU.EX2 per element (relative error ~2^-22; the arguments are
      // <= ~0, so no overflow; expf() costs ~4x the instructions and this epilogue is what bounds the scores GEMM)
      const float sub2 = st.r_sub * 1.4426950408889634f;
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) f[j] = exp2f(fmaf(f[j], 1.4426950408889634f, -sub2));
    }
    if (!full) {
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) if (col0 + j >= n_cols) f[j] = 0.f;
    }
    if (SOFTMAX && row_sum) {
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < CHUNK; ++j) a += f[j];
      st.r_sum += a;
    }
    if (GEN && gn_sums) {              // all 32 lanes take part; rows outside the image contribute zeros
      if (!st.valid) {
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) f[j] = 0.f;
      }
      group_stats<CHUNK>(st, f, col0);
    }
    if (!st.valid) return;
    if (PATCH && patch_sums) {
#pragma unroll
      for (int g = 0; g < CHUNK / 8; ++g) {
        if (col0 + 8 * g >= n_cols) break;
        const size_t cell = ((size_t)((c_off + col0 + 8 * g) >> 3) * og.rows_alloc + st.orow) * 8;
        float o[8];
        if (patch_other_bf16) {
          const uint4 raw = po[g];               // loaded ahead of the TMEM loads (tile()): this epilogue stalled on exactly this load
          const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) { o[2 * j] = __uint_as_float(w4[j] << 16); o[2 * j + 1] = __uint_as_float(w4[j] & 0xffff0000u); }
        } else {
          const float4 o0 = __ldg(reinterpret_cast<const float4*>(patch_other + cell));
          const float4 o1 = __ldg(reinterpret_cast<const float4*>(patch_other + cell + 4));
          o[0] = o0.x; o[1] = o0.y; o[2] = o0.z; o[3] = o0.w; o[4] = o1.x; o[5] = o1.y; o[6] = o1.z; o[7] = o1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = f[8 * g + j];
          st.p_ab = fmaf(a, o[j], st.p_ab); st.p_aa = fmaf(a, a, st.p_aa); st.p_bb = fmaf(o[j], o[j], st.p_bb);
        }
      }
    }
    if (ROWM && out_rowmajor) {
#pragma unroll
      for (int j = 0; j < CHUNK; j += 4)
        *reinterpret_cast<float4*>(out_rowmajor + rel * ld + col0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
    if (!(out_f32 || out_hi || (GEN && out_nchw))) return;
    if (GEN && pixel_shuffle) {
      if constexpr (CHUNK == 32) {     // 32 columns = 8 channels x (dy, dx)
#pragma unroll
        for (int sub = 0; sub < 4; ++sub) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = f[4 * j + sub];
          store_cell(st, v, c_off + (col0 >> 2), sub >> 1, sub & 1);
        }
      }
    } else if (GEN && phase_cols && (phase_cols & 31)) {
      // few channels per phase (the composed up-block + output conv: phase_cols = image channels): scalar NCHW scatter
      if (out_nchw) {
        const long long Wo = (long long)up * ag.w, plane = (long long)up * ag.h * Wo;
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) {
          const int col = col0 + j;
          if (col < n_cols) {
            const int ph = col / phase_cols, ch = col - ph * phase_cols;
            out_nchw[st.nchw0 + (long long)ch * plane + (long long)(ph >> 1) * Wo + (ph & 1)] = f[j];
          }
        }
      }
    } else if (PHASE && phase_cols) {  // the four parity phases of a ConvTranspose2d side by side along the columns
      const int ph = col0 / phase_cols, ch = col0 - ph * phase_cols;
#pragma unroll
      for (int g = 0; g < CHUNK / 8; ++g) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = f[8 * g + j];
        store_cell(st, v, c_off + ch + 8 * g, ph >> 1, ph & 1);
      }
    } else {
#pragma unroll
      for (int g = 0; g < CHUNK / 8; ++g) {
        if (!FULLCOLS && col0 + 8 * g >= n_cols) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = f[8 * g + j];
        store_cell(st, v, c_off + col0 + 8 * g, 0, 0);
      }
    }
  }

  __device__ __forceinline__ void tile(State& st, uint32_t tmem_acc, long long m_tile, int n_tile, int n_tiles, int row, int part) const {
    tile_span<BLOCK_N / (WARPS / 4)>(st, tmem_acc, m_tile, n_tile, n_tiles, row, part);
  }
  // the same with the columns split over the warps of ONE group (WARPS / GROUPS warps per tile)
  __device__ __forceinline__ void tile_group(State& st, uint32_t tmem_acc, long long m_tile, int n_tile, int n_tiles, int row, int part) const {
    tile_span<BLOCK_N / (WARPS / GROUPS / 4)>(st, tmem_acc, m_tile, n_tile, n_tiles, row, part);
  }
  template <int SPAN>                                    // SPAN = columns this warp covers
  __device__ __forceinline__ void tile_span(State& st, uint32_t tmem_acc, long long m_tile, int n_tile, int n_tiles, int row, int part) const {
    const long long rel = m_tile * gemm::BLOCK_M + row;
    if (!st.init) {
      st.init = true;
      st.valid = decode_row(ag, rel, st.img, st.y, st.x);
      if (st.valid) {
        if (GEN && bias_per_row) st.row_bias = __ldg(bias + (long long)st.y * ag.w + st.x);
        const int Y = up * st.y + py, X = up * st.x + px;
        st.orow = place_row(og, st.img, Y, X);
        const long long Wo = (long long)up * ag.w, Ho = (long long)up * ag.h;
        st.nchw0 = ((long long)st.img * nchw_c * Ho + Y) * Wo + X;
        if (SOFTMAX && row_max) st.r_sub = __ldg(row_max + rel);
        if (SOFTMAX && row_div) st.r_inv = 1.0f / __ldg(row_div + rel);
      }
    }
    constexpr int CHUNK = (SPAN >= 32 && BLOCK_N >= 128) ? 32 : 16;     // the narrow kernels run 18 warps: 16 columns at a time fit their registers
    const bool row_stats = SOFTMAX && (row_max_out || row_sum) && n_tile + (int)gridDim.y >= n_tiles;      // (evaluated before the loads below)
#pragma unroll 1
    for (int c0 = part * SPAN; c0 < (part + 1) * SPAN; c0 += CHUNK) {
      const int col0 = n_tile * BLOCK_N + c0;
      const bool work = (st.valid || (GEN && gn_sums)) && (FULLCOLS || col0 < n_cols);
      float bv[CHUNK];
      uint4 po[CHUNK / 8];
      if (work) {
        load_bias<CHUNK>(st, bv, col0);
        if (PATCH && patch_sums && patch_other_bf16 && st.valid) {
#pragma unroll
          for (int g = 0; g < CHUNK / 8; ++g) {
            const size_t cell = ((size_t)((c_off + col0 + 8 * g) >> 3) * og.rows_alloc + st.orow) * 8;
            po[g] = col0 + 8 * g < n_cols ? __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(patch_other) + cell))
                                          : make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      uint32_t r[CHUNK];
      if constexpr (CHUNK == 32) sm100::tmem_ld_32x32(tmem_acc + c0, r);
      else sm100::tmem_ld_32x16(tmem_acc + c0, r);
      if constexpr (PAIR) {
        uint32_t r2[CHUNK];
        if constexpr (CHUNK == 32) sm100::tmem_ld_32x32(tmem_acc + pair_off + c0, r2);
        else sm100::tmem_ld_32x16(tmem_acc + pair_off + c0, r2);
        sm100::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
      }
      sm100::tmem_ld_wait();
      if (work) chunk<CHUNK>(st, r, bv, po, col0, rel);
    }
    if (PATCH && patch_sums) flush_patch(st);
    // fused softmax statistics: one atomic per row once this CTA has swept its last column tile of the row tile
    if (row_stats && st.valid) {
      if (row_max_out) {                 // float max through the ordered-integer trick (the buffer starts at -1.7e38)
        if (st.r_max >= 0.f) atomicMax(reinterpret_cast<int*>(row_max_out + rel), __float_as_int(st.r_max));
        else atomicMin(reinterpret_cast<unsigned*>(row_max_out + rel), __float_as_uint(st.r_max));
      } else {
        atomicAdd(row_sum + rel, st.r_sum);
      }
    }
  }

  // The 32 lanes of a warp are 32 consecutive flat rows: pixels of the same patch row form contiguous runs.  A segmented
  // shuffle reduction leaves each run's total in its first lane, which adds it to the patch's three accumulators.
  __device__ __forceinline__ void flush_patch(State& st) const {
    const int lane = threadIdx.x & 31;
    const int ppr = ag.w / patch_size;
    const int run = st.valid ? (st.img * ag.h + st.y) * ppr + st.x / patch_size : -1;     // unique per (image row, patch column)
    float ab = st.p_ab, aa = st.p_aa, bb = st.p_bb;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int ro = __shfl_down_sync(0xffffffffu, run, off);
      const float t0 = __shfl_down_sync(0xffffffffu, ab, off), t1 = __shfl_down_sync(0xffffffffu, aa, off),
                  t2 = __shfl_down_sync(0xffffffffu, bb, off);
      if (lane + off < 32 && ro == run) { ab += t0; aa += t1; bb += t2; }
    }
    const int rprev = __shfl_up_sync(0xffffffffu, run, 1);
    if (run >= 0 && (lane == 0 || rprev != run)) {
      float* dst = patch_sums + ((size_t)(st.img * (ag.h / patch_size) + st.y / patch_size) * ppr + st.x / patch_size) * 3;
      atomicAdd(dst, ab); atomicAdd(dst + 1, aa); atomicAdd(dst + 2, bb);
    }
    st.p_ab = st.p_aa = st.p_bb = 0.f;
  }
};

int lean_mode(const gpemsr_igemm_desc_t& d, int block_n);

template <int BLOCK_N, int BLOCK_K, int SPLIT, int NSTAGE, int LEAN = 0>
int launch(const gemm::Operands& op, const gpemsr_igemm_desc_t& d, cudaStream_t s) {
  if constexpr (LEAN == 0 && BLOCK_N >= 64) {
    const int mode = lean_mode(d, BLOCK_N);
    if (mode == 1) return launch<BLOCK_N, BLOCK_K, SPLIT, NSTAGE, 1>(op, d, s);
    if (mode == 9) return launch<BLOCK_N, BLOCK_K, SPLIT, NSTAGE, 9>(op, d, s);
    if constexpr (BLOCK_N >= 128) {
      if (mode == 17) return launch<BLOCK_N, BLOCK_K, SPLIT, NSTAGE, 17>(op, d, s);
    }
  }
  using Cfg = gemm::Config<BLOCK_N, BLOCK_K, SPLIT, NSTAGE>;
  using Epi = EpiConv<BLOCK_N, false, LEAN>;
  Epi e;
  e.ag = to_geom(d.a_geom); e.og = to_geom(d.o_geom);
  e.n_cols = d.n_cols; e.scale = d.scale; e.bias = d.bias; e.bias_per_row = d.bias_per_row; e.act = d.act; e.slope = d.slope;
  e.residual = d.residual; e.up = d.up; e.py = d.py; e.px = d.px; e.pixel_shuffle = d.pixel_shuffle; e.phase_cols = d.phase_cols; e.c_off = d.c_off;
  e.out_f32 = d.out_f32; e.out_hi = (__nv_bfloat16*)d.out_hi; e.out_lo = (__nv_bfloat16*)d.out_lo;
  e.out_nchw = d.out_nchw; e.nchw_c = d.nchw_c; e.out_rowmajor = d.out_rowmajor; e.ld = d.ld;
  e.gn_sums = d.gn_sums; e.gn_cpg = d.gn_cpg;
  e.patch_other = (const float*)d.patch_other; e.patch_other_bf16 = d.patch_other_bf16; e.patch_sums = d.patch_sums; e.patch_size = d.patch_size;
  e.row_max_out = d.row_max_out; e.row_max = d.row_max; e.row_sum = d.row_sum; e.row_div = d.row_div;
  const int sms = gpemsr::num_sms();
  // CTA pairs: the cta_group::2 kernel for every tile width >= 64 whose operands can be described by tensor maps; else (N >= 128)
  // the cta_group::1 kernel with the B stages multicast
  const bool pair_ok = BLOCK_N >= 64 && d.b_packed <= 1 && op.batch_tiles == 0 && gpemsr::use_pair_mma() && gpemsr::use_tensor_maps();
  const bool clustered = (BLOCK_N >= 128 || pair_ok) && op.m_tiles >= 2 && gpemsr::use_clusters();
  // grid: gx persistent row-tile walkers x gy column-tile splitters.  Model: one CTA per SM, CTAs run in waves, a CTA's
  // time = its (row tile, column tile) count.  Take the cheapest shape (e.g. 50 row tiles x 25 column tiles: 50 x 5 CTAs
  // = 2 waves x 5 units, not 50 x 3 = 150 CTAs whose 2 stragglers double the time); ties go to fewer column splits
  // (every split re-reads the A tiles).
  long long gx = std::min<long long>(op.m_tiles, sms), gy = 1;
  {
    // fixed cost of a CTA (launch, TMEM allocation, pipeline fill, un-overlapped last epilogue) ~ 3 us, in tile units
    const double tile_us = (double)op.taps * (op.k / 16) * SPLIT * (BLOCK_N / 2) / 1900.0;
    const double overhead = 3.0 / std::max(tile_us, 0.05);
    double best = -1.0;
    const long long gx_max = std::min<long long>(op.m_tiles, sms);
    for (long long y = 1; y <= op.n_tiles; ++y)
      for (long long x = clustered ? 2 : 1; x <= (clustered ? (gx_max + 1) / 2 * 2 : gx_max); x += clustered ? 2 : 1) {
        const long long waves = (x * y + sms - 1) / sms;
        const long long per = ((op.m_tiles + x - 1) / x) * ((op.n_tiles + y - 1) / y);
        const double cost = (double)waves * ((double)per + overhead);
        if (best < 0 || cost < best - 1e-9 || (cost < best + 1e-9 && (y < gy || (y == gy && x > gx)))) { best = cost; gx = x; gy = y; }
      }
  }
  if constexpr (BLOCK_N >= 64) {
    // CTA-pair MMAs (cta_group::2), operands by tensor map: packed weights or a plain K8-blocked B (attention), no weight batches
    if (clustered && pair_ok) {
      constexpr int PLANES = SPLIT == 3 ? 2 : 1, KCH = BLOCK_K / 8;
      constexpr int STAGE = PLANES * (gemm::BLOCK_M + BLOCK_N / 2) * BLOCK_K * 2;
      constexpr int NST = (227 * 1024 - 1024) / STAGE >= 8 ? 8 : (227 * 1024 - 1024) / STAGE;
      constexpr int SMEM = NST * STAGE + 1024;
      int off_min = 0, off_max = 0;
      for (int t = 0; t < op.taps; ++t) { off_min = std::min(off_min, op.a_row_off[t]); off_max = std::max(off_max, op.a_row_off[t]); }
      gemm::PairMaps tm;
      tm.b_packed = d.b_packed;
      bool maps_ok = op.a_row0 + off_min >= 0 && 2 * (op.a_row0 + op.m_tiles * gemm::BLOCK_M + off_max) < (1LL << 31);
      if (maps_ok) {
        const unsigned long long adims[2] = {2ull * (unsigned long long)op.a_rows, (unsigned long long)(op.k / 8)};
        const unsigned long long astr[2] = {0, (unsigned long long)op.a_rows * 16};
        const unsigned abox[2] = {2 * gemm::BLOCK_M, (unsigned)KCH};
        const void* abase[2] = {op.a_hi, op.a_lo};
        for (int p = 0; p < PLANES && maps_ok; ++p) maps_ok = gpemsr::encode_u64_map(&tm.a[p], abase[p], 2, adims, astr, abox);
      }
      if (maps_ok && d.b_packed) {
        const unsigned long long blocks = (unsigned long long)op.n_tiles * op.taps * (op.k / BLOCK_K);
        const unsigned long long bdims[3] = {2ull * BLOCK_N, (unsigned long long)(PLANES * KCH), blocks};
        const unsigned long long bstr[3] = {0, (unsigned long long)BLOCK_N * 16, (unsigned long long)PLANES * KCH * BLOCK_N * 16};
        const unsigned bbox[3] = {(unsigned)BLOCK_N, (unsigned)(PLANES * KCH), 1};
        maps_ok = gpemsr::encode_u64_map(&tm.b[0], op.b_hi, 3, bdims, bstr, bbox);
      } else if (maps_ok) {
        const unsigned long long bdims[2] = {2ull * (unsigned long long)op.b_rows, (unsigned long long)op.taps * (op.k / 8)};
        const unsigned long long bstr[2] = {0, (unsigned long long)op.b_rows * 16};
        const unsigned bbox[2] = {(unsigned)BLOCK_N, (unsigned)KCH};
        const void* bbase[2] = {op.b_hi, op.b_lo};
        for (int p = 0; p < PLANES && maps_ok; ++p) maps_ok = gpemsr::encode_u64_map(&tm.b[p], bbase[p], 2, bdims, bstr, bbox);
      }
      if (maps_ok) {
        auto kern = gemm::gemm_pair_kernel<BLOCK_N, BLOCK_K, SPLIT, NST, Epi>;
        GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        GPEMSR_CUDA_OK(gpemsr::launch_cluster(kern, dim3((unsigned)gx, (unsigned)gy), dim3(gemm::num_threads<Epi>()), SMEM, s, 2, op, e, tm));
        gpemsr::count_launch();
        return GPEMSR_OK;
      }
    }
  }
  if (clustered && BLOCK_N >= 128) {
    // wide tiles are bound by operand traffic out of L2: pairs of CTAs share every B stage by multicast
    auto kern = gemm::gemm_kernel<BLOCK_N, BLOCK_K, SPLIT, NSTAGE, Epi, 2>;
    GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    GPEMSR_CUDA_OK(gpemsr::launch_cluster(kern, dim3((unsigned)gx, (unsigned)gy), dim3(gemm::num_threads<Epi>()), Cfg::SMEM_BYTES, s, 2, op, e));
    gpemsr::count_launch();
    return GPEMSR_OK;
  }
  auto kern = gemm::gemm_kernel<BLOCK_N, BLOCK_K, SPLIT, NSTAGE, Epi>;
  GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  kern<<<dim3((unsigned)gx, (unsigned)gy), gemm::num_threads<Epi>(), Cfg::SMEM_BYTES, s>>>(op, e);
  GPEMSR_LAUNCH_OK("gemm_kernel<EpiConv>");
  return GPEMSR_OK;
}

// Tensor maps of the A planes for the tap-fused kernel, in 8-byte elements (see encode_u64_map): a segment of seg_len cells of one
// k-cell column is 2 x seg_len elements, so the map is {2 * rows, 2 (stride = half a segment), segments (stride = one padded image
// row), k-cells} with box {seg_len, 2, n_seg, 2} -> shared memory [k-cell][segment][row][8 bf16].  The half-segment and segment
// dimensions OVERLAP the row dimension (shifted views of the same cells; the driver accepts that); the rows extent is shortened
// so the last segment stays inside the allocation.  use = 0 (bulk copies) when the segments are not evenly spaced, seg_len is
// odd, the driver refuses the shape, or GPEMSR_TMA=0.
void maps_tapfuse(gemm::TmaMaps& tm, const gemm::Operands& op, int planes) {
  tm.use = 0;
  if (!gpemsr::use_tensor_maps() || op.seg_len > 256 || (op.seg_len & 1)) return;
  long long seg_stride = op.a_rows;                     // n_seg == 1: any valid stride
  if (op.n_seg > 1) {
    seg_stride = op.seg_row_off[1] - op.seg_row_off[0];
    for (int i = 2; i < op.n_seg; ++i) if (op.seg_row_off[i] - op.seg_row_off[i - 1] != seg_stride) return;
    if (seg_stride <= 0) return;
  }
  const long long rows = op.a_rows - (op.n_seg > 1 ? (op.n_seg - 1) * seg_stride : 0) - op.seg_len / 2;
  if (rows <= 0 || op.a_row0 + op.seg_row_off[0] < 0 || 2 * (op.a_row0 + op.m_tiles * gemm::BLOCK_M + op.seg_row_off[0]) >= (1LL << 31)) return;
  const unsigned long long dims[4] = {2ull * rows, 2, (unsigned long long)op.n_seg, (unsigned long long)(op.k / 8)};
  const unsigned long long strides[4] = {0, (unsigned long long)op.seg_len * 8, (unsigned long long)seg_stride * 16, (unsigned long long)op.a_rows * 16};
  const unsigned box[4] = {(unsigned)op.seg_len, 2, (unsigned)op.n_seg, 2u * (unsigned)op.kslabs};
  const void* base[2] = {op.a_hi, op.a_lo};
  for (int p = 0; p < planes; ++p)
    if (!gpemsr::encode_u64_map(&tm.a[p], base[p], 4, dims, strides, box)) return;
  tm.use = 1;
}

// dy-fused kernel: {2 * rows, 2, k-cells}, box {seg_len, 2, 2}
void maps_dyfuse(gemm::TmaMaps& tm, const gemm::Operands& op, int planes) {
  tm.use = 0;
  if (!gpemsr::use_tensor_maps() || op.seg_len > 256 || (op.seg_len & 1)) return;
  if (op.a_row0 + op.dy_row_off[0] < 0 || 2 * (op.a_row0 + op.m_tiles * gemm::BLOCK_M + op.dy_row_off[op.n_dy - 1]) >= (1LL << 31)) return;
  const long long rows = op.a_rows - op.seg_len / 2;
  const unsigned long long dims[3] = {2ull * rows, 2, (unsigned long long)(op.k / 8)};
  const unsigned long long strides[3] = {0, (unsigned long long)op.seg_len * 8, (unsigned long long)op.a_rows * 16};
  const unsigned box[3] = {(unsigned)op.seg_len, 2, 2};
  const void* base[2] = {op.a_hi, op.a_lo};
  for (int p = 0; p < planes; ++p)
    if (!gpemsr::encode_u64_map(&tm.a[p], base[p], 3, dims, strides, box)) return;
  tm.use = 1;
}

// the specialised epilogues (EpiConv<..., LEAN>, a feature mask): 1 = scale, per-column bias, ReLU / LeakyReLU, residual, fp32 /
// plane stores of whole column tiles; 3 = + patch correlation (VGG mask); 5 = + partial column tiles / row-major stores; 9 = +
// parity-phase scatter; 17 = + fused softmax; 0 = the generic path
int lean_mode(const gpemsr_igemm_desc_t& d, int block_n) {
  static const int enabled = [] { const char* e = getenv("GPEMSR_LEAN"); return (e && e[0] == '0') ? 0 : 1; }();      // GPEMSR_LEAN=0: A/B, tests
  if (!enabled || d.gn_sums || d.pixel_shuffle || d.out_nchw || d.bias_per_row) return 0;
  const bool softmax = d.row_max_out || d.row_max || d.row_sum || d.row_div || d.act == GPEMSR_ACT_EXP;
  const bool partial = d.out_rowmajor || d.n_cols % block_n != 0;
  const int special = (softmax ? 1 : 0) + (partial ? 1 : 0) + (d.phase_cols ? 1 : 0) + (d.patch_sums ? 1 : 0);
  if (special > 1) return 0;                          // one special feature at a time has a specialised epilogue
  if (softmax) return 17;                             // attention GEMMs: row-max pre-pass, scores + exp + row sums, P v^T / row sum
  if (d.phase_cols) return (d.phase_cols % 32 == 0 && (d.out_f32 || d.out_hi)) ? 9 : 0;      // merged ConvTranspose2d phases, whole cells per phase
  if (partial) return (d.out_f32 || d.out_hi || d.out_rowmajor) ? 5 : 0;                      // the tap GEMMs of the few-output convs
  if (d.patch_sums) return 3;
  return (d.out_f32 || d.out_hi) ? 1 : 0;
}
bool lean_epilogue(const gpemsr_igemm_desc_t& d, int block_n) { return lean_mode(d, block_n) == 1; }

template <int BLOCK_N, int SPLIT, int LEAN = 0>
int launch_fused(const gemm::Operands& op, const gpemsr_igemm_desc_t& d, size_t smem_bytes, cudaStream_t s) {
  if constexpr (LEAN == 0) {
    const int mode = lean_mode(d, BLOCK_N);
    if (mode == 1) return launch_fused<BLOCK_N, SPLIT, 1>(op, d, smem_bytes, s);
    if constexpr (BLOCK_N == 64 && SPLIT == 1) {
      if (mode == 3) return launch_fused<BLOCK_N, SPLIT, 3>(op, d, smem_bytes, s);       // VGG conv1_2 of the second image
    }
    if (mode == 5) return launch_fused<BLOCK_N, SPLIT, 5>(op, d, smem_bytes, s);
  }
  using Epi = EpiConv<BLOCK_N, SPLIT == 3, LEAN>;
  Epi e;
  e.ag = to_geom(d.a_geom); e.og = to_geom(d.o_geom);
  e.n_cols = d.n_cols; e.scale = d.scale; e.bias = d.bias; e.bias_per_row = d.bias_per_row; e.act = d.act; e.slope = d.slope;
  e.residual = d.residual; e.up = d.up; e.py = d.py; e.px = d.px; e.pixel_shuffle = d.pixel_shuffle; e.phase_cols = d.phase_cols; e.c_off = d.c_off;
  e.out_f32 = d.out_f32; e.out_hi = (__nv_bfloat16*)d.out_hi; e.out_lo = (__nv_bfloat16*)d.out_lo;
  e.out_nchw = d.out_nchw; e.nchw_c = d.nchw_c; e.out_rowmajor = d.out_rowmajor; e.ld = d.ld;
  e.gn_sums = d.gn_sums; e.gn_cpg = d.gn_cpg;
  e.patch_other = (const float*)d.patch_other; e.patch_other_bf16 = d.patch_other_bf16; e.patch_sums = d.patch_sums; e.patch_size = d.patch_size;
  e.row_max_out = d.row_max_out; e.row_max = d.row_max; e.row_sum = d.row_sum; e.row_div = d.row_div;
  auto kern = gemm::gemm_tapfuse_kernel<BLOCK_N, SPLIT, Epi>;
  GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const long long gx = std::min<long long>(op.m_tiles, gpemsr::num_sms());
  gemm::TmaMaps tm;
  maps_tapfuse(tm, op, SPLIT == 3 ? 2 : 1);
  kern<<<(unsigned)gx, gemm::num_threads<Epi>(), smem_bytes, s>>>(op, e, tm);
  GPEMSR_LAUNCH_OK("gemm_tapfuse_kernel<EpiConv>");
  return GPEMSR_OK;
}

// dy-fused plan (b_packed == 2): the taps must be a full n_dy x n_dx grid in dy-major order with contiguous offsets.
// Returns the dynamic smem size or 0 when the shape does not qualify.
size_t plan_dyfuse(gemm::Operands& op, const gpemsr_igemm_desc_t& d, int block_n, int wp) {
  const int planes = d.split == 3 ? 2 : 1;
  if (d.taps < 2) return 0;
  int n_dx = 1;
  while (n_dx < d.taps && d.tap_dy[n_dx] == d.tap_dy[0]) ++n_dx;
  if (d.taps % n_dx) return 0;
  const int n_dy = d.taps / n_dx;
  if (n_dy > 8) return 0;
  for (int t = 0; t < d.taps; ++t)
    if (d.tap_dy[t] != d.tap_dy[0] + t / n_dx || d.tap_dx[t] != d.tap_dx[0] + t % n_dx) return 0;
  op.n_dy = n_dy; op.n_dx = n_dx;
  op.seg_len = gemm::BLOCK_M + n_dx - 1;
  for (int i = 0; i < n_dy; ++i) op.dy_row_off[i] = (d.tap_dy[0] + i) * wp + d.tap_dx[0];
  const size_t stage_bytes = (size_t)planes * (((2 * (size_t)op.seg_len * 16 + 127) & ~(size_t)127) + (size_t)n_dx * 2 * block_n * 16);
  const size_t budget = 227 * 1024 - 1024;
  if (3 * stage_bytes > budget) return 0;
  op.nstage = (int)std::min<size_t>(8, budget / stage_bytes);
  return (size_t)op.nstage * stage_bytes + 1024;
}

template <int BLOCK_N, int SPLIT, int LEAN = 0>
int launch_dyfuse(const gemm::Operands& op, const gpemsr_igemm_desc_t& d, size_t smem_bytes, cudaStream_t s) {
  if constexpr (LEAN == 0) {
    const int mode = lean_mode(d, BLOCK_N);
    if (mode == 1) return launch_dyfuse<BLOCK_N, SPLIT, 1>(op, d, smem_bytes, s);
    if (mode == 5) return launch_dyfuse<BLOCK_N, SPLIT, 5>(op, d, smem_bytes, s);
  }
  using Epi = EpiConv<BLOCK_N, SPLIT == 3, LEAN>;
  Epi e;
  e.ag = to_geom(d.a_geom); e.og = to_geom(d.o_geom);
  e.n_cols = d.n_cols; e.scale = d.scale; e.bias = d.bias; e.bias_per_row = d.bias_per_row; e.act = d.act; e.slope = d.slope;
  e.residual = d.residual; e.up = d.up; e.py = d.py; e.px = d.px; e.pixel_shuffle = d.pixel_shuffle; e.phase_cols = d.phase_cols; e.c_off = d.c_off;
  e.out_f32 = d.out_f32; e.out_hi = (__nv_bfloat16*)d.out_hi; e.out_lo = (__nv_bfloat16*)d.out_lo;
  e.out_nchw = d.out_nchw; e.nchw_c = d.nchw_c; e.out_rowmajor = d.out_rowmajor; e.ld = d.ld;
  e.gn_sums = d.gn_sums; e.gn_cpg = d.gn_cpg;
  e.patch_other = (const float*)d.patch_other; e.patch_other_bf16 = d.patch_other_bf16; e.patch_sums = d.patch_sums; e.patch_size = d.patch_size;
  e.row_max_out = d.row_max_out; e.row_max = d.row_max; e.row_sum = d.row_sum; e.row_div = d.row_div;
  auto kern = gemm::gemm_dyfuse_kernel<BLOCK_N, SPLIT, Epi>;
  GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const long long gx = std::min<long long>(op.m_tiles, gpemsr::num_sms());
  gemm::TmaMaps tm;
  maps_dyfuse(tm, op, SPLIT == 3 ? 2 : 1);
  kern<<<(unsigned)gx, gemm::num_threads<Epi>(), smem_bytes, s>>>(op, e, tm);
  GPEMSR_LAUNCH_OK("gemm_dyfuse_kernel<EpiConv>");
  return GPEMSR_OK;
}

// Tap-fused plan: segments = distinct tap dy, each covering the dx range of all taps.  Returns the dynamic smem size, or 0
// when the resident B plus >= 3 A stages do not fit.
size_t plan_tapfuse(gemm::Operands& op, const gpemsr_igemm_desc_t& d, int block_n, int wp) {
  const int planes = d.split == 3 ? 2 : 1;
  int dx_min = 1 << 20, dx_max = -(1 << 20), dys[3], n_seg = 0;
  for (int t = 0; t < d.taps; ++t) {
    dx_min = std::min(dx_min, d.tap_dx[t]); dx_max = std::max(dx_max, d.tap_dx[t]);
    int sidx = -1;
    for (int i = 0; i < n_seg; ++i) if (dys[i] == d.tap_dy[t]) sidx = i;
    if (sidx < 0) { if (n_seg == 3) return 0; sidx = n_seg; dys[n_seg++] = d.tap_dy[t]; }
    op.tap_seg[t] = sidx;
  }
  op.n_seg = n_seg;
  op.seg_len = gemm::BLOCK_M + dx_max - dx_min;
  for (int i = 0; i < n_seg; ++i) op.seg_row_off[i] = dys[i] * wp + dx_min;
  for (int t = 0; t < d.taps; ++t) op.tap_dx[t] = d.tap_dx[t] - dx_min;
  const size_t b_bytes = ((size_t)planes * d.taps * (d.k_pad / 8) * block_n * 16 + 1023) & ~(size_t)1023;
  const size_t budget = 227 * 1024 - 1024;
  // k-slabs per stage: as many as leave room for three stages (fewer ring hand-shakes per tile; the single-plane 64-channel
  // layers fit a whole tile -- 4 slabs -- per stage, the (hi, lo) ones one slab)
  size_t stage_bytes = 0;
  for (int ks = 4; ks >= 1; ks >>= 1) {
    if ((d.k_pad / 16) % ks) continue;
    stage_bytes = (size_t)planes * (((size_t)ks * n_seg * 2 * op.seg_len * 16 + 127) & ~(size_t)127);
    op.kslabs = ks;
    if (b_bytes + 3 * stage_bytes <= budget) break;
  }
  if (b_bytes + 3 * stage_bytes > budget) return 0;
  op.nstage = (int)std::min<size_t>(8, (budget - b_bytes) / stage_bytes);
  return b_bytes + (size_t)op.nstage * stage_bytes + 1024;
}

// ------------------------------------------------------------------------------------------------ pack / unpack
__global__ void act_pack_nchw_kernel(const float* __restrict__ x, int c, Geom g, int c_off, float* __restrict__ f32,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long hw = (long long)g.h * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (img, cell, pixel), pixel fastest
  const int cells = (c + 7) / 8;
  if (t >= (long long)g.n * cells * hw) return;
  const long long p = t % hw;
  const int cc = (int)((t / hw) % cells), img = (int)(t / (hw * cells));
  const int y = (int)(p / g.w), xx = (int)(p % g.w);
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cc * 8 + j;
    v[j] = ch < c ? __ldg(x + ((long long)img * c + ch) * hw + p) : 0.f;
  }
  const size_t cell = ((size_t)((c_off >> 3) + cc) * g.rows_alloc + place_row(g, img, y, xx)) * 8;
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

__global__ void act_unpack_nchw_kernel(const float* __restrict__ f32, int c, Geom g, int c_off, float* __restrict__ x) {
  const long long hw = (long long)g.h * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cells = (c + 7) / 8;
  if (t >= (long long)g.n * cells * hw) return;
  const long long p = t % hw;
  const int cc = (int)((t / hw) % cells), img = (int)(t / (hw * cells));
  const int y = (int)(p / g.w), xx = (int)(p % g.w);
  const size_t cell = ((size_t)((c_off >> 3) + cc) * g.rows_alloc + place_row(g, img, y, xx)) * 8;
  const float4 a = *reinterpret_cast<const float4*>(f32 + cell), b = *reinterpret_cast<const float4*>(f32 + cell + 4);
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cc * 8 + j;
    if (ch < c) x[((long long)img * c + ch) * hw + p] = v[j];
  }
}

// ------------------------------------------------------------------------------------------------ DCNv2 (deformable im2col)
// torchvision.ops.deform_conv2d (modulated, 3x3, stride 1, pad 1, dilation 1) as a gather + a plain GEMM.  One thread = one
// (pixel, tap k, deformable group g): the group's 8 channels are ONE 32-byte fp32 cell of the input, sampled bilinearly at
// (y - 1 + ky + off_y, x - 1 + kx + off_x) with torchvision's border rule (a sample outside (-1, H) x (-1, W) is 0; corners
// outside the image contribute 0), multiplied by sigmoid(mask) and written as operand cell (k * groups + g) of the
// [pixels x 9 * C] im2col matrix (hi, lo planes).  The weights then act as a 1x1 convolution over those 9 * C channels.
// `om` is the conv_offset output in NCHW fp32 [n, 3 * groups * 9, h, w]: offset pair (y, x) of (g, k) = channels
// 2 * (g * 9 + k), + 1 ; mask logit of (g, k) = channel 2 * groups * 9 + g * 9 + k  (BasicSR DCNv2Pack: cat(o1, o2), mask).
__global__ void deform_im2col_kernel(const float* __restrict__ xf32, Geom gx, int groups, const float* __restrict__ om,
                                     uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, Geom go) {
  const long long hw = (long long)gx.h * gx.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (img, k, g, pixel): pixel fastest
  if (t >= (long long)gx.n * 9 * groups * hw) return;
  const long long p = t % hw;
  const int g = (int)((t / hw) % groups), k = (int)((t / (hw * groups)) % 9), img = (int)(t / (hw * groups * 9));
  const int y = (int)(p / gx.w), x = (int)(p % gx.w);
  const float* o = om + (long long)img * 3 * groups * 9 * hw + p;
  const int gk = g * 9 + k;
  const float off_y = __ldg(o + (long long)(2 * gk) * hw), off_x = __ldg(o + (long long)(2 * gk + 1) * hw);
  const float ml = __ldg(o + (long long)(2 * groups * 9 + gk) * hw);
  const float mask = 1.0f / (1.0f + expf(-ml));
  const float sy = (float)(y - 1 + k / 3) + off_y, sx = (float)(x - 1 + k % 3) + off_x;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (sy > -1.f && sx > -1.f && sy < (float)gx.h && sx < (float)gx.w) {
    const int y0 = (int)floorf(sy), x0 = (int)floorf(sx);
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float w4[4] = {hy * hx, hy * lx, ly * hx, ly * lx};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int yy = y0 + (c >> 1), xx = x0 + (c & 1);
      if (yy >= 0 && yy < gx.h && xx >= 0 && xx < gx.w) {
        const float* cell = xf32 + ((size_t)g * gx.rows_alloc + place_row(gx, img, yy, xx)) * 8;
        const float4 a = __ldg(reinterpret_cast<const float4*>(cell)), b = __ldg(reinterpret_cast<const float4*>(cell + 4));
        v[0] = fmaf(w4[c], a.x, v[0]); v[1] = fmaf(w4[c], a.y, v[1]); v[2] = fmaf(w4[c], a.z, v[2]); v[3] = fmaf(w4[c], a.w, v[3]);
        v[4] = fmaf(w4[c], b.x, v[4]); v[5] = fmaf(w4[c], b.y, v[5]); v[6] = fmaf(w4[c], b.z, v[6]); v[7] = fmaf(w4[c], b.w, v[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= mask;
  }
  uint4 h4, l4;
  split8(v, h4, l4);
  const size_t dst = (size_t)(k * groups + g) * go.rows_alloc + place_row(go, img, y, x);
  out_hi[dst] = h4;
  if (out_lo) out_lo[dst] = l4;
}

// ------------------------------------------------------------------------------------------------ SpyNet helpers
// ATen upsample_bilinear2d (the arithmetic of F.interpolate(mode='bilinear')) with the source index computed the way ATen
// does: align_corners ? dst * (in-1)/(out-1) : max((dst + 0.5) * (in/out or 1/scale_factor) - 0.5, 0).
//   out[n, co, y, x] (=|+=) mul[co] * (bilinear(x[n, ci(co), ...]) - sub[co]) / div[co]
// ci(co) = co % c_in (a one-channel frame broadcast to three channels), sub/div/mul NULL = identity.  rep_h / rep_w > 0:
// output rows / columns >= rep_h / rep_w repeat the last natural one (F.pad(..., mode='replicate') after the upsampling,
// SpyNet's odd-size case).  rh / rw are the host-computed fp32 scales (ATen's area_pixel_compute_scale).
__global__ void resize_bilinear_kernel(const float* __restrict__ x, int n, int c_in, int h, int w, int c_out, int ho, int wo,
                                       int align_corners, float rh, float rw, int rep_h, int rep_w,
                                       const float* __restrict__ sub, const float* __restrict__ div,
                                       const float* __restrict__ mul, int accumulate, float* __restrict__ out,
                                       float* __restrict__ out_nhwc) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n * c_out * ho * wo) return;
  const int ox0 = (int)(t % wo), oy0 = (int)((t / wo) % ho), co = (int)((t / ((long long)wo * ho)) % c_out);
  const int img = (int)(t / ((long long)wo * ho * c_out));
  const int oy = rep_h > 0 ? min(oy0, rep_h - 1) : oy0, ox = rep_w > 0 ? min(ox0, rep_w - 1) : ox0;
  float sy, sx;
  if (align_corners) { sy = rh * (float)oy; sx = rw * (float)ox; }
  else { sy = fmaxf(rh * ((float)oy + 0.5f) - 0.5f, 0.f); sx = fmaxf(rw * ((float)ox + 0.5f) - 0.5f, 0.f); }
  const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const float* p = x + ((long long)img * c_in + co % c_in) * h * w;
  float v = hy * (hx * __ldg(p + (long long)y0 * w + x0) + lx * __ldg(p + (long long)y0 * w + x1)) +
            ly * (hx * __ldg(p + (long long)y1 * w + x0) + lx * __ldg(p + (long long)y1 * w + x1));
  if (sub) v = v - sub[co];
  if (div) v = v / div[co];
  if (mul) v = v * mul[co];
  if (out) { if (accumulate) out[t] += v; else out[t] = v; }
  if (out_nhwc) out_nhwc[(((long long)img * ho + oy0) * wo + ox0) * c_out + co] = v;
}

// F.avg_pool2d(x, 2, 2, count_include_pad=False) on even sizes: the mean of each 2 x 2 block, ATen's order of additions
__global__ void avg_pool2_kernel(const float* __restrict__ x, long long planes, int h, int w, float* __restrict__ out) {
  const int ho = h / 2, wo = w / 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= planes * ho * wo) return;
  const int ox = (int)(t % wo), oy = (int)((t / wo) % ho);
  const float* p = x + (t / ((long long)wo * ho)) * h * w + (long long)(2 * oy) * w + 2 * ox;
  out[t] = (((__ldg(p) + __ldg(p + 1)) + __ldg(p + w)) + __ldg(p + w + 1)) / 4.0f;
}

// channel concatenation of up to three NCHW fp32 tensors straight into one 8-channel cell column (hi, lo planes):
// SpyNet's torch.cat([ref, warped, flow], 1) (3 + 3 + 2 channels) fused with the operand packing.
__global__ void pack_concat3_kernel(const float* __restrict__ a, int ca, const float* __restrict__ b, int cb,
                                    const float* __restrict__ c, int cc, Geom g, uint4* __restrict__ hi, uint4* __restrict__ lo) {
  const long long hw = (long long)g.h * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)g.n * hw) return;
  const int img = (int)(t / hw);
  const long long p = t % hw;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float val = 0.f;
    if (j < ca) val = __ldg(a + ((long long)img * ca + j) * hw + p);
    else if (j < ca + cb) val = __ldg(b + ((long long)img * cb + (j - ca)) * hw + p);
    else if (j < ca + cb + cc) val = __ldg(c + ((long long)img * cc + (j - ca - cb)) * hw + p);
    v[j] = val;
  }
  uint4 h4, l4;
  split8(v, h4, l4);
  const size_t cell = (size_t)place_row(g, img, (int)(p / g.w), (int)(p % g.w));
  hi[cell] = h4;
  if (lo) lo[cell] = l4;
}

// ---- VGG19 conv1_1 on a one-channel image (the three input channels of the reference are copies of each other, so their
// weights are summed on the host): thread = pixel, 9 loads shared by all output channels, one 16-byte (hi, lo) cell pair per 8
// channels.  Weights / bias are broadcast reads from shared memory.
__global__ void __launch_bounds__(256)
conv3x3_c1_relu_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ bias, int co, Geom g,
                       uint4* __restrict__ out_hi, uint4* __restrict__ out_lo) {
  __shared__ __align__(16) float sw[64 * 9 + 64];
  for (int i = threadIdx.x; i < co * 9; i += blockDim.x) sw[i] = w1[i];
  for (int i = threadIdx.x; i < co; i += blockDim.x) sw[64 * 9 + i] = bias[i];
  __syncthreads();
  const long long hw = (long long)g.h * g.w;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)g.n * hw) return;
  const int img = (int)(t / hw), y = (int)((t % hw) / g.w), xx = (int)(t % g.w);
  const float* p = x + (long long)img * hw;
  float v[9];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int yy = y + dy, xc = xx + dx;
      v[(dy + 1) * 3 + dx + 1] = (yy >= 0 && yy < g.h && xc >= 0 && xc < g.w) ? __ldg(p + (long long)yy * g.w + xc) : 0.f;
    }
  const long long row = place_row(g, img, y, xx);
  for (int cc = 0; cc < co / 8; ++cc) {
    // the 72 weights of a cell as 18 broadcast LDS.128 instead of 72 LDS.32 (the kernel is issue bound: 576 MACs per pixel)
    float wv[72];
    const float4* w4 = reinterpret_cast<const float4*>(sw + cc * 72);
#pragma unroll
    for (int q = 0; q < 18; ++q) { const float4 f4 = w4[q]; wv[4 * q] = f4.x; wv[4 * q + 1] = f4.y; wv[4 * q + 2] = f4.z; wv[4 * q + 3] = f4.w; }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float a = sw[64 * 9 + cc * 8 + j];
#pragma unroll
      for (int k = 0; k < 9; ++k) a = fmaf(v[k], wv[j * 9 + k], a);
      o[j] = fmaxf(a, 0.f);
    }
    uint4 hi, lo;
    split8(o, hi, lo);
    out_hi[(size_t)cc * g.rows_alloc + row] = hi;
    if (out_lo) out_lo[(size_t)cc * g.rows_alloc + row] = lo;
  }
}

__global__ void patch_cosine_kernel(const float* __restrict__ s, long long n, float eps, float* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ab = s[3 * i], na = fmaxf(sqrtf(s[3 * i + 1]), eps), nb = fmaxf(sqrtf(s[3 * i + 2]), eps);
  mask[i] = ab / (na * nb);
}

// space-to-depth of the bf16 operand planes (stride-2 convolutions, model/blocks.py:41-47 DownBlock): output pixel (y, x) of
// channel block (p*2 + q) holds input pixel (2y + p, 2x + q); whole 16-byte cells move, pixels past an odd edge are zero.
__global__ void space_to_depth_kernel(const uint4* __restrict__ in_hi, const uint4* __restrict__ in_lo, Geom gi, int cells,
                                      uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, Geom go) {
  const long long hw = (long long)go.h * go.w, total = (long long)go.n * 4 * cells * hw;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // (img, phase, cell, pixel): pixel fastest
  if (t >= total) return;
  int cc, ph, img, y, x;
  if (total < (1LL << 32)) {                 // 32-bit index arithmetic whenever the shape allows (always, in the model)
    const unsigned tt = (unsigned)t, uhw = (unsigned)hw, q = tt / uhw, px = tt - q * uhw, q2 = q / (unsigned)cells;
    cc = (int)(q - q2 * (unsigned)cells); ph = (int)(q2 & 3u); img = (int)(q2 >> 2);
    y = (int)(px / (unsigned)go.w); x = (int)(px - (unsigned)y * (unsigned)go.w);
  } else {
    const long long px = t % hw;
    cc = (int)((t / hw) % cells); ph = (int)((t / (hw * cells)) % 4); img = (int)(t / (hw * cells * 4));
    y = (int)(px / go.w); x = (int)(px % go.w);
  }
  const int sy = 2 * y + (ph >> 1), sx = 2 * x + (ph & 1);
  const bool in = sy < gi.h && sx < gi.w;
  const size_t src = (size_t)cc * gi.rows_alloc + (in ? place_row(gi, img, sy, sx) : 0);
  const size_t dst = (size_t)(ph * cells + cc) * go.rows_alloc + place_row(go, img, y, x);
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  out_hi[dst] = in ? in_hi[src] : zero;
  if (out_lo) out_lo[dst] = in ? in_lo[src] : zero;
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int n, int k, long long n_stride, long long k_stride, int taps,
                                    const int* __restrict__ tap_src, int b_rows, int k_pad, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (tap, kc, row)
  const int kcs = k_pad / 8;
  if (t >= (long long)taps * kcs * b_rows) return;
  const int row = (int)(t % b_rows), kc = (int)((t / b_rows) % kcs), tap = (int)(t / ((long long)b_rows * kcs));
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int kk = kc * 8 + j;
    v[j] = (row < n && kk < k) ? __ldg(w + row * n_stride + kk * k_stride + tap_src[tap]) : 0.f;
  }
  uint4 h, l;
  split8(v, h, l);
  *reinterpret_cast<uint4*>(hi + (size_t)t * 8) = h;
  if (lo) *reinterpret_cast<uint4*>(lo + (size_t)t * 8) = l;
}

// [n_tile][tap][k-chunk][plane][k-cell][block_n][8]
__global__ void pack_weights_tiled_kernel(const float* __restrict__ w, int n, int k, long long n_stride, long long k_stride, int taps,
                                          const int* __restrict__ tap_src, int block_n, int kch, int planes, int kchunks, int n_tiles,
                                          __nv_bfloat16* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_tiles * taps * kchunks * planes * kch * block_n;
  if (t >= total) return;
  long long r = t;
  const int nn = (int)(r % block_n); r /= block_n;
  const int kc = (int)(r % kch); r /= kch;
  const int plane = (int)(r % planes); r /= planes;
  const int kchunk = (int)(r % kchunks); r /= kchunks;
  const int tap = (int)(r % taps); r /= taps;
  const int row = (int)r * block_n + nn;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int kk = (kchunk * kch + kc) * 8 + j;
    v[j] = (row < n && kk < k) ? __ldg(w + row * n_stride + kk * k_stride + tap_src[tap]) : 0.f;
  }
  uint4 h, l;
  split8(v, h, l);
  *reinterpret_cast<uint4*>(out + (size_t)t * 8) = plane ? l : h;
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// per-channel sum / sum of squares over an image's rows (ring and tail rows are zero by construction)
__global__ void gn_stats_kernel(const float* __restrict__ x, int c, Geom g, double* __restrict__ sums) {
  const int cc = blockIdx.x, img = blockIdx.y;
  const long long r0 = g.m0 + (long long)img * g.r_img;
  const float* base = x + ((size_t)cc * g.rows_alloc + r0) * 8;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long r = (long long)blockIdx.z * blockDim.x + threadIdx.x; r < g.r_img; r += (long long)gridDim.z * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(base + r * 8), b = *reinterpret_cast<const float4*>(base + r * 8 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
  }
  __shared__ double red[8][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double ds = s[j], dq = q[j];
#pragma unroll
    for (int o = 16; o; o >>= 1) { ds += __shfl_xor_sync(0xffffffffu, ds, o); dq += __shfl_xor_sync(0xffffffffu, dq, o); }
    if (lane == 0) { red[warp][2 * j] = ds; red[warp][2 * j + 1] = dq; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0;
    for (int wv = 0; wv < (blockDim.x >> 5); ++wv) t += red[wv][threadIdx.x];
    const int ch = cc * 8 + (threadIdx.x >> 1);
    if (ch < c) atomicAdd(sums + ((size_t)img * c + ch) * 2 + (threadIdx.x & 1), t);
  }
}

__global__ void gn_scale_shift_kernel(const double* __restrict__ sums, int per_group, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, int n, int c, int groups, double count, float eps,
                                      float* __restrict__ ss) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * c) return;
  const int img = t / c, ch = t % c, cpg = c / groups, g0 = (ch / cpg) * cpg;
  double s = 0, q = 0;
  if (per_group) {
    s = sums[((size_t)img * groups + ch / cpg) * 2]; q = sums[((size_t)img * groups + ch / cpg) * 2 + 1];
  } else {
    for (int j = 0; j < cpg; ++j) { s += sums[((size_t)img * c + g0 + j) * 2]; q += sums[((size_t)img * c + g0 + j) * 2 + 1]; }
  }
  const double cnt = count * cpg, mean = s / cnt;
  double var = q / cnt - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[ch] * rstd;
  ss[(size_t)t * 2] = sc;
  ss[(size_t)t * 2 + 1] = beta[ch] - (float)mean * sc;
}

// y = act(x * scale + shift) (+ residual); re-rowed into `og` when it differs from `g`
__global__ void affine_act_kernel(const float* __restrict__ x, int c, Geom g, const float* __restrict__ ss, int act, float slope,
                                  const float* __restrict__ residual, Geom og, float* __restrict__ out_f32,
                                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_nchw) {
  // grid: x = pixels of one image, y = (img, cell) -- the kernel is HBM bound only if the index arithmetic stays out of the way:
  // the flat (img, cell, pixel) index of round 1 cost three 64-bit divisions per 64 bytes moved
  const unsigned hw = (unsigned)g.h * (unsigned)g.w;
  const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= hw) return;
  const int cells = (c + 7) / 8;
  const int cc = (int)(blockIdx.y % (unsigned)cells), img = (int)(blockIdx.y / (unsigned)cells);
  const int y = (int)(p / (unsigned)g.w), xx = (int)(p - (unsigned)y * (unsigned)g.w);
  const size_t cell = ((size_t)cc * g.rows_alloc + place_row(g, img, y, xx)) * 8;
  const float4 a = *reinterpret_cast<const float4*>(x + cell), b = *reinterpret_cast<const float4*>(x + cell + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = cc * 8 + j;
    if (ss && ch < c) {
      const float2 s2 = __ldg(reinterpret_cast<const float2*>(ss) + (size_t)img * c + ch);
      v[j] = fmaf(v[j], s2.x, s2.y);
    }
    v[j] = ch < c ? apply_act(v[j], act, slope) : 0.f;
  }
  const size_t ocell = ((size_t)cc * og.rows_alloc + place_row(og, img, y, xx)) * 8;
  if (residual) {
    const float4 r0 = *reinterpret_cast<const float4*>(residual + ocell), r1 = *reinterpret_cast<const float4*>(residual + ocell + 4);
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
  if (out_f32) {
    *reinterpret_cast<float4*>(out_f32 + ocell) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(out_f32 + ocell + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (out_hi) {
    uint4 h, l;
    split8(v, h, l);
    *reinterpret_cast<uint4*>(out_hi + ocell) = h;
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + ocell) = l;
  }
  if (out_nchw) {                        // the reference-layout copy returned to the caller (lanes <-> consecutive pixels)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = cc * 8 + j;
      if (ch < c) out_nchw[((long long)img * c + ch) * hw + p] = v[j];
    }
  }
}

// ------------------------------------------------------------------------------------------------ softmax
// ---- softmax over scores stored as K8-blocked fp32 cells [t_pad/8][rows_alloc][8] (row i = query, cell = 8 consecutive keys).
// Lanes <-> consecutive rows, so every cell access of a warp is one contiguous 1 KB run; no transpose is needed because the
// probabilities are written back in the same cell order (that is the A-operand layout of the P v^T GEMM).
// pass 1: per (row, part) online (max, sum exp) over a slice of the key cells
__global__ void softmax_cells_partial_kernel(const float* __restrict__ s, long long t, long long rows_alloc, int cells_per_part,
                                             float* __restrict__ part) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t) return;
  const long long ncell = (t + 7) / 8;
  const long long c0 = (long long)blockIdx.y * cells_per_part, c1 = min(ncell, c0 + cells_per_part);
  float m = -INFINITY, sum = 0.f;
  for (long long jc = c0; jc < c1; ++jc) {
    const float* p = s + (jc * rows_alloc + i) * 8;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) { if (jc * 8 + j >= t) v[j] = -INFINITY; cm = fmaxf(cm, v[j]); }
    const float mn = fmaxf(m, cm);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += expf(v[j] - mn);
    sum = sum * expf(m - mn) + acc;
    m = mn;
  }
  part[((long long)blockIdx.y * t + i) * 2] = m;
  part[((long long)blockIdx.y * t + i) * 2 + 1] = sum;
}
// pass 2: combine the parts -> stats[i] = (max, sum exp)
__global__ void softmax_cells_combine_kernel(const float* __restrict__ part, long long t, int parts, float* __restrict__ stats) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= t) return;
  float m = -INFINITY;
  for (int p = 0; p < parts; ++p) m = fmaxf(m, part[((long long)p * t + i) * 2]);
  float sum = 0.f;
  for (int p = 0; p < parts; ++p) sum += part[((long long)p * t + i) * 2 + 1] * expf(part[((long long)p * t + i) * 2] - m);
  stats[i * 2] = m;
  stats[i * 2 + 1] = sum;
}
// pass 3: p = exp(s - max) / sum -> (hi, lo) cells in place order
__global__ void softmax_cells_write_kernel(const float* __restrict__ s, long long t, long long rows_alloc, long long t_pad,
                                           const float* __restrict__ stats, __nv_bfloat16* __restrict__ p_hi,
                                           __nv_bfloat16* __restrict__ p_lo) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (jc, i), i fastest
  if (idx >= (t_pad / 8) * t_pad) return;
  const long long i = idx % t_pad, jc = idx / t_pad;
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (i < t && jc * 8 < t) {
    const float mx = stats[i * 2], den = stats[i * 2 + 1];
    const float* p = s + (jc * rows_alloc + i) * 8;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (jc * 8 + j < t) ? expf(x[j] - mx) / den : 0.f;
  }
  uint4 h, l;
  split8(v, h, l);
  *reinterpret_cast<uint4*>(p_hi + (size_t)idx * 8) = h;
  if (p_lo) *reinterpret_cast<uint4*>(p_lo + (size_t)idx * 8) = l;
}

// ------------------------------------------------------------------------------------------------ bilinear base
// ATen upsample_bilinear2d, align_corners=False, scale_factor given: src = (dst + 0.5) / scale - 0.5, clamped at 0
__global__ void add_bilinear_base_kernel(const float* __restrict__ xc, int n, int h, int w, int scale, float* __restrict__ out) {
  const long long Ho = (long long)h * scale, Wo = (long long)w * scale;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * Ho * Wo) return;
  const int ox = (int)(t % Wo), oy = (int)((t / Wo) % Ho), img = (int)(t / (Wo * Ho));
  const float rs = 1.0f / (float)scale;
  float sy = rs * ((float)oy + 0.5f) - 0.5f, sx = rs * ((float)ox + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy; sx = sx < 0.f ? 0.f : sx;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const float* p = xc + (long long)img * h * w;
  const float v = hy * (hx * p[(long long)y0 * w + x0] + lx * p[(long long)y0 * w + x1]) +
                  ly * (hx * p[(long long)y1 * w + x0] + lx * p[(long long)y1 * w + x1]);
  out[t] += v;
}

int pick_block_n(const gpemsr_igemm_desc_t& d) {
  return d.n_cols <= 16 && !d.pixel_shuffle ? 16 : d.n_cols <= 64 ? 64 : d.n_cols <= 128 ? 128 : 256;
}

// Exact evaluation of a 4-phase, 3x3-tap, position-class-dependent linear map on the one-pixel border ring of the 2x
// output (the composed up-block + output conv has different weights where the output conv's zero padding cuts taps off).
//   out[img, ch, 2a+py, 2b+px] = bias[cls][ch] + sum_ci sum_{dy,dx} x[ci, a+dy, b+dx] * wc[cls][ph][ch][ci][dy+1][dx+1]
// cls = ry * 3 + rx with ry in {0: top row, 1: interior, 2: bottom row} of the OUTPUT, likewise rx.
__global__ void border_phase_conv_kernel(const float* __restrict__ xf32, int cin, Geom g, const float* __restrict__ wc,
                                         const float* __restrict__ bias, int cout, float* __restrict__ out) {
  const int Ho = 2 * g.h, Wo = 2 * g.w;
  const int per_img = 2 * Wo + 2 * (Ho - 2);                      // ring pixels of one output image
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)g.n * per_img * cout) return;
  const int ch = (int)(t % cout);
  const long long r = t / cout;
  const int img = (int)(r / per_img);
  int q = (int)(r % per_img), oy, ox;
  if (q < Wo) { oy = 0; ox = q; }
  else if (q < 2 * Wo) { oy = Ho - 1; ox = q - Wo; }
  else { q -= 2 * Wo; oy = 1 + (q >> 1); ox = (q & 1) ? Wo - 1 : 0; }
  const int ry = oy == 0 ? 0 : (oy == Ho - 1 ? 2 : 1), rx = ox == 0 ? 0 : (ox == Wo - 1 ? 2 : 1);
  const int a = oy >> 1, b = ox >> 1, ph = (oy & 1) * 2 + (ox & 1);
  const float* wp = wc + ((((size_t)(ry * 3 + rx) * 4 + ph) * cout + ch) * cin) * 9;
  float acc = bias[(ry * 3 + rx) * cout + ch];
  for (int ci = 0; ci < cin; ++ci) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const long long row = place_row(g, img, a + dy, b + dx);   // the zero ring supplies out-of-image inputs
        acc = fmaf(xf32[((size_t)(ci >> 3) * g.rows_alloc + row) * 8 + (ci & 7)], __ldg(wp + ci * 9 + (dy + 1) * 3 + dx + 1), acc);
      }
  }
  out[(((long long)img * cout + ch) * Ho + oy) * Wo + ox] = acc;
}

}  // namespace

extern "C" {

int gpemsr_igemm(const gpemsr_igemm_desc_t* dp, gpemsr_stream_t stream) {
  using namespace gpemsr;
  if (!dp) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: null descriptor");
  const gpemsr_igemm_desc_t& d = *dp;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if ((rc = check_geom(d.a_geom, "igemm(a)")) != GPEMSR_OK) return rc;
  if ((d.out_f32 || d.out_hi || d.residual) && (rc = check_geom(d.o_geom, "igemm(out)")) != GPEMSR_OK) return rc;
  if (d.taps < 1 || d.taps > gemm::MAX_TAPS || d.k_pad <= 0 || d.k_pad % (d.b_packed == 2 ? 16 : d.split == 3 ? 32 : 64) || d.n_cols <= 0)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: taps=%d k_pad=%d n_cols=%d", d.taps, d.k_pad, d.n_cols);
  if (d.split != 1 && d.split != 3) return set_error(GPEMSR_ERR_UNSUPPORTED, "igemm: split must be 1 or 3");
  if (!d.a_hi || !d.b_hi || (d.split == 3 && (!d.a_lo || (!d.b_lo && !d.b_packed)))) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: null operand");
  if (d.up != 1 && d.up != 2) return set_error(GPEMSR_ERR_UNSUPPORTED, "igemm: up must be 1 or 2");
  if (d.pixel_shuffle && (d.up != 2 || d.n_cols % 32)) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: pixel_shuffle needs up=2 and n_cols %% 32 == 0");
  if (d.phase_cols && (d.up != 2 || d.n_cols != 4 * d.phase_cols || d.pixel_shuffle ||
                       ((d.phase_cols % 32) && (d.out_f32 || d.out_hi || !d.out_nchw || d.n_cols > 16))))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: phase_cols needs up=2, n_cols == 4 * phase_cols and either phase_cols %% 32 == 0 "
                     "or an NCHW-only output with n_cols <= 16");
  if (d.c_off % 8) return set_error(GPEMSR_ERR_BAD_ALIGN, "igemm: c_off must be a multiple of 8");
  if (d.out_rowmajor && (d.ld % 4)) return set_error(GPEMSR_ERR_BAD_ALIGN, "igemm: ld must be a multiple of 4");
  if (!d.err_flag) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: err_flag is required");
  if (d.gn_sums) {
    const int g = d.gn_cpg;
    if (!(g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32) || d.n_cols % g || d.pixel_shuffle || d.phase_cols || d.up != 1)
      return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: gn_sums needs gn_cpg in {1,2,4,8,16,32} dividing n_cols and a same-resolution output");
    GPEMSR_CUDA_OK(cudaMemsetAsync(d.gn_sums, 0, (size_t)d.a_geom.n * (d.n_cols / g) * 2 * sizeof(double), (cudaStream_t)stream));
  }

  if (d.patch_sums) {
    const int ps = d.patch_size;
    if (!d.patch_other || ps <= 0 || d.a_geom.h % ps || d.a_geom.w % ps || d.up != 1 || d.pixel_shuffle || d.phase_cols)
      return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: patch_sums needs patch_other, a same-resolution output and h, w divisible by patch_size");
    if ((rc = check_geom(d.o_geom, "igemm(patch)")) != GPEMSR_OK) return rc;
    GPEMSR_CUDA_OK(cudaMemsetAsync(d.patch_sums, 0, (size_t)d.a_geom.n * (d.a_geom.h / ps) * (d.a_geom.w / ps) * 3 * sizeof(float),
                                   (cudaStream_t)stream));
  }

  if (d.act == GPEMSR_ACT_EXP && !d.row_max) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: GPEMSR_ACT_EXP needs row_max");
  if (d.row_max_out || d.row_sum) {
    if (d.row_max_out && d.row_sum) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: row_max_out (pre-pass) excludes row_sum");
    if (d.up != 1 || d.pixel_shuffle || d.phase_cols) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: row statistics need a same-resolution output");
    const size_t rows = (size_t)d.a_geom.n * d.a_geom.r_img;
    if (d.row_max_out) GPEMSR_CUDA_OK(cudaMemsetAsync(d.row_max_out, 0xFE, rows * sizeof(float), (cudaStream_t)stream));   // -1.7e38
    else GPEMSR_CUDA_OK(cudaMemsetAsync(d.row_sum, 0, rows * sizeof(float), (cudaStream_t)stream));
  }

  const int block_n = pick_block_n(d);
  gemm::Operands op{};
  op.a_hi = (const __nv_bfloat16*)d.a_hi; op.a_lo = (const __nv_bfloat16*)d.a_lo;
  op.b_hi = (const __nv_bfloat16*)d.b_hi; op.b_lo = (const __nv_bfloat16*)d.b_lo;
  op.a_rows = d.a_geom.rows_alloc; op.b_rows = d.b_rows; op.b_packed = d.b_packed; op.k = d.k_pad; op.taps = d.taps;
  const int wp = d.a_geom.w + 2 * d.a_geom.padded;
  for (int t = 0; t < d.taps; ++t) {
    if (abs(d.tap_dy[t]) > d.a_geom.padded || abs(d.tap_dx[t]) > d.a_geom.padded)
      return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: tap (%d, %d) needs a zero ring of that width (geometry has %d)", d.tap_dy[t],
                       d.tap_dx[t], d.a_geom.padded);
    op.a_row_off[t] = d.tap_dy[t] * wp + d.tap_dx[t];
  }
  op.m_tiles = (long long)d.a_geom.n * d.a_geom.r_img / gemm::BLOCK_M;
  op.n_tiles = (d.n_cols + block_n - 1) / block_n;
  if (!d.b_packed && d.b_rows < op.n_tiles * block_n) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: b_rows=%d < %d", d.b_rows, op.n_tiles * block_n);
  op.a_row0 = d.a_geom.m0; op.err_flag = d.err_flag;
  cudaStream_t s = (cudaStream_t)stream;
  if (d.b_packed == 2) {                           // large tap grids on narrow layers: weights streamed per (k-slab, dy)
    const int bn = d.n_cols <= 16 ? 16 : d.n_cols <= 32 ? 32 : 64;
    if (d.n_cols > 64 || d.pixel_shuffle || d.phase_cols) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: dy-fused weights need n_cols <= 64 and a plain output");
    const size_t smem = plan_dyfuse(op, d, bn, wp);
    if (!smem) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: dy-fused weights need a full dy-major tap grid that fits shared memory");
    if (d.b_rows != bn) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: dy-fused weights were packed for %d columns, this launch uses %d", d.b_rows, bn);
    if (d.split == 3) return bn == 16 ? launch_dyfuse<16, 3>(op, d, smem, s) : bn == 32 ? launch_dyfuse<32, 3>(op, d, smem, s) : launch_dyfuse<64, 3>(op, d, smem, s);
    return bn == 16 ? launch_dyfuse<16, 1>(op, d, smem, s) : bn == 32 ? launch_dyfuse<32, 1>(op, d, smem, s) : launch_dyfuse<64, 1>(op, d, smem, s);
  }
  if (block_n <= 64 && op.n_tiles == 1) {          // narrow outputs: B resident in smem, taps share one A fetch
    const size_t smem = plan_tapfuse(op, d, block_n, wp);
    if (smem && d.b_packed) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm: this shape runs the tap-fused kernel, which needs "
                                             "gpemsr_pack_weights() weights (see gpemsr_igemm_plan)");
    if (smem) {
      if (d.split == 3) return block_n == 16 ? launch_fused<16, 3>(op, d, smem, s) : launch_fused<64, 3>(op, d, smem, s);
      return block_n == 16 ? launch_fused<16, 1>(op, d, smem, s) : launch_fused<64, 1>(op, d, smem, s);
    }
  }
  if (d.split == 3) {
    switch (block_n) {
      case 16: return launch<16, 32, 3, 6>(op, d, s);
      case 64: return launch<64, 32, 3, 6>(op, d, s);
      case 128: return launch<128, 32, 3, 5>(op, d, s);
      default: return launch<256, 32, 3, 4>(op, d, s);
    }
  }
  switch (block_n) {
    case 16: return launch<16, 64, 1, 6>(op, d, s);
    case 64: return launch<64, 64, 1, 6>(op, d, s);
    case 128: return launch<128, 64, 1, 5>(op, d, s);
    default: return launch<256, 64, 1, 4>(op, d, s);
  }
}

int gpemsr_igemm_plan(const gpemsr_igemm_desc_t* dp, int32_t* block_n, int32_t* tapfused) {
  using namespace gpemsr;
  if (!dp || !block_n || !tapfused) return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm_plan: null argument");
  if (dp->taps < 1 || dp->taps > gemm::MAX_TAPS || dp->k_pad <= 0 || dp->k_pad % (dp->split == 3 ? 32 : 64) || dp->n_cols <= 0)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "igemm_plan: taps=%d k_pad=%d n_cols=%d", dp->taps, dp->k_pad, dp->n_cols);
  const int bn = pick_block_n(*dp);
  *block_n = bn;
  gemm::Operands op{};
  const int n_tiles = (dp->n_cols + bn - 1) / bn;
  *tapfused = (bn <= 64 && n_tiles == 1 && plan_tapfuse(op, *dp, bn, /*wp (irrelevant for the fit)*/ 0) != 0) ? 1 : 0;
  return GPEMSR_OK;
}

size_t gpemsr_pack_weights_tiled_bytes(int n, int k_pad, int taps, int block_n, int split) {
  if (n <= 0 || k_pad <= 0 || taps <= 0 || block_n <= 0) return 0;
  const int planes = split == 3 ? 2 : 1;
  const size_t n_tiles = (size_t)(n + block_n - 1) / block_n;
  return n_tiles * taps * (size_t)k_pad * planes * block_n * 2;
}

int gpemsr_pack_weights_tiled(const float* w, int n, int k, int64_t n_stride, int64_t k_stride, int taps, const int32_t* tap_src,
                              int block_n, int k_pad, int split, void* out, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  const int block_k = split == 3 ? 32 : 64;
  if (!w || !out || !tap_src || n <= 0 || k <= 0 || taps <= 0 || k_pad < k || k_pad % block_k || (split != 1 && split != 3) ||
      (block_n != 16 && block_n != 64 && block_n != 128 && block_n != 256))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "pack_weights_tiled: bad arguments");
  const int planes = split == 3 ? 2 : 1, kch = block_k / 8, kchunks = k_pad / block_k, n_tiles = (n + block_n - 1) / block_n;
  const long long total = (long long)n_tiles * taps * kchunks * planes * kch * block_n;
  pack_weights_tiled_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, n, k, n_stride, k_stride, taps, tap_src, block_n, kch, planes, kchunks, n_tiles, (__nv_bfloat16*)out);
  GPEMSR_LAUNCH_OK("pack_weights_tiled_kernel");
  return GPEMSR_OK;
}

int gpemsr_act_pack_nchw(const float* x, int c, const gpemsr_geom_t* g, int c_off, float* f32, void* hi, void* lo,
                         gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!g || !x || c <= 0 || c_off % 8) return set_error(GPEMSR_ERR_BAD_SHAPE, "act_pack_nchw: bad arguments");
  if ((rc = check_geom(*g, "act_pack_nchw")) != GPEMSR_OK) return rc;
  const long long total = (long long)g->n * ((c + 7) / 8) * g->h * g->w;
  act_pack_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, c, to_geom(*g), c_off, f32,
                                                                                        (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  GPEMSR_LAUNCH_OK("act_pack_nchw_kernel");
  return GPEMSR_OK;
}

int gpemsr_act_unpack_nchw(const float* f32, int c, const gpemsr_geom_t* g, int c_off, float* x, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!g || !x || !f32 || c <= 0 || c_off % 8) return set_error(GPEMSR_ERR_BAD_SHAPE, "act_unpack_nchw: bad arguments");
  if ((rc = check_geom(*g, "act_unpack_nchw")) != GPEMSR_OK) return rc;
  const long long total = (long long)g->n * ((c + 7) / 8) * g->h * g->w;
  act_unpack_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(f32, c, to_geom(*g), c_off, x);
  GPEMSR_LAUNCH_OK("act_unpack_nchw_kernel");
  return GPEMSR_OK;
}

int gpemsr_deform_im2col(const float* x_f32, const gpemsr_geom_t* gx, int channels, int deform_groups, const float* offset_mask,
                         void* out_hi, void* out_lo, const gpemsr_geom_t* go, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_f32 || !gx || !go || !offset_mask || !out_hi || deform_groups <= 0 || channels != 8 * deform_groups)
    return set_error(GPEMSR_ERR_UNSUPPORTED, "deform_im2col: built for channels == 8 * deformable_groups (one cell per group; "
                     "GPEMSR: 64 channels, 8 groups), 3x3, stride 1, padding 1");
  if ((rc = check_geom(*gx, "deform_im2col(x)")) != GPEMSR_OK || (rc = check_geom(*go, "deform_im2col(out)")) != GPEMSR_OK) return rc;
  if (go->n != gx->n || go->h != gx->h || go->w != gx->w) return set_error(GPEMSR_ERR_BAD_SHAPE, "deform_im2col: geometries differ");
  const long long total = (long long)gx->n * 9 * deform_groups * gx->h * gx->w;
  deform_im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x_f32, to_geom(*gx), deform_groups, offset_mask, (uint4*)out_hi, (uint4*)out_lo, to_geom(*go));
  GPEMSR_LAUNCH_OK("deform_im2col_kernel");
  return GPEMSR_OK;
}

int gpemsr_resize_bilinear(const float* x, int n, int c_in, int h, int w, int c_out, int ho, int wo, int align_corners,
                           float rh, float rw, int rep_h, int rep_w, const float* sub, const float* div, const float* mul,
                           int accumulate, float* out, float* out_nhwc, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x || (!out && !out_nhwc) || n <= 0 || c_in <= 0 || c_out <= 0 || h <= 0 || w <= 0 || ho <= 0 || wo <= 0)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "resize_bilinear: bad arguments");
  const long long total = (long long)n * c_out * ho * wo;
  resize_bilinear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, n, c_in, h, w, c_out, ho, wo, align_corners, rh, rw, rep_h, rep_w, sub, div, mul, accumulate, out, out_nhwc);
  GPEMSR_LAUNCH_OK("resize_bilinear_kernel");
  return GPEMSR_OK;
}

int gpemsr_avg_pool2(const float* x, int64_t planes, int h, int w, float* out, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x || !out || planes <= 0 || h < 2 || w < 2 || (h & 1) || (w & 1))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "avg_pool2: needs even h, w");
  const long long total = planes * (h / 2) * (w / 2);
  avg_pool2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, planes, h, w, out);
  GPEMSR_LAUNCH_OK("avg_pool2_kernel");
  return GPEMSR_OK;
}

int gpemsr_pack_concat3(const float* a, int ca, const float* b, int cb, const float* c, int cc, const gpemsr_geom_t* g,
                        void* out_hi, void* out_lo, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!a || !g || !out_hi || ca <= 0 || cb < 0 || cc < 0 || ca + cb + cc > 8 || (cb && !b) || (cc && !c))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "pack_concat3: at most 8 channels in total");
  if ((rc = check_geom(*g, "pack_concat3")) != GPEMSR_OK) return rc;
  const long long total = (long long)g->n * g->h * g->w;
  pack_concat3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, ca, b, cb, c, cc, to_geom(*g),
                                                                                       (uint4*)out_hi, (uint4*)out_lo);
  GPEMSR_LAUNCH_OK("pack_concat3_kernel");
  return GPEMSR_OK;
}

int gpemsr_conv3x3_c1_relu(const float* x, const float* w1, const float* bias, int co, const gpemsr_geom_t* g, void* out_hi,
                           void* out_lo, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x || !w1 || !bias || !g || !out_hi || co <= 0 || co % 8 || co > 64)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "conv3x3_c1_relu: bad arguments (co must be a multiple of 8, <= 64)");
  if ((rc = check_geom(*g, "conv3x3_c1_relu")) != GPEMSR_OK) return rc;
  const long long total = (long long)g->n * g->h * g->w;
  conv3x3_c1_relu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, w1, bias, co, to_geom(*g),
                                                                                          (uint4*)out_hi, (uint4*)out_lo);
  GPEMSR_LAUNCH_OK("conv3x3_c1_relu_kernel");
  return GPEMSR_OK;
}

int gpemsr_patch_cosine(const float* patch_sums, int64_t n, float eps, float* mask, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (n == 0) return GPEMSR_OK;
  if (!patch_sums || !mask || n < 0) return set_error(GPEMSR_ERR_BAD_SHAPE, "patch_cosine: bad arguments");
  patch_cosine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(patch_sums, n, eps, mask);
  GPEMSR_LAUNCH_OK("patch_cosine_kernel");
  return GPEMSR_OK;
}

int gpemsr_space_to_depth(const void* in_hi, const void* in_lo, const gpemsr_geom_t* gi, int c, void* out_hi, void* out_lo,
                          const gpemsr_geom_t* go, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!in_hi || !out_hi || !gi || !go || c <= 0 || c % 8 || (out_lo && !in_lo))
    return set_error(GPEMSR_ERR_BAD_SHAPE, "space_to_depth: bad arguments (channels must be a multiple of 8)");
  if ((rc = check_geom(*gi, "space_to_depth(in)")) != GPEMSR_OK || (rc = check_geom(*go, "space_to_depth(out)")) != GPEMSR_OK) return rc;
  if (go->n != gi->n || go->h != (gi->h + 1) / 2 || go->w != (gi->w + 1) / 2)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "space_to_depth: output geometry must be ceil(h/2) x ceil(w/2)");
  const long long total = (long long)go->n * 4 * (c / 8) * go->h * go->w;
  space_to_depth_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)in_hi, (const uint4*)in_lo, to_geom(*gi), c / 8, (uint4*)out_hi, (uint4*)out_lo, to_geom(*go));
  GPEMSR_LAUNCH_OK("space_to_depth_kernel");
  return GPEMSR_OK;
}

int gpemsr_pack_weights(const float* w, int n, int k, int64_t n_stride, int64_t k_stride, int taps, const int32_t* tap_src,
                        int b_rows, int k_pad, void* hi, void* lo, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!w || !hi || !tap_src || n <= 0 || k <= 0 || taps <= 0 || b_rows < n || k_pad < k || k_pad % 8)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "pack_weights: bad arguments");
  const long long total = (long long)taps * (k_pad / 8) * b_rows;
  pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, n, k, n_stride, k_stride, taps, tap_src, b_rows, k_pad, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  GPEMSR_LAUNCH_OK("pack_weights_kernel");
  return GPEMSR_OK;
}

int gpemsr_gn_stats(const float* x_f32, int c, const gpemsr_geom_t* g, double* chan_sums, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_f32 || !g || !chan_sums || c <= 0) return set_error(GPEMSR_ERR_BAD_SHAPE, "gn_stats: bad arguments");
  if ((rc = check_geom(*g, "gn_stats")) != GPEMSR_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  GPEMSR_CUDA_OK(cudaMemsetAsync(chan_sums, 0, (size_t)g->n * c * 2 * sizeof(double), s));
  const int cells = (c + 7) / 8;
  long long splits = std::max<long long>(1, std::min<long long>((g->r_img + 256 * 8 - 1) / (256 * 8),
                                                                (4LL * num_sms() + (long long)cells * g->n - 1) / ((long long)cells * g->n)));
  gn_stats_kernel<<<dim3(cells, g->n, (unsigned)splits), 256, 0, s>>>(x_f32, c, to_geom(*g), chan_sums);
  GPEMSR_LAUNCH_OK("gn_stats_kernel");
  return GPEMSR_OK;
}

int gpemsr_gn_scale_shift(const double* chan_sums, int sums_per_group, const float* gamma, const float* beta, int n, int c,
                          int groups, double count_per_channel, float eps, float* scale_shift, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!chan_sums || !gamma || !beta || !scale_shift || n <= 0 || c <= 0 || groups <= 0 || c % groups)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "gn_scale_shift: bad arguments (c=%d groups=%d)", c, groups);
  gn_scale_shift_kernel<<<(n * c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(chan_sums, sums_per_group, gamma, beta, n, c, groups,
                                                                            count_per_channel, eps, scale_shift);
  GPEMSR_LAUNCH_OK("gn_scale_shift_kernel");
  return GPEMSR_OK;
}

int gpemsr_affine_act(const float* x_f32, int c, const gpemsr_geom_t* g, const float* scale_shift, int act, float slope,
                      const float* residual, const gpemsr_geom_t* og, float* out_f32, void* out_hi, void* out_lo,
                      float* out_nchw, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_f32 || !g || !og || c <= 0) return set_error(GPEMSR_ERR_BAD_SHAPE, "affine_act: bad arguments");
  if ((rc = check_geom(*g, "affine_act(in)")) != GPEMSR_OK || (rc = check_geom(*og, "affine_act(out)")) != GPEMSR_OK) return rc;
  if (g->n != og->n || g->h != og->h || g->w != og->w) return set_error(GPEMSR_ERR_BAD_SHAPE, "affine_act: geometries differ in shape");
  const long long hw = (long long)g->h * g->w, planes = (long long)g->n * ((c + 7) / 8);
  if (hw >= (1LL << 31) || planes > 65535) return set_error(GPEMSR_ERR_BAD_SHAPE, "affine_act: %lld pixels x %lld (image, cell) planes exceed the grid", hw, planes);
  affine_act_kernel<<<dim3((unsigned)((hw + 255) / 256), (unsigned)planes), 256, 0, (cudaStream_t)stream>>>(
      x_f32, c, to_geom(*g), scale_shift, act, slope, residual, to_geom(*og), out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
      out_nchw);
  GPEMSR_LAUNCH_OK("affine_act_kernel");
  return GPEMSR_OK;
}

int gpemsr_softmax_cells_blocked(const float* s_cells, int64_t t, int64_t rows_alloc, int64_t t_pad, float* scratch, void* p_hi,
                                 void* p_lo, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!s_cells || !scratch || !p_hi || t <= 0 || rows_alloc < t_pad || t_pad < t || t_pad % 8)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "softmax_cells_blocked: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long ncell = (t + 7) / 8;
  const int parts = (int)std::max<long long>(1, std::min<long long>(16, ncell / 16));
  const int cpp = (int)((ncell + parts - 1) / parts);
  float* part = scratch + 2 * t;                   // scratch: [t][2] stats, then [parts][t][2] partials
  softmax_cells_partial_kernel<<<dim3((unsigned)((t + 127) / 128), parts), 128, 0, st>>>(s_cells, t, rows_alloc, cpp, part);
  GPEMSR_LAUNCH_OK("softmax_cells_partial_kernel");
  softmax_cells_combine_kernel<<<(unsigned)((t + 127) / 128), 128, 0, st>>>(part, t, parts, scratch);
  GPEMSR_LAUNCH_OK("softmax_cells_combine_kernel");
  const long long total = (t_pad / 8) * t_pad;
  softmax_cells_write_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(s_cells, t, rows_alloc, t_pad, scratch,
                                                                           (__nv_bfloat16*)p_hi, (__nv_bfloat16*)p_lo);
  GPEMSR_LAUNCH_OK("softmax_cells_write_kernel");
  return GPEMSR_OK;
}

int gpemsr_border_phase_conv(const float* x_f32, int cin, const gpemsr_geom_t* g, const float* wc, const float* bias, int cout,
                             float* out_nchw, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_f32 || !g || !wc || !bias || !out_nchw || cin <= 0 || cout <= 0 || !g->padded)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "border_phase_conv: bad arguments");
  if ((rc = check_geom(*g, "border_phase_conv")) != GPEMSR_OK) return rc;
  const long long total = (long long)g->n * (4LL * g->w + 2 * (2LL * g->h - 2)) * cout;
  border_phase_conv_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(x_f32, cin, to_geom(*g), wc, bias, cout, out_nchw);
  GPEMSR_LAUNCH_OK("border_phase_conv_kernel");
  return GPEMSR_OK;
}

int gpemsr_add_bilinear_base(const float* x_center, int n, int h, int w, int scale, float* out, gpemsr_stream_t stream) {
  using namespace gpemsr;
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  if (!x_center || !out || n <= 0 || h <= 0 || w <= 0 || scale <= 0) return set_error(GPEMSR_ERR_BAD_SHAPE, "add_bilinear_base: bad arguments");
  const long long total = (long long)n * h * scale * w * scale;
  add_bilinear_base_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_center, n, h, w, scale, out);
  GPEMSR_LAUNCH_OK("add_bilinear_base_kernel");
  return GPEMSR_OK;
}

}  // extern "C"
