// a-1 / a-2: codebook lookup = GEMM [rows x D] . [D x K] + per-row arg-extremum + code-vector gather.
//
// Replaces Codebook.forward (model/codebook.py:15-32: NHWC copy, d = |z|^2 + |e|^2 - 2 z.e^T, argmin, embedding
// gather, NCHW copy) and Indexer.embedding + Codebook.inference_lr (model/indexer.py:47,53 / 96,100 and
// model/codebook.py:34-43: Linear(512->1024) + softmax + top-1 + gather).  The [rows x K] distance / logit matrix
// never reaches HBM: it lives in TMEM and is reduced by the GEMM epilogue.
//
// Exactness.  tcgen05 has no fp32-input MMA, so the GEMM runs ONE bf16 pass and the epilogue keeps,
// per row, every code whose approximate score is within a rigorous error margin of the best one:
//     |z.e - bf16(z).bf16(e)| <= (2^-8 + 2^-18) |z|_2 |e|_2                 (two roundings of 2^-9 each, Cauchy-Schwarz)
// Rows with a single survivor are decided in the epilogue; the others (their candidate lists are tiny) are re-scored on
// CUDA cores with fp32 FMA accumulation in the reference's association (|z|^2 + |e_k|^2) - 2 z.e_k (codebook.py:19-21),
// lowest index on ties (codebook.py:23).  The result is therefore the arg-min of fp32-accumulated distances.
//
// Pipeline of one call (all on the caller's stream, no host sync):
//   vq_prep_codes   e fp32 [K,D]   -> K8-blocked bf16 B operand, c_k (= |e_k|^2 or -bias_k), max |e_k|
//   vq_prep_rows    z fp32 NCHW    -> K8-blocked bf16 A operand (the NCHW->NHWC permute of :16 fused with the bf16
//                                     conversion), |z|^2 and the candidate margin per row; ONE pass over z
//   gemm_kernel<EpiArgExtremum>    -> a CTA owns 128 rows and sweeps all code tiles; the epilogue thread of a row carries
//                                     the running minimum and a <= 8 entry candidate list across the tiles
//   vq_rescore      32-row blocks: ambiguous rows only; fp32 z staged transposed in smem chunk by chunk, (row, code) pairs
//                   re-scored by warps round-robin -> final int64 indices
//   vq_gather       e[idx] -> NCHW z_q through a smem transpose (coalesced code reads and coalesced stores)
#include "capi_common.h"
#include "gemm_core.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace {

constexpr int CMAX = 32;             // candidate-list slots per row (two sub-lists of CMAX/2, one per epilogue warp of the
                                     // row); more appends -> exact scan of every code for that row
constexpr int CSUB = CMAX / 2;
constexpr int VQ_BLOCK_N = 256;
constexpr int VQ_BLOCK_K = 32;       // k-chunk of the GEMM stages; the operands are padded to a multiple of it
constexpr int VQ_STAGES = 5;
constexpr uint32_t OVERFLOW = 0xFFFFFFFFu;
constexpr int FIN_ROWS = 32;         // rows per finalize block
constexpr int FIN_THREADS = 512;

inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

struct Workspace {
  __nv_bfloat16* a;       // [d_pad/8][rows_pad][8]
  __nv_bfloat16* b;       // [k_pad/256][d_pad/8][256][8] (tiled: one contiguous block per GEMM stage)
  float* c;               // [k_pad]   |e_k|^2 (or -bias_k); +inf for padding codes
  float* zz;              // [rows_pad]
  float* margin;          // [rows_pad]
  uint32_t* cand_cnt;     // [rows_pad][2] appended candidates per (row, column half), OVERFLOW when > CSUB
  uint2* cand;            // [rows_pad][CMAX] (code, approx score), ascending code order
  float* runmin;          // [rows_pad][2] approximate minimum over each column half
  float* emax;            // [2] max_k |e_k|_2 and max_k |e_k - bf16(e_k)|_2 (float bits, written with atomicMax)
  int* err;               // [1] GEMM pipeline error flag
  size_t bytes;
};

Workspace carve(void* base, long long rows, int d, int k) {
  const long long rows_pad = round_up(std::max<long long>(rows, 1), gemm::BLOCK_M);
  const long long d_pad = round_up(d + 3, VQ_BLOCK_K), k_pad = round_up(k, VQ_BLOCK_N);      // + 3: the folded constant's k slots
  uintptr_t p = (uintptr_t)base;
  auto take = [&](size_t n) { uintptr_t r = p; p += (uintptr_t)round_up((long long)n, 256); return (void*)r; };
  Workspace w;
  w.a = (__nv_bfloat16*)take((size_t)rows_pad * d_pad * 2);
  w.b = (__nv_bfloat16*)take((size_t)k_pad * d_pad * 2);
  w.c = (float*)take((size_t)k_pad * 4);
  w.zz = (float*)take((size_t)rows_pad * 4);
  w.margin = (float*)take((size_t)rows_pad * 4);
  w.cand_cnt = (uint32_t*)take((size_t)rows_pad * 2 * 4);
  w.cand = (uint2*)take((size_t)rows_pad * CMAX * 8);
  w.runmin = (float*)take((size_t)rows_pad * 2 * 4);
  w.emax = (float*)take(8);
  w.err = (int*)take(4);
  w.bytes = (size_t)(p - (uintptr_t)base);
  return w;
}

// ---------------------------------------------------------------------------------------------------------------
// codes: one warp per code k.  c_k = |e_k|^2 (mode 0) or -bias_k (mode 1); B operand cells; max norms.
// The per-code constant is FOLDED INTO THE GEMM: the three k slots after the real data hold t_k = c_k / alpha split into
// three bf16 pieces (hi + lo + lo2 = t_k to 2^-24) and the A operand holds 1.0 there, so the accumulator is
//   acc_k = bf16(z).bf16(e_k) + c_k / alpha,   score_k = alpha * acc_k,   arg-min score = arg-max acc   (alpha < 0)
// and the epilogue needs no per-column constant load or FFMA at all.
__global__ void vq_prep_codes(const float* __restrict__ e, const float* __restrict__ bias, int k, int d, int k_pad, int d_pad,
                              int mode, float alpha, __nv_bfloat16* __restrict__ b, float* __restrict__ c, float* __restrict__ emax) {
  const int code = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (code >= k_pad) return;
  const bool real = code < k;
  float ss = 0.f, sd = 0.f;                 // |e_k|^2 and |e_k - bf16(e_k)|^2
  if (real)
    for (int i = lane; i < d; i += 32) {
      const float v = e[(size_t)code * d + i], r = v - sm100::bf16_round(v);
      ss = fmaf(v, v, ss);
      sd = fmaf(r, r, sd);
    }
#pragma unroll
  for (int o = 16; o; o >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, o); sd += __shfl_xor_sync(0xffffffffu, sd, o); }
  const float ck = real ? (mode == 0 ? ss : -bias[code]) : INFINITY;
  const float t = ck / alpha;                                   // alpha is -2 or -1: exact
  float t_hi = sm100::bf16_round(t), t_lo = 0.f, t_lo2 = 0.f;
  if (real) { t_lo = sm100::bf16_round(t - t_hi); t_lo2 = sm100::bf16_round((t - t_hi) - t_lo); }
  for (int kc = lane; kc < d_pad / 8; kc += 32) {
    uint32_t h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int dd = kc * 8 + 2 * j + u;
        v[u] = dd < d ? (real ? e[(size_t)code * d + dd] : 0.f) : (dd == d ? t_hi : dd == d + 1 ? t_lo : dd == d + 2 ? t_lo2 : 0.f);
      }
      h[j] = sm100::pack_bf16x2(v[0], v[1]);
    }
    // tiled layout [code tile][k-cell][256 codes][8]: every GEMM stage (VQ_BLOCK_K / 8 consecutive cells) is one contiguous copy
    const size_t cell = ((size_t)(code / VQ_BLOCK_N) * (d_pad / 8) + kc) * VQ_BLOCK_N + (code % VQ_BLOCK_N);
    *reinterpret_cast<uint4*>(b + cell * 8) = make_uint4(h[0], h[1], h[2], h[3]);
  }
  if (lane == 0) {
    c[code] = ck;
    if (real) {                                         // non-negative floats order like ints
      atomicMax(reinterpret_cast<int*>(emax), __float_as_int(sqrtf(ss) * 1.0000002f));
      atomicMax(reinterpret_cast<int*>(emax) + 1, __float_as_int(sqrtf(sd) * 1.0000002f));
    }
  }
}

// rows: one thread per row; consecutive threads <-> consecutive hw, so the NCHW reads (stride hw between the 8 values of
// a cell) and the blocked 16-byte cell writes are both coalesced.  One pass over z produces the bf16 operand, |z|^2 and
// the candidate margin.  With dz = z - bf16(z), de_k = e_k - bf16(e_k) the approximate score of code k is off by
//   |alpha| |z.e_k - bf16(z).bf16(e_k)| = |alpha| |dz.e_k + bf16(z).de_k| <= |alpha| (|dz| |e_k| + |bf16(z)| |de_k|)
// (Cauchy-Schwarz; |dz| is measured, not bounded by 2^-9 |z|, which tightens the margin ~1.7x on typical data), so
//   margin = 2 * 1.05 * |alpha| * (|dz| max|e| + |z| (1 + 2^-8) max|de|)   (two scores; x1.05 covers the tensor core's
//            fp32 accumulation, <= 512 * 2^-23 relative)  +  2^-20 * (|z|^2 + max|e|^2 + 1)  (fp32 quantisation of the
//            reference's own (|z|^2 + |e|^2) - 2 z.e and summation-order noise)
__global__ void vq_prep_rows(const float* __restrict__ z, long long rows, long long hw, int d, int d_pad, long long rows_pad,
                             float alpha_abs, const float* __restrict__ emax, __nv_bfloat16* __restrict__ a,
                             float* __restrict__ zz, float* __restrict__ margin) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows_pad) return;
  const bool live = r < rows;
  const long long bi = live ? r / hw : 0, p = live ? r % hw : 0;
  const float* src = z + bi * d * hw + p;
  float s = 0.f, sdz = 0.f;
  for (int kc = 0; kc < d_pad / 8; ++kc) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int dd = kc * 8 + j;
      v[j] = (live && dd < d) ? __ldg(src + (long long)dd * hw) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s = fmaf(v[j], v[j], s);
      const float rj = v[j] - sm100::bf16_round(v[j]);
      sdz = fmaf(rj, rj, sdz);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {                    // the three k slots that multiply the folded constant (see vq_prep_codes)
      const int dd = kc * 8 + j;
      if (dd >= d && dd < d + 3) v[j] = 1.0f;
    }
    *reinterpret_cast<uint4*>(a + ((size_t)kc * rows_pad + r) * 8) =
        make_uint4(sm100::pack_bf16x2(v[0], v[1]), sm100::pack_bf16x2(v[2], v[3]), sm100::pack_bf16x2(v[4], v[5]),
                   sm100::pack_bf16x2(v[6], v[7]));
  }
  if (live) {
    zz[r] = s;
    const float em = emax[0], dem = emax[1];
    const float bound = sqrtf(sdz) * 1.0000002f * em + sqrtf(s) * (1.0f + 0x1p-8f) * dem;
    margin[r] = alpha_abs * 2.0f * 1.05f * bound + 0x1p-20f * (s + em * em + 1.0f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GEMM epilogue.  The accumulator already contains the per-code constant (acc_k = dot_k + c_k / alpha), so with alpha < 0
// the best code is the arg-MAX of the raw accumulators.  Per code tile: pass 1 raises the row's running maximum (one
// FMNMX per column), pass 2 appends every code within margin / |alpha| of it straight to the row's candidate list in
// global memory (a predicated store + counter bump: the lanes of a warp are different rows, so anything heavier would
// serialise).  The list is a superset of the codes within `margin` of the FINAL optimum because the running maximum only
// increases; vq_rescore filters it against the final value.  Scores are stored as alpha * acc ("smaller is better").
// (A single-pass variant with a seeded running maximum was measured: no faster, and its longer lists slow vq_rescore.)
struct EpiArgExtremum {
  const float* margin;
  uint32_t* cand_cnt;      // [rows][2]
  uint2* cand;             // [rows][2][CSUB] (code, score bits)
  float* runmin_out;       // [rows][2]
  long long rows;
  float alpha;

  static constexpr int WARPS = 8;        // two warps per TMEM lane quarter: each row is scanned by two threads, one per
                                         // half of the tile's columns, with its own running maximum and sub-list
  struct State {
    float runmax = -INFINITY;
    uint32_t cnt = 0;
  };

  __device__ __forceinline__ float max32(const uint32_t (&r)[32], float mx) const {
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]))));
    return mx;
  }
  __device__ __forceinline__ void scan32(State& st, const uint32_t (&r)[32], uint32_t code0, float thr, uint2* __restrict__ list) const {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]), v2 = __uint_as_float(r[j + 2]), v3 = __uint_as_float(r[j + 3]);
      if (fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)) >= thr) {
        const float v[4] = {v0, v1, v2, v3};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (v[u] >= thr) {
            if (st.cnt < CSUB) list[st.cnt] = make_uint2(code0 + j + u, __float_as_uint(alpha * v[u]));
            ++st.cnt;
          }
        }
      }
    }
  }

  // TMEM loads are software-pipelined one 32-column chunk ahead of the arithmetic.
  __device__ __forceinline__ void tile(State& st, uint32_t tmem_acc, long long m_tile, int n_tile, int n_tiles, int row, int part) const {
    constexpr int SPAN = VQ_BLOCK_N / 2;                       // columns per epilogue warp
    const long long gr = m_tile * gemm::BLOCK_M + row;
    const bool live = gr < rows;
    const int cbase = part * SPAN;
    const float mg = live ? __ldg(margin + gr) / fabsf(alpha) : 0.f;
    uint2* list = cand + ((size_t)(live ? gr : 0) * 2 + part) * CSUB;
    const uint32_t tm = tmem_acc + cbase;
    uint32_t ra[32], rb[32];
    float mx = st.runmax;
    sm100::tmem_ld_32x32(tm, ra);
#pragma unroll 1
    for (int c0 = 0; c0 < SPAN; c0 += 64) {
      sm100::tmem_ld_wait();
      sm100::tmem_ld_32x32(tm + c0 + 32, rb);
      mx = max32(ra, mx);
      sm100::tmem_ld_wait();
      sm100::tmem_ld_32x32(tm + ((c0 + 64) & (SPAN - 1)), ra);       // wraps to the first column for pass 2
      mx = max32(rb, mx);
    }
    st.runmax = mx;
    const float thr = live ? mx - mg : INFINITY;       // padding rows never append
    const uint32_t code_base = (uint32_t)(n_tile * VQ_BLOCK_N + cbase);
#pragma unroll 1
    for (int c0 = 0; c0 < SPAN; c0 += 64) {
      sm100::tmem_ld_wait();
      sm100::tmem_ld_32x32(tm + c0 + 32, rb);
      scan32(st, ra, code_base + c0, thr, list);
      sm100::tmem_ld_wait();
      if (c0 + 64 < SPAN) sm100::tmem_ld_32x32(tm + c0 + 64, ra);
      scan32(st, rb, code_base + c0 + 32, thr, list);
    }
    if (n_tile == n_tiles - 1 && live) {
      cand_cnt[gr * 2 + part] = st.cnt > CSUB ? OVERFLOW : st.cnt;
      runmin_out[gr * 2 + part] = alpha * mx;
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// vq_rescore: a block owns FIN_ROWS consecutive rows of one image (consecutive hw) and decides their indices.
//   mode 0 (Codebook.forward):  d_k = (|z|^2 + |e_k|^2) - 2 * dot_k, minimise, lowest index on ties
//   mode 1 (inference_lr):      l_k = dot_k + bias_k, maximise, lowest index on ties
// (0) the warps filter each row's candidate sub-lists (one coalesced read per row) against the row's final minimum; rows with one survivor are decided,
// the others contribute (row, code) pairs to a block-wide work list (blocks without pairs exit without touching z);
// (1) the fp32 z tile is staged transposed through shared memory in 128-channel chunks (coalesced reads along hw) and
// warps take pairs round-robin, one coalesced fp32 partial dot product per (pair, chunk), so the load is balanced
// whatever the mix of rows; (2) the exact scores pick the winner.  The small per-chunk tile keeps many blocks resident.
__device__ __forceinline__ float score_of(float dot, int mode, float zz, float ck) {
  if (mode == 0) return __fsub_rn(__fadd_rn(zz, ck), __fmul_rn(2.0f, dot));
  return -__fadd_rn(dot, -ck);          // ck = -bias_k ; negated logit so that "smaller is better"
}

constexpr int FIN_MAXPAIRS = FIN_ROWS * CMAX;
constexpr int FIN_DCH = 128;                     // channels staged per chunk

__global__ void __launch_bounds__(FIN_THREADS)
vq_rescore(const float* __restrict__ z, const float* __restrict__ w, long long hw, int blocks_per_image, int d, int k, int mode,
           const float* __restrict__ c, const float* __restrict__ zz, const float* __restrict__ margin,
           const float* __restrict__ runmin, const uint32_t* __restrict__ cand_cnt, const uint2* __restrict__ cand,
           long long* __restrict__ idx_out) {
  constexpr int NW = FIN_THREADS / 32;
  constexpr int LD = FIN_ROWS + 1;
  __shared__ float zt[FIN_DCH * LD];
  __shared__ int s_idx[FIN_ROWS];
  __shared__ int s_first[FIN_ROWS + 1];          // first pair of each row (prefix sum)
  __shared__ uint32_t s_pair[FIN_MAXPAIRS];      // (row << 16) | code
  __shared__ float s_score[FIN_MAXPAIRS];        // fp32 dot products, accumulated chunk by chunk
  __shared__ int s_over[FIN_ROWS];
  __shared__ int s_nover, s_npairs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bi = blockIdx.x / blocks_per_image;
  const long long p0 = (long long)(blockIdx.x % blocks_per_image) * FIN_ROWS;
  const int nrows = (int)min((long long)FIN_ROWS, hw - p0);
  const long long row0 = bi * hw + p0;

  // ---- step 0: every warp filters two rows; a row's two 16-entry sub-lists are one coalesced 256-byte read
  static_assert(CSUB == 16, "one candidate entry per lane");
  __shared__ uint16_t s_tmp[FIN_ROWS][2 * CSUB];
  __shared__ int s_cnt[FIN_ROWS];
  __shared__ unsigned s_overmask;
  if (threadIdx.x == 0) s_overmask = 0;
  __syncthreads();
  for (int r = warp; r < FIN_ROWS; r += NW) {
    int cnt = 0;
    if (r < nrows) {
      const long long gr = row0 + r;
      const uint32_t n0 = cand_cnt[gr * 2], n1 = cand_cnt[gr * 2 + 1];
      if (n0 == OVERFLOW || n1 == OVERFLOW) {
        if (lane == 0) atomicOr(&s_overmask, 1u << r);
      } else {
        const float thr = fminf(runmin[gr * 2], runmin[gr * 2 + 1]) + margin[gr];
        const uint2 e = cand[(size_t)gr * 2 * CSUB + lane];
        const bool keep = (uint32_t)(lane & (CSUB - 1)) < (lane < CSUB ? n0 : n1) && __uint_as_float(e.y) <= thr;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        cnt = __popc(m);
        if (keep) s_tmp[r][__popc(m & ((1u << lane) - 1))] = (uint16_t)e.x;
        if (cnt <= 1) {                                                   // decided (0 survivors only for NaN rows)
          const int only = __shfl_sync(0xffffffffu, (int)e.x, m ? __ffs(m) - 1 : 0);
          if (lane == 0) s_idx[r] = cnt ? only : 0;
        }
      }
    }
    if (lane == 0) s_cnt[r] = cnt >= 2 ? cnt : 0;
  }
  __syncthreads();
  if (warp == 0) {
    const int np = s_cnt[lane];
    int incl = np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    s_first[lane] = incl - np;
    if (lane == 31) { s_first[32] = incl; s_npairs = incl; }
    const unsigned om = s_overmask;
    if ((om >> lane) & 1u) s_over[__popc(om & ((1u << lane) - 1))] = lane;
    if (lane == 0) s_nover = __popc(om);
  }
  __syncthreads();
  for (int r = warp; r < FIN_ROWS; r += NW) {
    const int np = s_cnt[r], f = s_first[r];
    if (lane < np) s_pair[f + lane] = ((uint32_t)r << 16) | s_tmp[r][lane];
  }
  __syncthreads();
  const int npairs = s_npairs, nover = s_nover;

  if (npairs > 0) {
    const float* zb = z + bi * d * hw + p0 + (lane < nrows ? lane : 0);
    constexpr int PER = FIN_DCH / NW;
    float v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) { const int dd = warp + u * NW; v[u] = dd < d ? __ldg(zb + (long long)dd * hw) : 0.f; }
    for (int d0 = 0; d0 < d; d0 += FIN_DCH) {
      const int dn = min(FIN_DCH, d - d0);
      // stage z[d0 .. d0+dn) x 32 rows, transposed; the NEXT chunk's loads are issued before the dot products below
#pragma unroll
      for (int u = 0; u < PER; ++u) zt[(warp + u * NW) * LD + lane] = v[u];
      __syncthreads();
      if (d0 + FIN_DCH < d) {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int dd = d0 + FIN_DCH + warp + u * NW;
          v[u] = dd < d ? __ldg(zb + (long long)dd * hw) : 0.f;
        }
      }
      for (int p = warp; p < npairs; p += NW) {
        const uint32_t pr = s_pair[p];
        const int r = (int)(pr >> 16), kk = (int)(pr & 0xFFFFu);
        const float* wk = w + (size_t)kk * d + d0;
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < FIN_DCH / 32; ++u) {
          const int dd = lane + 32 * u;
          acc = fmaf(zt[dd * LD + r], dd < dn ? __ldg(wk + dd) : 0.f, acc);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_score[p] = d0 ? s_score[p] + acc : acc;
      }
      __syncthreads();
    }
    if (threadIdx.x < nrows) {
      const int f = s_first[threadIdx.x], l = s_first[threadIdx.x + 1];
      if (l > f) {
        const float zzr = zz[row0 + threadIdx.x];
        float bs = INFINITY; int best = 0;
        for (int p = f; p < l; ++p) {                                // lowest index wins ties (the two sub-lists interleave)
          const int kk = (int)(s_pair[p] & 0xFFFFu);
          const float sc = score_of(s_score[p], mode, zzr, c[kk]);
          if (sc < bs || (sc == bs && kk < best)) { bs = sc; best = kk; }
        }
        s_idx[threadIdx.x] = best;
      }
    }
  }
  // rows whose sub-lists overflowed (many duplicated / near-tied codes; rare): exact scan of ALL codes, z read in place
  for (int o = 0; o < nover; ++o) {
    __shared__ float s_wbest[NW];
    __shared__ int s_wbestk[NW];
    const int r = s_over[o];
    const float* zr = z + bi * d * hw + p0 + r;
    const float zzr = zz[row0 + r];
    float bs = INFINITY;
    int bk = 0x7fffffff;
    for (int kk = warp; kk < k; kk += NW) {
      float acc = 0.f;
      for (int i = lane; i < d; i += 32) acc = fmaf(__ldg(zr + (long long)i * hw), __ldg(w + (size_t)kk * d + i), acc);
#pragma unroll
      for (int of = 16; of; of >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, of);
      const float sc = score_of(acc, mode, zzr, c[kk]);
      if (sc < bs) { bs = sc; bk = kk; }
    }
    if (lane == 0) { s_wbest[warp] = bs; s_wbestk[warp] = bk; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float b = INFINITY; int bkk = 0x7fffffff;
      for (int wv = 0; wv < NW; ++wv)
        if (s_wbest[wv] < b || (s_wbest[wv] == b && s_wbestk[wv] < bkk)) { b = s_wbest[wv]; bkk = s_wbestk[wv]; }
      s_idx[r] = bkk == 0x7fffffff ? 0 : bkk;
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x < nrows) idx_out[row0 + threadIdx.x] = s_idx[threadIdx.x];
}

// vq_gather: z_q[b, :, hw] = table[idx[b, hw], :] -- the embedding gather and the NHWC->NCHW permute of codebook.py:24,30
// in one pass: code vectors are read coalesced along d, transposed through shared memory and stored coalesced along hw.
// Optionally accumulates sum((z_q - z)^2) for the loss (codebook.py:26).
constexpr int GA_DCH = 128;
__global__ void __launch_bounds__(256)
vq_gather(const float* __restrict__ table, const long long* __restrict__ idx, long long hw, int blocks_per_image, int dq,
          float* __restrict__ zq, const float* __restrict__ z, float* __restrict__ sq_err) {
  constexpr int LD = FIN_ROWS + 1;
  __shared__ float t[GA_DCH * LD];
  __shared__ int s_idx[FIN_ROWS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bi = blockIdx.x / blocks_per_image;
  const long long p0 = (long long)(blockIdx.x % blocks_per_image) * FIN_ROWS;
  const int nrows = (int)min((long long)FIN_ROWS, hw - p0);
  const int d0 = blockIdx.y * GA_DCH, dn = min(GA_DCH, dq - d0);
  if (threadIdx.x < FIN_ROWS) s_idx[threadIdx.x] = threadIdx.x < nrows ? (int)idx[bi * hw + p0 + threadIdx.x] : 0;
  __syncthreads();
#pragma unroll
  for (int rr = 0; rr < FIN_ROWS / 8; ++rr) {
    const int r = warp + 8 * rr;
    const float* src = table + (size_t)s_idx[r] * dq + d0;
    float v[GA_DCH / 32];
#pragma unroll
    for (int u = 0; u < GA_DCH / 32; ++u) v[u] = (lane + 32 * u) < dn ? __ldg(src + lane + 32 * u) : 0.f;
#pragma unroll
    for (int u = 0; u < GA_DCH / 32; ++u) t[(lane + 32 * u) * LD + r] = v[u];
  }
  __syncthreads();
  float err = 0.f;
  if (lane < nrows) {
    float* qb = zq + (bi * dq + d0) * hw + p0 + lane;
    const float* zb = z ? z + (bi * dq + d0) * hw + p0 + lane : nullptr;
    for (int dd = warp; dd < dn; dd += 8) {
      const float v = t[dd * LD + lane];
      __stcs(qb + (long long)dd * hw, v);
      if (zb) { const float dl = v - __ldg(zb + (long long)dd * hw); err = fmaf(dl, dl, err); }
    }
  }
  if (sq_err) {
#pragma unroll
    for (int o = 16; o; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
    if (lane == 0 && err != 0.f) atomicAdd(sq_err, err);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// vq_finish: vq_rescore + vq_gather in ONE kernel for 16-byte friendly shapes (d % 4 == 0, dq % 4 == 0, aligned bases) --
// the path every model shape takes.  Same decisions as vq_rescore, with a much leaner inner loop:
//   * the block's whole fp32 z tile [32 rows][<= 512 channels] is staged ROW-major in shared memory (coalesced reads along
//     hw, float4 stores; row stride 516 floats keeps every quarter-warp float4 access conflict-free), so a (row, code)
//     pair costs 4 x (LDS.128 + LDG.128 + 4 FFMA) and ONE warp reduction instead of 16 scalar triples and 4 reductions;
//   * the winners' code vectors are then transposed through the same buffer and stored along hw (NCHW), so the indices
//     never round-trip through HBM and the tile's z (needed by the optional loss) is still in L2.
constexpr int FZ_D = 512;                        // channels per staged tile
constexpr int FZ_LD = FZ_D + 4;
constexpr int FZ_SMEM = FIN_ROWS * FZ_LD * 4;

__global__ void __launch_bounds__(FIN_THREADS, 3)
vq_finish(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ table, long long hw,
          int blocks_per_image, int d, int k, int dq, int mode, const float* __restrict__ c, const float* __restrict__ zz,
          const float* __restrict__ margin, const float* __restrict__ runmin, const uint32_t* __restrict__ cand_cnt,
          const uint2* __restrict__ cand, long long* __restrict__ idx_out, float* __restrict__ zq, float* __restrict__ sq_err) {
  constexpr int NW = FIN_THREADS / 32;
  static_assert(NW == 16 && FIN_ROWS == 32 && CSUB == 16, "thread mapping below");
  extern __shared__ __align__(16) float zt[];                  // [FIN_ROWS][FZ_LD]; first used as the step-0 scratch
  __shared__ int s_idx[FIN_ROWS];
  __shared__ int s_first[FIN_ROWS + 1];
  __shared__ uint32_t s_pair[FIN_MAXPAIRS];      // (row << 16) | code
  __shared__ float s_score[FIN_MAXPAIRS];
  __shared__ int s_over[FIN_ROWS];
  __shared__ int s_cnt[FIN_ROWS];
  __shared__ int s_nover, s_npairs;
  __shared__ unsigned s_overmask;
  uint16_t (*s_tmp)[2 * CSUB] = reinterpret_cast<uint16_t (*)[2 * CSUB]>(zt);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bi = blockIdx.x / blocks_per_image;
  const long long p0 = (long long)(blockIdx.x % blocks_per_image) * FIN_ROWS;
  const int nrows = (int)min((long long)FIN_ROWS, hw - p0);
  const long long row0 = bi * hw + p0;

  if (threadIdx.x == 0) s_overmask = 0;
  if (threadIdx.x < FIN_ROWS) s_idx[threadIdx.x] = 0;
  __syncthreads();
  for (int r = warp; r < FIN_ROWS; r += NW) {      // step 0: filter the candidate sub-lists against the final minimum
    int cnt = 0;
    if (r < nrows) {
      const long long gr = row0 + r;
      const uint32_t n0 = cand_cnt[gr * 2], n1 = cand_cnt[gr * 2 + 1];
      if (n0 == OVERFLOW || n1 == OVERFLOW) {
        if (lane == 0) atomicOr(&s_overmask, 1u << r);
      } else {
        const float thr = fminf(runmin[gr * 2], runmin[gr * 2 + 1]) + margin[gr];
        const uint2 e = cand[(size_t)gr * 2 * CSUB + lane];
        const bool keep = (uint32_t)(lane & (CSUB - 1)) < (lane < CSUB ? n0 : n1) && __uint_as_float(e.y) <= thr;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        cnt = __popc(m);
        if (keep) s_tmp[r][__popc(m & ((1u << lane) - 1))] = (uint16_t)e.x;
        if (cnt <= 1) {
          const int only = __shfl_sync(0xffffffffu, (int)e.x, m ? __ffs(m) - 1 : 0);
          if (lane == 0) s_idx[r] = cnt ? only : 0;
        }
      }
    }
    if (lane == 0) s_cnt[r] = cnt >= 2 ? cnt : 0;
  }
  __syncthreads();
  if (warp == 0) {
    const int np = s_cnt[lane];
    int incl = np;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    s_first[lane] = incl - np;
    if (lane == 31) { s_first[32] = incl; s_npairs = incl; }
    const unsigned om = s_overmask;
    if ((om >> lane) & 1u) s_over[__popc(om & ((1u << lane) - 1))] = lane;
    if (lane == 0) s_nover = __popc(om);
  }
  __syncthreads();
  for (int r = warp; r < FIN_ROWS; r += NW) {
    const int np = s_cnt[r], f = s_first[r];
    if (lane < np) s_pair[f + lane] = ((uint32_t)r << 16) | s_tmp[r][lane];
  }
  __syncthreads();
  const int npairs = s_npairs, nover = s_nover;

  // rows whose sub-lists overflowed (many near-tied codes: the tail of the append-count distribution) are scanned against
  // ALL codes; their partial dot products live in s_ovdot and are accumulated tile by tile like the pair scores
  __shared__ float s_ovbest[NW];
  __shared__ int s_ovbestk[NW];
  if (npairs > 0 || nover > 0) {
    for (int d0 = 0; d0 < d; d0 += FZ_D) {
      const int dn = min(FZ_D, d - d0);
      const float* zb = z + (bi * d + d0) * hw + p0 + (lane < nrows ? lane : 0);
      if (d0) __syncthreads();
      // stage the tile: thread (lane = row, quad q) loads 4 channels along coalesced hw lines and stores one float4.
      // (4-byte cp.async copies straight into a transposed tile were measured: 0.5 ms slower.)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int dd = 4 * (warp + NW * (u + 4 * half));
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (dd < dn) {
            const float* q = zb + (long long)dd * hw;
            v[u].x = __ldg(q); v[u].y = __ldg(q + hw); v[u].z = __ldg(q + 2 * hw); v[u].w = __ldg(q + 3 * hw);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<float4*>(zt + lane * FZ_LD + 4 * (warp + NW * (u + 4 * half))) = v[u];
      }
      __syncthreads();
      for (int p = warp; p < npairs; p += NW) {
        const uint32_t pr = s_pair[p];
        const int r = (int)(pr >> 16), kk = (int)(pr & 0xFFFFu);
        const float* wk = w + (size_t)kk * d + d0;
        const float* zr = zt + r * FZ_LD;
        float acc = 0.f;
#pragma unroll
        for (int u = 0; u < FZ_D / 128; ++u) {
          const int dd = 128 * u + 4 * lane;
          if (dd < dn) {
            const float4 a = *reinterpret_cast<const float4*>(zr + dd);
            const float4 b = __ldg(reinterpret_cast<const float4*>(wk + dd));
            acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_score[p] = d0 ? s_score[p] + acc : acc;
      }
      if (d <= FZ_D) {                      // single tile (every model shape): overflow rows scan all codes from the staged tile
        for (int o = 0; o < nover; ++o) {
          const int r = s_over[o];
          const float zzr = zz[row0 + r];
          const float* zr = zt + r * FZ_LD;
          float bs = INFINITY;
          int bk = 0x7fffffff;
          for (int kk = warp; kk < k; kk += NW) {          // ascending per warp: the first minimum is the lowest index
            const float* wk = w + (size_t)kk * d;
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < FZ_D / 128; ++u) {
              const int dd = 128 * u + 4 * lane;
              if (dd < dn) {
                const float4 a4 = *reinterpret_cast<const float4*>(zr + dd);
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(wk + dd));
                acc = fmaf(a4.x, b4.x, acc); acc = fmaf(a4.y, b4.y, acc); acc = fmaf(a4.z, b4.z, acc); acc = fmaf(a4.w, b4.w, acc);
              }
            }
#pragma unroll
            for (int of = 16; of; of >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, of);
            const float sc = score_of(acc, mode, zzr, c[kk]);
            if (sc < bs) { bs = sc; bk = kk; }
          }
          if (lane == 0) { s_ovbest[warp] = bs; s_ovbestk[warp] = bk; }
          __syncthreads();
          if (threadIdx.x == 0) {
            float b = INFINITY; int bkk = 0x7fffffff;
            for (int wv = 0; wv < NW; ++wv)
              if (s_ovbest[wv] < b || (s_ovbest[wv] == b && s_ovbestk[wv] < bkk)) { b = s_ovbest[wv]; bkk = s_ovbestk[wv]; }
            s_idx[r] = bkk == 0x7fffffff ? 0 : bkk;
          }
          __syncthreads();
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < nrows) {
      const int f = s_first[threadIdx.x], l = s_first[threadIdx.x + 1];
      if (l > f) {
        const float zzr = zz[row0 + threadIdx.x];
        float bs = INFINITY; int best = 0;
        for (int p = f; p < l; ++p) {                                // lowest index wins ties (the two sub-lists interleave)
          const int kk = (int)(s_pair[p] & 0xFFFFu);
          const float sc = score_of(s_score[p], mode, zzr, c[kk]);
          if (sc < bs || (sc == bs && kk < best)) { bs = sc; best = kk; }
        }
        s_idx[threadIdx.x] = best;
      }
    }
  }
  // multi-tile channel counts (d > FZ_D; no model shape): overflow rows read z in place
  for (int o = 0; o < (d > FZ_D ? nover : 0); ++o) {
    __shared__ float s_wbest[NW];
    __shared__ int s_wbestk[NW];
    const int r = s_over[o];
    const float* zr = z + bi * d * hw + p0 + r;
    const float zzr = zz[row0 + r];
    float bs = INFINITY;
    int bk = 0x7fffffff;
    for (int kk = warp; kk < k; kk += NW) {
      float acc = 0.f;
      for (int i = lane; i < d; i += 32) acc = fmaf(__ldg(zr + (long long)i * hw), __ldg(w + (size_t)kk * d + i), acc);
#pragma unroll
      for (int of = 16; of; of >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, of);
      const float sc = score_of(acc, mode, zzr, c[kk]);
      if (sc < bs) { bs = sc; bk = kk; }
    }
    if (lane == 0) { s_wbest[warp] = bs; s_wbestk[warp] = bk; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float b = INFINITY; int bkk = 0x7fffffff;
      for (int wv = 0; wv < NW; ++wv)
        if (s_wbest[wv] < b || (s_wbest[wv] == b && s_wbestk[wv] < bkk)) { b = s_wbest[wv]; bkk = s_wbestk[wv]; }
      s_idx[r] = bkk == 0x7fffffff ? 0 : bkk;
    }
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x < nrows) idx_out[row0 + threadIdx.x] = s_idx[threadIdx.x];

  // ---- gather: z_q[bi, :, p0 + r] = table[idx[r], :]
  float err = 0.f;
  for (int q0 = 0; q0 < dq; q0 += FZ_D) {
    const int dn = min(FZ_D, dq - q0);
    if (q0) __syncthreads();
#pragma unroll
    for (int rr = 0; rr < FIN_ROWS / NW; ++rr) {
      const int r = warp + NW * rr;
      const float* src = table + (size_t)s_idx[r] * dq + q0;
#pragma unroll
      for (int u = 0; u < FZ_D / 128; ++u) {
        const int dd = 128 * u + 4 * lane;
        if (dd < dn) *reinterpret_cast<float4*>(zt + r * FZ_LD + dd) = __ldg(reinterpret_cast<const float4*>(src + dd));
      }
    }
    __syncthreads();
    if (lane < nrows) {
      float* qb = zq + (bi * dq + q0) * hw + p0 + lane;
      const float* zb = sq_err ? z + (bi * dq + q0) * hw + p0 + lane : nullptr;
#pragma unroll
      for (int u = 0; u < FZ_D / (4 * NW); ++u) {
        const int dd = 4 * (warp + NW * u);
        if (dd < dn) {
          const float4 t = *reinterpret_cast<const float4*>(zt + lane * FZ_LD + dd);
          float* q = qb + (long long)dd * hw;
          __stcs(q, t.x); __stcs(q + hw, t.y); __stcs(q + 2 * hw, t.z); __stcs(q + 3 * hw, t.w);
          if (zb) {
            const float* zs = zb + (long long)dd * hw;
            const float e0 = t.x - __ldg(zs), e1 = t.y - __ldg(zs + hw), e2 = t.z - __ldg(zs + 2 * hw), e3 = t.w - __ldg(zs + 3 * hw);
            err = fmaf(e0, e0, err); err = fmaf(e1, e1, err); err = fmaf(e2, e2, err); err = fmaf(e3, e3, err);
          }
        }
      }
    }
  }
  if (sq_err) {
#pragma unroll
    for (int o = 16; o; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
    if (lane == 0 && err != 0.f) atomicAdd(sq_err, err);
  }
}

// argmax over materialised logits + gather (Codebook.inference_lr on its own): one warp per row
__global__ void argmax_gather_kernel(const float* __restrict__ p, const float* __restrict__ table, long long rows, long long hw,
                                     int k, int dq, float* __restrict__ zq, long long* __restrict__ idx_out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* row = p + r * k;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < k; i += 32) {
    const float v = __ldg(row + i);
    if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;
  if (lane == 0) idx_out[r] = bi;
  const long long b = r / hw, q = r % hw;
  for (int i = lane; i < dq; i += 32) zq[(b * dq + i) * hw + q] = __ldg(table + (size_t)bi * dq + i);
}

int run_lookup(const float* z, const float* w, const float* bias, const float* table, int b, int d, long long hw, int k, int dq,
               int mode, float* zq, long long* idx, float* sq_err_sum, void* ws, size_t ws_bytes, cudaStream_t s) {
  using namespace gpemsr;
  if (b < 0 || d <= 0 || hw < 0 || k <= 0 || dq <= 0)
    return set_error(GPEMSR_ERR_BAD_SHAPE, "vq lookup: bad shape b=%d d=%d hw=%lld k=%d dq=%d", b, d, hw, k, dq);
  if (k > 65535) return set_error(GPEMSR_ERR_BAD_SHAPE, "vq lookup: more than 65535 codes");
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  const long long rows = (long long)b * hw;
  if (rows == 0) return GPEMSR_OK;
  if (!z || !w || !table || !zq || !idx) return set_error(GPEMSR_ERR_BAD_SHAPE, "vq lookup: null pointer");
  if (mode == 1 && !bias) return set_error(GPEMSR_ERR_BAD_SHAPE, "logits_argmax_gather: bias is required");
  const size_t need = gpemsr_vq_workspace_bytes(rows, d, k);
  if (!ws || ws_bytes < need) return set_error(GPEMSR_ERR_WORKSPACE, "vq lookup: workspace %zu B < required %zu B", ws_bytes, need);
  if (reinterpret_cast<uintptr_t>(ws) & 255) return set_error(GPEMSR_ERR_BAD_ALIGN, "vq lookup: workspace must be 256-byte aligned");

  Workspace W = carve(ws, rows, d, k);
  const long long rows_pad = round_up(rows, gemm::BLOCK_M);
  const int d_pad = (int)round_up(d + 3, VQ_BLOCK_K), k_pad = (int)round_up(k, VQ_BLOCK_N), n_tiles = k_pad / VQ_BLOCK_N;
  const float alpha = mode == 0 ? -2.0f : -1.0f;

  GPEMSR_CUDA_OK(cudaMemsetAsync(W.emax, 0, 512, s));       // emax and err (adjacent 256-byte slots)
  vq_prep_codes<<<(k_pad + 7) / 8, 256, 0, s>>>(w, bias, k, d, k_pad, d_pad, mode, alpha, W.b, W.c, W.emax);
  GPEMSR_LAUNCH_OK("vq_prep_codes");
  vq_prep_rows<<<(unsigned)((rows_pad + 127) / 128), 128, 0, s>>>(z, rows, hw, d, d_pad, rows_pad, fabsf(alpha), W.emax, W.a,
                                                                W.zz, W.margin);
  GPEMSR_LAUNCH_OK("vq_prep_rows");
  {
    gemm::Operands op{};
    op.a_hi = W.a; op.a_lo = nullptr; op.b_hi = W.b; op.b_lo = nullptr;
    op.a_rows = rows_pad; op.b_rows = k_pad; op.b_packed = 1; op.k = d_pad; op.taps = 1; op.a_row_off[0] = 0;
    op.m_tiles = rows_pad / gemm::BLOCK_M; op.n_tiles = n_tiles; op.a_row0 = 0; op.err_flag = W.err;
    EpiArgExtremum epi{W.margin, W.cand_cnt, W.cand, W.runmin, rows, alpha};
    using Cfg = gemm::Config<VQ_BLOCK_N, VQ_BLOCK_K, 1, 6>;
    const bool pair = op.m_tiles >= 2 && use_clusters();   // CTA pairs share every codebook stage by multicast (halves L2 traffic)
    const int ares_smem = gemm::ares_smem_bytes<VQ_BLOCK_N, VQ_BLOCK_K, VQ_STAGES>(d_pad);
    const bool a_resident = ares_smem <= 227 * 1024 && d_pad / VQ_BLOCK_K <= gemm::ARES_MAX_CHUNKS && !getenv("GPEMSR_VQ_STREAM_A");
    if (a_resident && pair) {
      // the [128 x d_pad] A tile stays in shared memory for all code tiles (gemm_ares_kernel): A leaves L2 once, not 4 times
      auto kern = gemm::gemm_ares_kernel<VQ_BLOCK_N, VQ_BLOCK_K, VQ_STAGES, EpiArgExtremum, 2>;
      GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ares_smem));
      const int grid = (int)((std::min<long long>(op.m_tiles, num_sms()) + 1) / 2 * 2);
      GPEMSR_CUDA_OK(launch_cluster(kern, dim3(grid), dim3(gemm::num_threads_ares<EpiArgExtremum>()), ares_smem, s, 2, op, epi));
    } else if (a_resident) {
      auto kern = gemm::gemm_ares_kernel<VQ_BLOCK_N, VQ_BLOCK_K, VQ_STAGES, EpiArgExtremum, 1>;
      GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ares_smem));
      const int grid = (int)std::min<long long>(op.m_tiles, num_sms());
      kern<<<grid, gemm::num_threads_ares<EpiArgExtremum>(), ares_smem, s>>>(op, epi);
    } else if (pair) {
      auto kern = gemm::gemm_kernel<VQ_BLOCK_N, VQ_BLOCK_K, 1, 6, EpiArgExtremum, 2>;
      GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      const int grid = (int)((std::min<long long>(op.m_tiles, num_sms()) + 1) / 2 * 2);
      GPEMSR_CUDA_OK(launch_cluster(kern, dim3(grid), dim3(gemm::num_threads<EpiArgExtremum>()), Cfg::SMEM_BYTES, s, 2, op, epi));
    } else {
      auto kern = gemm::gemm_kernel<VQ_BLOCK_N, VQ_BLOCK_K, 1, 6, EpiArgExtremum>;
      GPEMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
      const int grid = (int)std::min<long long>(op.m_tiles, num_sms());
      kern<<<grid, gemm::num_threads<EpiArgExtremum>(), Cfg::SMEM_BYTES, s>>>(op, epi);
    }
    GPEMSR_LAUNCH_OK("gemm_kernel<EpiArgExtremum>");
  }
  {
    const int blocks_per_image = (int)((hw + FIN_ROWS - 1) / FIN_ROWS);
    const bool want_err = mode == 0 && sq_err_sum != nullptr;
    if (want_err) GPEMSR_CUDA_OK(cudaMemsetAsync(sq_err_sum, 0, sizeof(float), s));
    const bool vec_ok = d % 4 == 0 && dq % 4 == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(table)) & 15) == 0;
    if (vec_ok) {
      GPEMSR_CUDA_OK(cudaFuncSetAttribute(vq_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM));
      vq_finish<<<(unsigned)((long long)b * blocks_per_image), FIN_THREADS, FZ_SMEM, s>>>(
          z, w, table, hw, blocks_per_image, d, k, dq, mode, W.c, W.zz, W.margin, W.runmin, W.cand_cnt, W.cand, idx, zq,
          want_err ? sq_err_sum : nullptr);
      GPEMSR_LAUNCH_OK("vq_finish");
    } else {                                   // odd channel counts / unaligned tables: the scalar two-kernel form
      vq_rescore<<<(unsigned)((long long)b * blocks_per_image), FIN_THREADS, 0, s>>>(
          z, w, hw, blocks_per_image, d, k, mode, W.c, W.zz, W.margin, W.runmin, W.cand_cnt, W.cand, idx);
      GPEMSR_LAUNCH_OK("vq_rescore");
      dim3 grid((unsigned)((long long)b * blocks_per_image), (unsigned)((dq + GA_DCH - 1) / GA_DCH));
      vq_gather<<<grid, 256, 0, s>>>(table, idx, hw, blocks_per_image, dq, zq, want_err ? z : nullptr, want_err ? sq_err_sum : nullptr);
      GPEMSR_LAUNCH_OK("vq_gather");
    }
  }
  return GPEMSR_OK;
}

}  // namespace

extern "C" {

size_t gpemsr_vq_workspace_bytes(int64_t n_rows, int d, int k) {
  if (n_rows < 0 || d <= 0 || k <= 0) return 0;
  return carve(nullptr, n_rows, d, k).bytes;
}

int gpemsr_vq_lookup_nchw(const float* z, const float* emb, int b, int d, int64_t hw, int k, float* zq, int64_t* idx,
                          float* sq_err_sum, void* ws, size_t ws_bytes, gpemsr_stream_t stream) {
  return run_lookup(z, emb, nullptr, emb, b, d, hw, k, d, 0, zq, (long long*)idx, sq_err_sum, ws, ws_bytes, (cudaStream_t)stream);
}

int gpemsr_logits_argmax_gather(const float* feat, const float* w, const float* bias, const float* emb, int b, int d, int64_t hw,
                                int k, int dq, float* zq, int64_t* idx, float* logits, void* ws, size_t ws_bytes,
                                gpemsr_stream_t stream) {
  if (logits)
    return gpemsr::set_error(GPEMSR_ERR_UNSUPPORTED, "logits_argmax_gather: materialising the logits is not built yet; pass "
                             "logits = NULL (the fused path never writes them)");
  return run_lookup(feat, w, bias, emb, b, d, hw, k, dq, 1, zq, (long long*)idx, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

int gpemsr_argmax_gather(const float* p, const float* emb, int b, int64_t hw, int k, int dq, float* zq, int64_t* idx,
                         gpemsr_stream_t stream) {
  using namespace gpemsr;
  if (b < 0 || hw < 0 || k <= 0 || dq <= 0) return set_error(GPEMSR_ERR_BAD_SHAPE, "argmax_gather: bad shape");
  int rc = check_device_current();
  if (rc != GPEMSR_OK) return rc;
  const long long rows = (long long)b * hw;
  if (rows == 0) return GPEMSR_OK;
  if (!p || !emb || !zq || !idx) return set_error(GPEMSR_ERR_BAD_SHAPE, "argmax_gather: null pointer");
  argmax_gather_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p, emb, rows, hw, k, dq, zq, (long long*)idx);
  GPEMSR_LAUNCH_OK("argmax_gather_kernel");
  return GPEMSR_OK;
}

}  // extern "C"
