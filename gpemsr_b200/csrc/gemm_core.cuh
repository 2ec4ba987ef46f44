// Warp-specialised tcgen05 GEMM core shared by the VQ lookup, the implicit-GEMM convolutions and the
// non-local attention:   D[m, n] = sum_t sum_k  A[m + a_row_off[t], k] * B_t[n, k]      (fp32 accumulate in TMEM)
//
// Operand format ("K8-blocked", bf16): a matrix X[rows, K] is stored as  [K/8][rows_alloc][8]  i.e. for every group
// of 8 consecutive k a column of 16-byte row cells.  A [rows x 8k] slab is therefore ONE contiguous run in HBM
// (fetched by one cp.async.bulk, no tensor map needed) and lands in shared memory exactly in the tcgen05
// K-major SWIZZLE_NONE canonical layout (core matrix = 8 rows x 16 B, SBO = 128 B, LBO = rows*16 B).
// Because rows are 16-byte cells at a uniform stride, a row SHIFT of the A operand is just a different start
// address -- that is what turns a 3x3 convolution over a zero-padded, flattened image into 9 shifted GEMMs
// with no im2col buffer (conv_igemm.cu).
//
// Precision: SPLIT == 1 -> one bf16 pass.  SPLIT == 3 -> operands carry (hi, lo) bf16 planes with
// hi = bf16(x), lo = bf16(x - hi) and three MMAs hi*hi + lo*hi + hi*lo are accumulated (error ~2^-16 relative):
// fp32-faithful convolutions on the bf16 tensor pipe at 1/3 of its rate (still ~5x the fp32 CUDA-core rate).
//
// Roles (192 threads): warp 0 = bulk-copy producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue
// (each owns the TMEM lane quarter (warp % 4)).  Pipelines: NSTAGE smem stages (full/empty mbarriers) and two
// TMEM accumulator buffers (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Persistent: CTA b owns the 128-row tiles b, b+grid, ... and sweeps ALL column tiles of each before moving on, so an
// epilogue can carry per-row state across the column tiles of a row (the VQ arg-min) and the A slab stays L2-hot.
// gridDim.y > 1 splits the column tiles across CTAs instead (stateless epilogues on problems with few row tiles).
#pragma once
#include "sm100.cuh"

// Profiling builds only (tools/ablate.py): bit 1 = no epilogue work, 2 = A tiles loaded once, 4 = no MMAs, in the tap-fused
// kernel.  The shipped library is built with 0 and none of the branches exist in it.
#ifndef GPEMSR_ABLATE
#define GPEMSR_ABLATE 0
#endif
#ifndef GPEMSR_L2_PREFETCH
#define GPEMSR_L2_PREFETCH 1
#endif

// bit 8: per-role cycle accounting (clock64 around every barrier wait), printed by CTA 0 and CTA 77 at kernel end
#if GPEMSR_ABLATE & 8
#include <cstdio>
__device__ __forceinline__ unsigned long long prof_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define GPEMSR_PROF_DECL long long prof_w[2] = {0, 0}, prof_work = 0, prof_t0 = 0, prof_t1 = clock64(), prof_n = 0; const long long prof_start = prof_t1; \
    const unsigned long long prof_ns0 = prof_ns();
#define GPEMSR_PROF_T0 prof_t0 = clock64();
#define GPEMSR_PROF_WAIT(i) prof_t1 = clock64(); prof_w[i] += prof_t1 - prof_t0;
#define GPEMSR_PROF_WORK prof_work += clock64() - prof_t1; ++prof_n;
#define GPEMSR_PROF_PRINT(a, b, c) if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77)) \
    printf("cta %3d warp %2d  total %8lld cycles in %7llu ns, %5lld trips | %s %8lld | %s %8lld | %s %8lld\n", (int)blockIdx.x, warp, \
           clock64() - prof_start, prof_ns() - prof_ns0, prof_n, a, prof_w[0], b, prof_w[1], c, prof_work);
#else
#define GPEMSR_PROF_DECL
#define GPEMSR_PROF_T0
#define GPEMSR_PROF_WAIT(i)
#define GPEMSR_PROF_WORK
#define GPEMSR_PROF_PRINT(a, b, c)
#endif

namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int NUM_THREADS = 192;          // 2 role warps + 4 epilogue warps (Epi::WARPS == 4)
template <class Epi> constexpr int num_threads() { return 64 + 32 * Epi::WARPS; }
constexpr int MAX_TAPS = 49;

struct Operands {
  const __nv_bfloat16* a_hi;   // [K/8][a_rows][8]
  const __nv_bfloat16* a_lo;   // same shape, only read when SPLIT == 3
  const __nv_bfloat16* b_hi;   // [taps][K/8][b_rows][8]
  const __nv_bfloat16* b_lo;
  long long a_rows;            // allocated rows of A (k-chunk stride = a_rows * 16 B)
  int b_rows;                  // allocated rows of B per tap (>= n_tiles * BLOCK_N)
  int b_packed;                // 1: B is [n_tile][tap][k-chunk][plane][k-cell][BLOCK_N][8] (one contiguous block per stage,
                               //    b_lo unused); 0: plain K8-blocked [tap][K/8][b_rows][8] per plane
  int k;                       // reduction length per tap, multiple of BLOCK_K
  int taps;                    // number of shifted GEMMs accumulated into one tile
  int a_row_off[MAX_TAPS];     // row shift of A for every tap
  long long m_tiles;           // number of 128-row tiles
  int n_tiles;                 // number of BLOCK_N column tiles
  long long a_row0;            // row of A that tile 0 / row 0 maps to (so shifts may be negative)
  int* err_flag;               // device int, set non-zero on a pipeline time-out
  long long b_batch_elems;     // batched (block-diagonal) GEMM: row tile m uses B + (m / batch_tiles) * b_batch_elems (bf16
  int batch_tiles;             //   elements, both planes); batch_tiles == 0: one B for every row tile.  Streaming kernel only.
  // ---- tap-fused mode (gemm_tapfuse_kernel): B resident in smem, A fetched once per k-slab as <= 3 row segments
  int n_seg;                   // distinct tap dy values
  int seg_row_off[3];          // first A row of segment s relative to the tile's first row (dy * wp + dx_min)
  int seg_len;                 // rows per segment = 128 + (dx_max - dx_min)
  int tap_seg[MAX_TAPS];       // segment of tap t
  int tap_dx[MAX_TAPS];        // row offset of tap t inside its segment (dx - dx_min)
  int nstage;                  // smem stages that fit next to the resident B
  int kslabs;                  // 16-wide k-slabs per stage (1, 2 or 4): fewer, larger stages when shared memory allows
  // ---- dy-fused mode (gemm_dyfuse_kernel): large tap grids (7x7), weights streamed per (k-slab, dy)
  int n_dy, n_dx;              // taps = n_dy * n_dx, ordered dy-major
  int dy_row_off[8];           // first A row of the dy-th segment relative to the tile's first row (dy * wp + dx_min)
};

template <int BLOCK_N, int BLOCK_K, int SPLIT, int NSTAGE>
struct Config {
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N");
  static_assert(BLOCK_K % 16 == 0, "UMMA K = 16 for bf16");
  static_assert(SPLIT == 1 || SPLIT == 3, "SPLIT");
  static constexpr int PLANES = SPLIT == 3 ? 2 : 1;
  static constexpr int KCH = BLOCK_K / 8;                           // 16-byte k-cells per stage
  static constexpr int A_PLANE_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_PLANE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = PLANES * (A_PLANE_BYTES + B_PLANE_BYTES);
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 1024 /* barriers, tmem ptr, epilogue scratch */;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

// Tensor maps of the A planes (hi, lo) of a tap-fused / dy-fused launch; use == 0: the kernel issues plain bulk copies.
struct TmaMaps {
  sm100::TensorMap a[2];
  int use;
};

struct Barriers {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t b_full;             // tap-fused mode: resident B has landed
  uint32_t tmem_base;
  uint32_t tap_tab[2 * MAX_TAPS];   // tap-fused mode: per-tap descriptor start-address offsets
};

// Epilogue concept:
//   struct Epi {
//     struct State { ... };     // per-thread (= per-row) state, default-constructed at the start of every 128-row tile
//     static constexpr int WARPS = 4 or 8;   // epilogue warps: with 8, two warps share a TMEM lane quarter and each takes
//                                            // half of the tile's columns (`part` 0 / 1) -- twice the epilogue issue rate
//     __device__ void tile(State&, uint32_t tmem_acc /*lane-adjusted*/, long long m_tile, int n_tile, int n_tiles,
//                          int row_in_tile /*0..127 = this thread's row*/, int part) const;
//   };
// The epilogue reads its accumulator with sm100::tmem_ld_32x32(tmem_acc + col, regs) and must finish with the
// loads retired (tmem_ld_wait) before returning.

// CLUSTER == 2: the two CTAs of a cluster work on neighbouring row tiles in lock-step and SHARE the B operand: each
// loads half of every B stage and multicasts it into both shared memories, so weight / codebook traffic out of L2
// (the binding resource of this kernel: ~6.3 KB/clk chip-wide) is halved.  A stage is recycled only when BOTH tensor
// cores have consumed it (tcgen05.commit multicast onto both CTAs' `empty` barriers, which therefore count 2).
template <int BLOCK_N, int BLOCK_K, int SPLIT, int NSTAGE, class Epi, int CLUSTER = 1>
__global__ void __launch_bounds__(64 + 32 * Epi::WARPS, 1)
gemm_kernel(const __grid_constant__ Operands op, const __grid_constant__ Epi epi) {
  using Cfg = Config<BLOCK_N, BLOCK_K, SPLIT, NSTAGE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NSTAGE * Cfg::STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kiters_per_tap = op.k / BLOCK_K;
  const int kiters = op.taps * kiters_per_tap;
  const uint32_t crank = CLUSTER > 1 ? sm100::cluster_ctarank() : 0;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CLUSTER) - 1);
  // every CTA of a cluster runs the same number of row-tile iterations; surplus ones recompute the last tile and
  // discard it, so the lock-step multicast protocol never has a missing partner
  const long long n_iter = (op.m_tiles + gridDim.x - 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { sm100::mbar_init(&bars->full[s], 1); sm100::mbar_init(&bars->empty[s], CLUSTER); }
    for (int b = 0; b < 2; ++b) { sm100::mbar_init(&bars->tmem_full[b], 1); sm100::mbar_init(&bars->tmem_empty[b], Epi::WARPS); }
    sm100::fence_mbar_init();
  }
  if (warp == 1) sm100::tmem_alloc<Cfg::TMEM_COLS>(&bars->tmem_base);
  sm100::tc_fence_before();
  if constexpr (CLUSTER > 1) sm100::cluster_sync_all(); else __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== producer: bulk copies HBM/L2 -> smem =====================
    // The producer warp is on the critical path (one warp feeds the whole tensor pipe), so its loop is kept lean:
    // every lane owns ONE copy slot whose source pointer is advanced incrementally, and packed weights arrive as a
    // single copy per stage (one half per CTA when the cluster shares B).
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    constexpr int NA = Cfg::PLANES * Cfg::KCH;                 // A copies per stage (one 16-byte-cell column each)
    const int nb = op.b_packed ? 1 : NA;                       // B copy slots per stage
    const bool is_a = lane < NA;
    // with a cluster, B slot s of an unpacked operand is issued by CTA (s % CLUSTER); a packed stage is split in halves
    const bool is_b = lane >= NA && lane < NA + nb && (CLUSTER == 1 || op.b_packed || ((lane - NA) % CLUSTER) == (int)crank);
    const int slot = is_a ? lane : lane - NA;
    const int plane = slot / Cfg::KCH, kc = slot % Cfg::KCH;
    constexpr uint32_t kBStage = Cfg::PLANES * Cfg::B_PLANE_BYTES;
    const uint32_t b_part = op.b_packed ? kBStage / CLUSTER : BLOCK_N * 16;     // bytes this CTA's B copy moves
    const uint32_t sm_off = is_a ? (uint32_t)(plane * Cfg::A_PLANE_BYTES + kc * (BLOCK_M * 16))
                                 : (uint32_t)(Cfg::PLANES * Cfg::A_PLANE_BYTES) +
                                       (op.b_packed ? crank * b_part : (uint32_t)(plane * Cfg::B_PLANE_BYTES + kc * (BLOCK_N * 16)));
    const uint32_t bytes = is_a ? BLOCK_M * 16 : b_part;
    const long long a_kstep = (long long)Cfg::KCH * op.a_rows * 8;           // elements per k-chunk advance of A
    const long long b_kstep = op.b_packed ? (long long)kBStage / 2 : (long long)Cfg::KCH * op.b_rows * 8;
    for (long long i = 0; i < n_iter && ok; ++i) {
      const long long m_tile = min(blockIdx.x + i * gridDim.x, op.m_tiles - 1);
      const __nv_bfloat16* a_tile = (plane ? op.a_lo : op.a_hi) + ((long long)kc * op.a_rows + op.a_row0 + m_tile * BLOCK_M) * 8;
      const long long b_batch = op.batch_tiles ? (m_tile / op.batch_tiles) * op.b_batch_elems : 0;
      for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
        const __nv_bfloat16* b_src;
        if (op.b_packed) b_src = op.b_hi + (long long)n_tile * op.taps * kiters_per_tap * b_kstep + crank * (b_part / 2);
        else b_src = (plane ? op.b_lo : op.b_hi) + ((long long)kc * op.b_rows + (long long)n_tile * BLOCK_N) * 8 + b_batch;
        for (int tap = 0; tap < op.taps && ok; ++tap) {
          const __nv_bfloat16* a_src = a_tile + (long long)op.a_row_off[tap] * 8;
          for (int kci = 0; kci < kiters_per_tap; ++kci) {
            ok = sm100::mbar_wait(&bars->empty[stage], phase ^ 1, op.err_flag, 1);
            if (!ok) break;
            uint8_t* st = stages + stage * Cfg::STAGE_BYTES;
            if (lane == 0) sm100::mbar_arrive_expect_tx(&bars->full[stage], Cfg::STAGE_BYTES);
            __syncwarp();
            if (is_a) sm100::bulk_g2s(st + sm_off, a_src, bytes, &bars->full[stage]);
            else if (is_b) {
              if constexpr (CLUSTER > 1) sm100::bulk_g2s_multicast(st + sm_off, b_src, bytes, &bars->full[stage], kAllCtas);
              else sm100::bulk_g2s(st + sm_off, b_src, bytes, &bars->full[stage]);
            }
            a_src += a_kstep;
            b_src += b_kstep;
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected thread runs the whole loop (no per-trip elect / reconvergence) =====
    if (sm100::elect_one()) {
    constexpr uint32_t idesc = sm100::idesc_bf16_f32(BLOCK_M, BLOCK_N);
    uint32_t stage = 0, phase = 0, acc_buf = 0, acc_phase = 0;
    bool ok = true;
    const uint64_t a_desc0 = sm100::smem_desc_kmajor_noswz(sm100::smem_u32(stages), BLOCK_M * 16, 128);
    const uint64_t b_desc0 = sm100::smem_desc_kmajor_noswz(sm100::smem_u32(stages) + Cfg::PLANES * Cfg::A_PLANE_BYTES, BLOCK_N * 16, 128);
    for (long long i = 0; i < n_iter && ok; ++i) {
      for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
        ok = sm100::mbar_wait(&bars->tmem_empty[acc_buf], acc_phase ^ 1, op.err_flag, 2);
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc_buf * BLOCK_N;
        for (int it = 0; it < kiters && ok; ++it) {
          ok = sm100::mbar_wait(&bars->full[stage], phase, op.err_flag, 3);
          if (!ok) break;
          sm100::tc_fence_after();
          {
            // descriptors = base descriptor + offset in the 14-bit start-address field (address >> 4): the issuing warp
            // is close to the critical path, so nothing is rebuilt per trip
            const uint64_t sa = a_desc0 + (uint64_t)(stage * (Cfg::STAGE_BYTES >> 4));
            const uint64_t sb = b_desc0 + (uint64_t)(stage * (Cfg::STAGE_BYTES >> 4));
#pragma unroll
            for (int k16 = 0; k16 < BLOCK_K / 16; ++k16) {
              // one UMMA consumes two 16-byte k-cells: advance the start address by 2 cell columns per step
              const uint64_t a_hi = sa + k16 * ((2 * BLOCK_M * 16) >> 4);
              const uint64_t b_hi = sb + k16 * ((2 * BLOCK_N * 16) >> 4);
              sm100::umma_bf16(tmem_acc, a_hi, b_hi, idesc, (it | k16) != 0);
              if constexpr (SPLIT == 3) {
                const uint64_t a_lo = a_hi + (Cfg::A_PLANE_BYTES >> 4);
                const uint64_t b_lo = b_hi + (Cfg::B_PLANE_BYTES >> 4);
                sm100::umma_bf16(tmem_acc, a_lo, b_hi, idesc, true);
                sm100::umma_bf16(tmem_acc, a_hi, b_lo, idesc, true);
              }
            }
            // smem stage reusable once these MMAs retire (in BOTH CTAs when the cluster shares B)
            if constexpr (CLUSTER > 1) sm100::umma_commit_multicast(&bars->empty[stage], kAllCtas);
            else sm100::umma_commit(&bars->empty[stage]);
            if (it == kiters - 1) sm100::umma_commit(&bars->tmem_full[acc_buf]);   // accumulator complete
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
      }
    }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (TMEM -> registers -> HBM) =====================
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    uint32_t acc_buf = 0, acc_phase = 0;
    bool ok = true;
    for (long long i = 0; i < n_iter && ok; ++i) {
      const long long m_tile = blockIdx.x + i * gridDim.x;
      const bool live = m_tile < op.m_tiles;
      typename Epi::State st{};
      for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
        ok = sm100::mbar_wait(&bars->tmem_full[acc_buf], acc_phase, op.err_flag, 4);
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc_buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
        if (live) epi.tile(st, tmem_acc, m_tile, n_tile, op.n_tiles, row, (warp - 2) >> 2);
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&bars->tmem_empty[acc_buf]);
        if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
      }
    }
  }

  sm100::tc_fence_before();
  if constexpr (CLUSTER > 1) sm100::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant of the streaming kernel (`tcgen05.mma.cta_group::2`, M = 256): the two CTAs of a cluster work on neighbouring
// row tiles and ONE instruction of the leader drives both tensor cores.  Each CTA keeps its own A tile and only HALF of the B
// rows (N / 2) in shared memory, so per stage a CTA receives 32 KB instead of 48 KB and the tensor core reads 8 KB instead of
// 12 KB of operands per MMA: shared memory is what the single-CTA kernel saturates (per stage 6 MMAs x 12 KB read + 48 KB
// written = 960 cycles at 128 B/clk against 768 cycles of math; profiles/r02_mma_probe.txt).  Six 32 KB stages instead of four 48 KB.
// Operands arrive by tensor-map TMA (8-byte elements, see encode_u64_map): per stage one box per A plane {128 rows x 4 cells}
// and one box for this CTA's half of the packed weight block {128 rows x (plane, cell)}: three instructions per stage.
// Protocol (leader = cluster rank 0):
//   full[s]       on the leader only: its producer expects the bytes of BOTH CTAs' stage; the peer's TMA loads count down the
//                 leader's barrier through its cluster address (cp.async.bulk.tensor ... .cta_group::2)
//   empty[s]      per CTA, released by the leader's tcgen05.commit multicast once the pair's MMAs on the stage have retired
//   tmem_full[b]  per CTA, same commit multicast after the last k-step
//   tmem_empty[b] on the leader: all epilogue warps of BOTH CTAs arrive there (the peer's through the cluster address)
struct PairMaps {
  sm100::TensorMap a[2];       // A planes: {2 * rows, k-cells}, box {256, BLOCK_K / 8}
  sm100::TensorMap b[2];       // b_packed: b[0] = packed weights {2 * BLOCK_N, planes * cells, stage blocks}, box {BLOCK_N, planes * cells, 1};
                               // else one map per plane of the K8-blocked operand {2 * b_rows, taps * K / 8}, box {BLOCK_N, BLOCK_K / 8}
  int b_packed;
};

template <int BLOCK_N, int BLOCK_K, int SPLIT, int NSTAGE, class Epi>
__global__ void __launch_bounds__(64 + 32 * Epi::WARPS, 1)
gemm_pair_kernel(const __grid_constant__ Operands op, const __grid_constant__ Epi epi, const __grid_constant__ PairMaps tm) {
  constexpr int PLANES = SPLIT == 3 ? 2 : 1;
  constexpr int KCH = BLOCK_K / 8;
  constexpr int A_PLANE_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int BH_PLANE_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;           // this CTA's half of the weight rows
  constexpr int STAGE_BYTES = PLANES * (A_PLANE_BYTES + BH_PLANE_BYTES);
  constexpr int TMEM_COLS = 2 * BLOCK_N <= 256 ? 256 : 512;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "pair MMA: N multiple of 16 per CTA half");
  static_assert(NSTAGE <= 8, "Barriers holds 8 stages");
  static_assert(NSTAGE * STAGE_BYTES + 1024 <= 227 * 1024, "shared memory budget");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stages = smem;
  Barriers* bars = reinterpret_cast<Barriers*>(smem + NSTAGE * STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kiters_per_tap = op.k / BLOCK_K;
  const int kiters = op.taps * kiters_per_tap;
  const uint32_t crank = sm100::cluster_ctarank();
  const bool leader = crank == 0;
  const long long n_iter = (op.m_tiles + gridDim.x - 1) / gridDim.x;    // identical in both CTAs of a pair (surplus tiles are discarded)

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { sm100::mbar_init(&bars->full[s], 1); sm100::mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { sm100::mbar_init(&bars->tmem_full[b], 1); sm100::mbar_init(&bars->tmem_empty[b], 2 * Epi::WARPS); }
    sm100::fence_mbar_init();
  }
  if (warp == 1) sm100::tmem_alloc_pair<TMEM_COLS>(&bars->tmem_base);
  sm100::tc_fence_before();
  sm100::cluster_sync_all();
  sm100::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== producer (one elected thread): 2 A boxes + 1 weight box per stage =====================
    if (sm100::elect_one()) {
      uint32_t stage = 0, phase = 0;
      bool ok = true;
#pragma unroll
      for (int plane = 0; plane < PLANES; ++plane) sm100::tma_prefetch_desc(&tm.a[plane]);
#pragma unroll
      for (int plane = 0; plane < PLANES; ++plane) sm100::tma_prefetch_desc(&tm.b[plane]);
      GPEMSR_PROF_DECL
      const int kcells = op.k / 8;
      for (long long i = 0; i < n_iter && ok; ++i) {
        const long long m_tile = min(blockIdx.x + i * gridDim.x, op.m_tiles - 1);
        const long long row0 = op.a_row0 + m_tile * BLOCK_M;
        for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
          int blk = n_tile * op.taps * kiters_per_tap;                 // packed weight block of (n_tile, tap 0, k-chunk 0)
          for (int tap = 0; tap < op.taps && ok; ++tap) {
            const int a_c0 = (int)(2 * (row0 + op.a_row_off[tap]));
            for (int kci = 0; kci < kiters_per_tap; ++kci, ++blk) {
              GPEMSR_PROF_T0
              ok = sm100::mbar_wait(&bars->empty[stage], phase ^ 1, op.err_flag, 1);
              GPEMSR_PROF_WAIT(0)
              if (!ok) break;
              uint8_t* st = stages + stage * STAGE_BYTES;
              const uint32_t full_leader = sm100::cluster_map(&bars->full[stage], 0);
              if (leader) sm100::mbar_arrive_expect_tx(&bars->full[stage], 2 * STAGE_BYTES);
#pragma unroll
              for (int plane = 0; plane < PLANES; ++plane)
                sm100::tma_load_2d_pair(st + plane * A_PLANE_BYTES, &tm.a[plane], a_c0, kci * KCH, full_leader);
              if (tm.b_packed) {
                sm100::tma_load_3d_pair(st + PLANES * A_PLANE_BYTES, &tm.b[0], (int)crank * BLOCK_N, 0, blk, full_leader);
              } else {
                // this CTA's BLOCK_N / 2 rows of column tile n_tile, KCH cells of (tap, k-chunk)
#pragma unroll
                for (int plane = 0; plane < PLANES; ++plane)
                  sm100::tma_load_2d_pair(st + PLANES * A_PLANE_BYTES + plane * BH_PLANE_BYTES, &tm.b[plane],
                                          (2 * n_tile + (int)crank) * BLOCK_N, tap * kcells + kci * KCH, full_leader);
              }
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
              GPEMSR_PROF_WORK
            }
          }
        }
      }
      GPEMSR_PROF_PRINT("pair producer: wait empty", "-", "issue loads")
    }
    __syncwarp();
  } else if (warp == 1) {
    if (sm100::elect_one()) {
      uint32_t stage = 0, phase = 0, acc_buf = 0, acc_phase = 0;
      bool ok = true;
      GPEMSR_PROF_DECL
      if (leader) {
        // ===================== MMA issuer of the pair =====================
        constexpr uint32_t idesc = sm100::idesc_bf16_f32(2 * BLOCK_M, BLOCK_N);
        const uint64_t a_desc0 = sm100::smem_desc_kmajor_noswz(sm100::smem_u32(stages), BLOCK_M * 16, 128);
        const uint64_t b_desc0 = sm100::smem_desc_kmajor_noswz(sm100::smem_u32(stages) + PLANES * A_PLANE_BYTES, (BLOCK_N / 2) * 16, 128);
        for (long long i = 0; i < n_iter && ok; ++i) {
          for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
            GPEMSR_PROF_T0
            ok = sm100::mbar_wait(&bars->tmem_empty[acc_buf], acc_phase ^ 1, op.err_flag, 2);
            GPEMSR_PROF_WAIT(1)
            if (!ok) break;
            sm100::tc_fence_after();
            const uint32_t tmem_acc = tmem_base + acc_buf * BLOCK_N;
            for (int it = 0; it < kiters && ok; ++it) {
              GPEMSR_PROF_T0
              ok = sm100::mbar_wait(&bars->full[stage], phase, op.err_flag, 3);
              GPEMSR_PROF_WAIT(0)
              if (!ok) break;
              sm100::tc_fence_after();
              const uint64_t sa = a_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
              const uint64_t sb = b_desc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
#pragma unroll
              for (int k16 = 0; k16 < BLOCK_K / 16; ++k16) {
                const uint64_t a_hi = sa + k16 * ((2 * BLOCK_M * 16) >> 4);
                const uint64_t b_hi = sb + k16 * ((2 * (BLOCK_N / 2) * 16) >> 4);
                sm100::umma_bf16_pair(tmem_acc, a_hi, b_hi, idesc, (it | k16) != 0);
                if constexpr (SPLIT == 3) {
                  const uint64_t a_lo = a_hi + (A_PLANE_BYTES >> 4);
                  const uint64_t b_lo = b_hi + (BH_PLANE_BYTES >> 4);
                  sm100::umma_bf16_pair(tmem_acc, a_lo, b_hi, idesc, true);
                  sm100::umma_bf16_pair(tmem_acc, a_hi, b_lo, idesc, true);
                }
              }
              sm100::umma_commit_pair(&bars->empty[stage], 3);                               // stage reusable in both CTAs
              if (it == kiters - 1) sm100::umma_commit_pair(&bars->tmem_full[acc_buf], 3);   // both accumulators complete
              if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
              GPEMSR_PROF_WORK
            }
            if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
          }
        }
        GPEMSR_PROF_PRINT("pair mma: wait full", "wait tmem_empty", "issue + commit")
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (each CTA drains its own 128 rows) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t acc_buf = 0, acc_phase = 0;
    bool ok = true;
    for (long long i = 0; i < n_iter && ok; ++i) {
      const long long m_tile = blockIdx.x + i * gridDim.x;
      const bool live = m_tile < op.m_tiles;
      typename Epi::State st{};
      for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
        ok = sm100::mbar_wait(&bars->tmem_full[acc_buf], acc_phase, op.err_flag, 4);
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc_buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
        if (live) epi.tile(st, tmem_acc, m_tile, n_tile, op.n_tiles, row, (warp - 2) >> 2);
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive_cluster(&bars->tmem_empty[acc_buf], 0);
        if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
      }
    }
  }

  sm100::tc_fence_before();
  sm100::cluster_sync_all();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// A-resident variant (single bf16 pass, one tap, tiled B): the whole [128 x K] A tile of a row tile lives in shared memory
// and is fetched ONCE, while the B tiles of all column tiles stream through the stage ring.  The streaming kernel re-reads A
// for every column tile; for the VQ lookup (4 code tiles) that is 4 x 147 KB of the 1.2 MB a row tile pulls out of L2, and
// L2 -> SM operand bandwidth is what bounds that kernel.  There is no room for a second A tile (K = 544: 136 KB), so the
// tile is recycled chunk by chunk: when the MMAs of the LAST column tile have consumed k-chunk c (tcgen05.commit on
// a_empty[c]) a dedicated A-producer warp refills it with the next row tile's chunk c -- about one column tile of MMA
// time before the next row tile needs it.
// Roles: warp 0 = B producer, warp 1 = MMA issuer, warp 2 = A producer, warps 3.. = epilogue.
constexpr int ARES_MAX_CHUNKS = 24;
struct BarriersAres {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t a_full[ARES_MAX_CHUNKS];
  uint64_t a_empty[ARES_MAX_CHUNKS];
  uint32_t tmem_base;
};
template <class Epi> constexpr int num_threads_ares() { return 96 + 32 * Epi::WARPS; }
template <int BLOCK_N, int BLOCK_K, int NSTAGE>
constexpr int ares_smem_bytes(int k) { return k * BLOCK_M * 2 + NSTAGE * BLOCK_N * BLOCK_K * 2 + 1024; }

template <int BLOCK_N, int BLOCK_K, int NSTAGE, class Epi, int CLUSTER = 1>
__global__ void __launch_bounds__(96 + 32 * Epi::WARPS, 1)
gemm_ares_kernel(const __grid_constant__ Operands op, const __grid_constant__ Epi epi) {
  constexpr int KCH = BLOCK_K / 8;                              // 16-byte k-cells per chunk
  constexpr int A_CHUNK_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  constexpr int TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256) ? 256 : 512;
  static_assert(sizeof(BarriersAres) <= 1024, "barrier block");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int nch = op.k / BLOCK_K;                               // A chunks (<= ARES_MAX_CHUNKS, checked by the host)
  uint8_t* a_res = smem;
  uint8_t* stages = smem + nch * A_CHUNK_BYTES;
  BarriersAres* bars = reinterpret_cast<BarriersAres*>(stages + NSTAGE * B_STAGE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CLUSTER > 1 ? sm100::cluster_ctarank() : 0;
  constexpr uint16_t kAllCtas = (uint16_t)((1u << CLUSTER) - 1);
  const long long n_iter = (op.m_tiles + gridDim.x - 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { sm100::mbar_init(&bars->full[s], 1); sm100::mbar_init(&bars->empty[s], CLUSTER); }
    for (int b = 0; b < 2; ++b) { sm100::mbar_init(&bars->tmem_full[b], 1); sm100::mbar_init(&bars->tmem_empty[b], Epi::WARPS); }
    for (int c = 0; c < nch; ++c) { sm100::mbar_init(&bars->a_full[c], 1); sm100::mbar_init(&bars->a_empty[c], 1); }
    sm100::fence_mbar_init();
  }
  if (warp == 1) sm100::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  sm100::tc_fence_before();
  if constexpr (CLUSTER > 1) sm100::cluster_sync_all(); else __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== B producer: one contiguous (half-)stage copy per k-chunk =====================
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    constexpr uint32_t b_part = B_STAGE_BYTES / CLUSTER;
    const long long b_kstep = B_STAGE_BYTES / 2;                // elements per k-chunk of the tiled B
    for (long long i = 0; i < n_iter && ok; ++i) {
      for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
        const __nv_bfloat16* b_src = op.b_hi + (long long)n_tile * nch * b_kstep + crank * (b_part / 2);
        for (int kci = 0; kci < nch; ++kci) {
          ok = sm100::mbar_wait(&bars->empty[stage], phase ^ 1, op.err_flag, 1);
          if (!ok) break;
          if (lane == 0) {
            sm100::mbar_arrive_expect_tx(&bars->full[stage], B_STAGE_BYTES);
            uint8_t* dst = stages + stage * B_STAGE_BYTES + crank * b_part;
            if constexpr (CLUSTER > 1) sm100::bulk_g2s_multicast(dst, b_src, b_part, &bars->full[stage], kAllCtas);
            else sm100::bulk_g2s(dst, b_src, b_part, &bars->full[stage]);
          }
          __syncwarp();
          b_src += b_kstep;
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== A producer: refill chunk c as soon as the previous row tile has released it ==============
    bool ok = true;
    for (long long i = 0; i < n_iter && ok; ++i) {
      const long long m_tile = min(blockIdx.x + i * gridDim.x, op.m_tiles - 1);
      const __nv_bfloat16* a_tile = op.a_hi + (op.a_row0 + m_tile * BLOCK_M) * 8;
      const uint32_t par = (uint32_t)(i & 1);
      for (int c = 0; c < nch && ok; ++c) {
        ok = sm100::mbar_wait(&bars->a_empty[c], par ^ 1, op.err_flag, 5);
        if (!ok) break;
        if (lane == 0) sm100::mbar_arrive_expect_tx(&bars->a_full[c], A_CHUNK_BYTES);
        __syncwarp();
        if (lane < KCH)
          sm100::bulk_g2s(a_res + c * A_CHUNK_BYTES + lane * (BLOCK_M * 16),
                          a_tile + (long long)(c * KCH + lane) * op.a_rows * 8, BLOCK_M * 16, &bars->a_full[c]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: ONE elected thread runs the whole loop (no per-trip elect / reconvergence) =====
    if (sm100::elect_one()) {
    constexpr uint32_t idesc = sm100::idesc_bf16_f32(BLOCK_M, BLOCK_N);
    uint32_t stage = 0, phase = 0, acc_buf = 0, acc_phase = 0;
    bool ok = true;
    // the column tiles this CTA sweeps: first / last decide when A chunks are awaited / released
    const int n_first = blockIdx.y;
    const int n_last = n_first + ((op.n_tiles - 1 - n_first) / (int)gridDim.y) * (int)gridDim.y;
    const uint64_t a_desc0 = sm100::smem_desc_kmajor_noswz(sm100::smem_u32(a_res), BLOCK_M * 16, 128);
    const uint64_t b_desc0 = sm100::smem_desc_kmajor_noswz(sm100::smem_u32(stages), BLOCK_N * 16, 128);
    for (long long i = 0; i < n_iter && ok; ++i) {
      const uint32_t par = (uint32_t)(i & 1);
      for (int n_tile = n_first; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
        ok = sm100::mbar_wait(&bars->tmem_empty[acc_buf], acc_phase ^ 1, op.err_flag, 2);
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc_buf * BLOCK_N;
        // The single issuing warp is itself close to the critical path here (2 MMAs = 256 tensor cycles per trip): the trip
        // is kept lean -- descriptors are a base plus a 14-bit start-address offset, no per-trip descriptor rebuild.
        // (Issuing two chunks per trip was measured slower: it holds two of the five B stages twice as long.)
        for (int c = 0; c < nch && ok; ++c) {
          if (n_tile == n_first) {
            ok = sm100::mbar_wait(&bars->a_full[c], par, op.err_flag, 6);
            if (!ok) break;
          }
          ok = sm100::mbar_wait(&bars->full[stage], phase, op.err_flag, 3);
          if (!ok) break;
          sm100::tc_fence_after();
          {
            const uint64_t da = a_desc0 + (uint64_t)(c * (A_CHUNK_BYTES >> 4));
            const uint64_t db = b_desc0 + (uint64_t)(stage * (B_STAGE_BYTES >> 4));
#pragma unroll
            for (int k16 = 0; k16 < BLOCK_K / 16; ++k16)
              sm100::umma_bf16(tmem_acc, da + k16 * ((2 * BLOCK_M * 16) >> 4), db + k16 * ((2 * BLOCK_N * 16) >> 4), idesc, (c | k16) != 0);
            if constexpr (CLUSTER > 1) sm100::umma_commit_multicast(&bars->empty[stage], kAllCtas);
            else sm100::umma_commit(&bars->empty[stage]);
            if (n_tile == n_last) sm100::umma_commit(&bars->a_empty[c]);       // chunk c may be refilled for the next row tile
            if (c == nch - 1) sm100::umma_commit(&bars->tmem_full[acc_buf]);
          }
          if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
        }
        if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
      }
    }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t acc_buf = 0, acc_phase = 0;
    bool ok = true;
    for (long long i = 0; i < n_iter && ok; ++i) {
      const long long m_tile = blockIdx.x + i * gridDim.x;
      const bool live = m_tile < op.m_tiles;
      typename Epi::State st{};
      for (int n_tile = blockIdx.y; n_tile < op.n_tiles && ok; n_tile += gridDim.y) {
        ok = sm100::mbar_wait(&bars->tmem_full[acc_buf], acc_phase, op.err_flag, 4);
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc_buf * BLOCK_N + ((uint32_t)(q * 32) << 16);
        if (live) epi.tile(st, tmem_acc, m_tile, n_tile, op.n_tiles, row, (warp - 3) >> 2);
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&bars->tmem_empty[acc_buf]);
        if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
      }
    }
  }

  sm100::tc_fence_before();
  if constexpr (CLUSTER > 1) sm100::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tap-fused variant for narrow outputs (one column tile, B small enough to live in shared memory).
//   * B (all taps, all k) is loaded ONCE per CTA and stays resident;
//   * per 16-wide k-slab the producer fetches the A rows of the tile's neighbourhood once, as one segment per distinct
//     tap dy (rows [r0 + dy*wp + dx_min, +128 + dx_max - dx_min)); every tap then reads its operand from the same
//     stage at a 16-byte row offset (the canonical layout has uniform 16-byte rows, so a shift is a start address).
// A 3x3 convolution thus reads each activation row 3x instead of 9x and never re-reads weights: the wide-image 64-channel
// layers go from L2-operand-bound to shared-memory-operand / epilogue bound.
template <int BLOCK_N, int SPLIT, class Epi>
__global__ void __launch_bounds__(64 + 32 * Epi::WARPS, 1)
gemm_tapfuse_kernel(const __grid_constant__ Operands op, const __grid_constant__ Epi epi, const __grid_constant__ TmaMaps tm) {
  constexpr int PLANES = SPLIT == 3 ? 2 : 1;
  constexpr int KCH = 2;                                   // one UMMA k-step (16 bf16) per stage
  // SPLIT == 3 pairs the two products that share the A operand: a_hi * [w_hi | w_lo] is ONE MMA with N = 2 * BLOCK_N (the hi and
  // lo weight rows of a (tap, k-cell) sit next to each other in shared memory), a_lo * w_hi the second (N = BLOCK_N, into the
  // first half).  The accumulator is 2 * BLOCK_N columns and the epilogue adds column c + BLOCK_N to column c (same TMEM lane).
  // Isolated cost of one M = 128, K = 16 MMA (tools/mma_probe.cu, profiles/r02_mma_probe.txt): max(math, operand fetch at 128 B/clk)
  // = 39 / 42 / 49 / 64 / 128 cycles for N = 16 / 32 / 64 / 128 / 256, so the pair N = 128 + N = 64 is 112 cycles against 3 x 49.
  constexpr int NMUL = SPLIT == 3 ? 2 : 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int kcells = op.k / 8;
  const uint32_t b_tap_bytes = (uint32_t)kcells * BLOCK_N * 16;               // one tap, one plane
  const uint32_t b_bytes = (uint32_t)PLANES * op.taps * b_tap_bytes;
  const uint32_t seg_bytes = (uint32_t)op.seg_len * 16;                       // one 16-byte k-cell column of a segment
  // a stage plane is [k-cell][segment][row][8] (the box order of the tensor map: ascending global stride), padded to the 128-byte
  // alignment a TMA destination needs
  // A stage carries S = op.kslabs k-slabs (2 * S cell columns): the ring hand-shake (commit -> empty -> refill -> full) costs the
  // issuing threads a few hundred cycles per stage, as much as the 9 N = 64 MMAs of a single-plane slab take (432 cycles).
  const int S = op.kslabs;
  const uint32_t slab_bytes = (uint32_t)op.n_seg * KCH * seg_bytes;           // one k-slab of one plane
  const uint32_t a_box_bytes = (uint32_t)S * slab_bytes;
  const uint32_t a_plane_bytes = (a_box_bytes + 127u) & ~127u;
  const uint32_t stage_bytes = PLANES * a_plane_bytes;
  uint8_t* b_res = smem;
  uint8_t* stages = smem + ((b_bytes + 1023) & ~1023u);
  Barriers* bars = reinterpret_cast<Barriers*>(stages + (size_t)op.nstage * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kiters = op.k / 16 / S;
  const int nstage = op.nstage;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) { sm100::mbar_init(&bars->full[s], 1); sm100::mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { sm100::mbar_init(&bars->tmem_full[b], 1); sm100::mbar_init(&bars->tmem_empty[b], Epi::WARPS / Epi::GROUPS); }
    sm100::mbar_init(&bars->b_full, 1);
    sm100::fence_mbar_init();
  }
  constexpr int TMEM_COLS = (2 * NMUL * BLOCK_N <= 32) ? 32 : (2 * NMUL * BLOCK_N <= 64) ? 64 : (2 * NMUL * BLOCK_N <= 128) ? 128 : 256;
  if (warp == 1) sm100::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== producer =====================
    bool ok = true;
    // ONE elected thread runs the whole refill loop.  UBLKCP takes its addresses from uniform registers: issued from a converged
    // single-thread region it is a plain instruction, while "one copy per lane" compiles to an ELECT / R2UR.BROADCAST / BRA.U.ANY
    // loop that serialises the lanes at ~50 cycles each -- 12 copies per stage were ~730 cycles of the refill latency, and for
    // the single-plane (split 1) layers more than the 432 tensor cycles the stage feeds (profiles/r02_tapfuse_roles.txt).
    if (sm100::elect_one()) {
      if (tm.use) {
#pragma unroll
        for (int plane = 0; plane < PLANES; ++plane) sm100::tma_prefetch_desc(&tm.a[plane]);
      }
      // the resident weights are fetched by the epilogue warps (below); this thread only registers the byte count
      if (blockIdx.x < op.m_tiles) sm100::mbar_arrive_expect_tx(&bars->b_full, b_bytes);
      uint32_t stage = 0, phase = 0;
      const long long a_kstep = (long long)S * KCH * op.a_rows * 8;
      GPEMSR_PROF_DECL
      for (long long m_tile = blockIdx.x; m_tile < op.m_tiles && ok; m_tile += gridDim.x) {
        const long long row0 = op.a_row0 + m_tile * BLOCK_M;
        for (int it = 0; it < kiters && ok; ++it) {
          GPEMSR_PROF_T0
          ok = sm100::mbar_wait(&bars->empty[stage], phase ^ 1, op.err_flag, 1);
          GPEMSR_PROF_WAIT(0)
          if (!ok) break;
          uint8_t* sa = stages + (size_t)stage * stage_bytes;
#if GPEMSR_ABLATE & 2
          if (m_tile != blockIdx.x) sm100::mbar_arrive(&bars->full[stage]); else     // profiling build: A loaded for the first tile only
#endif
          {
            sm100::mbar_arrive_expect_tx(&bars->full[stage], PLANES * a_box_bytes);
            if (tm.use) {
              // one box per plane, 8-byte elements: {seg_len, 2 halves, n_seg segments, 2 * S cells} at (2 * first row of segment 0, 0, 0, k-cell)
#pragma unroll
              for (int plane = 0; plane < PLANES; ++plane)
                sm100::tma_load_4d(sa + plane * a_plane_bytes, &tm.a[plane], (int)(2 * (row0 + op.seg_row_off[0])), 0, 0, it * S * KCH, &bars->full[stage]);
#if GPEMSR_L2_PREFETCH
              // the same k-slab of this CTA's NEXT tile, DRAM -> L2: the stage ring holds ~2 stages (50 KB) in flight, not enough to
              // cover the DRAM latency at the rate the MMAs consume them; from L2 it is
              if (m_tile + gridDim.x < op.m_tiles) {
#pragma unroll
                for (int plane = 0; plane < PLANES; ++plane)
                  sm100::tma_prefetch_l2_4d(&tm.a[plane], (int)(2 * (row0 + (long long)gridDim.x * BLOCK_M + op.seg_row_off[0])), 0, 0, it * S * KCH);
              }
#endif
            } else {
#pragma unroll
              for (int plane = 0; plane < PLANES; ++plane) {
                const __nv_bfloat16* pb = (plane ? op.a_lo : op.a_hi) + (long long)it * a_kstep;
                for (int kc = 0; kc < S * KCH; ++kc) {
#pragma unroll
                  for (int seg = 0; seg < 3; ++seg) {
                    if (seg < op.n_seg)
                      sm100::bulk_g2s(sa + plane * a_plane_bytes + (uint32_t)(kc * op.n_seg + seg) * seg_bytes,
                                      pb + ((long long)kc * op.a_rows + row0 + op.seg_row_off[seg]) * 8, seg_bytes, &bars->full[stage]);
                  }
                }
              }
            }
          }
          if (++stage == (uint32_t)nstage) { stage = 0; phase ^= 1; }
          GPEMSR_PROF_WORK
        }
      }
      GPEMSR_PROF_PRINT("producer: wait empty", "-", "issue copies")
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = sm100::idesc_bf16_f32(BLOCK_M, BLOCK_N);
    constexpr uint32_t idesc_pair = sm100::idesc_bf16_f32(BLOCK_M, NMUL * BLOCK_N);
    uint32_t stage = 0, phase = 0, acc_buf = 0, acc_phase = 0;
    bool ok = true;
    if (blockIdx.x < op.m_tiles) ok = sm100::mbar_wait(&bars->b_full, 0, op.err_flag, 5);
    const uint32_t sb0 = sm100::smem_u32(b_res);
    // per-tap start-address offsets (16-byte units): [2t] = A (segment + dx shift), [2t+1] = B (tap block)
    uint32_t* tap_tab = bars->tap_tab;
    for (int t = lane; t < op.taps; t += 32) {
      tap_tab[2 * t] = ((uint32_t)op.tap_seg[t] * seg_bytes + (uint32_t)op.tap_dx[t] * 16) >> 4;
      tap_tab[2 * t + 1] = ((uint32_t)t * PLANES * b_tap_bytes) >> 4;
    }
    __syncwarp();
    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);                   // SBO = 128 B, descriptor version 1
    const uint32_t a_desc_lo = ((((uint32_t)op.n_seg * seg_bytes) >> 4) & 0x3FFFu) << 16;    // LBO = the next k-cell: n_seg segments further
    const uint32_t b_desc_lo = (((uint32_t)PLANES * BLOCK_N * 16 >> 4) & 0x3FFFu) << 16;      // LBO: the next k-cell's [plane][n] block
    // a full 3x3 tap grid in dy-major order: segments of BLOCK_M + 2 rows, tap (dy, dx) at segment dy, row dx
    bool is3x3 = op.taps == 9 && op.n_seg == 3 && op.seg_len == BLOCK_M + 2;
    for (int t = 0; t < 9 && is3x3; ++t) is3x3 = op.tap_seg[t] == t / 3 && op.tap_dx[t] == t % 3;
    // ONE elected thread runs the whole loop.  The tensor pipe queues only a few MMAs, so whatever the issuing thread does
    // between the last MMA of a stage and the first of the next is a bubble: the barrier of the NEXT stage (and, on the last
    // k-slab, the next accumulator buffer) is probed while this stage's MMAs are still being issued, and the blocking wait runs
    // only when that probe failed.
    if (sm100::elect_one()) {
    GPEMSR_PROF_DECL
    bool full_ready = false, acc_ready = false;
    for (long long m_tile = blockIdx.x; m_tile < op.m_tiles && ok; m_tile += gridDim.x) {
      GPEMSR_PROF_T0
      if (!acc_ready) ok = sm100::mbar_wait(&bars->tmem_empty[acc_buf], acc_phase ^ 1, op.err_flag, 2);
      GPEMSR_PROF_WAIT(1)
      if (!ok) break;
      sm100::tc_fence_after();
      acc_ready = false;
      const uint32_t tmem_acc = tmem_base + acc_buf * (NMUL * BLOCK_N);
      for (int it = 0; it < kiters && ok; ++it) {
        GPEMSR_PROF_T0
        if (!full_ready) ok = sm100::mbar_wait(&bars->full[stage], phase, op.err_flag, 3);
        GPEMSR_PROF_WAIT(0)
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t nstg = stage + 1 == (uint32_t)nstage ? 0 : stage + 1, nph = stage + 1 == (uint32_t)nstage ? phase ^ 1 : phase;
        for (int sl = 0; sl < S; ++sl) {
          // descriptors differ from a per-slab base only in the 14-bit start-address field (16-byte units): two adds per MMA
          const uint32_t a_lo0 = a_desc_lo + ((sm100::smem_u32(stages + (size_t)stage * stage_bytes) + (uint32_t)sl * slab_bytes) >> 4);
          const uint32_t b_lo0 = b_desc_lo + ((sb0 + (uint32_t)((it * S + sl) * KCH) * (PLANES * BLOCK_N * 16)) >> 4);
          const bool last_slab = sl == S - 1;
#if GPEMSR_ABLATE & 4
          if (false) {                                     // profiling build: no MMAs, only the commits
#else
          if (is3x3) {
#endif
            // The 3x3 case (every 64-channel convolution of the model): the nine start-address offsets are compile-time
            // constants (segment dy at 130 cell rows, dx = one row), so the fully unrolled loop is an add-immediate per
            // descriptor.
            const uint32_t b_tap16 = ((uint32_t)PLANES * b_tap_bytes) >> 4, a_pl16 = a_plane_bytes >> 4;
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const uint32_t a_lo = a_lo0 + (uint32_t)((t / 3) * (BLOCK_M + 2) + t % 3), b_lo = b_lo0 + (uint32_t)t * b_tap16;
              const uint64_t a_hi_d = ((uint64_t)kDescHi << 32) | a_lo, b_hi_d = ((uint64_t)kDescHi << 32) | b_lo;
              if constexpr (SPLIT == 3) {
                const uint64_t a_lo_d = ((uint64_t)kDescHi << 32) | (a_lo + a_pl16);
                sm100::umma_bf16(tmem_acc, a_hi_d, b_hi_d, idesc_pair, (it | sl | t) != 0);    // a_hi * [w_hi | w_lo]
                sm100::umma_bf16(tmem_acc, a_lo_d, b_hi_d, idesc, true);                       // a_lo * w_hi
              } else {
                sm100::umma_bf16(tmem_acc, a_hi_d, b_hi_d, idesc, (it | sl | t) != 0);
              }
              if (t == 5 && last_slab) full_ready = sm100::mbar_try_wait(&bars->full[nstg], nph);
            }
          } else {
          for (int t = 0; t < (GPEMSR_ABLATE & 4 ? 0 : op.taps); ++t) {
            const uint32_t a_lo = a_lo0 + tap_tab[2 * t], b_lo = b_lo0 + tap_tab[2 * t + 1];
            const uint64_t a_hi_d = ((uint64_t)kDescHi << 32) | a_lo, b_hi_d = ((uint64_t)kDescHi << 32) | b_lo;
            if constexpr (SPLIT == 3) {
              const uint64_t a_lo_d = ((uint64_t)kDescHi << 32) | (a_lo + (a_plane_bytes >> 4));
              sm100::umma_bf16(tmem_acc, a_hi_d, b_hi_d, idesc_pair, (it | sl | t) != 0);      // a_hi * [w_hi | w_lo]
              sm100::umma_bf16(tmem_acc, a_lo_d, b_hi_d, idesc, true);                         // a_lo * w_hi
            } else {
              sm100::umma_bf16(tmem_acc, a_hi_d, b_hi_d, idesc, (it | sl | t) != 0);
            }
          }
          if (last_slab) full_ready = sm100::mbar_try_wait(&bars->full[nstg], nph);
          }
          if (!last_slab) continue;
          sm100::umma_commit(&bars->empty[stage]);
          if (it == kiters - 1) {
            sm100::umma_commit(&bars->tmem_full[acc_buf]);
            acc_ready = sm100::mbar_try_wait(&bars->tmem_empty[acc_buf ^ 1], acc_buf == 1 ? acc_phase : acc_phase ^ 1);
          }
        }
        stage = nstg; phase = nph;
        GPEMSR_PROF_WORK
      }
      if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
    }
    GPEMSR_PROF_PRINT("mma: wait full", "wait tmem_empty", "issue + commit")
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // The resident weights first: one 1 KB .. 4 KB block per (tap, k-cell, plane) into the [tap][kc][plane][n] layout.  A bulk copy
    // costs its issuing thread ~65 cycles, and 144 of them from the producer alone were ~5 us before the first MMA of EVERY
    // launch (most launches of the model are small: ~50 us); the epilogue warps are idle until the first tile is done, so their
    // elected lanes issue a slice each.  (complete_tx may precede the producer's expect_tx: the phase cannot complete before it.)
    if (blockIdx.x < op.m_tiles && sm100::elect_one()) {
      const int nblk = op.taps * kcells;
      for (int r = warp - 2; r < nblk; r += Epi::WARPS) {                    // r = tap * kcells + kc
#pragma unroll
        for (int plane = 0; plane < PLANES; ++plane)
          sm100::bulk_g2s(b_res + ((size_t)r * PLANES + plane) * BLOCK_N * 16, (plane ? op.b_lo : op.b_hi) + (long long)r * op.b_rows * 8,
                          BLOCK_N * 16, &bars->b_full);
      }
    }
    __syncwarp();
    // Epi::GROUPS == 2: group g owns accumulator buffer g and takes every other tile (two tile epilogues in flight)
    constexpr int WPG = Epi::WARPS / Epi::GROUPS;
    const int q = warp & 3, grp = (warp - 2) / WPG, part = ((warp - 2) % WPG) >> 2;
    const int row = q * 32 + lane;
    uint32_t acc_buf = 0, acc_phase = 0;
    bool ok = true;
    GPEMSR_PROF_DECL
    for (long long m_tile = blockIdx.x; m_tile < op.m_tiles && ok; m_tile += gridDim.x) {
      if (Epi::GROUPS == 1 || (int)acc_buf == grp) {
        typename Epi::State st{};
        GPEMSR_PROF_T0
        ok = sm100::mbar_wait(&bars->tmem_full[acc_buf], acc_phase, op.err_flag, 4);
        GPEMSR_PROF_WAIT(0)
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc_buf * (NMUL * BLOCK_N) + ((uint32_t)(q * 32) << 16);
        if (!(GPEMSR_ABLATE & 1)) epi.tile_group(st, tmem_acc, m_tile, 0, 1, row, part);     // profiling build: no epilogue work
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&bars->tmem_empty[acc_buf]);
        GPEMSR_PROF_WORK
      }
      if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
    }
    if (warp == 2 || warp == 2 + Epi::WARPS - 1) { GPEMSR_PROF_PRINT("epilogue: wait tmem_full", "-", "tile") }
  }

  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// dy-fused variant for LARGE tap grids on narrow layers (SpyNet's 7x7 convolutions, 8..64 channels).  The streaming kernel
// fetches the A tile once per tap (49x) and the tap-fused kernel needs every tap's weights resident (401 KB for 64x32x49 in
// (hi, lo) planes: impossible).  Here a pipeline stage is one (16-wide k-slab, tap row dy): ONE A segment of 128 + 2P rows --
// the n_dx taps of that row are 16-byte start-address shifts inside it -- plus the n_dx weight blocks of that (slab, dy),
// which arrive as one contiguous copy from a [slab][dy][dx][cell][plane][BLOCK_N][8] packing (hi and lo weight rows of a
// (tap, k-cell) adjacent: the paired-N MMA of the SPLIT == 3 path reads them as one N = 2 * BLOCK_N operand).  A is read n_dy times instead
// of n_dy * n_dx times; a stage carries n_dx * SPLIT MMAs.
template <int BLOCK_N, int SPLIT, class Epi>
__global__ void __launch_bounds__(64 + 32 * Epi::WARPS, 1)
gemm_dyfuse_kernel(const __grid_constant__ Operands op, const __grid_constant__ Epi epi, const __grid_constant__ TmaMaps tm) {
  constexpr int PLANES = SPLIT == 3 ? 2 : 1;
  constexpr int NMUL = SPLIT == 3 ? 2 : 1;                 // paired-N MMAs, see gemm_tapfuse_kernel
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t seg_bytes = (uint32_t)op.seg_len * 16;                       // one 16-byte k-cell column of the segment
  const uint32_t a_plane_bytes = (2 * seg_bytes + 127u) & ~127u;              // [cell][row][8], padded: a TMA destination is 128-byte aligned
  const uint32_t a_bytes = PLANES * a_plane_bytes;
  const uint32_t b_tap_bytes = PLANES * 2 * BLOCK_N * 16;                     // one tap, one k-slab: [cell][plane][n][8]
  const uint32_t b_bytes = (uint32_t)op.n_dx * b_tap_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* stages = smem;
  Barriers* bars = reinterpret_cast<Barriers*>(stages + (size_t)op.nstage * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kiters = (op.k / 16) * op.n_dy;
  const int nstage = op.nstage;

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) { sm100::mbar_init(&bars->full[s], 1); sm100::mbar_init(&bars->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { sm100::mbar_init(&bars->tmem_full[b], 1); sm100::mbar_init(&bars->tmem_empty[b], Epi::WARPS / Epi::GROUPS); }
    sm100::fence_mbar_init();
  }
  constexpr int TMEM_COLS = (2 * NMUL * BLOCK_N <= 32) ? 32 : (2 * NMUL * BLOCK_N <= 64) ? 64 : (2 * NMUL * BLOCK_N <= 128) ? 128 : 256;
  if (warp == 1) sm100::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================== producer (one elected thread, see gemm_tapfuse_kernel): 2 * PLANES A cell columns + the weight block
    if (sm100::elect_one()) {
      bool ok = true;
      uint32_t stage = 0, phase = 0;
      const long long a_slab = 2LL * op.a_rows * 8;                               // elements per k-slab advance of A
      for (long long m_tile = blockIdx.x; m_tile < op.m_tiles && ok; m_tile += gridDim.x) {
        const __nv_bfloat16* b_src = op.b_hi;
        const long long row0 = op.a_row0 + m_tile * BLOCK_M;
        int slab = 0, dyi = 0;
        for (int it = 0; it < kiters && ok; ++it) {
          ok = sm100::mbar_wait(&bars->empty[stage], phase ^ 1, op.err_flag, 1);
          if (!ok) break;
          uint8_t* st = stages + (size_t)stage * stage_bytes;
          sm100::mbar_arrive_expect_tx(&bars->full[stage], PLANES * 2 * seg_bytes + b_bytes);
          if (tm.use) {
            // one box per plane, 8-byte elements: {seg_len, 2 halves, 2 cells} at (2 * first row of this dy, 0, first cell of the slab)
#pragma unroll
            for (int plane = 0; plane < PLANES; ++plane)
              sm100::tma_load_3d(st + plane * a_plane_bytes, &tm.a[plane], (int)(2 * (row0 + op.dy_row_off[dyi])), 0, slab * 2, &bars->full[stage]);
          } else {
            const long long a_off = slab * a_slab + (row0 + op.dy_row_off[dyi]) * 8;
#pragma unroll
            for (int plane = 0; plane < PLANES; ++plane) {
#pragma unroll
              for (int cell = 0; cell < 2; ++cell)
                sm100::bulk_g2s(st + plane * a_plane_bytes + (uint32_t)cell * seg_bytes, (plane ? op.a_lo : op.a_hi) + a_off + (long long)cell * op.a_rows * 8,
                                seg_bytes, &bars->full[stage]);
            }
          }
          sm100::bulk_g2s(st + a_bytes, b_src, b_bytes, &bars->full[stage]);
          b_src += b_bytes / 2;
          if (++dyi == op.n_dy) { dyi = 0; ++slab; }
          if (++stage == (uint32_t)nstage) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    if (sm100::elect_one()) {
      constexpr uint32_t idesc = sm100::idesc_bf16_f32(BLOCK_M, BLOCK_N);
      constexpr uint32_t idesc_pair = sm100::idesc_bf16_f32(BLOCK_M, NMUL * BLOCK_N);
      uint32_t stage = 0, phase = 0, acc_buf = 0, acc_phase = 0;
      bool ok = true;
      // descriptors differ only in their low word (LBO << 16 | start address >> 4); the high word (SBO = 128 B, version 1) is
      // a constant, so the issuing thread does 32-bit adds only.  An M = 128, K = 16 MMA costs max(math, (4 KB + N * 32 B) /
      // 128 B/clk) -- 39 .. 49 cycles for N = 16 .. 64 (profiles/r02_mma_probe.txt) -- so these narrow layers run far below the
      // tensor rate; keeping the (small) weight sets resident instead of streaming them was measured and changes nothing.
      constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_base = (((seg_bytes >> 4) & 0x3FFFu) << 16) | (sm100::smem_u32(stages) >> 4);
      const uint32_t b_lo_base = ((((uint32_t)PLANES * BLOCK_N * 16 >> 4) & 0x3FFFu) << 16) | ((sm100::smem_u32(stages) + a_bytes) >> 4);
      const uint32_t a_pl = a_plane_bytes >> 4, b_tap = b_tap_bytes >> 4;
      // the next stage's barrier (and the next accumulator's) is probed while this stage's MMAs are in flight, see gemm_tapfuse_kernel
      bool full_ready = false, acc_ready = false;
      for (long long m_tile = blockIdx.x; m_tile < op.m_tiles && ok; m_tile += gridDim.x) {
        if (!acc_ready) ok = sm100::mbar_wait(&bars->tmem_empty[acc_buf], acc_phase ^ 1, op.err_flag, 2);
        if (!ok) break;
        sm100::tc_fence_after();
        acc_ready = false;
        const uint32_t tmem_acc = tmem_base + acc_buf * (NMUL * BLOCK_N);
        for (int it = 0; it < kiters && ok; ++it) {
          if (!full_ready) ok = sm100::mbar_wait(&bars->full[stage], phase, op.err_flag, 3);
          if (!ok) break;
          sm100::tc_fence_after();
          const uint32_t nstg = stage + 1 == (uint32_t)nstage ? 0 : stage + 1, nph = stage + 1 == (uint32_t)nstage ? phase ^ 1 : phase;
          uint32_t a = a_lo_base + stage * (stage_bytes >> 4);
          uint32_t b = b_lo_base + stage * (stage_bytes >> 4);
          // one tap = one 16-byte row of A and one weight block of B.  The issuing thread bounds the narrow kernels, so the two tap
          // counts the model uses (3: the K > 64 3x3 convs, 7: SpyNet) are fully unrolled: add-immediate descriptors, no loop
          auto issue = [&](int dx) {
            const uint32_t ad = a + (uint32_t)dx, bd = b + (uint32_t)dx * b_tap;
            const uint64_t a_hi = ((uint64_t)kDescHi << 32) | ad, b_hi = ((uint64_t)kDescHi << 32) | bd;
            if constexpr (SPLIT == 3) {
              sm100::umma_bf16(tmem_acc, a_hi, b_hi, idesc_pair, (it | dx) != 0);                               // a_hi * [w_hi | w_lo]
              sm100::umma_bf16(tmem_acc, ((uint64_t)kDescHi << 32) | (ad + a_pl), b_hi, idesc, true);           // a_lo * w_hi
            } else {
              sm100::umma_bf16(tmem_acc, a_hi, b_hi, idesc, (it | dx) != 0);
            }
          };
          if (op.n_dx == 3) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) { issue(dx); if (dx == 1) full_ready = sm100::mbar_try_wait(&bars->full[nstg], nph); }
          } else if (op.n_dx == 7) {
#pragma unroll
            for (int dx = 0; dx < 7; ++dx) { issue(dx); if (dx == 4) full_ready = sm100::mbar_try_wait(&bars->full[nstg], nph); }
          } else {
            for (int dx = 0; dx < op.n_dx; ++dx) issue(dx);
            full_ready = sm100::mbar_try_wait(&bars->full[nstg], nph);
          }
          sm100::umma_commit(&bars->empty[stage]);
          if (it == kiters - 1) {
            sm100::umma_commit(&bars->tmem_full[acc_buf]);
            acc_ready = sm100::mbar_try_wait(&bars->tmem_empty[acc_buf ^ 1], acc_buf == 1 ? acc_phase : acc_phase ^ 1);
          }
          stage = nstg; phase = nph;
        }
        if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    // Epi::GROUPS == 2: group g owns accumulator buffer g and takes every other tile, see gemm_tapfuse_kernel
    constexpr int WPG = Epi::WARPS / Epi::GROUPS;
    const int q = warp & 3, grp = (warp - 2) / WPG, part = ((warp - 2) % WPG) >> 2;
    const int row = q * 32 + lane;
    uint32_t acc_buf = 0, acc_phase = 0;
    bool ok = true;
    for (long long m_tile = blockIdx.x; m_tile < op.m_tiles && ok; m_tile += gridDim.x) {
      if (Epi::GROUPS == 1 || (int)acc_buf == grp) {
        typename Epi::State st{};
        ok = sm100::mbar_wait(&bars->tmem_full[acc_buf], acc_phase, op.err_flag, 4);
        if (!ok) break;
        sm100::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + acc_buf * (NMUL * BLOCK_N) + ((uint32_t)(q * 32) << 16);
        epi.tile_group(st, tmem_acc, m_tile, 0, 1, row, part);
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&bars->tmem_empty[acc_buf]);
      }
      if (++acc_buf == 2) { acc_buf = 0; acc_phase ^= 1; }
    }
  }

  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    sm100::tc_fence_after();
    sm100::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace gemm
