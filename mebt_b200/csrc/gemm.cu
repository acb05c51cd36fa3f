// K2/K4: persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )        bf16 operands, fp32 accumulation in TMEM
//
// Replaces the cuBLAS addmm/mm calls behind nn.Linear on the MeBT hot path
// (reference: mebt/modules/gpt.py:126-128 q/k/v, :140 proj, :150-155 mlp, :248 head) and, with the
// MN-major operand modes, their dgrad/wgrad counterparts that autograd would dispatch.
//
// Structure (one CTA per SM, 192 threads):
//   warp 0    : TMA producer  - cp.async.bulk.tensor tiles into a STAGES-deep smem ring (128B swizzle)
//   warp 1    : MMA issuer    - one thread issues tcgen05.mma (M=128, N=BN, K=16) into TMEM
//   warps 2-5 : epilogue      - tcgen05.ld the fp32 accumulator, bias / GELU / residual, store
// Two TMEM accumulator buffers let the epilogue of tile i overlap the main loop of tile i+1.
//
// Operand majors: "K-major" = reduction dim contiguous (A is [M,K] row-major, B is [N,K] row-major,
// i.e. torch Linear weights).  "MN-major" = the M (or N) dim contiguous (A given as [K,M], B as [K,N]).
#include "common.cuh"

namespace mebt {

namespace {

constexpr int MEBT_GEMM_INTERNAL_ARGMIN = 1 << 30;   // not part of the C ABI
constexpr int MEBT_GEMM_INTERNAL_SAMPLE = 1 << 29;
struct SampleRequest { bool active; float scale; uint32_t k0, k1; };
thread_local SampleRequest g_sample_req = {false, 0.f, 0u, 0u};
constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;

// Epilogue staging: a ring of EPI_SLOTS slots of [128 rows x 128 bytes] (64 bf16 or 32 fp32 columns of the tile's
// 128 accumulator rows, 128B-swizzled; each epilogue warp fills its 32 rows).  Results leave through ONE TMA store per
// slot (full 128-byte lines instead of 32 scattered 16-byte pieces per store instruction; the TMA unit costs ~190 clk
// per box on top of ~1 clk per row, so few large boxes) and a residual / saved pre-activation operand arrives through
// TMA loads issued up to three slots ahead; both were what the MMA issuer ended up waiting for before.
constexpr int RASTER_M = 16;           // row blocks (pair mode: row-block pairs) per rasterisation band
constexpr int EPI_SLOTS = 4;
constexpr int EPI_SLOT_BYTES = 128 * 128;

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  const float* bias;                   // [N] or nullptr
  const __nv_bfloat16* residual;       // [M, ldres] or nullptr
  int ldres;
  void* C;
  int ldc;
  int gelu;
  int out_fp32;
  int accumulate;                      // C += result (fp32 output only)
  __nv_bfloat16* aux;                  // gelu: optional pre-activation output; dgelu: pre-activation input
  int ldaux;
  int dgelu;                           // result *= gelu'(aux)
  int in_kind;                         // operand read by the epilogue through TMA: 0 none, 1 residual, 2 dgelu aux,
                                       // 3 attention output O (the epilogue also emits delta = rowsum_head(C .* O))
  float* delta_out;                    // in_kind 3: [B, H, NQ] fp32, rows of C are (b, q), 64-column groups are heads
  int delta_nq, delta_h;
  int fp16_in;                         // operands are fp16 (kind::f16 A/B format 0) instead of bf16
  // nearest-code search (K9): no C; per row the running minimum of d = (row_sq[row] - 2 acc) + bias[col] over all
  // columns, merged across tiles by a 64-bit atomicMin on (orderable d bits << 32 | col)
  unsigned long long* argmin_out;
  const float* row_sq;
  // Gumbel-max sampling (K6 fused into the head GEMM): no C; per row the running maximum over all columns of
  // acc * sample_scale - lg2(-lg2(u)), u = a counter hash of (sample key, row, column) in (0, 1), merged across tiles
  // like the nearest-code search (argmin_out holds ~ordered(value) << 32 | column).  sample_scale = log2(e) / T.
  int sample_mode;
  float sample_scale;
  uint32_t sample_k0, sample_k1;
  // split-K: `splits` CTAs share one output tile; each writes its fp32 partial accumulator to `partials`
  // ([tile][split][128][BN]) and the last one to arrive (per-tile counter) sums them in split order and runs the
  // epilogue.  Deterministic: the summation order does not depend on arrival order.
  int splits;
  int kb_per_split;
  float* partials;
  int* counters;
  int pair;                            // host-side only: launch the cta_group::2 (SM pair) variant
  int dual;                            // host-side only: launch the two-issuer variant (DUAL, 128-wide tiles)
  DropKey drop;                        // residual-site dropout applied to (acc + bias) before the residual add (thr 0: off)
#ifdef MEBT_GEMM_TRACE
  long long* trace;                    // [grid][8] cycle counters (tools/gemm_bench.cu)
#endif
};

// Role-level cycle accounting for tools/gemm_bench.cu (compiled out of the library build).
#ifdef MEBT_GEMM_TRACE
#define TR_DECL long long tr_t = 0, tr_start = clock64(), tr_acc[6] = {0, 0, 0, 0, 0, 0}
#define TR_BEGIN tr_t = clock64()
#define TR_END(i) tr_acc[i] += clock64() - tr_t
#define TR_FLUSH(base, n)                                                                  \
  do {                                                                                     \
    if (p.trace != nullptr) {                                                              \
      for (int _i = 0; _i < (n); ++_i) p.trace[blockIdx.x * 8 + (base) + _i] = tr_acc[_i]; \
      p.trace[blockIdx.x * 8 + (base) + (n)] = clock64() - tr_start;                       \
    }                                                                                      \
  } while (0)
long long* g_gemm_trace = nullptr;
// absolute timestamps (SM clock) of one CTA's milestones: trace[148 * 13 + blockIdx.x * 8 + i]
#define TS(i) do { if (p.trace != nullptr) p.trace[148 * 13 + blockIdx.x * 16 + (i)] = clock64(); } while (0)
#else
#define TS(i)
#define TR_DECL
#define TR_BEGIN
#define TR_END(i)
#define TR_FLUSH(base, n)
#endif

template <int BN, int STAGES, bool PAIR = false>
struct SmemLayout {
  static constexpr int B_TILE_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // pair mode: each CTA holds half of the B tile
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;                         // epilogue staging (TMA store / load)
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_SLOTS * EPI_SLOT_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 512;                           // fp32 bias slice of the current tile
  static constexpr int TOTAL = BIAS_OFFSET + BN * 4 + 1024;                       // + alignment slack
  static_assert(TOTAL <= 232448, "shared memory budget");
};

// PAIR: the two CTAs of a cluster (an SM pair) execute ONE tcgen05.mma.cta_group::2 of shape M=256 x N=BN: each CTA
// stages its own 128 rows of A and HALF of the B tile, the leader CTA's single thread issues the MMAs, and each CTA
// receives its 128 accumulator rows in its own TMEM.  Per SM and k-block this moves A + B/2 instead of A + B
// through shared memory — at 128x256 tiles the 1-CTA form needs ~190 B/clk of shared-memory bandwidth (TMA fill +
// operand reads) against 128 B/clk available, which is what capped it at ~65 % tensor-pipe utilisation.
// Handshake: both CTAs' TMA loads complete on the LEADER's full barrier; the leader's tcgen05.commit (multicast)
// releases the stage in both CTAs and publishes the accumulator to both epilogues; both epilogues arrive on the
// leader's tmem_empty barrier.
// DUAL (128-wide tiles, one CTA per tile): TWO MMA-issuing threads (warp 1 and warp 6).  One issuer spends ~520 clk per
// k-block on its barrier wait + tcgen05.commit while a 128-wide k-block is 256 clk of tensor work (tools/mma_probe.cu), so
// the main loop is hand-off-bound.  Issuer j takes the k-blocks with (index in the tile) % 2 == j and accumulates them in
// ITS OWN TMEM accumulator; the epilogue adds the two halves.  The summation order is fixed (even k-blocks in order, odd
// k-blocks in order, one final add), so results stay bitwise reproducible - two issuers on ONE accumulator would make the
// order depend on their interleaving.
template <int BN, bool A_MN, bool B_MN, int STAGES, bool PAIR, bool DUAL = false>
__global__ void __launch_bounds__(DUAL ? GEMM_THREADS + 32 : GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_aux,
                 const __grid_constant__ CUtensorMap tma_in, const GemmParams p) {
  using L = SmemLayout<BN, STAGES, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* epi_bar = tmem_empty_bar + 2;                       // [EPI_SLOTS]: staged input operand landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(epi_bar + EPI_SLOTS);

  if (threadIdx.x == 0) TS(0);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work decomposition: plain = one 128xBN tile (x split) per item, strided by the grid;
  //                     pair  = one 256xBN tile pair per item, strided by the number of clusters
  const int pair_rank = PAIR ? int(blockIdx.x & 1) : 0;
  const int pair_m = (p.num_m_blocks + 1) >> 1;
  const int num_work = PAIR ? pair_m * p.num_n_blocks : p.num_m_blocks * p.num_n_blocks * p.splits;
  const int work0 = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int work_stride = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);
  // Rasterisation: bands of RASTER_M row blocks (2048 rows), inside a band the row block runs fastest, then the
  // column block.  The tiles in flight at any time then touch one or two bands of A (<= 16 MiB at K = 4096) and the
  // whole of B, which stay in L2; sweeping all of M per column block streamed A from HBM once per column block
  // (16x at 131072 x 4096 x 1024, which made that GEMM HBM-bound).
  auto raster = [&](int t, int rows_m, int& mi, int& ni) {
    const int band_tiles = RASTER_M * p.num_n_blocks;
    const int band = t / band_tiles;
    const int in_band = t - band * band_tiles;
    const int band_rows = min(RASTER_M, rows_m - band * RASTER_M);
    ni = in_band / band_rows;
    mi = band * RASTER_M + (in_band - ni * band_rows);
  };
  auto tile_origin = [&](int work, int& m0, int& n0, int& tile) {
    int mi, ni;
    if (PAIR) {
      tile = work;
      raster(work, pair_m, mi, ni);
      m0 = (2 * mi + pair_rank) * BM;
    } else {
      tile = work / p.splits;
      raster(tile, p.num_m_blocks, mi, ni);
      m0 = mi * BM;
    }
    n0 = ni * BN;
  };
  int* split_flag = reinterpret_cast<int*>(tmem_ptr_smem + 1);
  static_assert(!DUAL || (!PAIR && BN <= 128 && STAGES % 2 == 0), "DUAL: one CTA per tile, two accumulator halves, even ring");
  constexpr uint32_t ACC_STRIDE = DUAL ? 2 * BN : BN;      // TMEM columns per accumulator buffer
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;           // 128, 256 or 512 (power of two >= 32)

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tma_a);
    prefetch_tensormap(&tma_b);
    prefetch_tensormap(&tma_c);
    for (int s = 0; s < EPI_SLOTS; ++s) mbar_init(&epi_bar[s], 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], DUAL ? 2 : 1);       // DUAL: each issuer commits its half
      mbar_init(&tmem_empty_bar[s], PAIR ? 256 : 128);   // pair: the leader waits for both CTAs' epilogues
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) { tmem_alloc_2sm(tmem_ptr_smem, TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_ptr_smem, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();        // the peer signals this CTA's barriers: they must exist first
  // shared::cluster addresses of the leader's barriers (identity for the leader itself)
  uint32_t full_bar_leader = 0, tmem_empty_leader = 0;
  if constexpr (PAIR) {
    full_bar_leader = mapa_u32(smem_u32(full_bar), 0);
    tmem_empty_leader = mapa_u32(smem_u32(tmem_empty_bar), 0);
  }
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) TS(1);
  griddep_wait();
  if (threadIdx.x == 0) TS(2);

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TR_DECL;
      for (int work = work0; work < num_work; work += work_stride) {
        int m0, n0, tile;
        tile_origin(work, m0, n0, tile);
        const int split = PAIR ? 0 : work % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          TR_BEGIN;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          TR_END(0);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + A_TILE_BYTES;
          if constexpr (PAIR) {
            // both CTAs' bytes complete on the leader's barrier; the leader arms it for the pair
            const uint32_t bar = full_bar_leader + stage * 8;
            if (pair_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d_2sm(sA, &tma_a, bar, kb * BK, m0);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d_2sm(sA + i * (BK * 128), &tma_a, bar, m0 + i * 64, kb * BK);
            }
            if constexpr (!B_MN) {                       // this CTA's half of the B tile: rows n0 + rank * BN/2 ...
              tma_load_2d_2sm(sB, &tma_b, bar, kb * BK, n0 + pair_rank * (BN / 2));
            } else {
#pragma unroll
              for (int i = 0; i < BN / 128; ++i)
                tma_load_2d_2sm(sB + i * (BK * 128), &tma_b, bar, n0 + (pair_rank * (BN / 128) + i) * 64, kb * BK);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d(sA, &tma_a, &full_bar[stage], kb * BK, m0);               // box [64 k][128 m]
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)                                      // box [64 m][64 k]
              tma_load_2d(sA + i * (BK * 128), &tma_a, &full_bar[stage], m0 + i * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sB, &tma_b, &full_bar[stage], kb * BK, n0);               // box [64 k][BN n]
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)                                      // box [64 n][64 k]
              tma_load_2d(sB + i * (BK * 128), &tma_b, &full_bar[stage], n0 + i * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      TR_FLUSH(0, 1);
    }
  } else if (DUAL && (warp == 1 || warp == 6)) {
    // ================= the two MMA issuers of the DUAL variant =================
    if (lane == 0) {
      const uint32_t j = warp == 6 ? 1u : 0u;                 // this issuer's k-block parity and accumulator half
      const uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0) &
                             (p.fp16_in ? ~((1u << 7) | (1u << 10)) : ~0u);
      constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), A_MN ? BK * 128 : 16);
      const uint32_t b_off = uint32_t(A_TILE_BYTES >> 4) + ((uint32_t((B_MN ? BK * 128 : 16) >> 4) - uint32_t((A_MN ? BK * 128 : 16) >> 4)) << 16);
      const uint32_t nkb = uint32_t(p.num_k_blocks);          // no split-K in this variant
      uint32_t cnt = 0;                                       // k-blocks of this CTA's earlier tiles: ring position of k-block 0
      int it = 0;
      for (int work = work0; work < num_work; work += work_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);       // the epilogue drained both halves of this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc) * ACC_STRIDE + j * uint32_t(BN);
        // Issuer j takes the k-blocks whose RING position g is of parity j: with an even number of stages a stage then always
        // belongs to the same issuer, who waits for its phases strictly in order.  (Splitting by the index inside the tile let
        // one issuer reach a stage a whole ring ahead of the other: a parity wait for phase k + 1 on a barrier still in an
        // incomplete phase k succeeds at once - stale operands, seen as rare wrong results / launch failures under stress.)
        for (uint32_t i = ((cnt & 1u) == j) ? 0u : 1u; i < nkb; i += 2) {
          const uint32_t g = cnt + i;
          const uint32_t stage = g % uint32_t(STAGES), phase = (g / uint32_t(STAGES)) & 1u;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * uint32_t(L::STAGE_BYTES >> 4);
          umma_bf16_ss_x4<false>(tmem_d, a_lo, a_lo + b_off, A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4,
                                 B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4, desc_hi, desc_hi, idesc, i >= 2 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);                     // the producer may refill the stage once these MMAs retire
        }
        umma_commit(&tmem_full_bar[acc]);                     // this half is complete (the barrier counts both issuers)
        cnt += nkb;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && pair_rank == 0 && !DUAL) {       // pair mode: only the leader CTA issues
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0) &
                             (p.fp16_in ? ~((1u << 7) | (1u << 10)) : ~0u);            // A/B format: 1 = bf16, 0 = fp16
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      TR_DECL;
      static_assert(BK / UMMA_K == 4, "umma_bf16_ss_x4 issues the four K = 16 steps of a 64-wide k-block");
      constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem), A_MN ? BK * 128 : 16);
      // B tile: A_TILE_BYTES behind A in the stage; its LBO field may differ from A's
      const uint32_t b_off = uint32_t(A_TILE_BYTES >> 4) + ((uint32_t((B_MN ? BK * 128 : 16) >> 4) - uint32_t((A_MN ? BK * 128 : 16) >> 4)) << 16);
      for (int work = work0; work < num_work; work += work_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int split = PAIR ? 0 : work % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
        TR_BEGIN;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);     // epilogue drained this accumulator
        TR_END(1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc) * ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          TR_BEGIN;
          mbar_wait(&full_bar[stage], phase);
          TR_END(0);
          if (it == 0 && kb == kb0) TS(3);
          tc_fence_after();
          // descriptor low words of the stage's first K = 16 step (see umma_bf16_ss_x4); per step:
          // K-major : advance 16 elements (32 B) inside the 128 B swizzle row; LBO field 1, SBO = 8 rows * 128 B
          // MN-major: advance 16 k-rows (2048 B); LBO = next 64-wide MN atom (BK rows * 128 B), SBO = 8 k-rows
          const uint32_t a_lo = a_lo0 + uint32_t(stage) * uint32_t(L::STAGE_BYTES >> 4);
          umma_bf16_ss_x4<PAIR>(tmem_d, a_lo, a_lo + b_off, A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4,
                                B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4, desc_hi, desc_hi, idesc,
                                kb > kb0 ? 1u : 0u);
          if constexpr (PAIR) {
            umma_commit_2sm_mc(&empty_bar[stage], 3);                      // frees the stage in both CTAs
            if (kb == kb1 - 1) umma_commit_2sm_mc(&tmem_full_bar[acc], 3);  // accumulator ready in both CTAs
          } else {
            umma_commit(&empty_bar[stage]);                  // smem slot reusable once these MMAs retire
            if (kb == kb1 - 1) umma_commit(&tmem_full_bar[acc]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      TS(4);
      TR_FLUSH(2, 2);
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    TR_DECL;
    const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
    uint8_t* slots = smem + L::EPI_OFFSET;
    uint64_t* in_full = epi_bar;
    const bool epi_t0 = threadIdx.x == 64;                    // issues the epilogue's TMA traffic
    const bool tma_epi = p.splits == 1;
    const bool f32 = p.out_fp32 != 0;
    const bool aux_out = p.gelu && p.aux != nullptr;
    const bool k_gelu = p.gelu != 0;
    const int k_in = p.in_kind;
    const int och_per_tile = f32 ? BN / 32 : BN / 64;         // 128-byte output chunks per tile row
    const int my_tiles = work0 < num_work ? (num_work - work0 + work_stride - 1) / work_stride : 0;
    const int total_och = my_tiles * och_per_tile;
    // global coordinates of output chunk g (chunks are numbered across the CTA's tiles)
    auto och_coords = [&](int g, int& r0, int& c0) {
      int m0, n0, tile;
      tile_origin(work0 + (g / och_per_tile) * work_stride, m0, n0, tile);
      r0 = m0;
      c0 = n0 + (g % och_per_tile) * (f32 ? 32 : 64);
    };
    auto issue_in = [&](int g) {                              // epi_t0: fetch the input operand of chunk g into its slot
      int r0, c0;
      och_coords(g, r0, c0);
      mbar_arrive_expect_tx(&in_full[g & 3], EPI_SLOT_BYTES);
      tma_load_2d(slots + (g & 3) * EPI_SLOT_BYTES, &tma_in, &in_full[g & 3], c0, r0);
    };
    if (tma_epi && k_in != 0 && epi_t0)
      for (int g = 0; g < 3 && g < total_och; ++g) issue_in(g);
    int it = 0;
    for (int work = work0; work < num_work; work += work_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      int m0, n0, tile;
      tile_origin(work, m0, n0, tile);
      const int split = PAIR ? 0 : work % p.splits;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const bool k_drop = p.drop.thr != 0;
      const uint32_t drop_rk = drop_row_key(p.drop, uint32_t(row));
      // keep factors of one 32-column unit of this thread's row (same (seed, site, row, column) function as
      // dropout_rows_kernel, so that backward regenerates the mask the fused forward used)
      auto apply_drop = [&](float (&v)[32], int col0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float f0, f1;
          drop_pair(p.drop, drop_rk, uint32_t((col0 >> 1) + j), f0, f1);
          v[2 * j] *= f0; v[2 * j + 1] *= f1;
        }
      };
      // stage the tile's bias slice in shared memory once (one L2 round trip per tile instead of one per chunk)
      float* s_bias = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
      TR_BEGIN;
      if (p.bias != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");            // the previous tile's readers are done
        const int t = threadIdx.x - 64;
        for (int cidx = t; cidx < BN; cidx += 128) s_bias[cidx] = __ldg(p.bias + n0 + cidx);
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      TR_END(5);
      TR_BEGIN;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      TR_END(0);
      if (threadIdx.x == 64 && it == 0) TS(5);
      tc_fence_after();
      const float* my_partials = nullptr;
      if (p.splits > 1) {
        // park the raw accumulator, release TMEM, and find out whether this CTA completes the tile
        // thread-major layout: float4 slot ((c * 8 + j) * 128 + t) -> every warp store / load is 512 contiguous bytes
        float4* part = reinterpret_cast<float4*>(p.partials + (size_t(tile) * p.splits + split) * BM * BN) + (q * 32 + lane);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * ACC_STRIDE + uint32_t(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(part + (c * 8 + j) * 128, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                        __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {                               // first epilogue thread
          const int old = atomicAdd(p.counters + tile, 1);
          const int last = old == p.splits - 1;
          if (last) p.counters[tile] = 0;                      // ready for the next launch
          *split_flag = last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*split_flag == 0) continue;
        __threadfence();
        my_partials = p.partials + size_t(tile) * p.splits * BM * BN + size_t(q * 32 + lane) * 4;
      }
      if (p.sample_mode) {
        // categorical draw by the Gumbel-max rule: argmax_v (x_v / T + G_v), G = -ln(-ln u); in the log2 domain and up
        // to a positive factor, argmax_v (x_v log2(e) / T - lg2(-lg2 u_v)).  The logits never leave the accumulator.
        const uint32_t t_acc = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * ACC_STRIDE;
        const uint32_t rk = mix32(p.sample_k0 ^ mix32(uint32_t(row) + p.sample_k1));
        float bestv = -INFINITY;
        int besti = 0x7fffffff;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_acc + uint32_t(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t col = uint32_t(n0 + c * 32 + j);
            const uint32_t hsh = mix32(rk + col * 0x9E3779B9u);
            const float u = fmaf(float(hsh >> 9), 1.1920928955078125e-07f, 5.9604644775390625e-08f);   // (k + 0.5) 2^-23: in [2^-24, 1 - 2^-24], exact
            const float v = fmaf(__uint_as_float(r[j]), p.sample_scale, -lg2_approx(-lg2_approx(u)));
            if (v > bestv) { bestv = v; besti = int(col); }
          }
        }
        tc_fence_before();
        if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
        else mbar_arrive(&tmem_empty_bar[acc]);
        if (row_ok && besti != 0x7fffffff) {
          const uint32_t ub = __float_as_uint(bestv);
          const uint32_t key = ~((ub & 0x80000000u) ? ~ub : (ub | 0x80000000u));      // larger value -> smaller key
          atomicMin(p.argmin_out + row, (static_cast<unsigned long long>(key) << 32) | uint32_t(besti));
        }
        continue;
      }
      if (p.argmin_out != nullptr) {
        // K9 epilogue: nothing is stored but each row's best (distance, code) over this tile's columns
        const uint32_t t_acc = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * ACC_STRIDE;
        const float zq = row_ok ? __ldg(p.row_sq + row) : 0.f;
        float bestd = INFINITY;
        int besti = 0x7fffffff;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_acc + uint32_t(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            // the reference's expression order: (|z|^2 - 2 z.e) + |e|^2  (codebook.py:53-55)
            const float d = (zq - 2.0f * __uint_as_float(r[j])) + s_bias[c * 32 + j];
            if (d < bestd) { bestd = d; besti = n0 + c * 32 + j; }      // columns ascending: the first minimum wins
          }
        }
        tc_fence_before();
        if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
        else mbar_arrive(&tmem_empty_bar[acc]);
        if (row_ok && besti != 0x7fffffff) {
          const uint32_t u = __float_as_uint(bestd);
          const uint32_t key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
          atomicMin(p.argmin_out + row, (static_cast<unsigned long long>(key) << 32) | uint32_t(besti));
        }
        continue;
      }
      // One 32-column chunk of this thread's row: bias, GELU / GELU', residual, store.
      auto finish_chunk = [&](float (&v)[32], int c) {
        const int col = n0 + c * 32;
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        if (k_drop) apply_drop(v, col);
        if (p.gelu) {
          if (p.aux != nullptr && row_ok) {      // keep the pre-activation for the backward pass
            uint4* a4 = reinterpret_cast<uint4*>(p.aux + size_t(row) * p.ldaux + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              a4[j] = o;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 g = gelu_erf_x2(make_float2(v[2 * j], v[2 * j + 1]));
            v[2 * j] = g.x; v[2 * j + 1] = g.y;
          }
        }
        if (p.dgelu && row_ok) {
          const uint4* a4 = reinterpret_cast<const uint4*>(p.aux + size_t(row) * p.ldaux + col);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = __ldg(a4 + j);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 g = gelu_erf_grad_x2(unpack_bf16x2(w[t]));
              v[8 * j + 2 * t] *= g.x; v[8 * j + 2 * t + 1] *= g.y;
            }
          }
        }
        if (row_ok) {
          if (p.residual != nullptr) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + size_t(row) * p.ldres + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = __ldg(r4 + j);
              const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z),
                           d = unpack_bf16x2(u.w);
              v[8 * j + 0] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b.x; v[8 * j + 3] += b.y;
              v[8 * j + 4] += c2.x; v[8 * j + 5] += c2.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
            }
          }
          if (p.out_fp32) {
            float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + size_t(row) * p.ldc + col);
            if (p.accumulate) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 o = o4[j];
                o.x += v[4 * j + 0]; o.y += v[4 * j + 1]; o.z += v[4 * j + 2]; o.w += v[4 * j + 3];
                o4[j] = o;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
          } else {
            uint4* o4 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + size_t(row) * p.ldc + col);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              o4[j] = o;
            }
          }
        }
      };
      // One 32-column unit of this thread's row through the staging ring (TMA epilogue).
      float dacc = 0.f;                                        // in_kind 3: running dot product of the current head
      auto staged_unit = [&](const uint32_t (&r)[32], int u) {
        const int gu = it * (BN / 32) + u;                     // unit counter across this CTA's tiles
        const int g = f32 ? gu : gu >> 1;                      // output chunk it belongs to
        const int half = f32 ? 0 : (u & 1);
        const int s_c = aux_out ? (2 * g) & 3 : g & 3;
        const int sw = lane & 7;
        uint8_t* row_c = slots + s_c * EPI_SLOT_BYTES + (q * 32 + lane) * 128;
        if (threadIdx.x == 64 && it == 0 && u < 2) TS(8 + 4 * u);
        if ((f32 || half == 0) && k_in != 0) {
          TR_BEGIN;
          mbar_wait(&in_full[s_c], (g >> 2) & 1);              // this chunk's residual / pre-activation has landed
          TR_END(1);
        }
        if (threadIdx.x == 64 && it == 0 && u < 2) TS(9 + 4 * u);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        // every shared-memory operand of the unit is fetched up front (LDS, all in flight together); the option
        // switches sit outside the element loops so that each variant is straight-line code
        const uint32_t row_c_s = smem_u32(row_c);
        uint4 in4[4];
        if (k_in != 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) in4[j] = lds128(row_c_s + uint32_t(((half * 4 + j) ^ sw) << 4));
        }
        if (p.bias != nullptr) {
          const uint32_t b_s = smem_u32(s_bias + u * 32);
          uint4 b4[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) b4[j] = lds128(b_s + j * 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            v[4 * j + 0] += __uint_as_float(b4[j].x); v[4 * j + 1] += __uint_as_float(b4[j].y);
            v[4 * j + 2] += __uint_as_float(b4[j].z); v[4 * j + 3] += __uint_as_float(b4[j].w);
          }
        }
        if (k_drop) apply_drop(v, n0 + u * 32);
        if (k_gelu) {
          if (aux_out) {                                       // keep the pre-activation for the backward pass
            const uint32_t row_a_s = smem_u32(slots + ((2 * g + 1) & 3) * EPI_SLOT_BYTES + (q * 32 + lane) * 128);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
              o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
              sts128(row_a_s + uint32_t(((half * 4 + j) ^ sw) << 4), o);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 gl = gelu_erf_x2(make_float2(v[2 * j], v[2 * j + 1]));
            v[2 * j] = gl.x; v[2 * j + 1] = gl.y;
          }
        }
        if (k_in == 1) {                                       // residual
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {in4[j].x, in4[j].y, in4[j].z, in4[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 x = unpack_bf16x2(w[t]);
              v[8 * j + 2 * t] += x.x; v[8 * j + 2 * t + 1] += x.y;
            }
          }
        } else if (k_in == 2) {                                // gelu'(pre-activation)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {in4[j].x, in4[j].y, in4[j].z, in4[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 gg = gelu_erf_grad_x2(unpack_bf16x2(w[t]));
              v[8 * j + 2 * t] *= gg.x; v[8 * j + 2 * t + 1] *= gg.y;
            }
          }
        } else if (k_in == 3) {                                // delta = rowsum_head(C .* O)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {in4[j].x, in4[j].y, in4[j].z, in4[j].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 x = unpack_bf16x2(w[t]);
              dacc = fmaf(v[8 * j + 2 * t], x.x, dacc);
              dacc = fmaf(v[8 * j + 2 * t + 1], x.y, dacc);
            }
          }
          if (half == 1) {                                     // one 64-column chunk = one head of this row
            if (row_ok) {
              const int bq = row / p.delta_nq;
              const int hd = (n0 + u * 32) >> 6;
              p.delta_out[(size_t(bq) * p.delta_h + hd) * p.delta_nq + (row - bq * p.delta_nq)] = dacc;
            }
            dacc = 0.f;
          }
        }
        if (f32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(row_c_s + uint32_t((j ^ sw) << 4), make_uint4(__float_as_uint(v[4 * j + 0]), __float_as_uint(v[4 * j + 1]),
                                                                 __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            sts128(row_c_s + uint32_t(((half * 4 + j) ^ sw) << 4), o);
          }
        }
        if (threadIdx.x == 64 && it == 0 && u < 2) TS(10 + 4 * u);
        if (f32 || half == 1) {
          // One barrier per chunk: behind it every row of the slot is written (and fenced towards the async proxy), and
          // the slot the NEXT chunk writes has been read out by its previous store (epi_t0 checks before arriving).
          TR_BEGIN;
          fence_proxy_async_smem();
          if (epi_t0 && k_in == 0) { if (aux_out) tma_store_wait_read<0>(); else tma_store_wait_read<2>(); }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          TR_END(2);
          TR_BEGIN;
          if (epi_t0) {
            int r0, c0;
            och_coords(g, r0, c0);
            if (p.accumulate) tma_reduce_add_2d(&tma_c, slots + s_c * EPI_SLOT_BYTES, c0, r0);
            else tma_store_2d(&tma_c, slots + s_c * EPI_SLOT_BYTES, c0, r0);
            if (aux_out) tma_store_2d(&tma_aux, slots + ((2 * g + 1) & 3) * EPI_SLOT_BYTES, c0, r0);
            tma_store_commit();
            if (k_in != 0 && g + 3 < total_och) {
              tma_store_wait_read<1>();                        // chunk g-1's store has released the slot chunk g+3 reuses
              issue_in(g + 3);
            }
          }
          TR_END(3);
        }
        if (threadIdx.x == 64 && it == 0 && u < 2) TS(11 + 4 * u);
      };
      if (my_partials == nullptr) {
        // software pipeline over the accumulator: the TMEM load of unit u+1 is in flight while unit u is finished
        const uint32_t t_acc = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * ACC_STRIDE;
        uint32_t ra[32], rb[32];
        // DUAL: the odd k-blocks' half of the accumulator, BN columns further, is added in (fixed order: reproducible)
        auto add_half = [&](uint32_t (&r)[32], int c) {
          if constexpr (DUAL) {
            uint32_t r2[32];
            tmem_ld_32x32(t_acc + uint32_t(BN + c * 32), r2);
            tmem_ld_wait_regs(r2);
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          }
        };
        tmem_ld_32x32(t_acc, ra);
#pragma unroll 1
        for (int c = 0; c < BN / 32; c += 2) {
          TR_BEGIN;
          tmem_ld_wait_regs(ra);
          TR_END(4);
          add_half(ra, c);
          tmem_ld_32x32(t_acc + uint32_t((c + 1) * 32), rb);
          staged_unit(ra, c);
          TR_BEGIN;
          tmem_ld_wait_regs(rb);
          TR_END(4);
          add_half(rb, c + 1);
          if (c + 2 < BN / 32) tmem_ld_32x32(t_acc + uint32_t((c + 2) * 32), ra);
          else {                                   // accumulator fully read: hand it back before the last stores
            tc_fence_before();
            if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_leader + acc * 8);
            else mbar_arrive(&tmem_empty_bar[acc]);
          }
          staged_unit(rb, c + 1);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          for (int sp = 0; sp < p.splits; ++sp) {              // fixed order: bitwise reproducible
            const float4* i4 = reinterpret_cast<const float4*>(my_partials + size_t(sp) * BM * BN) + c * 8 * 128;
            float4 t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldcg(i4 + j * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[4 * j + 0] += t[j].x; v[4 * j + 1] += t[j].y; v[4 * j + 2] += t[j].z; v[4 * j + 3] += t[j].w;
            }
          }
          finish_chunk(v, c);
        }
      }
    }
    if (threadIdx.x == 64) TS(6);
    if (tma_epi && epi_t0) tma_store_wait_read<0>();          // the stores have read their slots (the writes drain with the grid)
    if (threadIdx.x == 64) TR_FLUSH(5, 1);
#ifdef MEBT_GEMM_TRACE
    if (threadIdx.x == 64 && p.trace != nullptr) for (int _i = 1; _i < 6; ++_i) p.trace[148 * 8 + blockIdx.x * 5 + _i - 1] = tr_acc[_i];
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();        // the peer may still be signalling this CTA's barriers
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
    if (lane == 0) TS(7);
  }
}

template <int BN, bool A_MN, bool B_MN, bool PAIR, bool DUAL = false>
int launch_gemm_impl(const void* A, const void* B, const GemmParams& p, int lda, int ldb, cudaStream_t stream) {
  // 64 KiB of the 227 KiB go to the epilogue staging ring; the rest is the operand ring (DUAL: an even number of stages)
  constexpr int STAGES = PAIR ? (BN == 256 ? 5 : 6) : ((BN == 256) ? 3 : (BN == 128 ? (DUAL ? 4 : 5) : 6));
  using L = SmemLayout<BN, STAGES, PAIR>;
  CUtensorMap ta, tb;
  int rc;
  if (!A_MN) rc = get_tensor_map_2d(&ta, A, 2, uint64_t(p.K), uint64_t(p.M), uint64_t(lda) * 2, BK, BM);
  else       rc = get_tensor_map_2d(&ta, A, 2, uint64_t(p.M), uint64_t(p.K), uint64_t(lda) * 2, 64, BK);
  if (rc) return rc;
  // K-major B in pair mode is fetched as two half-height boxes (one per CTA of the cluster)
  if (!B_MN) rc = get_tensor_map_2d(&tb, B, 2, uint64_t(p.K), uint64_t(p.N), uint64_t(ldb) * 2, BK, PAIR ? BN / 2 : BN);
  else       rc = get_tensor_map_2d(&tb, B, 2, uint64_t(p.N), uint64_t(p.K), uint64_t(ldb) * 2, 64, BK);
  if (rc) return rc;
  // epilogue operands: 128-byte-wide boxes of the tile's 128 rows (one per chunk)
  CUtensorMap tc = ta, taux = ta, tin = ta;
  if (p.splits == 1 && p.argmin_out == nullptr) {
    if (p.out_fp32) rc = get_tensor_map_2d(&tc, p.C, 4, uint64_t(p.N), uint64_t(p.M), uint64_t(p.ldc) * 4, 32, 128);
    else            rc = get_tensor_map_2d(&tc, p.C, 2, uint64_t(p.N), uint64_t(p.M), uint64_t(p.ldc) * 2, 64, 128);
    if (rc) return rc;
    if (p.gelu && p.aux != nullptr) {
      rc = get_tensor_map_2d(&taux, p.aux, 2, uint64_t(p.N), uint64_t(p.M), uint64_t(p.ldaux) * 2, 64, 128);
      if (rc) return rc;
    }
    if (p.in_kind == 1) rc = get_tensor_map_2d(&tin, p.residual, 2, uint64_t(p.N), uint64_t(p.M), uint64_t(p.ldres) * 2, 64, 128);
    if (p.in_kind == 2) rc = get_tensor_map_2d(&tin, p.aux, 2, uint64_t(p.N), uint64_t(p.M), uint64_t(p.ldaux) * 2, 64, 128);
    if (p.in_kind == 3) rc = get_tensor_map_2d(&tin, p.aux, 2, uint64_t(p.N), uint64_t(p.M), uint64_t(p.ldaux) * 2, 64, 128);
    if (rc) return rc;
  }
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, STAGES, PAIR, DUAL>;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  }
  LaunchScope ls(FAM_GEMM, 2.0 * double(p.M) * double(p.N) * double(p.K), stream);
  if constexpr (PAIR) {
    const int pairs = ((p.num_m_blocks + 1) / 2) * p.num_n_blocks;
    int clusters = grid_cap() / 2;
    if (clusters > pairs) clusters = pairs;
    MEBT_CUDA_OK(launch_pdl_cluster2(kern, dim3(2 * clusters), dim3(GEMM_THREADS), L::TOTAL, stream, ta, tb, tc, taux, tin, p));
  } else {
    const int tiles = p.num_m_blocks * p.num_n_blocks * p.splits;
    const int grid = tiles < grid_cap() ? tiles : grid_cap();
    const cudaError_t le = launch_pdl(kern, dim3(grid), dim3(DUAL ? GEMM_THREADS + 32 : GEMM_THREADS), L::TOTAL, stream, ta, tb, tc, taux, tin, p);
    if (le != cudaSuccess) {
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, kern);
      set_last_error("gemm launch failed: %s (regs %d, max threads %d, static smem %zu, dynamic smem %d, max dynamic %d)",
                     cudaGetErrorString(le), fa.numRegs, fa.maxThreadsPerBlock, fa.sharedSizeBytes, L::TOTAL,
                     fa.maxDynamicSharedSizeBytes);
      return MEBT_ERR_CUDA;
    }
  }
  MEBT_LAUNCH_OK("gemm_bf16_kernel");
  return MEBT_OK;
}

template <int BN, bool A_MN, bool B_MN>
int launch_gemm(const void* A, const void* B, const GemmParams& p, int lda, int ldb, cudaStream_t stream) {
  if constexpr (BN >= 128) {
    if (p.pair) return launch_gemm_impl<BN, A_MN, B_MN, true>(A, B, p, lda, ldb, stream);
  }
  if constexpr (BN == 128) {
    if (p.dual) return launch_gemm_impl<BN, A_MN, B_MN, false, true>(A, B, p, lda, ldb, stream);
  }
  return launch_gemm_impl<BN, A_MN, B_MN, false>(A, B, p, lda, ldb, stream);
}

template <int BN>
int dispatch_major(int a_mn, int b_mn, const void* A, const void* B, const GemmParams& p, int lda, int ldb,
                   cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(A, B, p, lda, ldb, stream);
  if (!a_mn && b_mn) return launch_gemm<BN, false, true>(A, B, p, lda, ldb, stream);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false>(A, B, p, lda, ldb, stream);
  return launch_gemm<BN, true, true>(A, B, p, lda, ldb, stream);
}

}  // namespace

int gemm_bf16_aux(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                  int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                  cudaStream_t stream);
int gemm_bf16_drop(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                   int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                   const DropKey* drop, cudaStream_t stream);
int gemm_bf16_ex(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                 int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                 const DropKey* drop, float* delta_out, int delta_nq, int delta_h, cudaStream_t stream);

// Split-K scratch: the only device memory the library owns (partials are consumed inside the launch that wrote
// them; the per-tile counters are returned to zero by the CTA that completes the tile).  Launches that use it are
// ordered by the caller's stream, like every other launch; it is sized once for the largest split-K problem seen.
struct SplitKWorkspace { float* partials; int* counters; cudaStream_t stream; };
// one scratch area per stream (the training engine overlaps weight-gradient GEMMs on a side stream)
static SplitKWorkspace* splitk_workspace(size_t need_bytes, int tiles, cudaStream_t stream) {
  static SplitKWorkspace pool[4] = {};
  static int used = 0;
  constexpr size_t kMax = size_t(96) << 20;
  if (need_bytes > kMax || tiles > 4096) return nullptr;
  for (int i = 0; i < used; ++i)
    if (pool[i].stream == stream) return &pool[i];
  if (used == 4) return nullptr;
  SplitKWorkspace& ws = pool[used];
  if (cudaMalloc(&ws.partials, kMax) != cudaSuccess) { ws.partials = nullptr; return nullptr; }
  if (cudaMalloc(&ws.counters, 4096 * sizeof(int)) != cudaSuccess) { ws.counters = nullptr; return nullptr; }
  cudaMemset(ws.counters, 0, 4096 * sizeof(int));
  ws.stream = stream;
  ++used;
  return &ws;
}

int gemm_bf16(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
              int K, const float* bias, const void* residual, int ldres, int flags, cudaStream_t stream) {
  return gemm_bf16_aux(A, lda, a_mn, B, ldb, b_mn, C, ldc, M, N, K, bias, residual, ldres, nullptr, 0, flags, stream);
}

int gemm_bf16_aux(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                  int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                  cudaStream_t stream) {
  return gemm_bf16_drop(A, lda, a_mn, B, ldb, b_mn, C, ldc, M, N, K, bias, residual, ldres, aux, ldaux, flags, nullptr,
                        stream);
}

// C = drop(A B^T + bias) + residual: the epilogue multiplies (acc + bias) by the keep factors of `drop` (nullptr or
// thr == 0: none) before the residual add - nn.Dropout on the proj / mlp outputs (gpt.py:140,154) without a launch.
int gemm_bf16_drop(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                   int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                   const DropKey* drop, cudaStream_t stream) {
  return gemm_bf16_ex(A, lda, a_mn, B, ldb, b_mn, C, ldc, M, N, K, bias, residual, ldres, aux, ldaux, flags, drop, nullptr, 0,
                      0, stream);
}

// + delta: C = A B^T as bf16 and, from the same fp32 accumulators, delta[b, h, q] = sum over the 64 columns of head h
// of C[(b, q), :] * O[(b, q), :] with O passed in `aux` - the row sums the attention backward needs (dO = C), without
// the separate pass over dO and O.
int gemm_bf16_ex(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                 int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                 const DropKey* drop, float* delta_out, int delta_nq, int delta_h, cudaStream_t stream) {
  MEBT_REQUIRE(M > 0 && N > 0 && K > 0, MEBT_ERR_SHAPE, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  MEBT_REQUIRE(N % 64 == 0, MEBT_ERR_SHAPE, "gemm: N=%d must be a multiple of 64", N);
  MEBT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, MEBT_ERR_SHAPE, "gemm: lda/ldb must be multiples of 8 elements");
  const bool out_fp32 = (flags & MEBT_GEMM_OUT_FP32) != 0;
  MEBT_REQUIRE(ldc % (out_fp32 ? 4 : 8) == 0, MEBT_ERR_SHAPE, "gemm: ldc=%d breaks 16-byte row alignment", ldc);
  MEBT_REQUIRE(!(flags & MEBT_GEMM_ACCUMULATE) || out_fp32, MEBT_ERR_UNSUPPORTED,
               "gemm: accumulate needs fp32 output");
  MEBT_REQUIRE(residual == nullptr || ldres % 8 == 0, MEBT_ERR_SHAPE, "gemm: ldres must be a multiple of 8");
  GemmParams p;
  p.M = M; p.N = N; p.K = K;
  p.bias = bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.ldres = ldres;
  p.C = C; p.ldc = ldc;
  p.gelu = (flags & MEBT_GEMM_GELU) ? 1 : 0;
  p.out_fp32 = out_fp32 ? 1 : 0;
  p.accumulate = (flags & MEBT_GEMM_ACCUMULATE) ? 1 : 0;
  p.aux = static_cast<__nv_bfloat16*>(aux);
  p.ldaux = ldaux;
  p.dgelu = (flags & MEBT_GEMM_DGELU) ? 1 : 0;
  MEBT_REQUIRE(!p.dgelu || (aux != nullptr && !p.gelu), MEBT_ERR_SHAPE, "gemm: DGELU needs the pre-activation in aux");
  MEBT_REQUIRE(aux == nullptr || ldaux % 8 == 0, MEBT_ERR_SHAPE, "gemm: ldaux must be a multiple of 8");
  MEBT_REQUIRE(!(p.dgelu && residual != nullptr) && !(out_fp32 && (p.dgelu || residual != nullptr)), MEBT_ERR_UNSUPPORTED,
               "gemm: the epilogue reads one bf16 operand (residual or gelu' input) and only with bf16 output");
  p.in_kind = (residual != nullptr && !(flags & MEBT_GEMM_INTERNAL_ARGMIN)) ? 1 : (p.dgelu ? 2 : 0);
  p.delta_out = delta_out; p.delta_nq = delta_nq; p.delta_h = delta_h;
  p.fp16_in = 0; p.argmin_out = nullptr; p.row_sq = nullptr;
  p.sample_mode = 0; p.sample_scale = 0.f; p.sample_k0 = p.sample_k1 = 0u;
  if (flags & MEBT_GEMM_INTERNAL_SAMPLE) {          // gemm_bf16_sample below: C carries the packed output
    MEBT_REQUIRE(g_sample_req.active, MEBT_ERR_UNSUPPORTED, "gemm: internal flag");
    p.sample_mode = 1;
    p.sample_scale = g_sample_req.scale;
    p.sample_k0 = g_sample_req.k0; p.sample_k1 = g_sample_req.k1;
    p.argmin_out = static_cast<unsigned long long*>(C);
    p.C = nullptr;
    flags |= MEBT_GEMM_NO_SPLITK;
  }
  if (flags & MEBT_GEMM_INTERNAL_ARGMIN) {          // gemm_f16_argmin below: C carries the packed output, residual the row norms
    p.fp16_in = 1;
    p.argmin_out = static_cast<unsigned long long*>(C);
    p.row_sq = reinterpret_cast<const float*>(residual);
    p.residual = nullptr;
    p.C = nullptr;
    flags |= MEBT_GEMM_NO_SPLITK;
  }
  if (delta_out != nullptr) {
    MEBT_REQUIRE(aux != nullptr && residual == nullptr && !p.gelu && !p.dgelu && !out_fp32 && bias == nullptr &&
                 delta_nq > 0 && M % delta_nq == 0 && N == 64 * delta_h, MEBT_ERR_UNSUPPORTED,
                 "gemm: the delta epilogue needs O in aux, a plain bf16 output of width 64 * H and whole batches of rows");
    p.in_kind = 3;
    flags |= MEBT_GEMM_NO_SPLITK;
  }
  p.drop = DropKey{0u, 0u, 0u, 1.f};
  if (drop != nullptr && drop->thr != 0) {
    MEBT_REQUIRE(!p.gelu && !p.dgelu && !out_fp32, MEBT_ERR_UNSUPPORTED, "gemm: dropout epilogue only on plain bf16 outputs");
    p.drop = *drop;
  }
#ifdef MEBT_GEMM_TRACE
  p.trace = g_gemm_trace;
#endif
  p.num_m_blocks = (M + BM - 1) / BM;
  p.num_k_blocks = (K + BK - 1) / BK;
  // Tile width: the widest that divides N (measured: BN=256 wins or ties at every shape of the path, because the
  // per-CTA cost is TMA latency per k-block, not MMA issue).  Problems with fewer tiles than SMs are split along K.
  // Measured (tools/gemm_probe.py): with >= one wave of 128x256 tiles the widest tile wins (least L2 traffic per
  // flop); below one wave a CTA's k-loop runs at its MMA rate regardless of how many CTAs run, so the best width is the
  // narrowest one that still fits a single wave (most SMs busy, no second wave).
  int bn = 64;
  {
    const int widths[3] = {256, 128, 64};
    bool chosen = false;
    if (N % 256 == 0 && int64_t(p.num_m_blocks) * (N / 256) >= sm_count()) { bn = 256; chosen = true; }
    for (int i = 2; i >= 0 && !chosen; --i)
      if (N % widths[i] == 0 && int64_t(p.num_m_blocks) * (N / widths[i]) <= sm_count()) { bn = widths[i]; chosen = true; }
    if (!chosen) bn = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64);
  }
  // Tried: 256-wide tiles split three ways along K for the K = 4096 latent MLP GEMMs (1536 x 1024 x 4096: 48 tiles x 3
  // splits instead of 96 unsplit 128-wide tiles): 27.6 us against 25.1 us - the partial round trip costs more than
  // the shorter k-loop saves.
  if (flags & MEBT_GEMM_FORCE_BN256) { MEBT_REQUIRE(N % 256 == 0, MEBT_ERR_SHAPE, "BN256 needs N%%256==0"); bn = 256; }
  if (flags & MEBT_GEMM_FORCE_BN128) { MEBT_REQUIRE(N % 128 == 0, MEBT_ERR_SHAPE, "BN128 needs N%%128==0"); bn = 128; }
  if (flags & MEBT_GEMM_FORCE_BN64) bn = 64;
  p.num_n_blocks = N / bn;
  p.splits = 1;
  p.kb_per_split = p.num_k_blocks;
  p.partials = nullptr;
  p.counters = nullptr;
  // at least two waves of tiles: pair up vertically adjacent tiles and share the B tile by TMA multicast
  // ... and 256-wide tiles already from 96 tiles up (measured, tools/gemm_probe.py: 1536 x 3072/4096 x 1024, 3072 x 2048 x 1024,
  // 3072 x 1024 x 4096 gain 3-10 % as pairs - each CTA stages half of B -; below that the 48-tile problems lose)
  const int64_t n_tiles = int64_t(p.num_m_blocks) * p.num_n_blocks;
  p.pair = (bn >= 128 && !(flags & MEBT_GEMM_NO_PAIR) && p.num_m_blocks >= 2 &&
            (n_tiles >= 2 * int64_t(sm_count()) || (bn == 256 && p.num_m_blocks % 2 == 0 && n_tiles >= 96))) ||
           (flags & MEBT_GEMM_FORCE_PAIR && bn >= 128);
  if (!p.pair && !(flags & MEBT_GEMM_NO_SPLITK)) {
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int want = sm_count() / tiles;                       // CTAs available per tile
    // measured: the partial round trip + completion handshake costs ~5 us, so splitting pays only for long
    // reductions (K >= 2048) and with >= 8 k-blocks left per split
    if (want > p.num_k_blocks / 8) want = p.num_k_blocks / 8;
    if (want > 8) want = 8;
    if (want >= 2 && p.num_k_blocks >= 32) {
      const int kpb = (p.num_k_blocks + want - 1) / want;
      const int splits = (p.num_k_blocks + kpb - 1) / kpb;
      const size_t need = size_t(tiles) * splits * BM * bn * sizeof(float);
      SplitKWorkspace* ws = splitk_workspace(need, tiles, stream);
      if (ws != nullptr && splits >= 2) {
        p.splits = splits;
        p.kb_per_split = kpb;
        p.partials = ws->partials;
        p.counters = ws->counters;
      }
    }
  }
  // two issuing threads for the 128-wide tiles (DUAL): plain TMA epilogue only, and enough k-blocks for both halves
  // Opt-in (flag MEBT_GEMM_DUAL or MEBT_GEMM_DUAL=1): with the even ring it needs (4 stages instead of 5) the training step
  // gains 0.4 % (11.31 -> 11.26 ms) - not worth a second code path by default; the convolution kernel, whose 64-wide
  // k-blocks are 184 clk of tensor work, keeps it on (csrc/conv3d.cu).
  static const int dual_env = [] { const char* e = getenv("MEBT_GEMM_DUAL"); return e != nullptr ? atoi(e) : 0; }();
  p.dual = (dual_env != 0 || (flags & MEBT_GEMM_DUAL)) && bn == 128 && !p.pair && p.splits == 1 && p.argmin_out == nullptr && !p.sample_mode &&
           p.num_k_blocks >= 4;
  switch (bn) {
    case 256: return dispatch_major<256>(a_mn, b_mn, A, B, p, lda, ldb, stream);
    case 128: return dispatch_major<128>(a_mn, b_mn, A, B, p, lda, ldb, stream);
    default: return dispatch_major<64>(a_mn, b_mn, A, B, p, lda, ldb, stream);
  }
}

// K6 fused into the head GEMM: packed[row] = (~ordered(best value) << 32 | argmax column) of the Gumbel-max draw over
// all N columns of A B^T / T (bf16 operands); `packed` must be initialised to all ones.  No logits are written.
int gemm_bf16_sample(const void* A, int lda, const void* B, int ldb, int M, int N, int K, float temperature,
                     unsigned long long seed, unsigned long long offset, unsigned long long* packed, cudaStream_t stream) {
  MEBT_REQUIRE(packed != nullptr && temperature > 0.f, MEBT_ERR_SHAPE, "gemm_bf16_sample: bad arguments");
  g_sample_req.active = true;
  g_sample_req.scale = float(1.4426950408889634 / (double(temperature) + 1e-8));
  unsigned long long sd = seed + (offset + 1) * 0x9E3779B97F4A7C15ull;
  sd ^= sd >> 30; sd *= 0xBF58476D1CE4E5B9ull; sd ^= sd >> 27; sd *= 0x94D049BB133111EBull; sd ^= sd >> 31;   // splitmix64
  g_sample_req.k0 = uint32_t(sd);
  g_sample_req.k1 = uint32_t(sd >> 32);
  const int rc = gemm_bf16_ex(A, lda, 0, B, ldb, 0, packed, 8, M, N, K, nullptr, nullptr, 0, nullptr, 0,
                              MEBT_GEMM_INTERNAL_SAMPLE, nullptr, nullptr, 0, 0, stream);
  g_sample_req.active = false;
  return rc;
}

// K9 on the tensor cores: packed[row] = min over codes n of (order(d) << 32 | n), d = (row_sq[row] - 2 A[row].B[n]) + col_sq[n],
// with fp16 operands (A [M, K] and B [N, K] K-major; see csrc/vq.cu for the split that makes the products fp32-accurate).
int gemm_f16_argmin(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* col_sq,
                    const float* row_sq, unsigned long long* packed, cudaStream_t stream) {
  MEBT_REQUIRE(col_sq != nullptr && row_sq != nullptr && packed != nullptr, MEBT_ERR_SHAPE, "gemm_f16_argmin: null operand");
  return gemm_bf16_ex(A, lda, 0, B, ldb, 0, packed, 8, M, N, K, col_sq, row_sq, 8, nullptr, 0, MEBT_GEMM_INTERNAL_ARGMIN, nullptr,
                      nullptr, 0, 0, stream);
}

}  // namespace mebt

#ifdef MEBT_GEMM_TRACE
extern "C" void mebt_gemm_set_trace(long long* buf) { mebt::g_gemm_trace = buf; }
#endif

extern "C" int mebt_gemm_bf16_aux(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major,
                                  void* C, int ldc, int M, int N, int K, const float* bias, const void* residual,
                                  int ldres, void* aux, int ldaux, int flags, void* stream) {
  return mebt::gemm_bf16_aux(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, bias, residual, ldres, aux, ldaux,
                             flags, static_cast<cudaStream_t>(stream));
}

extern "C" int mebt_gemm_bf16(const void* A, int lda, int a_mn_major, const void* B, int ldb, int b_mn_major, void* C,
                              int ldc, int M, int N, int K, const float* bias, const void* residual, int ldres,
                              int flags, void* stream) {
  return mebt::gemm_bf16(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, bias, residual, ldres, flags,
                         static_cast<cudaStream_t>(stream));
}
