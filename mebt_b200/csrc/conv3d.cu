// 3-D convolution of the VQGAN encoder / decoder as an implicit GEMM on the tensor cores:
// replaces nn.Conv3d / nn.ConvTranspose3d behind SamePadConv3d / SamePadConvTranspose3d (mebt/vqgan.py:358-405), the
// GroupNorm + SiLU that precedes every ResBlock convolution (:255-260, :336-356) and the replicate padding (F.pad, :381,404).
//
// Layout: activations channels-last bf16, X[b][t][h][w][c].  A convolution with taps (KT, KH, KW) and strides
// (ST, SH, SW) over the ALREADY PADDED input is
//     Y[b,t,h,w,co] = bias[co] + sum_{dt,dh,dw,c} Xp[b, t ST + dt + OT, h SH + dh + OH, w SW + dw + OW, c] * W[co][(dt,dh,dw)][c]
// = a GEMM with M = output positions, N = Cout, K = taps x Cin whose A operand needs no im2col buffer: one output tile is
// a patch of PT x PH x PW = 128 positions, and the A tile of (tap, 64-channel block) is ONE 5-D TMA box
// {64 c, PW, PH, PT, 1} of Xp at the tap's offset (element strides = the convolution strides), landing in shared memory
// as the 128 x 64 K-major, 128B-swizzled tile tcgen05.mma reads.  B = packed weights [Cout][taps * Cp] (Cp = channels
// rounded up to 64, zero padded; TMA zero-fills the activations' missing channels).  The transposed convolutions of the
// decoder run as one such stride-1 convolution per output parity class (2 taps per up-sampled dimension, origin offset
// = parity), written into the interleaved positions of the up-sampled tensor by the 5-D TMA store (element strides 2).
//
// Kernel = csrc/gemm_grouped.cu's structure: 192 threads (TMA producer warp, one MMA-issuing thread, 4 epilogue warps),
// 3-6 stage operand ring, two TMEM accumulators, epilogue (bias, residual, bf16) through a 4-slot staging ring and TMA store.
// pad_norm_act_kernel writes Xp: replicate padding fused with GroupNorm(32) / eval BatchNorm + SiLU of the source.
//
// Variants of the main loop (template parameters of conv3d_igemm_kernel, chosen in mebt_conv3d_ndhwc):
//   per-tap  (ROW = 0): a stage = one (tap, channel block); 64 / 128 / 256-wide tiles by output channels.  DUAL (<= 128-wide):
//            two MMA-issuing threads, an accumulator half each, stages owned by ring parity (even ring).
//   ROW = 1 : stride 1 along w, output rows of 128 positions, <= 64 output channels: a tile is one output row and ONE box of
//            128 + KW - 1 positions serves all KW taps along w through descriptor starts shifted by dw * 128 bytes.
//   ROW = 2 : the same with the CTA's rows taken two at a time: two activation boxes, one set of weight tiles per stage,
//            four TMEM accumulators, two staging slots.
//   window  (host side): cin = 64 > ldx - the K slice of a position runs on into the following positions of its row, i.e. the
//            taps along w of a few-channel input (the RGB video) packed into one k-block.
// Measured rules that shaped them (Findings): a 64-wide tile is bound by L2 -> shared-memory operand fills, not by
// the tensor pipe; a TMA box partly outside its tensor (or with 16-byte rows) is filled at a fraction of the normal rate, so
// operands come in whole boxes; two consumers of one barrier ring must each own their stages.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mebt {

int get_tensor_map_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);
int get_tensor_map_5d(CUtensorMap* out, const void* ptr, const uint64_t dims[5], const uint64_t strides_bytes[4],
                      const uint32_t box[5], const uint32_t elem_strides[5]);

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int CV_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int EPI_SLOTS = 4;
constexpr int EPI_SLOT_BYTES = 128 * 128;

// ROW mode (stride 1 along w, output rows of 128 positions): one tile = one output row, and the A operands of ALL the
// taps along w come from ONE box {64 c, 128 + KW - 1 positions}: tap dw is the same shared-memory rows shifted by dw
// (descriptor start address + dw * 128 bytes; the 128-byte swizzle is a function of the absolute shared-memory address,
// so a start that is not 1024-byte aligned reads what TMA wrote - measured: correct with the descriptor's matrix-base-offset
// field left 0, wrong with it set to (address >> 7) & 7).  A stage then holds KW k-blocks: a third of the L2 -> smem operand traffic of the 64-wide
// tiles (which is what bounds them) and one barrier hand-off per KW k-blocks instead of one per k-block.
constexpr int ROW_A_BYTES = 17 * 1024;   // up to 136 rows of 128 bytes
constexpr int ROW_MAX_KW = 4;

// ROW = 2 (row pairs): the CTA works on TWO of its output rows at a time - two activation boxes per stage, ONE set of weight
// tiles for both (the weights are the same for every tile and are 60 % of a ROW-mode stage's bytes), four TMEM accumulators.
template <int BN, int ROW = 0, bool DUAL = false>
struct ConvSmem {
  static constexpr int STAGES = ROW ? 3 : (BN == 256 ? 3 : (BN == 128 ? (DUAL ? 4 : 5) : 6));
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int ROW_KW = ROW == 2 ? 3 : ROW_MAX_KW;            // weight tiles a stage holds
  static constexpr int A_BYTES = ROW ? ROW * ROW_A_BYTES : A_TILE_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + (ROW ? ROW_KW : 1) * B_TILE_BYTES;
  static constexpr int EPI_SLOTS_N = ROW == 2 ? 2 : EPI_SLOTS;
  static constexpr int NACC = ROW == 2 ? 4 : 2;                        // TMEM accumulator buffers
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_SLOTS_N * EPI_SLOT_BYTES;
  static constexpr int BIAS_OFFSET = BAR_OFFSET + 256;
  static constexpr int TOTAL = BIAS_OFFSET + BN * 4 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
  static_assert(STAGE_BYTES % 1024 == 0, "stages keep the 1024-byte alignment of the swizzle pattern");
};

struct ConvParams {
  int PT, PH, PW;             // output patch of one tile (PT * PH * PW = 128)
  int nt_t, nt_h, nt_w;       // tiles per output dimension
  int tiles_m, num_n_blocks, total_tiles;
  int KT, KH, KW;             // taps
  int ST, SH, SW;             // input step per output position
  int OT, OH, OW;             // origin offset in the padded input
  int YT, YH, YW;             // output step per output position in the destination tensor (transposed convolution: 2) ...
  int Y0T, Y0H, Y0W;          // ... and its origin (the parity)
  int cblocks;                // Cp / 64
  int Cout;
  const float* bias;          // [Cout] or nullptr
  const __nv_bfloat16* resid; // dense [B, To, Ho, Wo, ldr] or nullptr (added after the bias)
  long long res_b, res_t, res_h, res_w;   // residual strides in elements
  __nv_bfloat16* y_direct;    // <= 8 output channels (the RGB reconstruction): plain 16-byte stores instead of the TMA store
  long long yd_b, yd_t, yd_h, yd_w;       // destination strides in elements
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// tile number -> (n block, batch element, patch origin in output positions)
__device__ __forceinline__ void locate(const ConvParams& p, int work, int& ni, int& b, int& t0, int& h0, int& w0) {
  ni = work / p.tiles_m;
  int m = work - ni * p.tiles_m;
  const int wi = m % p.nt_w; m /= p.nt_w;
  const int hi = m % p.nt_h; m /= p.nt_h;
  const int ti = m % p.nt_t; m /= p.nt_t;
  b = m;
  t0 = ti * p.PT; h0 = hi * p.PH; w0 = wi * p.PW;
}

// DUAL (per-tap form, <= 128-wide tiles): two MMA-issuing threads (warps 1 and 6), each accumulating every other k-block in
// its own TMEM half; the epilogue adds the halves (csrc/gemm.cu, DUAL: one issuer's barrier wait + commit cost ~520 clk per
// k-block against 184 / 256 clk of tensor work for a 64- / 128-wide k-block).
template <int BN, int ROW, bool DUAL>
__global__ void __launch_bounds__(DUAL ? CV_THREADS + 32 : CV_THREADS, 1)
conv3d_igemm_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                    const __grid_constant__ CUtensorMap tma_y, const ConvParams p) {
  using L = ConvSmem<BN, ROW, DUAL>;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  constexpr int NACC = L::NACC;
  uint64_t* tmem_empty_bar = tmem_full_bar + NACC;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + NACC);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  static_assert(!DUAL || (!ROW && BN <= 128 && STAGES % 2 == 0), "DUAL: per-tap form, two accumulator halves, even ring");
  constexpr uint32_t ACC_STRIDE = DUAL ? 2 * BN : BN;
  constexpr uint32_t TMEM_COLS = NACC * ACC_STRIDE;
  static_assert(ROW != 2 || (BN == 64 && !DUAL), "row pairs: 64-wide tiles");

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tma_x);
    prefetch_tensormap(&tma_w);
    prefetch_tensormap(&tma_y);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < NACC; ++s) {
      mbar_init(&tmem_full_bar[s], DUAL ? 2 : 1);
      mbar_init(&tmem_empty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr_smem, TMEM_COLS); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_wait();

  const int work0 = int(blockIdx.x), work_stride = int(gridDim.x);
  const int taps = p.KT * p.KH * p.KW;
  const int nkb = taps * p.cblocks;
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if constexpr (ROW == 2) {
        // row pairs: the CTA's tiles two at a time (work, work + stride): two activation boxes, one set of weight tiles
        for (int work = work0; work < p.total_tiles; work += 2 * work_stride) {
          const bool has_b = work + work_stride < p.total_tiles;
          int ni, b, t0, h0, w0, ni2 = 0, b2 = 0, t2 = 0, h2 = 0, w2 = 0;
          locate(p, work, ni, b, t0, h0, w0);
          if (has_b) locate(p, work + work_stride, ni2, b2, t2, h2, w2);
          const uint32_t stage_tx = uint32_t((has_b ? 2 : 1) * (128 + p.KW - 1) * 128 + p.KW * L::B_TILE_BYTES);
          for (int dt = 0; dt < p.KT; ++dt)
            for (int dh = 0; dh < p.KH; ++dh)
              for (int cb = 0; cb < p.cblocks; ++cb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * L::STAGE_BYTES;
                uint8_t* sB = sA + 2 * ROW_A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                tma_load_5d(sA, &tma_x, &full_bar[stage], cb * BK, w0 * p.SW + p.OW, h0 * p.SH + p.OH + dh, t0 * p.ST + p.OT + dt, b);
                if (has_b)
                  tma_load_5d(sA + ROW_A_BYTES, &tma_x, &full_bar[stage], cb * BK, w2 * p.SW + p.OW, h2 * p.SH + p.OH + dh,
                              t2 * p.ST + p.OT + dt, b2);
                for (int dw = 0; dw < p.KW; ++dw)
                  tma_load_2d(sB + dw * L::B_TILE_BYTES, &tma_w, &full_bar[stage],
                              (((dt * p.KH + dh) * p.KW + dw) * p.cblocks + cb) * BK, ni * BN);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
        }
      } else
      for (int work = work0; work < p.total_tiles; work += work_stride) {
        int ni, b, t0, h0, w0;
        locate(p, work, ni, b, t0, h0, w0);
        const int xt = t0 * p.ST + p.OT, xh = h0 * p.SH + p.OH, xw = w0 * p.SW + p.OW;
        if constexpr (ROW) {
          const uint32_t stage_tx = uint32_t((128 + p.KW - 1) * 128 + p.KW * L::B_TILE_BYTES);
          for (int dt = 0; dt < p.KT; ++dt)
            for (int dh = 0; dh < p.KH; ++dh)
              for (int cb = 0; cb < p.cblocks; ++cb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sA = smem + stage * L::STAGE_BYTES;
                uint8_t* sB = sA + ROW_A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                tma_load_5d(sA, &tma_x, &full_bar[stage], cb * BK, xw, xh + dh, xt + dt, b);       // box {64 c, 128 + KW - 1, 1, 1, 1}
                for (int dw = 0; dw < p.KW; ++dw)
                  tma_load_2d(sB + dw * L::B_TILE_BYTES, &tma_w, &full_bar[stage],
                              (((dt * p.KH + dh) * p.KW + dw) * p.cblocks + cb) * BK, ni * BN);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
        } else {
          int kb = 0;
          for (int dt = 0; dt < p.KT; ++dt)
            for (int dh = 0; dh < p.KH; ++dh)
              for (int dw = 0; dw < p.KW; ++dw)
                for (int cb = 0; cb < p.cblocks; ++cb, ++kb) {
                  mbar_wait(&empty_bar[stage], phase ^ 1);
                  uint8_t* sA = smem + stage * L::STAGE_BYTES;
                  uint8_t* sB = sA + A_TILE_BYTES;
                  mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                  tma_load_5d(sA, &tma_x, &full_bar[stage], cb * BK, xw + dw, xh + dh, xt + dt, b);   // box {64 c, PW, PH, PT, 1}
                  tma_load_2d(sB, &tma_w, &full_bar[stage], kb * BK, ni * BN);                        // box [64 k][BN n]
                  if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
        }
      }
    }
  } else if (DUAL && (warp == 1 || warp == 6)) {
    // ================= the two MMA issuers of the DUAL variant =================
    if (lane == 0) {
      const uint32_t j = warp == 6 ? 1u : 0u;
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
      uint32_t cnt = 0;
      int it = 0;
      for (int work = work0; work < p.total_tiles; work += work_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc) * ACC_STRIDE + j * uint32_t(BN);
        // Issuer j takes the k-blocks whose RING position g is of parity j: with an even number of stages a stage then always
        // belongs to the same issuer, who waits for its phases strictly in order.  (Splitting by the index inside the tile let
        // one issuer reach a stage a whole ring ahead of the other: a parity wait for phase k + 1 on a barrier still in an
        // incomplete phase k succeeds at once - stale operands, seen as rare wrong results / launch failures under stress.)
        for (uint32_t i = ((cnt & 1u) == j) ? 0u : 1u; i < uint32_t(nkb); i += 2) {
          const uint32_t g = cnt + i;
          const uint32_t stage = g % uint32_t(STAGES), phase = (g / uint32_t(STAGES)) & 1u;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(smem_u32(smem + stage * L::STAGE_BYTES), 16);
          umma_bf16_ss_x4<false>(tmem_d, a_lo, a_lo + uint32_t(A_TILE_BYTES >> 4), (UMMA_K * 2) >> 4, (UMMA_K * 2) >> 4, desc_hi,
                                 desc_hi, idesc, i >= 2 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
        }
        umma_commit(&tmem_full_bar[acc]);
        cnt += uint32_t(nkb);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t desc_hi = smem_desc_hi_sw128(1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if constexpr (ROW == 2) {
        const int nst = p.KT * p.KH * p.cblocks;
        for (int work = work0; work < p.total_tiles; work += 2 * work_stride, it += 2) {
          const bool has_b = work + work_stride < p.total_tiles;
          const int acc = it & 3;                             // tiles it and it + 1: accumulators acc and acc + 1
          const uint32_t acc_phase = (it >> 2) & 1;
          mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
          if (has_b) mbar_wait(&tmem_empty_bar[acc + 1], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t tmem_a = tmem_base + uint32_t(acc) * ACC_STRIDE, tmem_b = tmem_a + ACC_STRIDE;
          for (int st = 0; st < nst; ++st) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t base = smem_u32(smem + stage * L::STAGE_BYTES);
            for (int dw = 0; dw < p.KW; ++dw) {
              const uint32_t a_lo = smem_desc_lo(base + uint32_t(dw) * 128u, 16);
              const uint32_t b_lo = smem_desc_lo(base + uint32_t(2 * ROW_A_BYTES + dw * L::B_TILE_BYTES), 16);
              umma_bf16_ss_x4<false>(tmem_a, a_lo, b_lo, (UMMA_K * 2) >> 4, (UMMA_K * 2) >> 4, desc_hi, desc_hi, idesc,
                                     (st > 0 || dw > 0) ? 1u : 0u);
            }
            if (has_b)
              for (int dw = 0; dw < p.KW; ++dw) {
                const uint32_t a_lo = smem_desc_lo(base + uint32_t(ROW_A_BYTES) + uint32_t(dw) * 128u, 16);
                const uint32_t b_lo = smem_desc_lo(base + uint32_t(2 * ROW_A_BYTES + dw * L::B_TILE_BYTES), 16);
                umma_bf16_ss_x4<false>(tmem_b, a_lo, b_lo, (UMMA_K * 2) >> 4, (UMMA_K * 2) >> 4, desc_hi, desc_hi, idesc,
                                       (st > 0 || dw > 0) ? 1u : 0u);
              }
            umma_commit(&empty_bar[stage]);
            if (st == nst - 1) {
              umma_commit(&tmem_full_bar[acc]);
              if (has_b) umma_commit(&tmem_full_bar[acc + 1]);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      } else
      for (int work = work0; work < p.total_tiles; work += work_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);     // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc) * ACC_STRIDE;
        if constexpr (ROW) {
          const int nst = p.KT * p.KH * p.cblocks;
          for (int st = 0; st < nst; ++st) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t base = smem_u32(smem + stage * L::STAGE_BYTES);
            for (int dw = 0; dw < p.KW; ++dw) {
              // tap dw: the same rows, dw positions (128 bytes each) further; B: the tap's own [BN][64] tile
              const uint32_t a_lo = smem_desc_lo(base + uint32_t(dw) * 128u, 16);
              const uint32_t b_lo = smem_desc_lo(base + uint32_t(ROW_A_BYTES + dw * L::B_TILE_BYTES), 16);
              umma_bf16_ss_x4<false>(tmem_d, a_lo, b_lo, (UMMA_K * 2) >> 4, (UMMA_K * 2) >> 4, desc_hi, desc_hi, idesc,
                                     (st > 0 || dw > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (st == nst - 1) umma_commit(&tmem_full_bar[acc]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // K-major operands: a K = 16 step advances 32 bytes inside the 128-byte swizzle row (csrc/gemm.cu)
          const uint32_t a_lo = smem_desc_lo(smem_u32(smem + stage * L::STAGE_BYTES), 16);
          umma_bf16_ss_x4<false>(tmem_d, a_lo, a_lo + uint32_t(A_TILE_BYTES >> 4), (UMMA_K * 2) >> 4, (UMMA_K * 2) >> 4, desc_hi,
                                 desc_hi, idesc, kb > 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == nkb - 1) umma_commit(&tmem_full_bar[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (warps 2..5): TMEM -> (+ bias, + residual) -> bf16 -> staging slot -> 5-D TMA store
    const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
    uint8_t* slots = smem + L::EPI_OFFSET;
    float* s_bias = reinterpret_cast<float*>(smem + L::BIAS_OFFSET);
    const bool epi_t0 = threadIdx.x == 64;
    const int sw = lane & 7;
    const int r = q * 32 + lane;                              // row of the tile = position of the patch, w fastest
    const int pw = r % p.PW, ph = (r / p.PW) % p.PH, pt = r / (p.PW * p.PH);
    int it = 0;
    int gu = 0;                                               // stored 64-channel chunks so far: consecutive staging slots
    for (int work = work0; work < p.total_tiles; work += work_stride, ++it) {
      int ni, b, t0, h0, w0;
      locate(p, work, ni, b, t0, h0, w0);
      const int n0 = ni * BN;
      const int acc = it & (NACC - 1);
      const uint32_t acc_phase = (it / NACC) & 1;
      asm volatile("bar.sync 1, 128;" ::: "memory");          // the previous tile's readers of s_bias are done
      for (int c = threadIdx.x - 64; c < BN; c += 128) s_bias[c] = (p.bias != nullptr && n0 + c < p.Cout) ? __ldg(p.bias + n0 + c) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const __nv_bfloat16* res_row =
          p.resid != nullptr ? p.resid + b * p.res_b + (long long)(t0 + pt) * p.res_t + (long long)(h0 + ph) * p.res_h +
                                   (long long)(w0 + pw) * p.res_w : nullptr;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc) * ACC_STRIDE;
      // only the 64-channel chunks that hold real output channels are stored
      const int chunks = min(BN / 64, (p.Cout - n0 + 63) / 64);
#pragma unroll 1
      for (int ch = 0; ch < BN / 64; ++ch) {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(t_acc + uint32_t(ch * 64), ra);
        tmem_ld_32x32(t_acc + uint32_t(ch * 64 + 32), rb);
        tmem_ld_wait();
        if constexpr (DUAL) {                                  // + the odd k-blocks' half (fixed order: reproducible)
          uint32_t ra2[32], rb2[32];
          tmem_ld_32x32(t_acc + uint32_t(BN + ch * 64), ra2);
          tmem_ld_32x32(t_acc + uint32_t(BN + ch * 64 + 32), rb2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            ra[j] = __float_as_uint(__uint_as_float(ra[j]) + __uint_as_float(ra2[j]));
            rb[j] = __float_as_uint(__uint_as_float(rb[j]) + __uint_as_float(rb2[j]));
          }
        }
        if (ch == BN / 64 - 1) {                               // accumulator fully read: hand it back before the last stores
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[acc]);
        }
        if (ch >= chunks) continue;                            // block-uniform
        if (p.y_direct != nullptr) {
          // 8 output channels = 16 bytes per position: a TMA store would move 128 separate 16-byte rows per tile (measured:
          // the layer took 282 us against 172 us for its 64-channel twin); consecutive lanes hold consecutive positions of
          // the row, so plain stores coalesce into 512-byte segments
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ra[0]) + s_bias[0], __uint_as_float(ra[1]) + s_bias[1]);
          o.y = pack_bf16x2(__uint_as_float(ra[2]) + s_bias[2], __uint_as_float(ra[3]) + s_bias[3]);
          o.z = pack_bf16x2(__uint_as_float(ra[4]) + s_bias[4], __uint_as_float(ra[5]) + s_bias[5]);
          o.w = pack_bf16x2(__uint_as_float(ra[6]) + s_bias[6], __uint_as_float(ra[7]) + s_bias[7]);
          *reinterpret_cast<uint4*>(p.y_direct + b * p.yd_b + (long long)(t0 + pt) * p.yd_t + (long long)(h0 + ph) * p.yd_h +
                                    (long long)(w0 + pw) * p.yd_w) = o;
          continue;
        }
        const int s_c = gu & (L::EPI_SLOTS_N - 1);
        ++gu;
        uint8_t* row_c = slots + s_c * EPI_SLOT_BYTES + r * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t* rr = half == 0 ? ra : rb;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]) + s_bias[ch * 64 + half * 32 + j];
          if (res_row != nullptr) {
            const int c0 = n0 + ch * 64 + half * 32;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (c0 + 8 * j < p.Cout) {                       // channel counts are multiples of 8
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(res_row + c0 + 8 * j));
                const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
                v[8 * j + 0] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += bb.x; v[8 * j + 3] += bb.y;
                v[8 * j + 4] += c2.x; v[8 * j + 5] += c2.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
            o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(row_c + (((half * 4 + j) ^ sw) << 4)) = o;
          }
        }
        // one barrier per chunk: behind it every row of the slot is written (and fenced towards the async proxy), and
        // the slot the NEXT chunk writes has been read out by its previous store (epi_t0 checks before arriving)
        fence_proxy_async_smem();
        if (epi_t0) { if constexpr (L::EPI_SLOTS_N == 4) tma_store_wait_read<2>(); else tma_store_wait_read<0>(); }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_t0) {
          tma_store_5d(&tma_y, slots + s_c * EPI_SLOT_BYTES, n0 + ch * 64, w0 * p.YW + p.Y0W, h0 * p.YH + p.Y0H,
                       t0 * p.YT + p.Y0T, b);
          tma_store_commit();
        }
      }
    }
    if (epi_t0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- GroupNorm statistics: partial (sum, sum of squares) per (batch element, group, slab of positions) ---------------
// x channels-last [B, P, ldc] (P = T*H*W positions); grid (slabs, B); partial[b][g][slab][2].  HBM-bound: one thread reads
// 16 bytes (8 channels) of a position, a warp whole contiguous rows; the 256 / (ldc / 8) positions a block reads per
// iteration are summed per thread and channel pair, then per group in a fixed order (reproducible sums).
constexpr int GN_MAX_SLABS = 64;
constexpr int GN_MAX_C = 2048;
__global__ void __launch_bounds__(256) groupnorm_partial_kernel(const __nv_bfloat16* __restrict__ x, long long P, int ldc, int cg,
                                                                int G, int slabs, float* __restrict__ partial) {
  __shared__ float sm1[1024], sm2[1024];
  const int slab = blockIdx.x, b = blockIdx.y;
  const int cv = ldc >> 3, R = 256 / cv, Q = cv * 4;         // vectors per position, positions per iteration, channel pairs
  const int v = threadIdx.x % cv, r = threadIdx.x / cv;
  const long long p0 = P * slab / slabs, p1 = P * (slab + 1) / slabs;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (r < R) {
    const __nv_bfloat16* base = x + (size_t(b) * P) * ldc + v * 8;
    auto acc = [&](const uint4& u) {
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        s1[k] += f.x + f.y;
        s2[k] = fmaf(f.x, f.x, fmaf(f.y, f.y, s2[k]));
      }
    };
    long long pos = p0 + r;
    for (; pos + 3LL * R < p1; pos += 4LL * R) {              // four independent 16-byte loads in flight per thread
      const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(base + pos * ldc));
      const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(base + (pos + R) * ldc));
      const uint4 u2 = __ldg(reinterpret_cast<const uint4*>(base + (pos + 2LL * R) * ldc));
      const uint4 u3 = __ldg(reinterpret_cast<const uint4*>(base + (pos + 3LL * R) * ldc));
      acc(u0); acc(u1); acc(u2); acc(u3);
    }
    for (; pos < p1; pos += R) acc(__ldg(reinterpret_cast<const uint4*>(base + pos * ldc)));
#pragma unroll
    for (int k = 0; k < 4; ++k) { sm1[r * Q + v * 4 + k] = s1[k]; sm2[r * Q + v * 4 + k] = s2[k]; }
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += 256) {
    float a = 0.f, c = 0.f;
    const int q0 = g * (cg >> 1), q1 = q0 + (cg >> 1);
    for (int rr = 0; rr < R; ++rr)
      for (int q = q0; q < q1; ++q) { a += sm1[rr * Q + q]; c += sm2[rr * Q + q]; }
    float* o = partial + ((size_t(b) * G + g) * slabs + slab) * 2;
    o[0] = a; o[1] = c;
  }
}

// per (batch element, channel): y = x * a + s with a = rstd * gamma, s = beta - mean * a; coef[b][c] = (a, s)
__global__ void __launch_bounds__(256) groupnorm_coef_kernel(const float* __restrict__ partial, int C, int cg, int G, int slabs,
                                                             float inv_n, float eps, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float2* __restrict__ coef) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g = c / cg;
    const float2* pp = reinterpret_cast<const float2*>(partial) + (size_t(b) * G + g) * slabs;
    float s1 = 0.f, s2 = 0.f;
    for (int s = 0; s < slabs; ++s) { const float2 t = pp[s]; s1 += t.x; s2 += t.y; }       // slab order: reproducible
    const float mean = s1 * inv_n;
    const float rstd = rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + eps);
    const float a = rstd * gamma[c];
    coef[size_t(b) * C + c] = make_float2(a, fmaf(-mean, a, beta[c]));
  }
}

// ---- y = pad_replicate(act(norm(x))) : the operand of the next convolution -----------------------------------------------
// x [B, T, H, W, ldx]; y [B, T + pt0 + pt1, H + ph0 + ph1, W + pw0 + pw1, ldy] with C real channels (ldy >= C; the
// channels [C, ldy) are zeroed).  norm: 0 none, 1 GroupNorm (coef[b][c] from groupnorm_coef_kernel), 2 per-channel affine
// (eval BatchNorm folded into scale / shift).  act: 0 none, 1 SiLU.  One block = one padded (b, t, h) row of Wp positions,
// one thread = 8 channels of one position: 16-byte loads / stores, no per-element index arithmetic beyond one division.
struct PadParams {
  int B, T, H, W, C, ldx, ldy;
  int pt0, ph0, pw0, Tp, Hp, Wp;
  int norm, act;
  const float2* coef;        // norm 1: [B][C] (a, s)
  const float* gamma;        // norm 2: scale
  const float* beta;         // norm 2: shift
};
__global__ void __launch_bounds__(256) pad_norm_act_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                           const PadParams p) {
  extern __shared__ float2 cf[];                               // [C] (a, s) of this batch element
  const int row = blockIdx.x;
  const int h = row % p.Hp, t = (row / p.Hp) % p.Tp, b = row / (p.Hp * p.Tp);
  if (p.norm == 1) {
    for (int c = threadIdx.x; c < p.C; c += 256) cf[c] = p.coef[size_t(b) * p.C + c];
    __syncthreads();
  } else if (p.norm == 2) {
    for (int c = threadIdx.x; c < p.C; c += 256) cf[c] = make_float2(p.gamma[c], p.beta[c]);
    __syncthreads();
  }
  const int ts = min(max(t - p.pt0, 0), p.T - 1), hs = min(max(h - p.ph0, 0), p.H - 1);
  const __nv_bfloat16* xs = x + ((size_t(b) * p.T + ts) * p.H + hs) * p.W * size_t(p.ldx);
  __nv_bfloat16* yd = y + size_t(row) * p.Wp * p.ldy;
  const int cv = p.ldy >> 3, n = p.Wp * cv;
  const bool ragged = (p.C & 7) != 0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int w = i / cv, c0 = (i - w * cv) * 8;
    uint4 out = make_uint4(0, 0, 0, 0);
    if (c0 < p.C) {
      const int ws = min(max(w - p.pw0, 0), p.W - 1);
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xs + size_t(ws) * p.ldx + c0));
      if (p.norm == 0 && p.act == 0 && !ragged) {
        out = u;
      } else {
        float v[8];
        { const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
          v[0] = a.x; v[1] = a.y; v[2] = bb.x; v[3] = bb.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y; }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float xv = v[k];
          if (ragged && c0 + k >= p.C) { v[k] = 0.f; continue; }
          if (p.norm != 0) { const float2 as = cf[c0 + k]; xv = fmaf(xv, as.x, as.y); }
          if (p.act == 1) xv = __fdividef(xv, 1.f + __expf(-xv));
          v[k] = xv;
        }
        out = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
    }
    *reinterpret_cast<uint4*>(yd + size_t(i) * 8) = out;
  }
}

template <int BN, int ROW, bool DUAL = false>
int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const ConvParams& p, double flops,
                cudaStream_t st) {
  using L = ConvSmem<BN, ROW, DUAL>;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(conv3d_igemm_kernel<BN, ROW, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  }
  LaunchScope ls(FAM_GEMM, flops, st);
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  MEBT_CUDA_OK(launch_pdl(conv3d_igemm_kernel<BN, ROW, DUAL>, dim3(grid), dim3(DUAL ? CV_THREADS + 32 : CV_THREADS), L::TOTAL, st, tx, tw, ty, p));
  MEBT_LAUNCH_OK("conv3d_igemm_kernel");
  return MEBT_OK;
}

}  // namespace
}  // namespace mebt

extern "C" {

size_t mebt_groupnorm_workspace_bytes(int B, int groups) {
  // partial sums [B][groups][GN_MAX_SLABS][2] + per-channel coefficients [B][GN_MAX_C] float2
  return size_t(B) * groups * mebt::GN_MAX_SLABS * 2 * sizeof(float) + size_t(B) * mebt::GN_MAX_C * sizeof(float2);
}

int mebt_pad_norm_act(const void* x, int ldx, void* y, int ldy, int B, int T, int H, int W, int C, const int* pad6, int norm,
                      int act, int groups, float eps, const float* gamma, const float* beta, void* workspace,
                      size_t workspace_bytes, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && C > 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, MEBT_ERR_SHAPE,
               "pad_norm_act: bad shape (channel strides must be multiples of 8)");
  MEBT_REQUIRE(norm >= 0 && norm <= 2 && act >= 0 && act <= 1, MEBT_ERR_UNSUPPORTED, "pad_norm_act: norm %d act %d", norm, act);
  for (int i = 0; i < 6; ++i) MEBT_REQUIRE(pad6[i] >= 0 && pad6[i] <= 8, MEBT_ERR_SHAPE, "pad_norm_act: pad[%d] = %d", i, pad6[i]);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PadParams p;
  p.B = B; p.T = T; p.H = H; p.W = W; p.C = C; p.ldx = ldx; p.ldy = ldy;
  p.pt0 = pad6[0]; p.ph0 = pad6[2]; p.pw0 = pad6[4];
  p.Tp = T + pad6[0] + pad6[1]; p.Hp = H + pad6[2] + pad6[3]; p.Wp = W + pad6[4] + pad6[5];
  p.norm = norm; p.act = act;
  p.coef = nullptr; p.gamma = gamma; p.beta = beta;
  const long long rows = (long long)B * p.Tp * p.Hp;
  MEBT_REQUIRE(rows < (1LL << 31) && (long long)p.Wp * (ldy / 8) < (1LL << 31), MEBT_ERR_SHAPE, "pad_norm_act: tensor too large");
  if (norm == 1) {
    MEBT_REQUIRE(groups > 0 && groups <= 256 && C % groups == 0 && (C / groups) % 2 == 0 && gamma != nullptr && beta != nullptr,
                 MEBT_ERR_SHAPE, "pad_norm_act: GroupNorm needs C %% groups == 0 and an even group width (C=%d groups=%d)", C, groups);
    MEBT_REQUIRE(C <= GN_MAX_C && ldx <= GN_MAX_C, MEBT_ERR_UNSUPPORTED, "pad_norm_act: GroupNorm over at most %d channels", GN_MAX_C);
    MEBT_REQUIRE(workspace != nullptr && workspace_bytes >= mebt_groupnorm_workspace_bytes(B, groups), MEBT_ERR_WORKSPACE,
                 "pad_norm_act: workspace too small");
    const long long P = (long long)T * H * W;
    // slabs: enough blocks to keep every SM's memory pipe full, at least 256 positions each
    int slabs = int(std::min<long long>(GN_MAX_SLABS, std::max<long long>(1, (sm_count() * 4 + B - 1) / B)));
    slabs = int(std::max<long long>(1, std::min<long long>(slabs, P / 256)));
    float* partial = static_cast<float*>(workspace);
    float2* coef = reinterpret_cast<float2*>(partial + size_t(B) * groups * GN_MAX_SLABS * 2);
    LaunchScope ls(FAM_LAYERNORM, double(B) * T * H * W * C * 2.0, st);
    groupnorm_partial_kernel<<<dim3(slabs, B), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), P, ldx, C / groups, groups,
                                                             slabs, partial);
    MEBT_LAUNCH_OK("groupnorm_partial_kernel");
    groupnorm_coef_kernel<<<B, 256, 0, st>>>(partial, C, C / groups, groups, slabs, 1.f / (float(T) * H * W * (C / groups)), eps,
                                             gamma, beta, coef);
    MEBT_LAUNCH_OK("groupnorm_coef_kernel");
    p.coef = coef;
  } else if (norm == 2) {
    MEBT_REQUIRE(gamma != nullptr && beta != nullptr, MEBT_ERR_SHAPE, "pad_norm_act: affine norm needs scale and shift");
    MEBT_REQUIRE(C <= GN_MAX_C, MEBT_ERR_UNSUPPORTED, "pad_norm_act: affine norm over at most %d channels", GN_MAX_C);
  }
  const long long total = rows * p.Wp * (ldy / 8);
  LaunchScope ls(FAM_LAYERNORM, double(total) * 16.0 * 2.0, st);
  pad_norm_act_kernel<<<int(rows), 256, norm != 0 ? size_t(C) * sizeof(float2) : 0, st>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), p);
  MEBT_LAUNCH_OK("pad_norm_act_kernel");
  return MEBT_OK;
}

int mebt_conv3d_ndhwc(const void* xp, int ldx, const int* xdims4, const void* w, int cin, const float* bias,
                      const void* resid, int ldr, void* y, int ldy, const int* ydims4, int cout, const int* taps3,
                      const int* step3, const int* origin3, const int* ystep3, const int* yorigin3, const int* odims3,
                      void* stream) {
  // xp: padded input [B, Tp, Hp, Wp, ldx] (xdims4 = B, Tp, Hp, Wp); y: destination [B, Ty, Hy, Wy, ldy] (ydims4);
  // odims3 = (To, Ho, Wo) output positions computed by this launch; they land at y[b, t * ystep + yorigin, ...].
  using namespace mebt;
  const int B = xdims4[0];
  const int To = odims3[0], Ho = odims3[1], Wo = odims3[2];
  // WINDOW mode (cin = 64 > ldx): the 64-element K slice of a position runs on into the following positions of its row
  // (ldx channels each) - the taps along w of a few-channel input packed into ONE k-block: taps3[2] must be 1 and the
  // weights hold [cout][taps][(dw, c)] with dw < 64 / ldx.  The caller pads the row so that every window stays inside it.
  const bool window = cin > ldx;
  const int wpos = window ? cin / ldx : 1;
  MEBT_REQUIRE(B > 0 && To > 0 && Ho > 0 && Wo > 0 && cin > 0 && cout > 0 && ldx % 8 == 0 && ldy % 8 == 0 && ldy >= cout &&
               cout % 8 == 0 && ydims4[0] == B, MEBT_ERR_SHAPE, "conv3d: bad shape (channel counts / strides must be multiples of 8)");
  MEBT_REQUIRE(!window || (cin == 64 && taps3[2] == 1 && step3[2] == 1 && (odims3[2] - 1) + origin3[2] + wpos <= xdims4[3]),
               MEBT_ERR_SHAPE, "conv3d: ldx < cin is the window mode (cin = 64, one tap and unit step along w, rows padded by 64 / ldx - 1)");
  for (int d = 0; d < 3; ++d) {
    MEBT_REQUIRE(taps3[d] >= 1 && taps3[d] <= 4 && step3[d] >= 1 && step3[d] <= 2 && ystep3[d] >= 1 && ystep3[d] <= 2 &&
                 origin3[d] >= 0 && yorigin3[d] >= 0, MEBT_ERR_UNSUPPORTED, "conv3d: taps 1-4, steps 1-2");
    MEBT_REQUIRE((odims3[d] - 1) * step3[d] + origin3[d] + taps3[d] <= xdims4[1 + d], MEBT_ERR_SHAPE,
                 "conv3d: the input (dim %d: %d) is too small for %d outputs", d, xdims4[1 + d], odims3[d]);
    MEBT_REQUIRE((odims3[d] - 1) * ystep3[d] + yorigin3[d] < ydims4[1 + d], MEBT_ERR_SHAPE, "conv3d: the destination is too small");
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  // patch: as wide as possible along w, then h, then t (each a power of two dividing the output extent)
  auto pow2_div = [](int n, int cap) { int v = 1; while (v * 2 <= cap && n % (v * 2) == 0) v *= 2; return v; };
  const int bn = cout <= 64 ? 64 : (cout <= 128 ? 128 : 256);
  // ROW mode (see ConvSmem): whole output rows of 128 positions, the taps along w served by one box
  static const int row_env = [] { const char* e = getenv("MEBT_CONV_ROW"); return e != nullptr ? atoi(e) : 1; }();
  const bool row = row_env != 0 && bn == 64 && Wo % 128 == 0 && step3[2] == 1 && ystep3[2] == 1 && taps3[2] >= 1 &&
                   taps3[2] <= ROW_MAX_KW;
  if (row) {
    p.PW = 128; p.PH = 1; p.PT = 1;
  } else {
    p.PW = pow2_div(Wo, 16);
    p.PH = pow2_div(Ho, 128 / p.PW);
    p.PT = pow2_div(To, 128 / (p.PW * p.PH));
  }
  MEBT_REQUIRE(p.PT * p.PH * p.PW == 128, MEBT_ERR_UNSUPPORTED,
               "conv3d: the output extent %d x %d x %d does not tile into 128-position patches", To, Ho, Wo);
  p.nt_t = To / p.PT; p.nt_h = Ho / p.PH; p.nt_w = Wo / p.PW;
  p.tiles_m = B * p.nt_t * p.nt_h * p.nt_w;
  p.num_n_blocks = (cout + bn - 1) / bn;
  p.total_tiles = p.tiles_m * p.num_n_blocks;
  p.KT = taps3[0]; p.KH = taps3[1]; p.KW = taps3[2];
  p.ST = step3[0]; p.SH = step3[1]; p.SW = step3[2];
  p.OT = origin3[0]; p.OH = origin3[1]; p.OW = origin3[2];
  p.YT = ystep3[0]; p.YH = ystep3[1]; p.YW = ystep3[2];
  p.Y0T = yorigin3[0]; p.Y0H = yorigin3[1]; p.Y0W = yorigin3[2];
  p.cblocks = (cin + 63) / 64;
  p.Cout = cout;
  p.bias = bias;
  p.resid = static_cast<const __nv_bfloat16*>(resid);
  if (resid != nullptr) {
    MEBT_REQUIRE(ldr % 8 == 0 && ldr >= cout, MEBT_ERR_SHAPE, "conv3d: residual stride");
    p.res_w = ldr; p.res_h = (long long)Wo * ldr; p.res_t = (long long)Ho * Wo * ldr; p.res_b = (long long)To * Ho * Wo * ldr;
  }
  if (cout == 8 && ldy == 8 && resid == nullptr && ystep3[0] == 1 && ystep3[1] == 1 && ystep3[2] == 1 && yorigin3[0] == 0 &&
      yorigin3[1] == 0 && yorigin3[2] == 0) {
    p.y_direct = static_cast<__nv_bfloat16*>(y);
    p.yd_w = ldy; p.yd_h = (long long)ydims4[3] * ldy; p.yd_t = (long long)ydims4[2] * ydims4[3] * ldy;
    p.yd_b = (long long)ydims4[1] * ydims4[2] * ydims4[3] * ldy;
  }
  const int taps = p.KT * p.KH * p.KW;
  const uint64_t Kp = uint64_t(taps) * p.cblocks * 64;
  CUtensorMap tx, tw, ty;
  {
    const uint64_t dims[5] = {uint64_t(cin), uint64_t(xdims4[3] - (wpos - 1)), uint64_t(xdims4[2]), uint64_t(xdims4[1]), uint64_t(B)};
    const uint64_t strides[4] = {uint64_t(ldx) * 2, uint64_t(xdims4[3]) * ldx * 2, uint64_t(xdims4[2]) * xdims4[3] * ldx * 2,
                                 uint64_t(xdims4[1]) * xdims4[2] * xdims4[3] * ldx * 2};
    // ROW: one box = the row's 128 positions + the KW - 1 further ones its last taps read (h / t: a single line)
    const uint32_t box[5] = {64, uint32_t(row ? 128 + p.KW - 1 : p.PW * p.SW), uint32_t(row ? 1 : p.PH * p.SH),
                             uint32_t(row ? 1 : p.PT * p.ST), 1};
    const uint32_t es[5] = {1, uint32_t(row ? 1 : p.SW), uint32_t(row ? 1 : p.SH), uint32_t(row ? 1 : p.ST), 1};
    int rc = get_tensor_map_5d(&tx, xp, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    // rows in whole 64-wide tiles (zero rows behind cout): a box that is partly outside the tensor is filled by TMA at a
    // fraction of the rate of a plain box (measured: the 64 -> 3 / 64 -> 32 layers 213 us against 176 us for 64 -> 64)
    int rc = get_tensor_map_2d(&tw, w, 2, Kp, uint64_t((cout + 63) / 64 * 64), Kp * 2, 64, uint32_t(bn));
    if (rc) return rc;
  }
  {
    const uint64_t dims[5] = {uint64_t(cout), uint64_t(ydims4[3]), uint64_t(ydims4[2]), uint64_t(ydims4[1]), uint64_t(B)};
    const uint64_t strides[4] = {uint64_t(ldy) * 2, uint64_t(ydims4[3]) * ldy * 2, uint64_t(ydims4[2]) * ydims4[3] * ldy * 2,
                                 uint64_t(ydims4[1]) * ydims4[2] * ydims4[3] * ldy * 2};
    const uint32_t box[5] = {64, uint32_t(p.PW * p.YW), uint32_t(p.PH * p.YH), uint32_t(p.PT * p.YT), 1};
    const uint32_t es[5] = {1, uint32_t(p.YW), uint32_t(p.YH), uint32_t(p.YT), 1};
    int rc = get_tensor_map_5d(&ty, y, dims, strides, box, es);
    if (rc) return rc;
  }
  // window mode: counted as one tap of ldx channels per window (the zero-weighted slots are not work)
  const double flops = 2.0 * double(p.tiles_m) * 128.0 * double(cout) * double(taps) * double(window ? ldx : cin);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const int dual_env = [] { const char* e = getenv("MEBT_CONV_DUAL"); return e != nullptr ? atoi(e) : 1; }();
  const bool dual = dual_env != 0 && taps * p.cblocks >= 4;
  // row pairs once every SM has at least two rows (two tiles share the weight tiles of a stage)
  static const int row2_env = [] { const char* e = getenv("MEBT_CONV_ROW2"); return e != nullptr ? atoi(e) : 1; }();
  if (row && row2_env != 0 && p.KW <= 3 && p.total_tiles >= 2 * sm_count()) return launch_conv<64, 2>(tx, tw, ty, p, flops, st);
  if (row) return launch_conv<64, 1>(tx, tw, ty, p, flops, st);
  if (bn == 64) return dual ? launch_conv<64, 0, true>(tx, tw, ty, p, flops, st) : launch_conv<64, 0>(tx, tw, ty, p, flops, st);
  if (bn == 128) return dual ? launch_conv<128, 0, true>(tx, tw, ty, p, flops, st) : launch_conv<128, 0>(tx, tw, ty, p, flops, st);
  return launch_conv<256, 0>(tx, tw, ty, p, flops, st);
}

}  // extern "C"
