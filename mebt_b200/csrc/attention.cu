// K3: latent attention for the four MeBT block modes, head_dim 64, bf16 in / fp32 softmax.
//   O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] / sqrt(64) ) V[b,k,h,:]
// reference: CrossAttention.forward, mebt/modules/gpt.py:131-137 (bmm -> softmax -> bmm, materialising
// [B,h,NQ,NK] fp32), with the key sets of Block.forward (gpt.py:164-175):
//   latent_enc (256 x NC) | latent_self (256 x 256) | latent_dec (NT x 256) | lt2l (256 x (256 + NT)).
// lt2l's torch.cat([sos_emb, targets]) is never materialised: the kernel walks two K/V sources.
// NK == 0 (first draft step: no context) yields O = 0, like the empty softmax in the reference.
//
// Persistent CTAs (one per SM, 320 threads) walk a list of work items (batch element, head, PAIR of 128-query tiles):
//   warps 0-3 : softmax warpgroup 0 (query tile 2i)      warps 4-7 : softmax warpgroup 1 (query tile 2i+1)
//   warp 8    : TMA producer (the item's two Q tiles into a double buffer, K/V tiles of 128 keys through a 5-deep ring)
//   warp 9    : tcgen05.mma issuer: S_t = Q_t K^T (M128 N128 K64, both operands in smem) and
//               O_t += P_t V (M128 N64 K128, P read from TENSOR MEMORY, V from smem)
// (the two single-thread roles carry the highest warp ids: the sub-partition arbiter favours the highest eligible warp
// id, which keeps the MMA issuer from queueing behind the softmax warps it shares a scheduler with)
// Both warpgroups walk the SAME K/V tiles (each tile is loaded once and feeds four MMAs), each with its own S, P and O
// regions of tensor memory (128 + 64 + 64 columns); while one exponentiates, the other's S / PV round trip runs.
// Each softmax thread owns one query row (= its TMEM lane): it reads S with tcgen05.ld, forms P = 2^(s c - m c) in
// bf16 and writes it back with tcgen05.st in the packed layout the PV MMA takes its A operand in, so P never touches
// shared memory (with P in smem a 128 x 128 tile costs 80 KB of MMA operand reads + 32 KB of P writes against
// 128 B/clk).  O accumulates in TMEM across K/V tiles.
// The stabiliser m is LAZY: it is seeded from the first 32 scores of the row and only advanced when a score exceeds
// it by more than 8 in the log2 domain (P would exceed 256) - the exact result does not depend on the stabiliser,
// only overflow safety does - and only then is O rescaled in TMEM (warp-uniform branch, tcgen05.ld / st).
// The exponentials need 1024 clk of MUFU per tile (one per score, 16 / clk / SM) against 512 clk of nominal MMA time, so
// one exponential in four is evaluated on the FMA pipe (Cody-Waite split + degree-3 polynomial, relative error 7.5e-5,
// far below the bf16 rounding of P).  Measured (tools/attn_bench.cu): the kernel runs at about
// 2000 clk per tile and what bounds it is the issue rate of the small-N tcgen05.mma instructions (about 1430 clk per
// tile for the 4 + 8 MMAs even with the softmax disabled), not the MUFU or the shared-memory pipe.
#include <type_traits>
#include "common.cuh"

namespace mebt {
namespace {

constexpr int AT_BQ = 128;
constexpr int AT_BKV = 128;
constexpr int AT_HS = 64;
#ifndef MEBT_ATTN_TWO_ISSUERS
#define MEBT_ATTN_TWO_ISSUERS 1     // 1: S = Q K^T and O += P V are issued by two different threads (warps 9 and 10)
#endif
constexpr int AT_THREADS = MEBT_ATTN_TWO_ISSUERS ? 352 : 320;
constexpr int AT_TILE_BYTES = 128 * 64 * 2;      // 16 KiB: a [128 x 64] bf16 tile (Q, K or V)
constexpr int AT_KV_STAGES = 5;
constexpr int AT_MAX_SPLITS = 8;                 // split-KV fan-out (workspace = AT_MAX_SPLITS partial outputs)
constexpr int AT_SMEM_Q = 0;                                     // 2 buffers x 2 query tiles
constexpr int AT_SMEM_K = AT_SMEM_Q + 4 * AT_TILE_BYTES;
constexpr int AT_SMEM_V = AT_SMEM_K + AT_KV_STAGES * AT_TILE_BYTES;
constexpr int AT_SMEM_BAR = AT_SMEM_V + AT_KV_STAGES * AT_TILE_BYTES;
constexpr int AT_SMEM_TOTAL = AT_SMEM_BAR + 512;
constexpr uint32_t AT_TMEM_COLS = 512;           // S_t: 128 t    O_t: 256 + 64 t    P_t: 384 + 64 t
static_assert(AT_SMEM_TOTAL <= 232448, "shared memory budget");

#ifndef MEBT_ATTN_POLY
#define MEBT_ATTN_POLY 1      // 1: every fourth exponential on the FMA pipe
#endif

#ifdef MEBT_ATTN_TRACE
long long* g_attn_trace = nullptr;
#define ATR_BEGIN tr_t = clock64()
#define ATR_END(i) tr_acc[i] += clock64() - tr_t
#else
#define ATR_BEGIN
#define ATR_END(i)
#endif

struct AttnParams {
  int NQ, NK1, NK2, H, B;
  int q_col0, k1_col0, v1_col0, k2_col0, v2_col0;
  __nv_bfloat16* O;
  int ldo;
  float* lse;              // optional [B, H, NQ]
  float scale_log2;        // log2(e) / sqrt(hs)
  float scale;             // 1 / sqrt(hs)
  long long* trace;
  DropKey drop;            // attention-probability dropout (training); thr == 0: off
  // split-KV (few work items, long key lists: small-batch sampling): `splits` CTAs share an item, each walks `nts` of
  // its K/V tiles and leaves an UNNORMALISED fp32 O row plus its (stabiliser m, row sum l); attention_combine_kernel
  // merges them.  splits == 1: nts = all tiles, O is final.
  int splits, nts;
  float* part_o;           // [splits, B, H, NQ, 64]
  float* part_ml;          // [splits, B, H, NQ, 2]
};

// 2^x for x <= ~8 on the FMA pipe: n = round(x), f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial, the
// exponent added as an integer.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float xf = x + 12582912.f;               // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (xf - 12582912.f);
  float pl = fmaf(0.0551716685f, f, 0.2426111251f);
  pl = fmaf(pl, f, 0.6932609677f);
  pl = fmaf(pl, f, 0.9999280572f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(xf) << 23));
}

// the same for two values at once on the packed f32x2 FMA path (half the issue slots per element)
__device__ __forceinline__ float2 ex2_poly_x2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 xf = __fadd2_rn(x, magic);
  const float2 fl = __fadd2_rn(xf, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(fl, make_float2(-1.f, -1.f), x);
  float2 pl = __ffma2_rn(make_float2(0.0551716685f, 0.0551716685f), f, make_float2(0.2426111251f, 0.2426111251f));
  pl = __ffma2_rn(pl, f, make_float2(0.6932609677f, 0.6932609677f));
  pl = __ffma2_rn(pl, f, make_float2(0.9999280572f, 0.9999280572f));
  return make_float2(__int_as_float(__float_as_int(pl.x) + (__float_as_int(xf.x) << 23)),
                     __int_as_float(__float_as_int(pl.y) + (__float_as_int(xf.y) << 23)));
}

template <bool DROP>
__global__ void __launch_bounds__(AT_THREADS, 1)
latent_attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv1,
                            const __grid_constant__ CUtensorMap tm_kv2, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_SMEM_BAR);
  uint64_t* q_full = bars + 0;     // [2] Q buffers (a pair of query tiles each)
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;    // [AT_KV_STAGES]
  uint64_t* kv_empty = bars + 12;  // [AT_KV_STAGES]
  uint64_t* s_full = bars + 20;    // [2] per warpgroup
  uint64_t* p_full = bars + 22;    // [2]
  uint64_t* pv_done = bars + 24;   // [2] the warpgroup's latest PV has retired: O includes it, P may be overwritten
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles1 = (p.NK1 + AT_BKV - 1) / AT_BKV;
  const int tiles2 = (p.NK2 + AT_BKV - 1) / AT_BKV;
  const int nt = tiles1 + tiles2;
  const int q_tiles = (p.NQ + AT_BQ - 1) / AT_BQ;
  const int q_pairs = (q_tiles + 1) >> 1;
  const int n_items = p.B * p.H * q_pairs * p.splits;
  const int my_items = int(blockIdx.x) < n_items ? (n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;
  const int nts = p.nts;           // K/V tiles per item (= nt without split-KV)
  // item -> (b, h, K/V split, query-tile pair); the pair runs fastest so that consecutive CTAs share a (b, h)'s K/V in L2
  auto item_coords = [&](int n, int& b, int& h, int& qp, int& sp) {
    const int item = int(blockIdx.x) + n * int(gridDim.x);
    qp = item % q_pairs;
    const int rest = item / q_pairs;
    sp = rest % p.splits;
    const int bh = rest / p.splits;
    h = bh % p.H;
    b = bh / p.H;
  };

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm_q);
    prefetch_tensormap(&tm_kv1);
    prefetch_tensormap(&tm_kv2);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); mbar_init(&pv_done[s], 1);
    }
    for (int s = 0; s < AT_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_ptr_smem, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  griddep_wait();
  // the softmax warpgroups take the registers the single-thread roles do not need (S row + P row live per thread)

  if (warp == 8) {
    if (lane == 0 && nt > 0) {
      int gj = 0;                    // K/V tiles loaded by this CTA
      for (int n = 0; n < my_items; ++n) {
        int b, h, qp, sp;
        item_coords(n, b, h, qp, sp);
        const int qb = n & 1;
        mbar_wait_backoff(&q_empty[qb], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], 2 * AT_TILE_BYTES);
        // a pair's second tile may lie past the last query row: the rows it reads instead (the next batch element's,
        // or TMA zero fill past the tensor) are computed and never stored
        tma_load_2d(smem + AT_SMEM_Q + (2 * qb) * AT_TILE_BYTES, &tm_q, &q_full[qb], p.q_col0 + h * AT_HS,
                    b * p.NQ + (2 * qp) * AT_BQ);
        tma_load_2d(smem + AT_SMEM_Q + (2 * qb + 1) * AT_TILE_BYTES, &tm_q, &q_full[qb], p.q_col0 + h * AT_HS,
                    b * p.NQ + (2 * qp + 1) * AT_BQ);
        for (int jl = 0; jl < nts; ++jl, ++gj) {
          const int j = sp * nts + jl;                 // tile of the item's whole key list
          const int s = gj % AT_KV_STAGES;
          mbar_wait_backoff(&kv_empty[s], ((gj / AT_KV_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * AT_TILE_BYTES);
          if (j < tiles1) {
            const int r = b * p.NK1 + j * AT_BKV;
            tma_load_2d(smem + AT_SMEM_K + s * AT_TILE_BYTES, &tm_kv1, &kv_full[s], p.k1_col0 + h * AT_HS, r);
            tma_load_2d(smem + AT_SMEM_V + s * AT_TILE_BYTES, &tm_kv1, &kv_full[s], p.v1_col0 + h * AT_HS, r);
          } else {
            const int r = b * p.NK2 + (j - tiles1) * AT_BKV;
            tma_load_2d(smem + AT_SMEM_K + s * AT_TILE_BYTES, &tm_kv2, &kv_full[s], p.k2_col0 + h * AT_HS, r);
            tma_load_2d(smem + AT_SMEM_V + s * AT_TILE_BYTES, &tm_kv2, &kv_full[s], p.v2_col0 + h * AT_HS, r);
          }
        }
      }
    }
#if MEBT_ATTN_TWO_ISSUERS
  // Two MMA-issuing threads.  With one, that thread was busy ~83 % of the kernel (tools/attn_bench.cu: ~1700 clk per
  // (query tile, K/V tile) step, 1000 of them inside the eight PV instructions, whose issue competes for scheduler slots
  // with the two softmax warps of its sub-partition) while the softmax warpgroups spent 63 % of their time waiting for S.
  // S = Q K^T (warp 9) and O += P V (warp 10) touch different accumulators and are ordered through the warpgroups'
  // barriers, so the result does not depend on how the two threads interleave.
  } else if (warp == 9) {
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
      const uint32_t q_lo0 = smem_desc_lo(smem_u32(smem + AT_SMEM_Q), 16);          // Q, K: K-major, 32 B per K = 16 step
      const uint32_t k_lo0 = smem_desc_lo(smem_u32(smem + AT_SMEM_K), 16);
      const int total = my_items * nts * 2;       // step g = 2 * (K/V tile counter gj) + warpgroup t
      for (int g = 0; g < total; ++g) {
        const int gj = g >> 1, t = g & 1;
        const int n = gj / nts, j = gj - n * nts;
        if (t == 0) {
          if (j == 0) mbar_wait(&q_full[n & 1], (n >> 1) & 1);
          mbar_wait(&kv_full[gj % AT_KV_STAGES], (gj / AT_KV_STAGES) & 1);
        }
        if (gj > 0) mbar_wait(&p_full[t], (gj - 1) & 1);     // warpgroup t has consumed the S of tile gj-1
        tc_fence_after();
        const uint32_t q_lo = q_lo0 + uint32_t(2 * (n & 1) + t) * uint32_t(AT_TILE_BYTES >> 4);
        const uint32_t k_lo = k_lo0 + uint32_t(gj % AT_KV_STAGES) * uint32_t(AT_TILE_BYTES >> 4);
        umma_bf16_ss_x4<false>(tmem_base + t * 128, q_lo, k_lo, 2, 2, kDescHi, kDescHi, idesc_s, 0u);
        umma_commit(&s_full[t]);
        if (t == 1 && j == nts - 1) umma_commit(&q_empty[n & 1]);     // the item's last S: its Q buffer is free
      }
    }
  } else if (warp == 10) {
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // P (TMEM, K-major) x V (MN-major: hs contiguous)
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
      const uint32_t v_lo0 = smem_desc_lo(smem_u32(smem + AT_SMEM_V), 64 * 128);    // V: MN-major, 2048 B per 16 keys
#ifdef MEBT_ATTN_TRACE
      long long mt[3] = {0, 0, 0}, mt0 = 0;
      const long long mstart = clock64();
#define MTR_BEGIN mt0 = clock64()
#define MTR_END(i) mt[i] += clock64() - mt0
#else
#define MTR_BEGIN
#define MTR_END(i)
#endif
      const int total = my_items * nts * 2;
      for (int g = 0; g < total; ++g) {
        const int t = g & 1, gj = g >> 1;
        const int j = gj % nts;
        MTR_BEGIN;
        if (t == 0) mbar_wait(&kv_full[gj % AT_KV_STAGES], (gj / AT_KV_STAGES) & 1);     // (complete long ago: this thread's own view of V)
        MTR_END(1);
        // warpgroup t is done with tile gj: S consumed, P in TMEM (and, at j == 0, the previous item's O read)
        MTR_BEGIN;
        mbar_wait(&p_full[t], gj & 1);
        MTR_END(0);
        tc_fence_after();
        const uint32_t v_lo = v_lo0 + uint32_t(gj % AT_KV_STAGES) * uint32_t(AT_TILE_BYTES >> 4);
        MTR_BEGIN;
        umma_bf16_ts_x8(tmem_base + 256 + t * 64, tmem_base + 384 + t * 64, 8u, v_lo, 2048u >> 4, kDescHi, idesc_o,
                        j != 0 ? 1u : 0u);
        MTR_END(2);
        umma_commit(&pv_done[t]);
        // both warpgroups' PV of this K/V stage have retired, hence (through the warpgroups) their S as well
        if (t == 1) umma_commit(&kv_empty[gj % AT_KV_STAGES]);
      }
#ifdef MEBT_ATTN_TRACE
      if (p.trace != nullptr) { long long* o = p.trace + 148 * 8 + blockIdx.x * 4; o[0] = mt[0]; o[1] = mt[1]; o[2] = mt[2]; o[3] = clock64() - mstart; }
#endif
    }
#else
  } else if (warp == 9) {
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // P (TMEM, K-major) x V (MN-major: hs contiguous)
#ifdef MEBT_ATTN_TRACE
      long long mt[3] = {0, 0, 0}, mt0 = 0;
      const long long mstart = clock64();
#define MTR_BEGIN mt0 = clock64()
#define MTR_END(i) mt[i] += clock64() - mt0
#else
#define MTR_BEGIN
#define MTR_END(i)
#endif
      constexpr uint32_t kDescHi = smem_desc_hi_sw128(1024);
      const uint32_t q_lo0 = smem_desc_lo(smem_u32(smem + AT_SMEM_Q), 16);          // Q, K: K-major, 32 B per K = 16 step
      const uint32_t k_lo0 = smem_desc_lo(smem_u32(smem + AT_SMEM_K), 16);
      const uint32_t v_lo0 = smem_desc_lo(smem_u32(smem + AT_SMEM_V), 64 * 128);    // V: MN-major, 2048 B per 16 keys
      // step g = 2 * (K/V tile counter gj) + warpgroup t
      const int total = my_items * nts * 2;
      // S_t of tile gj: legal once warpgroup t has consumed the S of tile gj-1
      auto issue_s = [&](int g, bool commit) {
        const int gj = g >> 1, t = g & 1;
        const int n = gj / nts, j = gj - n * nts;
        if (t == 0) {
          MTR_BEGIN;
          if (j == 0) mbar_wait(&q_full[n & 1], (n >> 1) & 1);
          mbar_wait(&kv_full[gj % AT_KV_STAGES], (gj / AT_KV_STAGES) & 1);
          MTR_END(1);
          tc_fence_after();
        }
        // the issuing thread is what bounds this kernel (tools/mma_probe.cu: ~48 clk per instruction issued from one
        // block over precomputed descriptor words against ~97 with a descriptor rebuilt per instruction; ~165 clk per
        // commit, ~163 per barrier wait), so: four / eight instructions per asm block, descriptor low words advanced by adds
        const uint32_t q_lo = q_lo0 + uint32_t(2 * (n & 1) + t) * uint32_t(AT_TILE_BYTES >> 4);
        const uint32_t k_lo = k_lo0 + uint32_t(gj % AT_KV_STAGES) * uint32_t(AT_TILE_BYTES >> 4);
        umma_bf16_ss_x4<false>(tmem_base + t * 128, q_lo, k_lo, 2, 2, kDescHi, kDescHi, idesc_s, 0u);
        if (commit) umma_commit(&s_full[t]);
        if (t == 1 && j == nts - 1) umma_commit(&q_empty[n & 1]);     // the item's last S: its Q buffer is free
      };
      if (total > 0) { issue_s(0, true); issue_s(1, true); }
      for (int g = 0; g < total; ++g) {
        const int t = g & 1, gj = g >> 1;
        const int j = gj % nts;
        // warpgroup t is done with tile gj: S consumed, P in TMEM (and, at j == 0, the previous item's O read)
        MTR_BEGIN;
        mbar_wait(&p_full[t], gj & 1);
        MTR_END(0);
        tc_fence_after();
        // PV of this tile, then the warpgroup's next S, then ONE commit on s_full[t]: "S of tile gj+1 is ready" then also
        // means "PV of tile gj has retired" (O includes it, P may be overwritten), which saves a commit here (~165 clk of
        // this thread) and a barrier wait in the warpgroup (~160-250 clk) per tile; the kernel is bound by these
        // hand-offs, not by the tensor pipe (tools/mma_probe.cu).  The very last PV of a warpgroup signals pv_done.
        const uint32_t v_lo = v_lo0 + uint32_t(gj % AT_KV_STAGES) * uint32_t(AT_TILE_BYTES >> 4);
        MTR_BEGIN;
        umma_bf16_ts_x8(tmem_base + 256 + t * 64, tmem_base + 384 + t * 64, 8u, v_lo, 2048u >> 4, kDescHi, idesc_o,
                        j != 0 ? 1u : 0u);
        MTR_END(2);
        if (g + 2 < total) {
          issue_s(g + 2, false);
          umma_commit(&s_full[t]);
        } else {
          umma_commit(&pv_done[t]);
        }
        if (t == 1) umma_commit(&kv_empty[gj % AT_KV_STAGES]);   // both warpgroups' S and PV of this K/V stage are done
      }
#ifdef MEBT_ATTN_TRACE
      if (p.trace != nullptr) { long long* o = p.trace + 148 * 8 + blockIdx.x * 4; o[0] = mt[0]; o[1] = mt[1]; o[2] = mt[2]; o[3] = clock64() - mstart; }
#endif
    }
#endif
  } else {
    // ===== softmax warpgroups =====
    const int t = warp >> 2;                      // warpgroup: query tile 2 * pair + t
    const int q = warp & 3;                       // TMEM lane quarter
    const int row = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t tmem_s = tmem_base + t * 128 + lane_addr;
    const uint32_t tmem_o = tmem_base + 256 + t * 64 + lane_addr;
    const uint32_t tmem_p = tmem_base + 384 + t * 64 + lane_addr;
    int k = 0;                                    // K/V tiles this warpgroup has processed (across items)
#ifdef MEBT_ATTN_TRACE
    long long tr_t = 0, tr_acc[4] = {0, 0, 0, 0};
    const long long tr_start = clock64();
#endif
    for (int n = 0; n < my_items; ++n) {
      int b, h, qp, sp;
      item_coords(n, b, h, qp, sp);
      const int qt = 2 * qp + t;
      const bool active = qt < q_tiles;            // an odd tile count leaves warpgroup 1 without a tile in the last pair
      float m = 0.f, l = 0.f;
      const uint32_t drop_rk = DROP ? drop_row_key(p.drop, uint32_t((b * p.H + h) * p.NQ + qt * AT_BQ + row)) : 0u;
      for (int j = 0; j < nts; ++j, ++k) {
        const int jg = sp * nts + j;                   // tile of the item's whole key list
        const int rem = jg < tiles1 ? p.NK1 - jg * AT_BKV : p.NK2 - (jg - tiles1) * AT_BKV;
        const int valid = min(AT_BKV, rem);
        const uint32_t pair0 = jg < tiles1 ? uint32_t(jg * (AT_BKV / 2)) : (1u << 19) | uint32_t((jg - tiles1) * (AT_BKV / 2));
        const bool full = valid == AT_BKV;            // warp-uniform: only a source's last tile can be ragged
        ATR_BEGIN;
        mbar_wait(&s_full[t], k & 1);
        ATR_END(0);
        tc_fence_after();
        if (active) {
          // 32 scores -> 16 columns of packed bf16 P
          auto chunk = [&](auto full_tag, const uint32_t (&r)[32], int c, float mb, float (&l4)[4], float (&m4)[4]) {
            constexpr bool FULL = decltype(full_tag)::value;
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float pe[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float sv = __uint_as_float(r[i + e]);
                const bool ok = FULL || c * 32 + i + e < valid;
                if (ok) m4[(i + e) & 3] = fmaxf(m4[(i + e) & 3], sv);
                const float x = fmaf(sv, p.scale_log2, -mb);
                float v;
                if (MEBT_ATTN_POLY && ((i + e) & 3) == 3) v = ex2_poly(x); else v = ex2_approx(x);
                pe[e] = ok ? v : 0.f;
                l4[(i + e) & 3] += pe[e];
              }
              if (DROP) {              // the row sum keeps the undropped probabilities (dropout follows the softmax)
                float f0, f1;
                drop_pair(p.drop, drop_rk, pair0 + uint32_t(c * 16 + (i >> 1)), f0, f1);
                pe[0] *= f0; pe[1] *= f1;
              }
              pk[i >> 1] = pack_bf16x2(pe[0], pe[1]);
            }
            tmem_st_32x16(tmem_p + c * 16, pk);
          };
          // One pass over the S row (the next 32 columns are in flight while the current ones are exponentiated); the
          // first P store waits for the previous tile's PV (it reads P).
          auto row_pass = [&](auto full_tag, float mb, float& l_tile, float& mx_tile, bool seed) {
            float l4[4] = {0.f, 0.f, 0.f, 0.f};
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(tmem_s, ra);
            tmem_ld_32x32(tmem_s + 32, rb);
            tmem_ld_wait_regs(ra);
            if (seed) {                               // seed the stabiliser from the first 32 scores
              float mx = __uint_as_float(ra[0]);
#pragma unroll
              for (int i = 1; i < 32; ++i)
                if (i < valid) mx = fmaxf(mx, __uint_as_float(ra[i]));
              m = mx;
              mb = mx * p.scale_log2;
            }
#if MEBT_ATTN_TWO_ISSUERS
            if (k > 0) {                              // the previous tile's PV reads P: wait before overwriting it
              ATR_BEGIN;
              mbar_wait(&pv_done[t], (k - 1) & 1);
              ATR_END(1);
              tc_fence_after();
            }
#endif
            // (one issuer: s_full of this tile was committed behind the previous tile's PV)
            chunk(full_tag, ra, 0, mb, l4, m4);
            tmem_ld_wait_regs(rb);
            tmem_ld_32x32(tmem_s + 64, ra);
            chunk(full_tag, rb, 1, mb, l4, m4);
            tmem_ld_wait_regs(ra);
            tmem_ld_32x32(tmem_s + 96, rb);
            chunk(full_tag, ra, 2, mb, l4, m4);
            tmem_ld_wait_regs(rb);
            chunk(full_tag, rb, 3, mb, l4, m4);
            l_tile = (l4[0] + l4[1]) + (l4[2] + l4[3]);
            mx_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          };
          // Fast pass (full tiles without dropout - all but a source's last tile): the softmax warps are bound by their
          // own issue slots and the MUFU pipe (two warps share a scheduler: per K/V tile 2 x 128 elements x (~6.5 slots,
          // 8 MUFU clk for three in four)), so the arithmetic is PACKED: exponent argument and row sum in f32x2, three
          // pairs in eight exponentiated by the packed polynomial, and instead of the running maximum of the raw
          // scores the maximum of the bf16 P pairs (one HMNMX2 per pair): P > 2^8 is exactly the lazy stabiliser's
          // overflow test, and only then (rare) does the slow pass below run to find the new maximum.
          auto chunk_fast = [&](const uint32_t (&r)[32], int c, const float2 sc2, const float2 nmb2, float2& lsum,
                                __nv_bfloat162& pmax) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, nmb2);
              float2 pv;
              if (MEBT_ATTN_POLY && ((i & 7) == 1 || (i & 7) == 4 || (i & 7) == 6)) {
                pv = ex2_poly_x2(x);
              } else {
                pv.x = ex2_approx(x.x);
                pv.y = ex2_approx(x.y);
              }
              lsum = __fadd2_rn(lsum, pv);
              pk[i] = pack_bf16x2(pv.x, pv.y);
              pmax = __hmax2(pmax, *reinterpret_cast<const __nv_bfloat162*>(&pk[i]));
            }
            tmem_st_32x16(tmem_p + c * 16, pk);
          };
          auto row_pass_fast = [&](float& l_tile, bool seed) -> bool {
            uint32_t ra[32], rb[32];
            tmem_ld_32x32(tmem_s, ra);
            tmem_ld_32x32(tmem_s + 32, rb);
            tmem_ld_wait_regs(ra);
            if (seed) {                               // seed the stabiliser from the first 32 scores
              float mx = __uint_as_float(ra[0]);
#pragma unroll
              for (int i = 1; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(ra[i]));
              m = mx;
            }
            const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
            const float2 nmb2 = make_float2(-m * p.scale_log2, -m * p.scale_log2);
            float2 lsum = make_float2(0.f, 0.f);
            __nv_bfloat162 pmax = __float2bfloat162_rn(0.f);
#if MEBT_ATTN_TWO_ISSUERS
            if (k > 0) {
              ATR_BEGIN;
              mbar_wait(&pv_done[t], (k - 1) & 1);
              ATR_END(1);
              tc_fence_after();
            }
#endif
            chunk_fast(ra, 0, sc2, nmb2, lsum, pmax);
            tmem_ld_wait_regs(rb);
            tmem_ld_32x32(tmem_s + 64, ra);
            chunk_fast(rb, 1, sc2, nmb2, lsum, pmax);
            tmem_ld_wait_regs(ra);
            tmem_ld_32x32(tmem_s + 96, rb);
            chunk_fast(ra, 2, sc2, nmb2, lsum, pmax);
            tmem_ld_wait_regs(rb);
            chunk_fast(rb, 3, sc2, nmb2, lsum, pmax);
            l_tile = lsum.x + lsum.y;
            const float pm = fmaxf(__low2float(pmax), __high2float(pmax));
            return !(pm <= 256.0f);                   // also true for inf / NaN
          };
          float l_tile, mx_tile;
#ifdef MEBT_ATTN_TRACE
          const long long rp0 = clock64(), w0 = tr_acc[1];
#endif
          bool need_slow = true;
          if (full && !DROP) need_slow = __any_sync(0xffffffffu, row_pass_fast(l_tile, j == 0));
          mx_tile = m;
          if (need_slow) {
            // (after a fast pass that overflowed the stabiliser is already seeded: do not re-seed)
            const bool seed_here = j == 0 && !(full && !DROP);
            if (full) row_pass(std::true_type{}, m * p.scale_log2, l_tile, mx_tile, seed_here);
            else row_pass(std::false_type{}, m * p.scale_log2, l_tile, mx_tile, seed_here);
          }
#ifdef MEBT_ATTN_TRACE
          tr_acc[2] += (clock64() - rp0) - (tr_acc[1] - w0);
#endif
          const bool grow = (mx_tile - m) * p.scale_log2 > 8.0f;
          if (__any_sync(0xffffffffu, grow)) {        // rare: re-base on the new maximum, rescale O, redo this tile
            const float alpha = grow ? ex2_approx((m - mx_tile) * p.scale_log2) : 1.f;
            if (grow) { l *= alpha; m = mx_tile; }
            if (j > 0) {                              // (the PV of tile k-1 has retired: row_pass waited for it)
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint32_t ro[32];
                tmem_ld_32x32(tmem_o + c * 32, ro);
                tmem_ld_wait_regs(ro);
#pragma unroll
                for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
                tmem_st_32x32(tmem_o + c * 32, ro);
              }
            }
            if (full) row_pass(std::true_type{}, m * p.scale_log2, l_tile, mx_tile, false);
            else row_pass(std::false_type{}, m * p.scale_log2, l_tile, mx_tile, false);
          }
          l += l_tile;
          tmem_st_wait();
        }
#if MEBT_ATTN_TWO_ISSUERS
        // a warpgroup without a tile (odd tile count) must not run ahead of the PV issuer either: parity waits alias
        // when a barrier gets two phases ahead of a waiter
        if (!active && k > 0) mbar_wait(&pv_done[t], (k - 1) & 1);
#endif
        tc_fence_before();
        mbar_arrive(&p_full[t]);
      }
      const int qrow = qt * AT_BQ + row;
      if (nt > 0) {                     // the item's last PV: O complete
        ATR_BEGIN;
        // ... signalled with the next item's first S (same commit), or, behind this CTA's last tile, on pv_done
#if MEBT_ATTN_TWO_ISSUERS
        mbar_wait(&pv_done[t], (k - 1) & 1);
#else
        if (n + 1 < my_items) mbar_wait(&s_full[t], k & 1);
        else mbar_wait(&pv_done[t], 0);
#endif
        ATR_END(1);
        tc_fence_after();
      }
      if (active && p.splits > 1) {
        // split-KV: leave the unnormalised row and its (m, l) for the combine kernel
        const size_t prow = ((size_t(sp) * p.B + b) * p.H + h) * p.NQ + qrow;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t ro[32];
          tmem_ld_32x32(tmem_o + c * 32, ro);
          tmem_ld_wait_regs(ro);
          if (qrow < p.NQ) {
            float4* d4 = reinterpret_cast<float4*>(p.part_o + prow * AT_HS + c * 32);
#pragma unroll
            for (int gq = 0; gq < 8; ++gq)
              d4[gq] = make_float4(__uint_as_float(ro[4 * gq]), __uint_as_float(ro[4 * gq + 1]), __uint_as_float(ro[4 * gq + 2]),
                                   __uint_as_float(ro[4 * gq + 3]));
          }
        }
        if (qrow < p.NQ) *reinterpret_cast<float2*>(p.part_ml + prow * 2) = make_float2(m, l);
      } else if (active) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        uint4* dst = reinterpret_cast<uint4*>(p.O + (size_t(b) * p.NQ + qrow) * p.ldo + h * AT_HS);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t ro[32];
          if (nt > 0) {
            tmem_ld_32x32(tmem_o + c * 32, ro);
            tmem_ld_wait_regs(ro);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) ro[i] = 0u;
          }
          if (qrow < p.NQ) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(ro[8 * gq + 0]) * inv, __uint_as_float(ro[8 * gq + 1]) * inv);
              u.y = pack_bf16x2(__uint_as_float(ro[8 * gq + 2]) * inv, __uint_as_float(ro[8 * gq + 3]) * inv);
              u.z = pack_bf16x2(__uint_as_float(ro[8 * gq + 4]) * inv, __uint_as_float(ro[8 * gq + 5]) * inv);
              u.w = pack_bf16x2(__uint_as_float(ro[8 * gq + 6]) * inv, __uint_as_float(ro[8 * gq + 7]) * inv);
              dst[c * 4 + gq] = u;
            }
          }
        }
        if (qrow < p.NQ && p.lse != nullptr)
          p.lse[(size_t(b) * p.H + h) * p.NQ + qrow] = l > 0.f ? m * p.scale + logf(l) : -INFINITY;
      }
    }
#ifdef MEBT_ATTN_TRACE
    if (p.trace != nullptr && lane == 0 && q == 0) {
      long long* o = p.trace + (blockIdx.x * 2 + t) * 4;
      o[0] = tr_acc[0]; o[1] = tr_acc[1]; o[2] = tr_acc[2]; o[3] = clock64() - tr_start;
    }
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

// Merge of the split-KV partials: one warp per (b, h, q) row, lane = two of the 64 head dimensions.
//   M = max_s m_s,  w_s = 2^((m_s - M) c),  O = sum_s w_s O_s / sum_s w_s l_s        (c = log2(e) / sqrt(hs))
__global__ void __launch_bounds__(256) attention_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                                                int splits, int B, int H, int NQ, float scale_log2,
                                                                __nv_bfloat16* __restrict__ O, int ldo) {
  griddep_wait();
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * H * NQ;
  if (row >= rows) return;
  float M = -INFINITY;
  for (int sidx = 0; sidx < splits; ++sidx) {
    const float2 ml = *reinterpret_cast<const float2*>(part_ml + (sidx * rows + row) * 2);
    if (ml.y > 0.f) M = fmaxf(M, ml.x);
  }
  float o0 = 0.f, o1 = 0.f, L = 0.f;
  for (int sidx = 0; sidx < splits; ++sidx) {
    const float2 ml = *reinterpret_cast<const float2*>(part_ml + (sidx * rows + row) * 2);
    if (!(ml.y > 0.f)) continue;
    const float w = ex2_approx((ml.x - M) * scale_log2);
    const float2 v = *reinterpret_cast<const float2*>(part_o + (sidx * rows + row) * AT_HS + lane * 2);
    o0 = fmaf(w, v.x, o0);
    o1 = fmaf(w, v.y, o1);
    L = fmaf(w, ml.y, L);
  }
  const float inv = L > 0.f ? 1.f / L : 0.f;
  const int q = int(row % NQ);
  const long long bh = row / NQ;
  const int h = int(bh % H), b = int(bh / H);
  *reinterpret_cast<uint32_t*>(O + (size_t(b) * NQ + q) * ldo + h * AT_HS + lane * 2) = pack_bf16x2(o0 * inv, o1 * inv);
}

}  // namespace
}  // namespace mebt

#ifdef MEBT_ATTN_TRACE
extern "C" void mebt_attn_set_trace(long long* buf) { mebt::g_attn_trace = buf; }
#endif

namespace mebt {
size_t latent_attention_fwd_workspace_bytes(int B, int H, int NQ) {
  return size_t(AT_MAX_SPLITS) * B * H * NQ * (AT_HS + 2) * sizeof(float);
}

int latent_attention_fwd(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0, int v1_col0, int NK1,
                         const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2, void* O, int ldo, float* lse, int B,
                         int H, int NQ, int head_dim, float drop_p, unsigned long long drop_seed, void* workspace,
                         size_t workspace_bytes, void* stream) {
  MEBT_REQUIRE(head_dim == AT_HS, MEBT_ERR_UNSUPPORTED, "attention: head_dim %d unsupported (every MeBT config uses 64)",
               head_dim);
  MEBT_REQUIRE(B > 0 && H > 0 && NQ > 0 && NK1 >= 0 && NK2 >= 0, MEBT_ERR_SHAPE, "attention: bad shape");
  MEBT_REQUIRE(ldq % 8 == 0 && ldo % 8 == 0 && q_col0 % 8 == 0, MEBT_ERR_SHAPE, "attention: Q/O strides must be 16B aligned");
  MEBT_REQUIRE(NK1 == 0 || (KV1 != nullptr && ld1 % 8 == 0), MEBT_ERR_SHAPE, "attention: bad KV1");
  MEBT_REQUIRE(NK2 == 0 || (KV2 != nullptr && ld2 % 8 == 0), MEBT_ERR_SHAPE, "attention: bad KV2");
  MEBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MEBT_ERR_SHAPE, "attention: dropout p = %f outside [0, 1)", drop_p);
  MEBT_REQUIRE(NK1 < (1 << 20) && NK2 < (1 << 20), MEBT_ERR_SHAPE, "attention: more than 2^20 keys per source");
  CUtensorMap tq, t1, t2;
  int rc = get_tensor_map_2d(&tq, Q, 2, uint64_t(ldq), uint64_t(B) * NQ, uint64_t(ldq) * 2, 64, 128);
  if (rc) return rc;
  // an absent source still needs a valid descriptor object; alias the query map (never dereferenced: 0 tiles)
  t1 = tq;
  t2 = tq;
  if (NK1 > 0) {
    rc = get_tensor_map_2d(&t1, KV1, 2, uint64_t(ld1), uint64_t(B) * NK1, uint64_t(ld1) * 2, 64, AT_BKV);
    if (rc) return rc;
  }
  if (NK2 > 0) {
    rc = get_tensor_map_2d(&t2, KV2, 2, uint64_t(ld2), uint64_t(B) * NK2, uint64_t(ld2) * 2, 64, AT_BKV);
    if (rc) return rc;
  }
  AttnParams p;
  p.NQ = NQ; p.NK1 = NK1; p.NK2 = NK2; p.H = H; p.B = B;
  p.q_col0 = q_col0; p.k1_col0 = k1_col0; p.v1_col0 = v1_col0; p.k2_col0 = k2_col0; p.v2_col0 = v2_col0;
  p.O = static_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.lse = lse;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.drop = make_drop_key(drop_p, drop_seed, 0);
#ifdef MEBT_ATTN_TRACE
  p.trace = g_attn_trace;
#else
  p.trace = nullptr;
#endif
  const bool drop = p.drop.thr != 0;
  auto kernel = drop ? latent_attention_fwd_kernel<true> : latent_attention_fwd_kernel<false>;
  static bool attr[2][64] = {};
  if (first_use_on_device(attr[drop])) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_TOTAL));
  }
  const int base_items = B * H * (((NQ + AT_BQ - 1) / AT_BQ + 1) / 2);
  const int nt = (NK1 + AT_BKV - 1) / AT_BKV + (NK2 + AT_BKV - 1) / AT_BKV;
  // split-KV: with fewer items than half the SMs and a long key list (small-batch sampling: B <= 2 at 8192 keys), the
  // largest divisor of the tile count that still fits one wave and leaves >= 4 tiles per CTA
  int splits = 1;
  if (!drop && lse == nullptr && workspace != nullptr && 2 * base_items <= sm_count() && nt >= 8)
    for (int c = 2; c <= AT_MAX_SPLITS; ++c)
      if (nt % c == 0 && base_items * c <= sm_count() && nt / c >= 4 &&
          workspace_bytes >= size_t(c) * B * H * NQ * (AT_HS + 2) * sizeof(float))
        splits = c;
  p.splits = splits;
  p.nts = nt / splits;
  p.part_o = static_cast<float*>(workspace);
  p.part_ml = p.part_o != nullptr ? p.part_o + size_t(splits) * B * H * NQ * AT_HS : nullptr;
  const int n_items = base_items * splits;
  dim3 grid(n_items < grid_cap() ? n_items : grid_cap());
  {
    LaunchScope ls(FAM_ATTENTION, 4.0 * double(B) * H * double(NQ) * double(NK1 + NK2) * AT_HS,
                   static_cast<cudaStream_t>(stream));
    MEBT_CUDA_OK(launch_pdl(kernel, grid, dim3(AT_THREADS), AT_SMEM_TOTAL, static_cast<cudaStream_t>(stream), tq, t1, t2, p));
  }
  MEBT_LAUNCH_OK("latent_attention_fwd_kernel");
  if (splits > 1) {
    const long long rows = (long long)B * H * NQ;
    LaunchScope ls(FAM_ATTENTION, 0.0, static_cast<cudaStream_t>(stream));
    MEBT_CUDA_OK(launch_pdl(attention_combine_kernel, dim3(int((rows + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                            static_cast<const float*>(p.part_o), static_cast<const float*>(p.part_ml), splits, B, H, NQ,
                            p.scale_log2, p.O, ldo));
    MEBT_LAUNCH_OK("attention_combine_kernel");
  }
  return MEBT_OK;
}

// The keep factors (0 or 1/(1-p)) the kernels above/below apply, materialised as fp32 [B, H, NQ, NK1 + NK2]: test
// support (the oracle replays the same mask), not used on the product path.
__global__ void attention_dropout_mask_kernel(float* __restrict__ out, int B, int H, int NQ, int NK1, int NK2, DropKey key) {
  const long long total = (long long)B * H * NQ * (NK1 + NK2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = int(i % (NK1 + NK2));
    const uint32_t row_id = uint32_t(i / (NK1 + NK2));       // (b * H + h) * NQ + q
    const uint32_t kl = kk < NK1 ? uint32_t(kk) : (1u << 20) | uint32_t(kk - NK1);
    float f0, f1;
    drop_pair(key, drop_row_key(key, row_id), kl >> 1, f0, f1);
    out[i] = key.thr == 0 ? 1.f : ((kl & 1u) ? f1 : f0);
  }
}
}  // namespace mebt

extern "C" int mebt_latent_attention_fwd(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                         int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0,
                                         int NK2, void* O, int ldo, float* lse, int B, int H, int NQ, int head_dim,
                                         void* stream) {
  return mebt::latent_attention_fwd(Q, ldq, q_col0, KV1, ld1, k1_col0, v1_col0, NK1, KV2, ld2, k2_col0, v2_col0, NK2, O, ldo,
                                    lse, B, H, NQ, head_dim, 0.f, 0ull, nullptr, 0, stream);
}

extern "C" size_t mebt_latent_attention_fwd_workspace_bytes(int B, int H, int NQ) {
  return mebt::latent_attention_fwd_workspace_bytes(B, H, NQ);
}

extern "C" int mebt_latent_attention_fwd_ws(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                            int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0,
                                            int NK2, void* O, int ldo, float* lse, int B, int H, int NQ, int head_dim,
                                            void* workspace, size_t workspace_bytes, void* stream) {
  return mebt::latent_attention_fwd(Q, ldq, q_col0, KV1, ld1, k1_col0, v1_col0, NK1, KV2, ld2, k2_col0, v2_col0, NK2, O, ldo,
                                    lse, B, H, NQ, head_dim, 0.f, 0ull, workspace, workspace_bytes, stream);
}

extern "C" int mebt_latent_attention_fwd_dropout(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                                 int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0,
                                                 int v2_col0, int NK2, void* O, int ldo, float* lse, int B, int H, int NQ,
                                                 int head_dim, float p, unsigned long long seed, void* stream) {
  return mebt::latent_attention_fwd(Q, ldq, q_col0, KV1, ld1, k1_col0, v1_col0, NK1, KV2, ld2, k2_col0, v2_col0, NK2, O, ldo,
                                    lse, B, H, NQ, head_dim, p, seed, nullptr, 0, stream);
}

extern "C" int mebt_attention_dropout_mask(float* out, int B, int H, int NQ, int NK1, int NK2, float p,
                                           unsigned long long seed, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(out != nullptr && B > 0 && H > 0 && NQ > 0 && NK1 >= 0 && NK2 >= 0 && NK1 + NK2 > 0, MEBT_ERR_SHAPE,
               "attention_dropout_mask: bad shape");
  const DropKey key = make_drop_key(p, seed, 0);
  attention_dropout_mask_kernel<<<148 * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, B, H, NQ, NK1, NK2, key);
  MEBT_LAUNCH_OK("attention_dropout_mask_kernel");
  return MEBT_OK;
}
