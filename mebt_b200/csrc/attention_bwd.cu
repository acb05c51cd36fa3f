// Backward of K3 (latent attention), head_dim 64 — what autograd computes for the bmm/softmax/bmm of
// CrossAttention.forward (mebt/modules/gpt.py:131-137) in the training step.
//
//   P = exp(S * scale - lse),  S = Q K^T          (recomputed from the saved log-sum-exp, nothing [NQ,NK] is stored)
//   dV = P^T dO          dP = dO V^T          dS = P .* (dP - delta) * scale,  delta = rowsum(dO .* O)
//   dQ = dS K            dK = dS^T Q
//
// ONE launch (attn_bwd_kernel) with two kinds of CTA, both tcgen05/TMEM/TMA like the forward, no atomics (bitwise
// reproducible):
//   dkv CTA : one per (128-key tile, head, batch, source); loops over query tiles and keeps the dV / dK accumulators
//             in TMEM.  P^T and dS^T are never transposed in memory: the P / dS tiles written to shared memory as
//             [q][k] are consumed as MN-major A operands.
//   dq CTA  : one per (128-query tile, head, batch); walks the key tiles of both sources and keeps dQ in TMEM.
// S and dP are recomputed in both (7 instead of 5 tile products): attention is ~5 % of the step's FLOPs.  The kinds
// used to be separate launches (dq, dkv of source 1, dkv of source 2): at the training shapes each was less than two
// waves of short CTAs, so three launch boundaries cost more than the work; now the heavier kind is scheduled first
// and the lighter one fills the tail.
#include <cstdlib>

#include "common.cuh"

namespace mebt {

int attn_delta(const void* dO, int lddo, const void* O, int ldo, float* delta, int B, int H, int NQ, cudaStream_t st);

namespace {

// warp 0: TMA producer, warp 1: MMA issuer, then NG softmax warpgroups.  The groups split the 128 keys of a tile
// (a warp may only touch the TMEM lanes of its quarter, warp % 4, so the split is over columns): with one group the
// kernel ran one warp per scheduler and exposed every tcgen05.ld / MUFU latency (10 % of the tensor pipe, round 1).
// NG (template parameter of everything below) = number of groups: 2 (320 threads, 64 keys per thread) or 4 (576 threads).
constexpr int TILE = 128 * 64 * 2;    // 16 KiB, a [128 x 64] bf16 tile
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t kHi = smem_desc_hi_sw128(1024);

struct BwdParams {
  int NQ, H;
  int NK[2];                 // keys of the two sources
  int q_col0;
  int k_col0[2], v_col0[2];
  int do_col0;
  const float* lse;          // [B,H,NQ]
  const float* delta;        // [B,H,NQ]
  __nv_bfloat16* dQ; int lddq; int dq_col0;
  __nv_bfloat16* dKV[2]; int lddkv[2]; int dk_col0[2], dv_col0[2];
  float scale, scale_log2;
  DropKey drop;              // attention-probability dropout of the forward (thr == 0: none)
  int n_dq, n_dkv0;          // CTAs (per head and batch element) of each kind: query tiles, key tiles of source 0
  int dq_first;              // blockIdx.x order: 1 = dq tiles, then dkv source 0, then source 1; 0 = dkv tiles first
};

// One thread = one query row and 4 / NG 32-key chunks (from chunk c0) of the current [128 q x 128 k] tile pair (S and dP
// in TMEM).  Computes P and dS and stores them as bf16 into the 128B-swizzled [q][k] shared-memory tiles.  Arithmetic
// in f32x2 pairs; FULL = every row and key of the tile is valid; `reuse_bar` (if any) is waited on just before the first
// store: the previous tile's P / dS are still being read by its dV / dK (dQ) products while this tile's values are formed.
//   p  = exp2(s * scale_log2 - lse_l2) [* f]          f = dropout keep factor (0 or 1/(1-p)) of the forward
//   dS = p0 * (dP * f - delta) * scale                 p0 = the undropped probability
template <int NG, bool FULL, bool DROP, bool WANT_P>
__device__ __forceinline__ void softmax_bwd_row(uint32_t tmem_s, uint32_t tmem_dp, uint32_t lane_addr, int row, int c0,
                                                bool row_ok, int valid_keys, float nlse_l2, float ndelta_s, float scale,
                                                float scale_log2, uint8_t* sP, uint8_t* sdS, const DropKey& dk,
                                                uint32_t row_key, uint32_t pair0, uint64_t* reuse_bar, uint32_t reuse_parity) {
  const float fs = DROP ? dk.inv_keep * scale : scale;
  const uint32_t thr16 = dk.thr << 16;
#pragma unroll 1
  for (int cc = 0; cc < 4 / NG; ++cc) {
    const int c = c0 + cc;
    uint32_t rs[32], rd[32];
    tmem_ld_32x32(tmem_s + lane_addr + c * 32, rs);
    tmem_ld_32x32(tmem_dp + lane_addr + c * 32, rd);
    tmem_ld_wait();
    uint32_t pk[16], dsk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 x = __ffma2_rn(make_float2(__uint_as_float(rs[2 * i]), __uint_as_float(rs[2 * i + 1])), splat2(scale_log2),
                                  splat2(nlse_l2));
      float2 pv = make_float2(ex2_approx(x.x), ex2_approx(x.y));
      if (!FULL) {
        pv.x = (row_ok && c * 32 + 2 * i < valid_keys) ? pv.x : 0.f;
        pv.y = (row_ok && c * 32 + 2 * i + 1 < valid_keys) ? pv.y : 0.f;
      }
      float2 dp = make_float2(__uint_as_float(rd[2 * i]), __uint_as_float(rd[2 * i + 1]));
      float2 pf = pv;
      if (DROP) {      // the two 16-bit halves of one hash decide elements 2 * pair and 2 * pair + 1 (drop_pair)
        const uint32_t h = mix32(row_key + (pair0 + uint32_t(c * 16 + i)) * 0x9E3779B9u);
        const bool k0 = (h << 16) >= thr16, k1 = h >= thr16;
        dp.x = k0 ? dp.x : 0.f;
        dp.y = k1 ? dp.y : 0.f;
        if (WANT_P) {
          pf = __fmul2_rn(pv, splat2(dk.inv_keep));
          pf.x = k0 ? pf.x : 0.f;
          pf.y = k1 ? pf.y : 0.f;
        }
      }
      const float2 ds = __fmul2_rn(pv, __ffma2_rn(dp, splat2(fs), splat2(ndelta_s)));
      if (WANT_P) pk[i] = pack_bf16x2(pf.x, pf.y);
      dsk[i] = pack_bf16x2(ds.x, ds.y);
    }
    if (cc == 0 && reuse_bar != nullptr) mbar_wait(reuse_bar, reuse_parity);
    const int off = (c >> 1) * TILE + row * 128;     // half (64 keys) = one 16 KiB swizzle-atom column
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int chunk = (c & 1) * 4 + g;
      const int sw = (chunk ^ (row & 7)) << 4;
      if (WANT_P) *reinterpret_cast<uint4*>(sP + off + sw) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      *reinterpret_cast<uint4*>(sdS + off + sw) = make_uint4(dsk[4 * g], dsk[4 * g + 1], dsk[4 * g + 2], dsk[4 * g + 3]);
    }
  }
}

template <int NG, bool WANT_P>
__device__ __forceinline__ void softmax_bwd_dispatch(bool full, uint32_t tmem_s, uint32_t tmem_dp, uint32_t lane_addr, int row,
                                                     int c0, bool row_ok, int valid_keys, float nlse_l2, float ndelta_s,
                                                     float scale, float scale_log2, uint8_t* sP, uint8_t* sdS,
                                                     const DropKey& dk, uint32_t row_key, uint32_t pair0, uint64_t* reuse_bar,
                                                     uint32_t reuse_parity) {
#define MEBT_SM_ARGS tmem_s, tmem_dp, lane_addr, row, c0, row_ok, valid_keys, nlse_l2, ndelta_s, scale, scale_log2, sP, sdS, \
                     dk, row_key, pair0, reuse_bar, reuse_parity
  if (dk.thr != 0) {
    if (full) softmax_bwd_row<NG, true, true, WANT_P>(MEBT_SM_ARGS); else softmax_bwd_row<NG, false, true, WANT_P>(MEBT_SM_ARGS);
  } else {
    if (full) softmax_bwd_row<NG, true, false, WANT_P>(MEBT_SM_ARGS); else softmax_bwd_row<NG, false, false, WANT_P>(MEBT_SM_ARGS);
  }
#undef MEBT_SM_ARGS
}

// store 32 columns (from column c * 32) of a [128 rows x 64] fp32 TMEM tile (lanes = rows) as bf16 into global memory
__device__ __forceinline__ void store_tmem_chunk(uint32_t tmem_addr, uint32_t lane_addr, int c, __nv_bfloat16* dst, bool ok) {
  uint32_t r[32];
  tmem_ld_32x32(tmem_addr + lane_addr + c * 32, r);
  tmem_ld_wait();
  if (ok) {
    uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 u;
      u.x = pack_bf16x2(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
      u.y = pack_bf16x2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
      u.z = pack_bf16x2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
      u.w = pack_bf16x2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
      d4[g] = u;
    }
  }
}

// ================================================================================================
// dK / dV
// ================================================================================================
constexpr int DKV_SMEM_K = 0, DKV_SMEM_V = TILE, DKV_SMEM_Q = 2 * TILE /* 2 stages */, DKV_SMEM_DO = 4 * TILE /* 2 stages */,
              DKV_SMEM_P = 6 * TILE, DKV_SMEM_DS = 8 * TILE, DKV_SMEM_BAR = 10 * TILE, DKV_SMEM_TOTAL = DKV_SMEM_BAR + 128;

template <int NG>
__device__ __forceinline__ void attn_bwd_dkv_body(uint8_t* smem, const CUtensorMap& tm_q, const CUtensorMap& tm_do,
                                                  const CUtensorMap& tm_kv, const BwdParams& p, const int src,
                                                  const int jt, const int h, const int b) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_SMEM_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;    // [2]
  uint64_t* q_empty = bars + 3;   // [2]
  uint64_t* sp_full = bars + 5;
  uint64_t* pds_full = bars + 6;
  uint64_t* dkv_done = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (p.NQ + 127) / 128;
  const int NK = p.NK[src];
  const int valid_keys = min(128, NK - jt * 128);

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm_q); prefetch_tensormap(&tm_do); prefetch_tensormap(&tm_kv);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    mbar_init(sp_full, 1); mbar_init(pds_full, 128 * NG); mbar_init(dkv_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 320;
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      const int krow = b * NK + jt * 128;
      mbar_arrive_expect_tx(kv_full, 2 * TILE);
      tma_load_2d(smem + DKV_SMEM_K, &tm_kv, kv_full, p.k_col0[src] + h * 64, krow);
      tma_load_2d(smem + DKV_SMEM_V, &tm_kv, kv_full, p.v_col0[src] + h * 64, krow);
      for (int i = 0; i < nq; ++i) {
        const int s = i & 1;
        mbar_wait(&q_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], 2 * TILE);
        const int qrow = b * p.NQ + i * 128;
        tma_load_2d(smem + DKV_SMEM_Q + s * TILE, &tm_q, &q_full[s], p.q_col0 + h * 64, qrow);
        tma_load_2d(smem + DKV_SMEM_DO + s * TILE, &tm_do, &q_full[s], p.do_col0 + h * 64, qrow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);   // [q x d] . [k x d]^T
      constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);     // (P|dS)^T as MN-major A, (dO|Q) as MN-major B
      const uint32_t sK = smem_u32(smem + DKV_SMEM_K), sV = smem_u32(smem + DKV_SMEM_V);
      const uint32_t sP = smem_u32(smem + DKV_SMEM_P), sdS = smem_u32(smem + DKV_SMEM_DS);
      auto issue_s_dp = [&](int i) {
        const uint32_t sQ = smem_u32(smem + DKV_SMEM_Q + (i & 1) * TILE), sdO = smem_u32(smem + DKV_SMEM_DO + (i & 1) * TILE);
        // four instructions per asm block over split descriptor words (tools/mma_probe.cu: 48-64 clk per instruction
        // against ~97 when each rebuilds its descriptors)
        umma_bf16_ss_x4<false>(tmem_s, smem_desc_lo(sQ, 16), smem_desc_lo(sK, 16), 2, 2, kHi, kHi, idesc_qk, 0u);
        umma_bf16_ss_x4<false>(tmem_dp, smem_desc_lo(sdO, 16), smem_desc_lo(sV, 16), 2, 2, kHi, kHi, idesc_qk, 0u);
        umma_commit(sp_full);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_s_dp(0);
      for (int i = 0; i < nq; ++i) {
        if (i + 1 < nq) mbar_wait(&q_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
        if (i + 1 < nq) issue_s_dp(i + 1);
        const uint32_t sQ = smem_u32(smem + DKV_SMEM_Q + (i & 1) * TILE), sdO = smem_u32(smem + DKV_SMEM_DO + (i & 1) * TILE);
        // reduction over the 128 queries of this tile, 16 (= 2048 B of every operand) at a time.  A: the [q rows][k cols]
        // tile read as MN-major (M = keys): two 64-key atoms 16 KiB apart, 8 q-rows per 1 KiB
        {
          const uint32_t st = 2048u >> 4;
          const uint32_t p_lo = smem_desc_lo(sP, TILE), ds_lo = smem_desc_lo(sdS, TILE);
          const uint32_t do_lo = smem_desc_lo(sdO, TILE), q_lo = smem_desc_lo(sQ, TILE);
          umma_bf16_ss_x4<false>(tmem_dv, p_lo, do_lo, st, st, kHi, kHi, idesc_t, i != 0 ? 1u : 0u);
          umma_bf16_ss_x4<false>(tmem_dv, p_lo + 4 * st, do_lo + 4 * st, st, st, kHi, kHi, idesc_t, 1u);
          umma_bf16_ss_x4<false>(tmem_dk, ds_lo, q_lo, st, st, kHi, kHi, idesc_t, i != 0 ? 1u : 0u);
          umma_bf16_ss_x4<false>(tmem_dk, ds_lo + 4 * st, q_lo + 4 * st, st, st, kHi, kHi, idesc_t, 1u);
        }
        umma_commit(&q_empty[i & 1]);
        umma_commit(dkv_done);
      }
    }
  } else {
    const int q4 = warp & 3, grp = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_addr = uint32_t(q4 * 32) << 16;
    for (int i = 0; i < nq; ++i) {
      const int qrow = i * 128 + row;
      const bool row_ok = qrow < p.NQ;
      float nlse_l2 = 0.f, ndelta_s = 0.f;
      if (row_ok) {
        const size_t o = (size_t(b) * p.H + h) * p.NQ + qrow;
        nlse_l2 = -p.lse[o] * LOG2E;
        ndelta_s = -p.delta[o] * p.scale;
      }
      mbar_wait(sp_full, i & 1);
      tc_fence_after();
      // the previous tile's P / dS must no longer be read when this tile's are stored (waited on inside, before the stores)
      softmax_bwd_dispatch<NG, true>(valid_keys == 128 && (i + 1) * 128 <= p.NQ, tmem_s, tmem_dp, lane_addr, row,
                                 grp * (4 / NG), row_ok, valid_keys, nlse_l2, ndelta_s, p.scale, p.scale_log2,
                                 smem + DKV_SMEM_P, smem + DKV_SMEM_DS, p.drop,
                                 drop_row_key(p.drop, uint32_t((b * p.H + h) * p.NQ + qrow)),
                                 (uint32_t(src) << 19) | uint32_t(jt * 64), i > 0 ? dkv_done : nullptr, (i - 1) & 1);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(pds_full);
    }
    mbar_wait(dkv_done, (nq - 1) & 1);
    tc_fence_after();
    const bool ok = row < valid_keys;            // TMEM lanes are key rows here
    __nv_bfloat16* base = p.dKV[src] + (size_t(b) * NK + jt * 128 + row) * p.lddkv[src] + h * 64;
    // the four 32-column chunks of [dV | dK] are split over the groups
    for (int cc = 0; cc < 4 / NG; ++cc) {
      const int c = grp * (4 / NG) + cc;
      if (c < 2) store_tmem_chunk(tmem_dv, lane_addr, c, base + p.dv_col0[src], ok);
      else store_tmem_chunk(tmem_dk, lane_addr, c - 2, base + p.dk_col0[src], ok);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ================================================================================================
// dQ
// ================================================================================================
constexpr int DQ_SMEM_Q = 0, DQ_SMEM_DO = TILE, DQ_SMEM_K = 2 * TILE /* 2 stages */, DQ_SMEM_V = 4 * TILE /* 2 stages */,
              DQ_SMEM_DS = 6 * TILE, DQ_SMEM_BAR = 8 * TILE, DQ_SMEM_TOTAL = DQ_SMEM_BAR + 128;

template <int NG>
__device__ __forceinline__ void attn_bwd_dq_body(uint8_t* smem, const CUtensorMap& tm_q, const CUtensorMap& tm_do,
                                                 const CUtensorMap& tm_kv1, const CUtensorMap& tm_kv2,
                                                 const BwdParams& p, const int qt, const int h, const int b) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_SMEM_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* sp_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* dq_done = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NK1 = p.NK[0], NK2 = p.NK[1];
  const int tiles1 = (NK1 + 127) / 128, tiles2 = (NK2 + 127) / 128;
  const int nt = tiles1 + tiles2;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm_q); prefetch_tensormap(&tm_do); prefetch_tensormap(&tm_kv1); prefetch_tensormap(&tm_kv2);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(sp_full, 1); mbar_init(ds_full, 128 * NG); mbar_init(dq_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dq = tmem_base + 256;
  griddep_wait();

  if (warp == 0) {
    if (lane == 0 && nt > 0) {
      const int qrow = b * p.NQ + qt * 128;
      mbar_arrive_expect_tx(q_full, 2 * TILE);
      tma_load_2d(smem + DQ_SMEM_Q, &tm_q, q_full, p.q_col0 + h * 64, qrow);
      tma_load_2d(smem + DQ_SMEM_DO, &tm_do, q_full, p.do_col0 + h * 64, qrow);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE);
        if (j < tiles1) {
          const int r = b * NK1 + j * 128;
          tma_load_2d(smem + DQ_SMEM_K + s * TILE, &tm_kv1, &kv_full[s], p.k_col0[0] + h * 64, r);
          tma_load_2d(smem + DQ_SMEM_V + s * TILE, &tm_kv1, &kv_full[s], p.v_col0[0] + h * 64, r);
        } else {
          const int r = b * NK2 + (j - tiles1) * 128;
          tma_load_2d(smem + DQ_SMEM_K + s * TILE, &tm_kv2, &kv_full[s], p.k_col0[1] + h * 64, r);
          tma_load_2d(smem + DQ_SMEM_V + s * TILE, &tm_kv2, &kv_full[s], p.v_col0[1] + h * 64, r);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, 64, 0, 1);    // dS (K-major) x K (MN-major: d contiguous)
      const uint32_t sQ = smem_u32(smem + DQ_SMEM_Q), sdO = smem_u32(smem + DQ_SMEM_DO), sdS = smem_u32(smem + DQ_SMEM_DS);
      auto issue_s_dp = [&](int j) {
        const uint32_t sK = smem_u32(smem + DQ_SMEM_K + (j & 1) * TILE), sV = smem_u32(smem + DQ_SMEM_V + (j & 1) * TILE);
        umma_bf16_ss_x4<false>(tmem_s, smem_desc_lo(sQ, 16), smem_desc_lo(sK, 16), 2, 2, kHi, kHi, idesc_qk, 0u);
        umma_bf16_ss_x4<false>(tmem_dp, smem_desc_lo(sdO, 16), smem_desc_lo(sV, 16), 2, 2, kHi, kHi, idesc_qk, 0u);
        umma_commit(sp_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s_dp(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
        if (j + 1 < nt) issue_s_dp(j + 1);
        const uint32_t sK = smem_u32(smem + DQ_SMEM_K + (j & 1) * TILE);
        // dQ += dS K over the 128 keys of the tile: dS K-major (two 64-key halves 16 KiB apart), K MN-major
        umma_bf16_ss_x4<false>(tmem_dq, smem_desc_lo(sdS, 16), smem_desc_lo(sK, TILE), 2, 2048u >> 4, kHi, kHi, idesc_dq,
                               j != 0 ? 1u : 0u);
        umma_bf16_ss_x4<false>(tmem_dq, smem_desc_lo(sdS + TILE, 16), smem_desc_lo(sK, TILE) + 4 * (2048u >> 4), 2, 2048u >> 4,
                               kHi, kHi, idesc_dq, 1u);
        umma_commit(&kv_empty[j & 1]);
        umma_commit(dq_done);
      }
    }
  } else {
    const int q4 = warp & 3, grp = (warp - 2) >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_addr = uint32_t(q4 * 32) << 16;
    const int qrow = qt * 128 + row;
    const bool row_ok = qrow < p.NQ;
    float nlse_l2 = 0.f, ndelta_s = 0.f;
    if (row_ok) {
      const size_t o = (size_t(b) * p.H + h) * p.NQ + qrow;
      nlse_l2 = -p.lse[o] * LOG2E;
      ndelta_s = -p.delta[o] * p.scale;
    }
    const uint32_t row_key = drop_row_key(p.drop, uint32_t((b * p.H + h) * p.NQ + qrow));
    for (int j = 0; j < nt; ++j) {
      const int valid = j < tiles1 ? min(128, NK1 - j * 128) : min(128, NK2 - (j - tiles1) * 128);
      mbar_wait(sp_full, j & 1);
      tc_fence_after();
      softmax_bwd_dispatch<NG, false>(valid == 128 && (qt + 1) * 128 <= p.NQ, tmem_s, tmem_dp, lane_addr, row, grp * (4 / NG),
                                  row_ok, valid, nlse_l2, ndelta_s, p.scale, p.scale_log2, nullptr, smem + DQ_SMEM_DS, p.drop,
                                  row_key, j < tiles1 ? uint32_t(j * 64) : (1u << 19) | uint32_t((j - tiles1) * 64),
                                  j > 0 ? dq_done : nullptr, (j - 1) & 1);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(ds_full);
    }
    __nv_bfloat16* dst = p.dQ + (size_t(b) * p.NQ + qrow) * p.lddq + p.dq_col0 + h * 64;
    // the two 32-column chunks of dQ go to the first two of (up to four) groups' chunk slots
    constexpr int DQ_PER_GROUP = NG == 1 ? 2 : 1;
    if (nt > 0) {
      mbar_wait(dq_done, (nt - 1) & 1);
      tc_fence_after();
      for (int cc = 0; cc < DQ_PER_GROUP; ++cc) {
        const int c = grp * DQ_PER_GROUP + cc;
        if (c < 2) store_tmem_chunk(tmem_dq, lane_addr, c, dst, row_ok);
      }
    } else if (row_ok && grp == 0) {
      uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
      for (int g = 0; g < 8; ++g) d4[g] = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

constexpr int AB_SMEM_TOTAL = DKV_SMEM_TOTAL > DQ_SMEM_TOTAL ? DKV_SMEM_TOTAL : DQ_SMEM_TOTAL;

// grid (n_dq + n_dkv0 + n_dkv1, H, B): the kind of a CTA is a function of blockIdx.x alone (block-uniform branch)
template <int NG>
__global__ void __launch_bounds__(64 + 128 * NG, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_kv1, const __grid_constant__ CUtensorMap tm_kv2, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int h = blockIdx.y, b = blockIdx.z;
  const int n_dkv = int(gridDim.x) - p.n_dq;
  int x = blockIdx.x;
  bool is_dq;
  if (p.dq_first) { is_dq = x < p.n_dq; if (!is_dq) x -= p.n_dq; }
  else { is_dq = x >= n_dkv; if (is_dq) x -= n_dkv; }
  if (is_dq) {
    attn_bwd_dq_body<NG>(smem, tm_q, tm_do, tm_kv1, tm_kv2, p, x, h, b);
  } else if (x < p.n_dkv0) {
    attn_bwd_dkv_body<NG>(smem, tm_q, tm_do, tm_kv1, p, 0, x, h, b);
  } else {
    attn_bwd_dkv_body<NG>(smem, tm_q, tm_do, tm_kv2, p, 1, x - p.n_dkv0, h, b);
  }
}

}  // namespace
}  // namespace mebt

namespace mebt {
// delta_ready: `workspace` already holds delta[b,h,q] = rowsum(dO .* O) (written by the epilogue of the GEMM that
// produced dO, csrc/gemm.cu in_kind 3); otherwise the preprocess kernel computes it here.
int latent_attention_bwd_launch(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2,
                                const void* O, int ldo, const void* dO, int lddo, const float* lse, void* dQ,
                                int lddq, int dq_col0, void* dKV1, int ldd1, int dk1_col0, int dv1_col0, void* dKV2,
                                int ldd2, int dk2_col0, int dv2_col0, int B, int H, int NQ, int head_dim, float drop_p,
                                unsigned long long drop_seed, void* workspace, size_t workspace_bytes, int delta_ready,
                                void* stream) {
  MEBT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MEBT_ERR_SHAPE, "attention_bwd: dropout p = %f outside [0, 1)", drop_p);
  MEBT_REQUIRE(head_dim == 64, MEBT_ERR_UNSUPPORTED, "attention_bwd: head_dim %d unsupported", head_dim);
  MEBT_REQUIRE(B > 0 && H > 0 && NQ > 0 && NK1 >= 0 && NK2 >= 0, MEBT_ERR_SHAPE, "attention_bwd: bad shape");
  MEBT_REQUIRE(workspace != nullptr && workspace_bytes >= size_t(B) * H * NQ * 4, MEBT_ERR_WORKSPACE,
               "attention_bwd: workspace too small");
  MEBT_REQUIRE(ldq % 8 == 0 && lddo % 8 == 0 && ldo % 8 == 0 && lddq % 8 == 0, MEBT_ERR_SHAPE, "attention_bwd: strides");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* delta = static_cast<float*>(workspace);
  int rc = MEBT_OK;
  if (!delta_ready) {
    rc = attn_delta(dO, lddo, O, ldo, delta, B, H, NQ, st);
    if (rc) return rc;
  }
  CUtensorMap tq, tdo, t1, t2;
  rc = get_tensor_map_2d(&tq, Q, 2, uint64_t(ldq), uint64_t(B) * NQ, uint64_t(ldq) * 2, 64, 128);
  if (rc) return rc;
  rc = get_tensor_map_2d(&tdo, dO, 2, uint64_t(lddo), uint64_t(B) * NQ, uint64_t(lddo) * 2, 64, 128);
  if (rc) return rc;
  t1 = tq; t2 = tq;
  if (NK1 > 0) { rc = get_tensor_map_2d(&t1, KV1, 2, uint64_t(ld1), uint64_t(B) * NK1, uint64_t(ld1) * 2, 64, 128); if (rc) return rc; }
  if (NK2 > 0) { rc = get_tensor_map_2d(&t2, KV2, 2, uint64_t(ld2), uint64_t(B) * NK2, uint64_t(ld2) * 2, 64, 128); if (rc) return rc; }
  static const int groups = [] { const char* e = getenv("MEBT_ATTN_BWD_GROUPS"); return e != nullptr && atoi(e) == 4 ? 4 : 2; }();
  static bool attr[64] = {};
  if (first_use_on_device(attr)) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_TOTAL));
    MEBT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM_TOTAL));
  }
  BwdParams p;
  p.NQ = NQ; p.H = H; p.NK[0] = NK1; p.NK[1] = NK2;
  p.q_col0 = q_col0; p.do_col0 = 0;
  p.k_col0[0] = k1_col0; p.v_col0[0] = v1_col0; p.k_col0[1] = k2_col0; p.v_col0[1] = v2_col0;
  p.lse = lse; p.delta = delta;
  p.dQ = static_cast<__nv_bfloat16*>(dQ); p.lddq = lddq; p.dq_col0 = dq_col0;
  p.dKV[0] = static_cast<__nv_bfloat16*>(dKV1); p.lddkv[0] = ldd1; p.dk_col0[0] = dk1_col0; p.dv_col0[0] = dv1_col0;
  p.dKV[1] = static_cast<__nv_bfloat16*>(dKV2); p.lddkv[1] = ldd2; p.dk_col0[1] = dk2_col0; p.dv_col0[1] = dv2_col0;
  p.scale = 0.125f; p.scale_log2 = 0.125f * LOG2E;
  p.drop = make_drop_key(drop_p, drop_seed, 0);
  for (int src = 0; src < 2; ++src)
    MEBT_REQUIRE(p.NK[src] == 0 || (p.dKV[src] != nullptr && p.lddkv[src] % 8 == 0), MEBT_ERR_SHAPE,
                 "attention_bwd: bad dKV%d", src + 1);
  const int nqt = (NQ + 127) / 128, nkt0 = (NK1 + 127) / 128, nkt1 = (NK2 + 127) / 128;
  p.n_dq = nqt; p.n_dkv0 = nkt0;
  // per-CTA work: a dq CTA runs 3 tile products per key tile, a dkv CTA 4 per query tile; the longer kind goes first
  p.dq_first = 3 * (nkt0 + nkt1) >= 4 * nqt ? 1 : 0;
  const double flops_tile = 2.0 * 128 * 128 * 64;
  {
    LaunchScope ls(FAM_ATTENTION, 7.0 * flops_tile * double(B) * H * nqt * (nkt0 + nkt1), st);
    const dim3 grid(nqt + nkt0 + nkt1, H, B);
    if (groups == 4) MEBT_CUDA_OK(launch_pdl(attn_bwd_kernel<4>, grid, dim3(64 + 128 * 4), AB_SMEM_TOTAL, st, tq, tdo, t1, t2, p));
    else MEBT_CUDA_OK(launch_pdl(attn_bwd_kernel<2>, grid, dim3(64 + 128 * 2), AB_SMEM_TOTAL, st, tq, tdo, t1, t2, p));
  }
  MEBT_LAUNCH_OK("attn_bwd_kernel");
  return MEBT_OK;
}
}  // namespace mebt

extern "C" {

size_t mebt_latent_attention_bwd_workspace_bytes(int B, int H, int NQ) { return size_t(B) * H * NQ * sizeof(float); }

int mebt_latent_attention_bwd(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0, int v1_col0,
                              int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2, const void* O,
                              int ldo, const void* dO, int lddo, const float* lse, void* dQ, int lddq, int dq_col0,
                              void* dKV1, int ldd1, int dk1_col0, int dv1_col0, void* dKV2, int ldd2, int dk2_col0,
                              int dv2_col0, int B, int H, int NQ, int head_dim, void* workspace, size_t workspace_bytes,
                              void* stream) {
  return mebt_latent_attention_bwd_dropout(Q, ldq, q_col0, KV1, ld1, k1_col0, v1_col0, NK1, KV2, ld2, k2_col0, v2_col0, NK2,
                                           O, ldo, dO, lddo, lse, dQ, lddq, dq_col0, dKV1, ldd1, dk1_col0, dv1_col0, dKV2,
                                           ldd2, dk2_col0, dv2_col0, B, H, NQ, head_dim, 0.f, 0ull, workspace,
                                           workspace_bytes, stream);
}

int mebt_latent_attention_bwd_dropout(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                      int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0, int NK2,
                                      const void* O, int ldo, const void* dO, int lddo, const float* lse, void* dQ,
                                      int lddq, int dq_col0, void* dKV1, int ldd1, int dk1_col0, int dv1_col0, void* dKV2,
                                      int ldd2, int dk2_col0, int dv2_col0, int B, int H, int NQ, int head_dim, float drop_p,
                                      unsigned long long drop_seed, void* workspace, size_t workspace_bytes, void* stream) {
  return mebt::latent_attention_bwd_launch(Q, ldq, q_col0, KV1, ld1, k1_col0, v1_col0, NK1, KV2, ld2, k2_col0, v2_col0, NK2, O,
                                           ldo, dO, lddo, lse, dQ, lddq, dq_col0, dKV1, ldd1, dk1_col0, dv1_col0, dKV2, ldd2,
                                           dk2_col0, dv2_col0, B, H, NQ, head_dim, drop_p, drop_seed, workspace,
                                           workspace_bytes, 0, stream);
}

}  // extern "C"
