// K3: latent attention for the four MeBT block modes, head_dim 64, bf16 in / fp32 softmax.
//   O[b, q, h, :] = softmax_k( Q[b,q,h,:] . K[b,k,h,:] / sqrt(64) ) V[b,k,h,:]
// reference: CrossAttention.forward, mebt/modules/gpt.py:131-137 (bmm -> softmax -> bmm, materialising
// [B,h,NQ,NK] fp32), with the key sets of Block.forward (gpt.py:164-175):
//   latent_enc (256 x NC) | latent_self (256 x 256) | latent_dec (NT x 256) | lt2l (256 x (256 + NT)).
// lt2l's torch.cat([sos_emb, targets]) is never materialised: the kernel walks two K/V sources.
// NK == 0 (first draft step: no context) yields O = 0, like the empty softmax in the reference.
//
// One CTA per (128-query tile, head, batch element), 192 threads:
//   warp 0    : TMA producer (Q once, then K/V tiles of 128 keys through a 2-deep ring)
//   warp 1    : tcgen05.mma issuer: S = Q K^T (M128 N128 K64) and O_j = P_j V_j (M128 N64 K128) into TMEM
//   warps 2-5 : online softmax; one query row per thread, S read from TMEM twice (max pass, exp pass),
//               P written to shared memory in the 128B-swizzled K-major layout the PV MMA consumes,
//               O_j folded into a register accumulator with the usual running-max rescale.
#include "common.cuh"

namespace mebt {
namespace {

constexpr int AT_BQ = 128;
constexpr int AT_BKV = 128;
constexpr int AT_HS = 64;
constexpr int AT_THREADS = 192;
constexpr int AT_TILE_BYTES = 128 * 64 * 2;      // 16 KiB: a [128 x 64] bf16 tile (Q, K or V)
constexpr int AT_SMEM_Q = 0;
constexpr int AT_SMEM_K = AT_TILE_BYTES;                         // 2 stages
constexpr int AT_SMEM_V = AT_SMEM_K + 2 * AT_TILE_BYTES;         // 2 stages
constexpr int AT_SMEM_P = AT_SMEM_V + 2 * AT_TILE_BYTES;         // 32 KiB: [128 x 128] bf16
constexpr int AT_SMEM_BAR = AT_SMEM_P + 2 * AT_TILE_BYTES;
constexpr int AT_SMEM_TOTAL = AT_SMEM_BAR + 128;
constexpr uint32_t AT_TMEM_COLS = 256;           // S: [0,128)  O: [128,192)

struct AttnParams {
  int NQ, NK1, NK2, H;
  int q_col0, k1_col0, v1_col0, k2_col0, v2_col0;
  __nv_bfloat16* O;
  int ldo;
  float* lse;              // optional [B, H, NQ]
  float scale_log2;        // log2(e) / sqrt(hs)
  float scale;             // 1 / sqrt(hs)
};

__global__ void __launch_bounds__(AT_THREADS, 2)
latent_attention_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv1,
                            const __grid_constant__ CUtensorMap tm_kv2, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AT_SMEM_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tiles1 = (p.NK1 + AT_BKV - 1) / AT_BKV;
  const int tiles2 = (p.NK2 + AT_BKV - 1) / AT_BKV;
  const int nt = tiles1 + tiles2;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tm_q);
    prefetch_tensormap(&tm_kv1);
    prefetch_tensormap(&tm_kv2);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_o = tmem_base + 128;
  griddep_wait();

  if (warp == 0) {
    if (lane == 0 && nt > 0) {
      mbar_arrive_expect_tx(q_full, AT_TILE_BYTES);
      tma_load_2d(smem + AT_SMEM_Q, &tm_q, q_full, p.q_col0 + h * AT_HS, b * p.NQ + qt * AT_BQ);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
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
  } else if (warp == 1) {
    if (lane == 0 && nt > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);    // P (K-major) x V (MN-major: hs contiguous)
      const uint32_t sQ = smem_u32(smem + AT_SMEM_Q);
      const uint32_t sP = smem_u32(smem + AT_SMEM_P);
      auto issue_s = [&](int j) {
        const uint32_t sK = smem_u32(smem + AT_SMEM_K + (j & 1) * AT_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < AT_HS / 16; ++k)
          umma_bf16_ss(tmem_s, make_smem_desc_sw128(sQ + k * 32, 16, 1024), make_smem_desc_sw128(sK + k * 32, 16, 1024),
                       idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        mbar_wait(p_full, j & 1);               // softmax(j) done: S consumed, P_j in smem, O_{j-1} consumed
        tc_fence_after();
        if (j + 1 < nt) issue_s(j + 1);         // overlaps softmax(j+1) with PV_j
        const uint32_t sV = smem_u32(smem + AT_SMEM_V + (j & 1) * AT_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < AT_BKV / 16; ++kk) {
          const uint64_t da = make_smem_desc_sw128(sP + (kk >> 2) * AT_TILE_BYTES + (kk & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sV + kk * 2048, 64 * 128, 1024);
          umma_bf16_ss(tmem_o, da, db, idesc_o, kk != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[j & 1]);
        umma_commit(o_full);
      }
    }
  } else {
    // ===== softmax warps =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float o_acc[AT_HS];
#pragma unroll
    for (int i = 0; i < AT_HS; ++i) o_acc[i] = 0.f;
    uint8_t* sP = smem + AT_SMEM_P;

    for (int j = 0; j < nt; ++j) {
      const int valid = j < tiles1 ? min(AT_BKV, p.NK1 - j * AT_BKV) : min(AT_BKV, p.NK2 - (j - tiles1) * AT_BKV);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const bool full = valid == AT_BKV;            // warp-uniform: only a source's last tile can be ragged
      // One pass over the S row: p = 2^(s*c - m*c) with the CURRENT running maximum m (no separate max pass), the row
      // maximum of this tile as a by-product.  m is only advanced when a score exceeds it by more than 8 in the log2
      // domain (p would exceed 256): the exact result does not depend on the stabiliser, only overflow safety does,
      // so the usual per-tile max pass, and the rescale of O that follows every small increase of the maximum, are
      // skipped.  The first tile takes its maximum explicitly.
      auto row_pass = [&](float mb, float& l_tile, float& mx_tile) {
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_s + lane_addr + c * 32, r);
          tmem_ld_wait();
          float pv[32];
          if (full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float sv = __uint_as_float(r[i]);
              m4[i & 3] = fmaxf(m4[i & 3], sv);
              pv[i] = ex2_approx(fmaf(sv, p.scale_log2, -mb));
              l4[i & 3] += pv[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float sv = __uint_as_float(r[i]);
              const bool ok = c * 32 + i < valid;
              if (ok) m4[i & 3] = fmaxf(m4[i & 3], sv);
              pv[i] = ok ? ex2_approx(fmaf(sv, p.scale_log2, -mb)) : 0.f;
              l4[i & 3] += pv[i];
            }
          }
          // 32 keys = four 16-byte chunks of this row inside one 64-key swizzle atom
          uint8_t* base = sP + (c >> 1) * AT_TILE_BYTES + row * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(pv[8 * g + 0], pv[8 * g + 1]);
            u.y = pack_bf16x2(pv[8 * g + 2], pv[8 * g + 3]);
            u.z = pack_bf16x2(pv[8 * g + 4], pv[8 * g + 5]);
            u.w = pack_bf16x2(pv[8 * g + 6], pv[8 * g + 7]);
            const int chunk = (c & 1) * 4 + g;
            *reinterpret_cast<uint4*>(base + ((chunk ^ (row & 7)) << 4)) = u;
          }
        }
        l_tile = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        mx_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      };
      if (j == 0) {                                 // explicit maximum of the first tile
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_s + lane_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (full || c * 32 + i < valid) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(r[i]));
        }
        m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
        mbar_wait(o_full, (j - 1) & 1);             // PV_{j-1} done: O_{j-1} readable, P buffer reusable
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_o + lane_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(r[i]);
        }
      }
      float l_tile, mx_tile;
      row_pass(m * p.scale_log2, l_tile, mx_tile);
      if ((mx_tile - m) * p.scale_log2 > 8.0f) {    // rare: re-base on the new maximum and redo this row
        const float alpha = ex2_approx((m - mx_tile) * p.scale_log2);
#pragma unroll
        for (int i = 0; i < AT_HS; ++i) o_acc[i] *= alpha;
        l *= alpha;
        m = mx_tile;
        row_pass(m * p.scale_log2, l_tile, mx_tile);
      }
      l += l_tile;
      tc_fence_before();
      fence_proxy_async_smem();       // make the st.shared P tile visible to the tensor-core (async) proxy
      mbar_arrive(p_full);
    }
    if (nt > 0) {
      mbar_wait(o_full, (nt - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_o + lane_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(r[i]);
      }
    }
    const int qrow = qt * AT_BQ + row;
    if (qrow < p.NQ) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.O + (size_t(b) * p.NQ + qrow) * p.ldo + h * AT_HS);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 u;
        u.x = pack_bf16x2(o_acc[8 * g + 0] * inv, o_acc[8 * g + 1] * inv);
        u.y = pack_bf16x2(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
        u.z = pack_bf16x2(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
        u.w = pack_bf16x2(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
        dst[g] = u;
      }
      if (p.lse != nullptr)
        p.lse[(size_t(b) * p.H + h) * p.NQ + qrow] = l > 0.f ? m * p.scale + logf(l) : -INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AT_TMEM_COLS);
  }
}

}  // namespace
}  // namespace mebt

extern "C" int mebt_latent_attention_fwd(const void* Q, int ldq, int q_col0, const void* KV1, int ld1, int k1_col0,
                                         int v1_col0, int NK1, const void* KV2, int ld2, int k2_col0, int v2_col0,
                                         int NK2, void* O, int ldo, float* lse, int B, int H, int NQ, int head_dim,
                                         void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(head_dim == AT_HS, MEBT_ERR_UNSUPPORTED, "attention: head_dim %d unsupported (every MeBT config uses 64)",
               head_dim);
  MEBT_REQUIRE(B > 0 && H > 0 && NQ > 0 && NK1 >= 0 && NK2 >= 0, MEBT_ERR_SHAPE, "attention: bad shape");
  MEBT_REQUIRE(ldq % 8 == 0 && ldo % 8 == 0 && q_col0 % 8 == 0, MEBT_ERR_SHAPE, "attention: Q/O strides must be 16B aligned");
  MEBT_REQUIRE(NK1 == 0 || (KV1 != nullptr && ld1 % 8 == 0), MEBT_ERR_SHAPE, "attention: bad KV1");
  MEBT_REQUIRE(NK2 == 0 || (KV2 != nullptr && ld2 % 8 == 0), MEBT_ERR_SHAPE, "attention: bad KV2");
  CUtensorMap tq, t1, t2;
  int rc = get_tensor_map_2d(&tq, Q, 2, uint64_t(ldq), uint64_t(B) * NQ, uint64_t(ldq) * 2, 64, 128);
  if (rc) return rc;
  // an absent source still needs a valid descriptor object; alias the query map (never dereferenced: 0 tiles)
  t1 = tq;
  t2 = tq;
  if (NK1 > 0) {
    rc = get_tensor_map_2d(&t1, KV1, 2, uint64_t(ld1), uint64_t(B) * NK1, uint64_t(ld1) * 2, 64, 128);
    if (rc) return rc;
  }
  if (NK2 > 0) {
    rc = get_tensor_map_2d(&t2, KV2, 2, uint64_t(ld2), uint64_t(B) * NK2, uint64_t(ld2) * 2, 64, 128);
    if (rc) return rc;
  }
  AttnParams p;
  p.NQ = NQ; p.NK1 = NK1; p.NK2 = NK2; p.H = H;
  p.q_col0 = q_col0; p.k1_col0 = k1_col0; p.v1_col0 = v1_col0; p.k2_col0 = k2_col0; p.v2_col0 = v2_col0;
  p.O = static_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.lse = lse;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  static bool attr = false;
  if (!attr) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(latent_attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      AT_SMEM_TOTAL));
    attr = true;
  }
  dim3 grid((NQ + AT_BQ - 1) / AT_BQ, H, B);
  {
    LaunchScope ls(FAM_ATTENTION, 4.0 * double(B) * H * double(NQ) * double(NK1 + NK2) * AT_HS,
                   static_cast<cudaStream_t>(stream));
    MEBT_CUDA_OK(launch_pdl(latent_attention_fwd_kernel, grid, dim3(AT_THREADS), AT_SMEM_TOTAL,
                            static_cast<cudaStream_t>(stream), tq, t1, t2, p));
  }
  MEBT_LAUNCH_OK("latent_attention_fwd_kernel");
  return MEBT_OK;
}
