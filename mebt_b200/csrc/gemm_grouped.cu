// Grouped weight-gradient GEMM: up to 6 independent problems  dW_g[N_out, K_in] (+)= dY_g^T X_g  in ONE persistent
// tcgen05 launch (fp32 output, both operands MN-major: dY stored [rows, N_out], X stored [rows, K_in]).
//
// Replaces the per-Linear wgrad launches that torch autograd issues for the nn.Linear layers of one Block
// (reference: mebt/modules/gpt.py:126-128 q/k/v, :140 proj, :150-155 mlp; backward of Block.forward :159-195).
// At the 16-frame training shapes (rows = 1536 or 3072) each weight gradient alone is 32-128 output tiles with a
// reduction of only 24-48 k-blocks: five such launches per block were five partial waves plus five launch / pipeline
// fill / drain latencies, and while one of them held every SM (one 200 KiB CTA per SM) the data-gradient chain on the
// main stream could not be scheduled.  As one launch the block's 250-450 tiles run as ~2-3 full waves.
//
// Structure = csrc/gemm.cu without its options (192 threads: TMA producer warp, single-thread MMA issuer, 4 epilogue
// warps; 3-stage 128B-swizzled operand ring; two TMEM accumulators; epilogue through a 4-slot staging ring and TMA
// store / reduce-add).  Tiles are numbered problem after problem; inside a problem in bands of 16 row blocks.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mebt {

namespace {

constexpr int GG_MAX = 6;
constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 3;
constexpr int GG_THREADS = 192;
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int B_TILE_BYTES = BN * BK * 2;
constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
constexpr int EPI_SLOTS = 4;
constexpr int EPI_SLOT_BYTES = 128 * 128;
constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
constexpr int BAR_OFFSET = EPI_OFFSET + EPI_SLOTS * EPI_SLOT_BYTES;
constexpr int SMEM_TOTAL = BAR_OFFSET + 256 + 1024;
constexpr int RASTER_M = 16;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");

struct GroupProblem {
  int M, N, K;                               // output [M = N_out, N = K_in], reduction K = rows
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int kb_split;                              // k-blocks [kb_split, num_k_blocks) come from the second operand pair (a2, b2)
  int tile_begin;                            // first global tile number of this problem
  int accumulate;                            // C += result (TMA reduce-add) instead of C = result
  long long flat_off;                        // ADAM: element offset of this weight in the flat parameter buffers
  int ldw;                                   // ADAM: row stride of the weight (elements)
};
// ADAM: the optimizer step fused into the epilogue (mebt_stack_backward_fused): flat fp32 masters / moments, the bf16
// operand copy, the decay-flag table of mebt_adamw_flat and the step's scalars
struct AdamFuse {
  float* p; float* m; float* v; __nv_bfloat16* p16;
  const unsigned char* decay; int shift;
  AdamScalars a;
};
struct GroupParams {
  int n_problems, total_tiles;
  GroupProblem prob[GG_MAX];
};
struct GroupMaps {
  CUtensorMap a[GG_MAX], b[GG_MAX], c[GG_MAX];
  CUtensorMap a2[GG_MAX], b2[GG_MAX];        // second reduction segment: dW = dY^T X + dY2^T X2 (lt2l key|value rows)
};

__device__ __forceinline__ void locate(const GroupParams& p, int work, int& g, int& m0, int& n0) {
  g = 0;
#pragma unroll
  for (int i = 1; i < GG_MAX; ++i)
    if (i < p.n_problems && work >= p.prob[i].tile_begin) g = i;
  const GroupProblem& q = p.prob[g];
  const int t = work - q.tile_begin;
  const int band_tiles = RASTER_M * q.num_n_blocks;
  const int band = t / band_tiles;
  const int in_band = t - band * band_tiles;
  const int band_rows = min(RASTER_M, q.num_m_blocks - band * RASTER_M);
  const int ni = in_band / band_rows;
  const int mi = band * RASTER_M + (in_band - ni * band_rows);
  m0 = mi * BM;
  n0 = ni * BN;
}

template <bool ADAM>
__global__ void __launch_bounds__(GG_THREADS, 1)
gemm_grouped_wgrad_kernel(const __grid_constant__ GroupMaps maps, const GroupParams p, const AdamFuse opt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 2 * BN;

  if (threadIdx.x == 0) {
    for (int g = 0; g < p.n_problems; ++g) {
      prefetch_tensormap(&maps.a[g]);
      prefetch_tensormap(&maps.b[g]);
      prefetch_tensormap(&maps.c[g]);
      if (p.prob[g].kb_split < p.prob[g].num_k_blocks) { prefetch_tensormap(&maps.a2[g]); prefetch_tensormap(&maps.b2[g]); }
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
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
  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = work0; work < p.total_tiles; work += work_stride) {
        int g, m0, n0;
        locate(p, work, g, m0, n0);
        const int nkb = p.prob[g].num_k_blocks, split = p.prob[g].kb_split;
        for (int kb = 0; kb < nkb; ++kb) {
          const bool seg2 = kb >= split;
          const CUtensorMap* ta = seg2 ? &maps.a2[g] : &maps.a[g];
          const CUtensorMap* tb = seg2 ? &maps.b2[g] : &maps.b[g];
          const int k0 = (seg2 ? kb - split : kb) * BK;       // rows past a segment's end are zero-filled by TMA
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * STAGE_BYTES;
          uint8_t* sB = sA + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)                                      // box [64 m][64 k]
            tma_load_2d(sA + i * (BK * 128), ta, &full_bar[stage], m0 + i * 64, k0);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)                                      // box [64 n][64 k]
            tma_load_2d(sB + i * (BK * 128), tb, &full_bar[stage], n0 + i * 64, k0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int work = work0; work < p.total_tiles; work += work_stride, ++it) {
        int g, m0, n0;
        locate(p, work, g, m0, n0);
        const int nkb = p.prob[g].num_k_blocks;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);     // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // the four K = 16 steps of the k-block as one instruction sequence over split descriptor words (csrc/gemm.cu).
          // MN-major: advance 16 k-rows (2048 B) per step; LBO = next 64-wide MN atom (BK rows * 128 B), SBO = 8 k-rows
          const uint32_t a_lo = smem_desc_lo(smem_u32(smem + stage * STAGE_BYTES), BK * 128);
          umma_bf16_ss_x4<false>(tmem_d, a_lo, a_lo + uint32_t(A_TILE_BYTES >> 4), (UMMA_K * 128) >> 4, (UMMA_K * 128) >> 4,
                                 smem_desc_hi_sw128(1024), smem_desc_hi_sw128(1024), idesc, kb > 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == nkb - 1) umma_commit(&tmem_full_bar[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ================= epilogue (warps 2..5): TMEM -> registers -> swizzled staging slot -> TMA store / reduce-add
    const int q = warp & 3;                                   // TMEM lane quarter this warp may touch
    uint8_t* slots = smem + EPI_OFFSET;
    const bool epi_t0 = threadIdx.x == 64;
    const int sw = lane & 7;
    int it = 0;
    for (int work = work0; work < p.total_tiles; work += work_stride, ++it) {
      int g, m0, n0;
      locate(p, work, g, m0, n0);
      const int accumulate = p.prob[g].accumulate;
      const CUtensorMap* tc = &maps.c[g];
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if constexpr (ADAM) {
        // The optimizer state of a tile is 3 x 128 KiB that nothing has touched since the last step: request it into L2
        // one tile ahead (this thread's row: 1 KiB of p, m and v each), so that the HBM reads run under the k-loops and
        // the epilogue's loads below are L2 hits.
        auto prefetch_tile = [&](int w) {
          int g2, m2, n2;
          locate(p, w, g2, m2, n2);
          const GroupProblem& gq = p.prob[g2];
          const int r2 = m2 + q * 32 + lane;
          if (r2 < gq.M) {
            const long long e2 = gq.flat_off + (long long)r2 * gq.ldw + n2;
            l2_prefetch_bulk(opt.p + e2, BN * 4);
            l2_prefetch_bulk(opt.m + e2, BN * 4);
            l2_prefetch_bulk(opt.v + e2, BN * 4);
          }
        };
        if (it == 0) prefetch_tile(work);
        if (work + work_stride < p.total_tiles) prefetch_tile(work + work_stride);
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN);
      auto staged_unit = [&](const uint32_t (&r)[32], int u) {
        const int gu = it * (BN / 32) + u;                   // 128-byte output chunk counter across this CTA's tiles
        const int s_c = gu & 3;
        uint8_t* row_c = slots + s_c * EPI_SLOT_BYTES + (q * 32 + lane) * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(row_c + ((j ^ sw) << 4)) =
              make_float4(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        // one barrier per chunk: behind it every row of the slot is written (and fenced towards the async proxy), and
        // the slot the NEXT chunk writes has been read out by its previous store (epi_t0 checks before arriving)
        fence_proxy_async_smem();
        if (epi_t0) tma_store_wait_read<2>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (epi_t0) {
          if (accumulate) tma_reduce_add_2d(tc, slots + s_c * EPI_SLOT_BYTES, n0 + u * 32, m0);
          else tma_store_2d(tc, slots + s_c * EPI_SLOT_BYTES, n0 + u * 32, m0);
          tma_store_commit();
        }
      };
      if constexpr (ADAM) {
        // fused optimizer step: this thread owns row (m0 + q * 32 + lane) of the weight and 32 consecutive columns per unit;
        // the accumulator IS the gradient.  p / m / v of the unit (3 x 128 contiguous bytes) are requested before the
        // TMEM load is waited for; nothing passes through shared memory and no gradient is written.
        const GroupProblem& gp = p.prob[g];
        const int row = m0 + q * 32 + lane;
        const bool ok = row < gp.M;
        const long long e0 = gp.flat_off + (long long)(ok ? row : 0) * gp.ldw + n0;
        const float keep = (opt.decay[e0 >> opt.shift] & 1) ? 1.f - opt.a.lr * opt.a.wd : 1.f;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_acc + uint32_t(c * 32), r);
          float4 pv[8], mv[8], vv[8];
          const long long e = e0 + c * 32;
          if (ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              pv[j] = *reinterpret_cast<const float4*>(opt.p + e + 4 * j);
              mv[j] = *reinterpret_cast<const float4*>(opt.m + e + 4 * j);
              vv[j] = *reinterpret_cast<const float4*>(opt.v + e + 4 * j);
            }
          }
          tmem_ld_wait_regs(r);
          if (c == BN / 32 - 1) {                  // accumulator fully read: hand it back before the last stores
            tc_fence_before();
            mbar_arrive(&tmem_empty_bar[acc]);
          }
          if (ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              adamw_element(pv[j].x, __uint_as_float(r[4 * j + 0]), mv[j].x, vv[j].x, keep, opt.a);
              adamw_element(pv[j].y, __uint_as_float(r[4 * j + 1]), mv[j].y, vv[j].y, keep, opt.a);
              adamw_element(pv[j].z, __uint_as_float(r[4 * j + 2]), mv[j].z, vv[j].z, keep, opt.a);
              adamw_element(pv[j].w, __uint_as_float(r[4 * j + 3]), mv[j].w, vv[j].w, keep, opt.a);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              *reinterpret_cast<float4*>(opt.p + e + 4 * j) = pv[j];
              *reinterpret_cast<float4*>(opt.m + e + 4 * j) = mv[j];
              *reinterpret_cast<float4*>(opt.v + e + 4 * j) = vv[j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(opt.p16 + e + 8 * j) =
                  make_uint4(pack_bf16x2(pv[2 * j].x, pv[2 * j].y), pack_bf16x2(pv[2 * j].z, pv[2 * j].w),
                             pack_bf16x2(pv[2 * j + 1].x, pv[2 * j + 1].y), pack_bf16x2(pv[2 * j + 1].z, pv[2 * j + 1].w));
          }
        }
        continue;
      }
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(t_acc, ra);
#pragma unroll 1
      for (int c = 0; c < BN / 32; c += 2) {
        tmem_ld_wait_regs(ra);
        tmem_ld_32x32(t_acc + uint32_t((c + 1) * 32), rb);
        staged_unit(ra, c);
        tmem_ld_wait_regs(rb);
        if (c + 2 < BN / 32) tmem_ld_32x32(t_acc + uint32_t((c + 2) * 32), ra);
        else {                                   // accumulator fully read: hand it back before the last stores
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[acc]);
        }
        staged_unit(rb, c + 1);
      }
    }
    if (!ADAM && epi_t0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// One entry of a grouped weight-gradient launch: dW[n_out, k_in] (+)= dY[rows, n_out]^T X[rows, k_in]
struct WgradDesc {
  const void* dY; int ld_dy;
  const void* X; int ldx;
  float* dW; int ldw;
  int n_out, k_in, rows, accumulate;
};

int gemm_bf16_aux(const void* A, int lda, int a_mn, const void* B, int ldb, int b_mn, void* C, int ldc, int M, int N,
                  int K, const float* bias, const void* residual, int ldres, void* aux, int ldaux, int flags,
                  cudaStream_t stream);

// The problems must write disjoint outputs (they run concurrently).  Falls back to one launch per problem when a
// shape does not fit the grouped kernel's tile (k_in % 256, 16-byte row alignment).
// fuse != nullptr: the AdamW step of every problem's weight in the epilogue instead of the gradient store (see AdamFuse;
// grad_base = the flat gradient buffer the dW pointers point into).  Requires shapes the grouped kernel takes.
struct AdamFuseHost {
  const float* grad_base; float* p; float* m; float* v; void* p16; const unsigned char* decay; int shift;
  float lr, beta1, beta2, eps, wd; int step;
};
AdamScalars make_adam_scalars(float lr, float beta1, float beta2, float eps, float wd, int step);   // csrc/elementwise.cu

// a problem with an optional second reduction segment: dW (+)= dY^T X + dY2^T X2 (rows2 = 0: none)
struct WgradDescEx {
  const void* dY; int ld_dy;
  const void* X; int ldx;
  float* dW; int ldw;
  int n_out, k_in, rows, accumulate;
  const void* dY2; int ld_dy2;
  const void* X2; int ldx2;
  int rows2;
};
int gemm_grouped_wgrad_ex(const WgradDescEx* d, int n, const AdamFuseHost* fuse, cudaStream_t stream);
int gemm_grouped_wgrad(const WgradDesc* d, int n, cudaStream_t stream) {
  MEBT_REQUIRE(n >= 0 && n <= GG_MAX, MEBT_ERR_SHAPE, "grouped wgrad: %d problems (max %d)", n, GG_MAX);
  WgradDescEx e[GG_MAX];
  for (int i = 0; i < n; ++i)
    e[i] = WgradDescEx{d[i].dY, d[i].ld_dy, d[i].X, d[i].ldx, d[i].dW, d[i].ldw, d[i].n_out, d[i].k_in, d[i].rows, d[i].accumulate,
                       nullptr, 0, nullptr, 0, 0};
  return gemm_grouped_wgrad_ex(e, n, nullptr, stream);
}

int gemm_grouped_wgrad_ex(const WgradDescEx* d, int n, const AdamFuseHost* fuse, cudaStream_t stream) {
  MEBT_REQUIRE(n >= 0 && n <= GG_MAX, MEBT_ERR_SHAPE, "grouped wgrad: %d problems (max %d)", n, GG_MAX);
  if (n == 0) return MEBT_OK;
  bool fits = true;
  for (int i = 0; i < n; ++i) {
    MEBT_REQUIRE(d[i].n_out > 0 && d[i].k_in > 0 && d[i].rows > 0, MEBT_ERR_SHAPE, "grouped wgrad: empty problem %d", i);
    fits = fits && d[i].k_in % BN == 0 && d[i].ld_dy % 8 == 0 && d[i].ldx % 8 == 0 && d[i].ldw % 4 == 0;
    if (d[i].rows2 > 0) fits = fits && d[i].ld_dy2 % 8 == 0 && d[i].ldx2 % 8 == 0;
  }
  bool two_seg = false;
  for (int i = 0; i < n; ++i) two_seg = two_seg || d[i].rows2 > 0;
  MEBT_REQUIRE(!two_seg || fits, MEBT_ERR_UNSUPPORTED, "grouped wgrad: a two-segment problem needs k_in %% 256 == 0 and aligned rows");
  MEBT_REQUIRE(fuse == nullptr || fits, MEBT_ERR_UNSUPPORTED, "grouped wgrad: a fused optimizer step needs k_in %% 256 == 0 and aligned rows");
  if (fuse != nullptr)
    for (int i = 0; i < n; ++i)
      MEBT_REQUIRE(!d[i].accumulate && d[i].ldw % 8 == 0 && (d[i].dW - fuse->grad_base) % 8 == 0 && d[i].dW >= fuse->grad_base, MEBT_ERR_UNSUPPORTED,
                   "grouped wgrad: a fused optimizer step cannot accumulate and needs 32-byte aligned weights");
  if (fuse == nullptr && !two_seg && (!fits || n == 1)) {
    for (int i = 0; i < n; ++i) {
      int rc = gemm_bf16_aux(d[i].dY, d[i].ld_dy, 1, d[i].X, d[i].ldx, 1, d[i].dW, d[i].ldw, d[i].n_out, d[i].k_in,
                             d[i].rows, nullptr, nullptr, 0, nullptr, 0,
                             MEBT_GEMM_OUT_FP32 | (d[i].accumulate ? MEBT_GEMM_ACCUMULATE : 0), stream);
      if (rc) return rc;
    }
    return MEBT_OK;
  }
  // longest reductions first: the tail of the launch is then made of the short tiles
  int order[GG_MAX];
  for (int i = 0; i < n; ++i) order[i] = i;
  for (int i = 1; i < n; ++i)
    for (int j = i; j > 0 && d[order[j]].rows + d[order[j]].rows2 > d[order[j - 1]].rows + d[order[j - 1]].rows2; --j) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
  GroupMaps maps;
  GroupParams p;
  memset(&p, 0, sizeof(p));
  p.n_problems = n;
  int tiles = 0;
  double flops = 0.0;
  for (int s = 0; s < n; ++s) {
    const WgradDescEx& w = d[order[s]];
    GroupProblem& q = p.prob[s];
    q.M = w.n_out; q.N = w.k_in; q.K = w.rows;
    q.num_m_blocks = (q.M + BM - 1) / BM;
    q.num_n_blocks = q.N / BN;
    q.kb_split = (q.K + BK - 1) / BK;
    q.num_k_blocks = q.kb_split + (w.rows2 + BK - 1) / BK;
    q.tile_begin = tiles;
    q.accumulate = w.accumulate ? 1 : 0;
    q.flat_off = fuse != nullptr ? (long long)(w.dW - fuse->grad_base) : 0;
    q.ldw = w.ldw;
    tiles += q.num_m_blocks * q.num_n_blocks;
    flops += 2.0 * double(q.M) * double(q.N) * double(q.K + w.rows2);
    int rc = get_tensor_map_2d(&maps.a[s], w.dY, 2, uint64_t(q.M), uint64_t(q.K), uint64_t(w.ld_dy) * 2, 64, BK);
    if (rc) return rc;
    rc = get_tensor_map_2d(&maps.b[s], w.X, 2, uint64_t(q.N), uint64_t(q.K), uint64_t(w.ldx) * 2, 64, BK);
    if (rc) return rc;
    rc = get_tensor_map_2d(&maps.c[s], w.dW, 4, uint64_t(q.N), uint64_t(q.M), uint64_t(w.ldw) * 4, 32, 128);
    if (rc) return rc;
    maps.a2[s] = maps.a[s]; maps.b2[s] = maps.b[s];
    if (w.rows2 > 0) {
      rc = get_tensor_map_2d(&maps.a2[s], w.dY2, 2, uint64_t(q.M), uint64_t(w.rows2), uint64_t(w.ld_dy2) * 2, 64, BK);
      if (rc) return rc;
      rc = get_tensor_map_2d(&maps.b2[s], w.X2, 2, uint64_t(q.N), uint64_t(w.rows2), uint64_t(w.ldx2) * 2, 64, BK);
      if (rc) return rc;
    }
  }
  for (int s = n; s < GG_MAX; ++s) {
    maps.a[s] = maps.a[0]; maps.b[s] = maps.b[0]; maps.c[s] = maps.c[0]; maps.a2[s] = maps.a[0]; maps.b2[s] = maps.b[0];
  }
  p.total_tiles = tiles;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(gemm_grouped_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    MEBT_CUDA_OK(cudaFuncSetAttribute(gemm_grouped_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
  }
  LaunchScope ls(FAM_GEMM, flops, stream);
  static const int cta_cap = getenv("MEBT_WGRAD_CTAS") != nullptr ? atoi(getenv("MEBT_WGRAD_CTAS")) : 0;   // experiment knob
  const int cap = cta_cap > 0 ? cta_cap : sm_count();
  const int grid = tiles < cap ? tiles : cap;
  AdamFuse opt;
  memset(&opt, 0, sizeof(opt));
  if (fuse != nullptr) {
    opt.p = fuse->p; opt.m = fuse->m; opt.v = fuse->v; opt.p16 = static_cast<__nv_bfloat16*>(fuse->p16);
    opt.decay = fuse->decay; opt.shift = fuse->shift;
    opt.a = make_adam_scalars(fuse->lr, fuse->beta1, fuse->beta2, fuse->eps, fuse->wd, fuse->step);
    MEBT_CUDA_OK(launch_pdl(gemm_grouped_wgrad_kernel<true>, dim3(grid), dim3(GG_THREADS), SMEM_TOTAL, stream, maps, p, opt));
  } else {
    MEBT_CUDA_OK(launch_pdl(gemm_grouped_wgrad_kernel<false>, dim3(grid), dim3(GG_THREADS), SMEM_TOTAL, stream, maps, p, opt));
  }
  MEBT_LAUNCH_OK("gemm_grouped_wgrad_kernel");
  return MEBT_OK;
}

}  // namespace mebt

// C-ABI: mebt_wgrad_desc_t mirrors WgradDesc (include/mebt_b200.h)
extern "C" int mebt_gemm_grouped_wgrad(const mebt_wgrad_desc_t* problems, int n_problems, void* stream) {
  static_assert(sizeof(mebt_wgrad_desc_t) == sizeof(mebt::WgradDesc), "descriptor layouts must match");
  return mebt::gemm_grouped_wgrad(reinterpret_cast<const mebt::WgradDesc*>(problems), n_problems,
                                  static_cast<cudaStream_t>(stream));
}
