// K5 / K6: consumers of the 16384-way logit rows.  One CTA (256 threads) owns one row and keeps it in
// registers (64 values per thread), so each logit is read from HBM exactly once:
//   K5 masked cross-entropy (+ label smoothing, + rank of the target for top-1/top-5, + dlogits)
//      reference: F.cross_entropy(sum) at mebt/transformer.py:726 and accuracy() at mebt/utils.py:80-94
//   K6 temperature / top-k / softmax / Exp-race ("Gumbel") argmax / confidence score
//      reference: sample_from_logits + gumbel_sort, mebt/transformer.py:843-889 / :826-841
#include <cuda_fp16.h>

#include "common.cuh"

namespace mebt {
namespace {

constexpr int LT = 256;          // threads per row
constexpr int MAX_V4 = 16;       // float4 groups per thread  -> V <= 256 * 16 * 4 = 16384

template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < LT / 32; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < LT / 32; ++i) r += red[i];   // fixed order: deterministic
  __syncthreads();
  return r;
}
__device__ __forceinline__ int block_sum_int(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  int r = 0;
#pragma unroll
  for (int i = 0; i < LT / 32; ++i) r += red[i];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// K5
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LT) masked_ce_kernel(const T* __restrict__ logits, long long ld,
                                                       const int64_t* __restrict__ targets, int V, float smoothing,
                                                       float* __restrict__ row_loss, int* __restrict__ row_rank,
                                                       T* __restrict__ dlogits, long long ldd, float grad_scale) {
  __shared__ float red[LT / 32];
  __shared__ int redi[LT / 32];
  const long long row = blockIdx.x;
  const T* xr = logits + row * ld;
  const int tgt = int(targets[row]);
  float4 v[MAX_V4];
  float m = -INFINITY, total = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (threadIdx.x + LT * i) * 4;
    if (c < V) {
      v[i] = load4<T>(xr + c);
      m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      total += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  m = block_max(m, red);
  total = block_sum(total, red);
  const float xt = (tgt >= 0 && tgt < V) ? float(xr[tgt]) : 0.f;
  float se = 0.f;
  int rank = 0;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (threadIdx.x + LT * i) * 4;
    if (c < V) {
      se += (expf(v[i].x - m) + expf(v[i].y - m)) + (expf(v[i].z - m) + expf(v[i].w - m));
      rank += (v[i].x > xt) + (v[i].y > xt) + (v[i].z > xt) + (v[i].w > xt);
    }
  }
  se = block_sum(se, red);
  rank = block_sum_int(rank, redi);
  const float lse = m + logf(se);
  if (threadIdx.x == 0) {
    const float nll = lse - xt;
    const float smooth = lse - total / float(V);
    row_loss[row] = (1.f - smoothing) * nll + smoothing * smooth;
    if (row_rank != nullptr) row_rank[row] = rank;
  }
  if (dlogits != nullptr) {
    // d(sum CE)/dlogit_v = softmax_v - (1-eps) [v == t] - eps / V, times the upstream scale
    T* dr = dlogits + row * ldd;
    const float inv = 1.f / se, u = smoothing / float(V);
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      const int c = (threadIdx.x + LT * i) * 4;
      if (c < V) {
        float4 g;
        g.x = expf(v[i].x - m) * inv - u;
        g.y = expf(v[i].y - m) * inv - u;
        g.z = expf(v[i].z - m) * inv - u;
        g.w = expf(v[i].w - m) * inv - u;
        if (tgt >= c && tgt < c + 4) (&g.x)[tgt - c] -= (1.f - smoothing);
        g.x *= grad_scale; g.y *= grad_scale; g.z *= grad_scale; g.w *= grad_scale;
        store4<T>(dr + c, g);
      }
    }
  }
}

// K5, bf16 logits (the training path): 256 threads per row, the row kept PACKED in registers (8 x 16 bytes = 64 logits per
// thread, 32 registers) so that four rows are in flight per SM and their load / reduce / store phases overlap; one
// combined block reduction for (max, sum), exponentials on MUFU ex2 (the fp32 kernel above stays the generic path).
constexpr int CE_T = 256;
__global__ void __launch_bounds__(CE_T, 3) masked_ce_bf16_kernel(const __nv_bfloat16* __restrict__ logits, long long ld,
                                                              const int64_t* __restrict__ targets, int V, float smoothing,
                                                              float* __restrict__ row_loss, int* __restrict__ row_rank,
                                                              __nv_bfloat16* dlogits, long long ldd, float grad_scale) {
  __shared__ float red_a[CE_T / 32], red_b[CE_T / 32];
  __shared__ int red_i[CE_T / 32];
  const long long row = blockIdx.x;
  const __nv_bfloat16* xr = logits + row * ld;
  const int tgt = int(targets[row]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4 v[8];
  float m = -INFINITY, total = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (threadIdx.x + CE_T * i) * 8;
    if (c < V) {
      v[i] = *reinterpret_cast<const uint4*>(xr + c);
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        m = fmaxf(m, fmaxf(f.x, f.y));
        total += f.x + f.y;
      }
    }
  }
  m = warp_max(m);
  total = warp_sum(total);
  if (lane == 0) { red_a[warp] = m; red_b[warp] = total; }
  __syncthreads();
  m = red_a[0];
  total = red_b[0];
#pragma unroll
  for (int i = 1; i < CE_T / 32; ++i) { m = fmaxf(m, red_a[i]); total += red_b[i]; }
  __syncthreads();
  const float xt = (tgt >= 0 && tgt < V) ? __bfloat162float(xr[tgt]) : 0.f;
  const float ml2 = m * 1.4426950408889634f;
  float se = 0.f;
  int rank = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = (threadIdx.x + CE_T * i) * 8;
    if (c < V) {
      const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        se += ex2_approx(fmaf(f.x, 1.4426950408889634f, -ml2)) + ex2_approx(fmaf(f.y, 1.4426950408889634f, -ml2));
        rank += (f.x > xt) + (f.y > xt);
      }
    }
  }
  se = warp_sum(se);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  if (lane == 0) { red_a[warp] = se; red_i[warp] = rank; }
  __syncthreads();
  se = 0.f;
  rank = 0;
#pragma unroll
  for (int i = 0; i < CE_T / 32; ++i) { se += red_a[i]; rank += red_i[i]; }     // fixed order: deterministic
  const float lse = m + logf(se);
  if (threadIdx.x == 0) {
    const float nll = lse - xt;
    const float smooth = lse - total / float(V);
    row_loss[row] = (1.f - smoothing) * nll + smoothing * smooth;
    if (row_rank != nullptr) row_rank[row] = rank;
  }
  if (dlogits != nullptr) {
    // d(sum CE)/dlogit_v = softmax_v - (1-eps) [v == t] - eps / V, times the upstream scale (may overwrite the logits)
    __nv_bfloat16* dr = dlogits + row * ldd;
    const float inv = grad_scale / se, u = grad_scale * smoothing / float(V), hot = grad_scale * (1.f - smoothing);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = (threadIdx.x + CE_T * i) * 8;
      if (c < V) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[i]);
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(w[k]);
          float g0 = fmaf(ex2_approx(fmaf(f.x, 1.4426950408889634f, -ml2)), inv, -u);
          float g1 = fmaf(ex2_approx(fmaf(f.y, 1.4426950408889634f, -ml2)), inv, -u);
          if (tgt == c + 2 * k) g0 -= hot;
          if (tgt == c + 2 * k + 1) g1 -= hot;
          ow[k] = pack_bf16x2(g0, g1);
        }
        *reinterpret_cast<uint4*>(dr + c) = o;
      }
    }
  }
}

// K5, bf16 logits, V <= 16384: PERSISTENT CTAs (two per SM) that stream rows through a two-deep shared-memory ring filled
// by 1-D bulk copies, so that the load of row r+1 is in flight during the whole of row r's arithmetic and stores (the
// one-row-per-CTA kernel above has no load outstanding while it reduces, exponentiates and stores: 0.41 of the HBM
// peak).  Each exponential is computed once and kept as packed fp16 (the gradient is stored as bf16: 8 mantissa bits).
constexpr int CE_ROW_BYTES = 16384 * 2;
constexpr int CE_STAGES = 3;          // rows in flight per CTA (two CTAs per SM: 192 KiB of the 227)
__global__ void __launch_bounds__(CE_T, 2) masked_ce_bf16_stream_kernel(const __nv_bfloat16* __restrict__ logits, long long ld,
                                                                     const int64_t* __restrict__ targets, int V, int rows,
                                                                     float smoothing, float* __restrict__ row_loss,
                                                                     int* __restrict__ row_rank, __nv_bfloat16* dlogits,
                                                                     long long ldd, float grad_scale) {
  extern __shared__ __align__(128) uint8_t ce_smem[];
  __shared__ float red_a[CE_T / 32], red_b[CE_T / 32];
  __shared__ int red_i[CE_T / 32];
  __shared__ __align__(8) uint64_t bar[CE_STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t row_bytes = uint32_t(V) * 2u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < CE_STAGES; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
  }
  __syncthreads();
  griddep_wait();
  long long row = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int i = 0; i < CE_STAGES - 1; ++i) {
      const long long r = row + (long long)i * gridDim.x;
      if (r < rows) {
        mbar_arrive_expect_tx(&bar[i], row_bytes);
        bulk_load_1d(ce_smem + i * CE_ROW_BYTES, logits + r * ld, row_bytes, &bar[i]);
      }
    }
  }
  for (int it = 0; row < rows; row += gridDim.x, ++it) {
    const int buf = it % CE_STAGES;
    const long long ahead = row + (long long)(CE_STAGES - 1) * gridDim.x;
    const int abuf = (it + CE_STAGES - 1) % CE_STAGES;
    // buffer abuf was last read in iteration it-1, which copied its row into registers before that iteration's barriers
    if (threadIdx.x == 0 && ahead < rows) {
      mbar_arrive_expect_tx(&bar[abuf], row_bytes);
      bulk_load_1d(ce_smem + abuf * CE_ROW_BYTES, logits + ahead * ld, row_bytes, &bar[abuf]);
    }
    const int tgt = int(targets[row]);
    mbar_wait(&bar[buf], uint32_t(it / CE_STAGES) & 1u);
    const uint8_t* sr = ce_smem + buf * CE_ROW_BYTES;
    // Packed arithmetic throughout (the kernel is issue-bound next to the HBM stream): the maximum and the rank count in
    // bf16x2 (exact: a maximum of bf16 values; counts <= 64 per lane half), the exponent argument and the sums in f32x2,
    // the label-smoothing sum only when smoothing is on, the one-hot correction as a single store after the row.
    uint4 v[8];
    __nv_bfloat162 m2 = __float2bfloat162_rn(-INFINITY);
    float total = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = (threadIdx.x + CE_T * i) * 8;
      if (c < V) {
        v[i] = *reinterpret_cast<const uint4*>(sr + size_t(c) * 2);
        const __nv_bfloat162* w = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
#pragma unroll
        for (int k = 0; k < 4; ++k) m2 = __hmax2(m2, w[k]);
        if (smoothing != 0.f) {
          const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = unpack_bf16x2(u[k]);
            total += f.x + f.y;
          }
        }
      }
    }
    const bool tgt_ok = tgt >= 0 && tgt < V;
    const __nv_bfloat16 xt_b = tgt_ok ? reinterpret_cast<const __nv_bfloat16*>(sr)[tgt] : __float2bfloat16(0.f);
    const float xt = __bfloat162float(xt_b);
    float m = warp_max(fmaxf(__low2float(m2), __high2float(m2)));
    total = warp_sum(total);
    if (lane == 0) { red_a[warp] = m; red_b[warp] = total; }
    __syncthreads();
    m = red_a[0];
    total = red_b[0];
#pragma unroll
    for (int i = 1; i < CE_T / 32; ++i) { m = fmaxf(m, red_a[i]); total += red_b[i]; }
    __syncthreads();
    const float2 nml2 = make_float2(-m * 1.4426950408889634f, -m * 1.4426950408889634f);
    const float2 l2e = make_float2(1.4426950408889634f, 1.4426950408889634f);
    const __nv_bfloat162 xt2 = __bfloat162bfloat162(xt_b);
    float2 se2 = make_float2(0.f, 0.f);
    __nv_bfloat162 cnt2 = __float2bfloat162_rn(0.f);
    uint32_t e16[32];                                         // exp(x - max) as packed fp16 pairs
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = (threadIdx.x + CE_T * i) * 8;
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[i]);
      const __nv_bfloat162* w = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 e = make_float2(0.f, 0.f);
        if (c < V) {
          const float2 x = __ffma2_rn(unpack_bf16x2(u[k]), l2e, nml2);
          e.x = ex2_approx(x.x);
          e.y = ex2_approx(x.y);
          se2 = __fadd2_rn(se2, e);
          cnt2 = __hadd2(cnt2, __hgt2(w[k], xt2));            // 1.0 where the logit beats the target's
        }
        const __half2 h = __floats2half2_rn(e.x, e.y);
        e16[4 * i + k] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    float se = warp_sum(se2.x + se2.y);
    int rank = int(__low2float(cnt2) + __high2float(cnt2));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0) { red_a[warp] = se; red_i[warp] = rank; }
    __syncthreads();
    se = 0.f;
    rank = 0;
#pragma unroll
    for (int i = 0; i < CE_T / 32; ++i) { se += red_a[i]; rank += red_i[i]; }     // fixed order: deterministic
    if (threadIdx.x == 0) {
      const float lse = m + logf(se);
      const float nll = lse - xt;
      const float smooth = lse - total / float(V);
      row_loss[row] = (1.f - smoothing) * nll + (smoothing != 0.f ? smoothing * smooth : 0.f);
      if (row_rank != nullptr) row_rank[row] = rank;
    }
    if (dlogits != nullptr) {
      // d(sum CE)/dlogit_v = softmax_v - (1-eps) [v == t] - eps / V, times the upstream scale (may overwrite the logits:
      // the row was read completely, into shared memory, before its first store)
      __nv_bfloat16* dr = dlogits + row * ldd;
      const float inv = grad_scale / se, u_s = grad_scale * smoothing / float(V), hot = grad_scale * (1.f - smoothing);
      const float2 inv2 = make_float2(inv, inv), nu2 = make_float2(-u_s, -u_s);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = (threadIdx.x + CE_T * i) * 8;
        if (c < V) {
          uint4 o;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 g2 = __ffma2_rn(__half22float2(*reinterpret_cast<const __half2*>(&e16[4 * i + k])), inv2, nu2);
            ow[k] = pack_bf16x2(g2.x, g2.y);
          }
          *reinterpret_cast<uint4*>(dr + c) = o;
        }
      }
      // the target's entry also carries -(1 - eps): rewritten by the thread that owns its column, after its own store of
      // the surrounding 16 bytes (same thread, program order)
      if (tgt_ok && ((tgt >> 3) & (CE_T - 1)) == int(threadIdx.x)) {
        const float e_t = ex2_approx(fmaf(xt, 1.4426950408889634f, -m * 1.4426950408889634f));
        dr[tgt] = __float2bfloat16(fmaf(e_t, inv, -u_s) - hot);
      }
    }
    __syncthreads();                                          // red_* are reused by the next row
  }
}

// deterministic single-CTA reduction of the per-row results: out = {ce_sum, n_top1, n_top5}
__global__ void ce_reduce_kernel(const float* __restrict__ row_loss, const int* __restrict__ row_rank, int rows,
                                 float* __restrict__ out) {
  __shared__ double sd[1024];
  __shared__ int s1[1024], s5[1024];
  double acc = 0.0;
  int n1 = 0, n5 = 0;
  for (int i = threadIdx.x; i < rows; i += 1024) {
    acc += double(row_loss[i]);
    if (row_rank != nullptr) {
      const int r = row_rank[i];
      n1 += (r == 0);
      n5 += (r < 5);
    }
  }
  sd[threadIdx.x] = acc; s1[threadIdx.x] = n1; s5[threadIdx.x] = n5;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sd[threadIdx.x] += sd[threadIdx.x + o];
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s5[threadIdx.x] += s5[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = float(sd[0]);
    out[1] = float(s1[0]);
    out[2] = float(s5[0]);
  }
}

// ---------------------------------------------------------------------------------------------
// K6
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_key(float f) {   // monotone float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Philox4x32-10 (Salmon et al.), counter = (row, group, offset_lo, offset_hi), key = seed
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

template <typename T, bool RACE>
__global__ void __launch_bounds__(LT) sample_logits_kernel(const T* __restrict__ logits, long long ld, int V,
                                                           float temp_div, int top_k, float top_p,
                                                           const float* __restrict__ noise,
                                                           unsigned long long seed, unsigned long long offset,
                                                           int64_t* __restrict__ ids, float* __restrict__ scores,
                                                           float* __restrict__ probs_out) {
  __shared__ float red[LT / 32];
  __shared__ int hist[256];
  __shared__ uint32_t sel_prefix;
  __shared__ int sel_remaining;
  __shared__ float best_val[LT / 32];
  __shared__ int best_idx[LT / 32];
  const long long row = blockIdx.x;
  const T* xr = logits + row * ld;
  float4 v[MAX_V4];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (threadIdx.x + LT * i) * 4;
    if (c < V) {
      float4 t = load4<T>(xr + c);
      // logits.float() / (temperature + 1e-8), true division as on the CPU path (transformer.py:859-860)
      t.x = __fdiv_rn(t.x, temp_div); t.y = __fdiv_rn(t.y, temp_div);
      t.z = __fdiv_rn(t.z, temp_div); t.w = __fdiv_rn(t.w, temp_div);
      // NaN scrub (transformer.py:866-868), without the host sync
      if (t.x != t.x) t.x = -INFINITY;
      if (t.y != t.y) t.y = -INFINITY;
      if (t.z != t.z) t.z = -INFINITY;
      if (t.w != t.w) t.w = -INFINITY;
      v[i] = t;
    } else {
      v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
  }

  if (top_k > 0 && top_k < V) {
    // k-th largest value by 4-pass radix select on the order-preserving key (top_k_logits, transformer.py:891-895:
    // everything strictly below the k-th largest value becomes -inf; ties with it survive)
    if (threadIdx.x == 0) { sel_prefix = 0; sel_remaining = top_k; }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      hist[threadIdx.x] = 0;
      __syncthreads();
      const uint32_t prefix = sel_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
#pragma unroll
      for (int i = 0; i < MAX_V4; ++i) {
        const int c = (threadIdx.x + LT * i) * 4;
        if (c < V) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t key = float_order_key((&v[i].x)[j]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1);
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int remaining = sel_remaining, b = 255;
        for (; b > 0; --b) {
          if (hist[b] >= remaining) break;
          remaining -= hist[b];
        }
        sel_prefix = prefix | (uint32_t(b) << shift);
        sel_remaining = remaining;
      }
      __syncthreads();
    }
    const uint32_t kth = sel_prefix;
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (float_order_key((&v[i].x)[j]) < kth) (&v[i].x)[j] = -INFINITY;
  }

#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  m = block_max(m, red);
  float se = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    v[i].x = expf(v[i].x - m); v[i].y = expf(v[i].y - m); v[i].z = expf(v[i].z - m); v[i].w = expf(v[i].w - m);
    se += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  se = block_sum(se, red);

  if (top_p > 0.f && top_p < 1.f) {
    // Nucleus filter (top_p_probs, transformer.py:898-910): in descending order of probability keep every token
    // up to and including the one whose cumulative mass first reaches top_p, zero the rest, renormalise.  The
    // boundary probability is found by a 4-pass radix descent over the float bits with per-bin MASS histograms
    // (no sort); tokens tying with the boundary value are all kept (the reference's sort breaks such ties
    // arbitrarily).  v[] leaves this block holding the filtered, renormalised probabilities and se = 1.
    __shared__ float mass[256];
    __shared__ uint32_t np_prefix;
    __shared__ float np_above;
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      v[i].x = __fdiv_rn(v[i].x, se); v[i].y = __fdiv_rn(v[i].y, se);
      v[i].z = __fdiv_rn(v[i].z, se); v[i].w = __fdiv_rn(v[i].w, se);
    }
    if (threadIdx.x == 0) { np_prefix = 0; np_above = 0.f; }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      mass[threadIdx.x] = 0.f;
      __syncthreads();
      const uint32_t prefix = np_prefix;
      const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
#pragma unroll
      for (int i = 0; i < MAX_V4; ++i) {
        const int c = (threadIdx.x + LT * i) * 4;
        if (c < V) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float pj = (&v[i].x)[j];
            const uint32_t key = __float_as_uint(pj);            // p >= 0: the bit pattern is order preserving
            if ((key & pmask) == prefix && pj > 0.f) atomicAdd(&mass[(key >> shift) & 0xFF], pj);
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        float above = np_above;
        int b = 255;
        for (; b > 0; --b) {
          if (above + mass[b] >= top_p) break;
          above += mass[b];
        }
        np_prefix = prefix | (uint32_t(b) << shift);
        np_above = above;
      }
      __syncthreads();
    }
    const uint32_t boundary = np_prefix;
    float kept = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float& pj = (&v[i].x)[j];
        if (__float_as_uint(pj) < boundary) pj = 0.f;
        kept += pj;
      }
    kept = block_sum(kept, red);
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      v[i].x = __fdiv_rn(v[i].x, kept); v[i].y = __fdiv_rn(v[i].y, kept);
      v[i].z = __fdiv_rn(v[i].z, kept); v[i].w = __fdiv_rn(v[i].w, kept);
    }
    se = 1.0f;
  }

  if constexpr (!RACE) {
    // ---- fast mode: inverse-CDF draw.  One Philox uniform per row and a block-wide prefix sum select an id from
    // exactly the categorical distribution the reference's exponential race samples from, without generating
    // 16384 Exp(1) variates per row (the race needs 1 Philox round trip + 2 logs per logit; this needs none).
    __shared__ __align__(16) float wsum[MAX_V4][LT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MAX_V4; ++i) {
      const float sacc = warp_sum((v[i].x + v[i].y) + (v[i].z + v[i].w));
      if (lane == 0) wsum[i][wid] = sacc;
    }
    __syncthreads();
    // warp 0 locates the (chunk, warp) cell holding the target: 128 cell sums, 4 per lane, in vocabulary order
    __shared__ int sel_cell;
    __shared__ float sel_resid;
    if (wid == 0) {
      const uint4 r = philox4x32(make_uint4(uint32_t(row), uint32_t(row >> 32), uint32_t(offset), uint32_t(offset >> 32)),
                                 make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
      const float u = (float(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);     // (0, 1]
      const float4 part = *reinterpret_cast<const float4*>(&wsum[0][0] + lane * 4);
      const float own4 = (part.x + part.y) + (part.z + part.w);
      float inc = own4;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const float target = u * __shfl_sync(0xffffffffu, inc, 31);
      const unsigned hit = __ballot_sync(0xffffffffu, own4 > 0.f && inc >= target);
      const unsigned any = __ballot_sync(0xffffffffu, own4 > 0.f);
      const int sl = hit ? (__ffs(hit) - 1) : (any ? 31 - __clz(any) : 0);
      if (lane == sl) {
        const float rr = hit ? target - (inc - own4) : INFINITY;   // INFINITY: rounding pushed the target past the total
        const float e[4] = {part.x, part.y, part.z, part.w};
        int j_sel = -1, j_last = 0;
        float cum = 0.f, before = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (e[j] > 0.f) { j_last = j; if (j_sel < 0 && cum + e[j] >= rr) { j_sel = j; before = cum; } }
          cum += e[j];
        }
        if (j_sel < 0) { j_sel = j_last; before = -INFINITY; }
        sel_cell = lane * 4 + j_sel;
        sel_resid = rr - before;
      }
    }
    __syncthreads();
    const int ci = sel_cell / (LT / 32), wi = sel_cell % (LT / 32);
    const float resid = sel_resid;
    if (wid == wi) {
      float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < MAX_V4; ++i)
        if (i == ci) mine = v[i];
      const float my_own = (mine.x + mine.y) + (mine.z + mine.w);
      float my_incl = my_own;                      // inclusive scan over the lanes of the selected chunk only
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, my_incl, o);
        if (lane >= o) my_incl += t;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, my_own > 0.f && my_incl >= resid);
      const unsigned any = __ballot_sync(0xffffffffu, my_own > 0.f);
      const int sel = hit ? (__ffs(hit) - 1) : (31 - __clz(any));
      if (lane == sel) {
        const float rr = resid - (my_incl - my_own);
        const float e[4] = {mine.x, mine.y, mine.z, mine.w};
        int j_sel = -1, j_last = 0;
        float cum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cum += e[j];
          if (e[j] > 0.f) { j_last = j; if (j_sel < 0 && cum >= rr) j_sel = j; }
        }
        if (j_sel < 0) j_sel = j_last;
        ids[row] = (threadIdx.x + LT * ci) * 4 + j_sel;
        if (scores != nullptr) scores[row] = __fdiv_rn(e[j_sel], se);
      }
    }
    if (probs_out != nullptr) {
#pragma unroll
      for (int i = 0; i < MAX_V4; ++i) {
        const int c = (threadIdx.x + LT * i) * 4;
        if (c < V)
          *reinterpret_cast<float4*>(probs_out + row * V + c) =
              make_float4(__fdiv_rn(v[i].x, se), __fdiv_rn(v[i].y, se), __fdiv_rn(v[i].z, se), __fdiv_rn(v[i].w, se));
      }
    }
    return;
  }

  // ---- parity mode: the reference's exponential race on caller-supplied Exp(1) noise ----
  // probs = softmax; psum = probs.sum() as gumbel_sort renormalises (transformer.py:834)
  float psum = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    v[i].x = __fdiv_rn(v[i].x, se); v[i].y = __fdiv_rn(v[i].y, se);
    v[i].z = __fdiv_rn(v[i].z, se); v[i].w = __fdiv_rn(v[i].w, se);
    psum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  psum = block_sum(psum, red);

  float best = -1.f, best_p = 0.f;
  int besti = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < MAX_V4; ++i) {
    const int c = (threadIdx.x + LT * i) * 4;
    if (c < V) {
      if (probs_out != nullptr) *reinterpret_cast<float4*>(probs_out + row * V + c) = v[i];
      const float4 q = __ldg(reinterpret_cast<const float4*>(noise + row * V + c));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (&v[i].x)[j];
        // (p / sum p) / q, zeroed where p == 0 (transformer.py:834-838); first index wins ties
        const float race = p > 0.f ? __fdiv_rn(__fdiv_rn(p, psum), (&q.x)[j]) : 0.f;
        if (race > best) { best = race; besti = c + j; best_p = p; }
      }
    }
  }
  // block arg-max, lowest index on ties
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    const float op = __shfl_xor_sync(0xffffffffu, best_p, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; best_p = op; }
  }
  if ((threadIdx.x & 31) == 0) {
    best_val[threadIdx.x >> 5] = best;
    best_idx[threadIdx.x >> 5] = besti;
    red[threadIdx.x >> 5] = best_p;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < LT / 32; ++w)
      if (best_val[w] > best || (best_val[w] == best && best_idx[w] < besti)) {
        best = best_val[w]; besti = best_idx[w]; best_p = red[w];
      }
    ids[row] = besti;
    if (scores != nullptr) scores[row] = best_p;
  }
}


// ---------------------------------------------------------------------------------------------
// K6 fast path, streaming form: same inverse-CDF draw as sample_logits_kernel<T,false> but the row is NOT held
// in registers.  Pass 1 (HBM) computes the online-softmax statistics, pass 2 re-reads the 64 KiB row (an L2 hit)
// to build the 128 cell sums, and the selected warp re-reads one 16-byte vector.  ~40 registers per thread ->
// 8 CTAs per SM, which is what lets a one-pass, latency-bound reader approach the HBM roof.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(LT, 8) sample_stream_kernel(const T* __restrict__ logits, long long ld, int V,
                                                              float inv_temp, unsigned long long seed,
                                                              unsigned long long offset, int64_t* __restrict__ ids,
                                                              float* __restrict__ scores) {
  __shared__ float red_m[LT / 32], red_s[LT / 32];
  __shared__ __align__(16) float wsum[MAX_V4][LT / 32];
  __shared__ int sel_cell;
  __shared__ float sel_resid;
  const long long row = blockIdx.x;
  const T* xr = logits + row * ld;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  auto scaled = [&](float4 t) {
    t.x *= inv_temp; t.y *= inv_temp; t.z *= inv_temp; t.w *= inv_temp;
    if (t.x != t.x) t.x = -INFINITY;
    if (t.y != t.y) t.y = -INFINITY;
    if (t.z != t.z) t.z = -INFINITY;
    if (t.w != t.w) t.w = -INFINITY;
    return t;
  };
  // pass 1 (HBM): row maximum only — no transcendental work, so the pass runs at memory speed
  float m = -INFINITY;
  for (int c = threadIdx.x * 4; c < V; c += LT * 4) {
    const float4 t = scaled(load4<T>(xr + c));
    m = fmaxf(m, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
  }
  const float wm = warp_max(m);
  if (lane == 0) red_m[wid] = wm;
  for (int i = threadIdx.x; i < MAX_V4 * (LT / 32); i += LT) (&wsum[0][0])[i] = 0.f;
  __syncthreads();
  float M = red_m[0];
#pragma unroll
  for (int w = 1; w < LT / 32; ++w) M = fmaxf(M, red_m[w]);
  // pass 2: cell sums in vocabulary order (cell = 128 consecutive logits = one warp's float4s of one 1024-chunk)
  for (int c = threadIdx.x * 4, i = 0; c < V; c += LT * 4, ++i) {
    const float4 t = scaled(load4<T>(xr + c));
    const float e4 = (__expf(t.x - M) + __expf(t.y - M)) + (__expf(t.z - M) + __expf(t.w - M));
    const float cs = warp_sum(e4);
    if (lane == 0) wsum[i][wid] = cs;
  }
  __syncthreads();
  if (wid == 0) {
    const uint4 r = philox4x32(make_uint4(uint32_t(row), uint32_t(row >> 32), uint32_t(offset), uint32_t(offset >> 32)),
                               make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
    const float u = (float(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float4 part = *reinterpret_cast<const float4*>(&wsum[0][0] + lane * 4);
    const float own4 = (part.x + part.y) + (part.z + part.w);
    float inc = own4;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    const float row_total = __shfl_sync(0xffffffffu, inc, 31);      // = sum_v exp(x_v - M), the softmax denominator
    if (lane == 0) red_s[0] = row_total;
    const float target = u * row_total;
    const unsigned hit = __ballot_sync(0xffffffffu, own4 > 0.f && inc >= target);
    const unsigned any = __ballot_sync(0xffffffffu, own4 > 0.f);
    const int sl = hit ? (__ffs(hit) - 1) : (any ? 31 - __clz(any) : 0);
    if (lane == sl) {
      const float rr = hit ? target - (inc - own4) : INFINITY;
      const float e[4] = {part.x, part.y, part.z, part.w};
      int j_sel = -1, j_last = 0;
      float cum = 0.f, before = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (e[j] > 0.f) { j_last = j; if (j_sel < 0 && cum + e[j] >= rr) { j_sel = j; before = cum; } }
        cum += e[j];
      }
      if (j_sel < 0) { j_sel = j_last; before = -INFINITY; }
      sel_cell = lane * 4 + j_sel;
      sel_resid = rr - before;
    }
  }
  __syncthreads();
  const int ci = sel_cell / (LT / 32), wi = sel_cell % (LT / 32);
  if (wid == wi) {
    const int c = (threadIdx.x + LT * ci) * 4;
    float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < V) {
      const float4 t = scaled(load4<T>(xr + c));
      mine = make_float4(__expf(t.x - M), __expf(t.y - M), __expf(t.z - M), __expf(t.w - M));
    }
    const float my_own = (mine.x + mine.y) + (mine.z + mine.w);
    float my_incl = my_own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, my_incl, o);
      if (lane >= o) my_incl += t;
    }
    const float resid = sel_resid;
    const unsigned hit = __ballot_sync(0xffffffffu, my_own > 0.f && my_incl >= resid);
    const unsigned any = __ballot_sync(0xffffffffu, my_own > 0.f);
    const int sel = hit ? (__ffs(hit) - 1) : (any ? 31 - __clz(any) : 0);
    if (lane == sel) {
      const float rr = resid - (my_incl - my_own);
      const float e[4] = {mine.x, mine.y, mine.z, mine.w};
      int j_sel = -1, j_last = 0;
      float cum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cum += e[j];
        if (e[j] > 0.f) { j_last = j; if (j_sel < 0 && cum >= rr) j_sel = j; }
      }
      if (j_sel < 0) j_sel = j_last;
      ids[row] = c + j_sel;
      if (scores != nullptr) scores[row] = e[j_sel] / red_s[0];
    }
  }
}

}  // namespace
}  // namespace mebt

extern "C" {

int mebt_masked_ce(const void* logits, long long ld, int dtype, const int64_t* targets, int rows, int V,
                   float label_smoothing, float* row_loss, int* row_rank, void* dlogits, long long ld_d,
                   float grad_scale, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(rows >= 0 && V > 0 && V % 4 == 0 && V <= LT * MAX_V4 * 4, MEBT_ERR_SHAPE,
               "masked_ce: V=%d must be a multiple of 4 and <= %d", V, LT * MAX_V4 * 4);
  MEBT_REQUIRE(ld % 4 == 0 && (dlogits == nullptr || ld_d % 4 == 0), MEBT_ERR_SHAPE, "masked_ce: bad row stride");
  if (rows == 0) return MEBT_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double eb = dtype == MEBT_DTYPE_FP32 ? 4.0 : 2.0;
  LaunchScope ls(FAM_CE, double(rows) * (double(V) * eb * (dlogits != nullptr ? 2.0 : 1.0) + 16.0), st);
  if (dtype == MEBT_DTYPE_FP32)
    masked_ce_kernel<float><<<rows, LT, 0, st>>>(static_cast<const float*>(logits), ld, targets, V, label_smoothing,
                                                 row_loss, row_rank, static_cast<float*>(dlogits), ld_d, grad_scale);
  else if (dtype == MEBT_DTYPE_BF16 && V % 8 == 0 && V <= CE_T * 64 && ld % 8 == 0 && (dlogits == nullptr || ld_d % 8 == 0) &&
           rows >= 2 * sm_count()) {
    static bool attr_set[64] = {};
    if (first_use_on_device(attr_set)) {
      MEBT_CUDA_OK(cudaFuncSetAttribute(masked_ce_bf16_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CE_STAGES * CE_ROW_BYTES));
    }
    const int grid = rows < 2 * sm_count() ? rows : 2 * sm_count();
    masked_ce_bf16_stream_kernel<<<grid, CE_T, CE_STAGES * CE_ROW_BYTES, st>>>(static_cast<const __nv_bfloat16*>(logits), ld, targets,
                                                                      V, rows, label_smoothing, row_loss, row_rank,
                                                                      static_cast<__nv_bfloat16*>(dlogits), ld_d, grad_scale);
  } else if (dtype == MEBT_DTYPE_BF16 && V % 8 == 0 && V <= CE_T * 64 && ld % 8 == 0 && (dlogits == nullptr || ld_d % 8 == 0))
    masked_ce_bf16_kernel<<<rows, CE_T, 0, st>>>(static_cast<const __nv_bfloat16*>(logits), ld, targets, V, label_smoothing,
                                                 row_loss, row_rank, static_cast<__nv_bfloat16*>(dlogits), ld_d, grad_scale);
  else if (dtype == MEBT_DTYPE_BF16)
    masked_ce_kernel<__nv_bfloat16><<<rows, LT, 0, st>>>(static_cast<const __nv_bfloat16*>(logits), ld, targets, V,
                                                         label_smoothing, row_loss, row_rank,
                                                         static_cast<__nv_bfloat16*>(dlogits), ld_d, grad_scale);
  else
    MEBT_REQUIRE(false, MEBT_ERR_DTYPE, "masked_ce: unsupported dtype %d", dtype);
  MEBT_LAUNCH_OK("masked_ce_kernel");
  return MEBT_OK;
}

int mebt_ce_reduce(const float* row_loss, const int* row_rank, int rows, float* out3, void* stream) {
  MEBT_REQUIRE(rows >= 0, MEBT_ERR_SHAPE, "ce_reduce: bad rows");
  mebt::LaunchScope ls(mebt::FAM_OTHER, double(rows) * 8.0, static_cast<cudaStream_t>(stream));
  mebt::ce_reduce_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(row_loss, row_rank, rows, out3);
  MEBT_LAUNCH_OK("ce_reduce_kernel");
  return MEBT_OK;
}

int mebt_sample_logits(const void* logits, long long ld, int dtype, int rows, int V, float temperature, int top_k,
                       float top_p, const float* noise, unsigned long long seed, unsigned long long offset,
                       int64_t* ids, float* scores, float* probs, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(rows >= 0 && V > 0 && V % 4 == 0 && V <= LT * MAX_V4 * 4, MEBT_ERR_SHAPE,
               "sample_logits: V=%d must be a multiple of 4 and <= %d", V, LT * MAX_V4 * 4);
  MEBT_REQUIRE(ld % 4 == 0, MEBT_ERR_SHAPE, "sample_logits: bad row stride");
  if (rows == 0) return MEBT_OK;
  // python: temperature + 1e-8 in double, then cast to fp32 for the tensor division
  const float temp_div = float(double(temperature) + 1e-8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double eb = dtype == MEBT_DTYPE_FP32 ? 4.0 : 2.0;
  LaunchScope ls(FAM_SAMPLE, double(rows) * (double(V) * (eb + (noise != nullptr ? 4.0 : 0.0) + (probs != nullptr ? 4.0 : 0.0)) + 12.0), st);
  const bool nucleus = top_p > 0.f && top_p < 1.f;
  if (noise == nullptr && top_k <= 0 && !nucleus && probs == nullptr && V % 128 == 0) {
    // fast mode without filters: streaming inverse-CDF kernel (multiplies by 1/(T + 1e-8) instead of dividing)
    const float inv_temp = float(1.0 / (double(temperature) + 1e-8));
    if (dtype == MEBT_DTYPE_FP32)
      sample_stream_kernel<float><<<rows, LT, 0, st>>>(static_cast<const float*>(logits), ld, V, inv_temp, seed, offset,
                                                       ids, scores);
    else if (dtype == MEBT_DTYPE_BF16)
      sample_stream_kernel<__nv_bfloat16><<<rows, LT, 0, st>>>(static_cast<const __nv_bfloat16*>(logits), ld, V, inv_temp,
                                                               seed, offset, ids, scores);
    else
      MEBT_REQUIRE(false, MEBT_ERR_DTYPE, "sample_logits: unsupported dtype %d", dtype);
    MEBT_LAUNCH_OK("sample_stream_kernel");
    return MEBT_OK;
  }
  if (dtype == MEBT_DTYPE_FP32)
    (noise != nullptr ? sample_logits_kernel<float, true> : sample_logits_kernel<float, false>)<<<rows, LT, 0, st>>>(
        static_cast<const float*>(logits), ld, V, temp_div, top_k, top_p, noise, seed, offset, ids, scores, probs);
  else if (dtype == MEBT_DTYPE_BF16)
    (noise != nullptr ? sample_logits_kernel<__nv_bfloat16, true> : sample_logits_kernel<__nv_bfloat16, false>)
        <<<rows, LT, 0, st>>>(static_cast<const __nv_bfloat16*>(logits), ld, V, temp_div, top_k, top_p, noise, seed,
                              offset, ids, scores, probs);
  else
    MEBT_REQUIRE(false, MEBT_ERR_DTYPE, "sample_logits: unsupported dtype %d", dtype);
  MEBT_LAUNCH_OK("sample_logits_kernel");
  return MEBT_OK;
}

}  // extern "C"
