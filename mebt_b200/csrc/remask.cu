// K7: confidence-based re-masking between two forwards of the maskgit sampling loop.
//   order = sort_desc( (s / sum s) / q^ctemp ),  q ~ Exp(1)
//   next_context = cat[context, target[order[:n_new]]],  next_target = target[order[n_new:]]
// reference: MaskGen.gumbel_top_k (mebt/mask_sampler.py:178-187) + generate_next_mask (:189-236).
// One CTA per batch row; the whole row (<= 8192 keys) is sorted in shared memory with a bitonic network
// on (key, index) pairs.  Order: key descending, index ascending on equal keys (a total order, so the
// result is deterministic; torch.sort gives no guarantee on ties).
#include "common.cuh"

namespace mebt {
namespace {

constexpr int RT = 1024;

__device__ __forceinline__ bool comes_before(float ka, int ia, float kb, int ib) {
  return ka > kb || (ka == kb && ia < ib);
}

__device__ __forceinline__ uint4 philox4x32_r(uint4 ctr, uint2 key) {
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

__global__ void __launch_bounds__(RT) remask_sort_kernel(const float* __restrict__ score, const float* __restrict__ noise,
                                                         float ctemp, const int64_t* __restrict__ ctx_idx,
                                                         int ctx_stride, const int64_t* __restrict__ tgt_idx,
                                                         int tgt_stride, int NC, int NT, int n_new, int n_pow2,
                                                         unsigned long long seed, unsigned long long offset,
                                                         int64_t* __restrict__ next_ctx, int64_t* __restrict__ next_tgt,
                                                         int64_t* __restrict__ order_out) {
  extern __shared__ uint8_t sm[];
  float* key = reinterpret_cast<float*>(sm);
  int* idx = reinterpret_cast<int*>(sm + size_t(n_pow2) * 4);
  __shared__ float red[RT / 32];
  const int b = blockIdx.x;
  const float* s = score + (long long)b * NT;

  // sum of scores in a fixed order (prob / prob.sum(-1), mask_sampler.py:180)
  float part = 0.f;
  for (int i = threadIdx.x; i < NT; i += RT) part += s[i];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  float total = 0.f;
  for (int w = 0; w < RT / 32; ++w) total += red[w];

  for (int i = threadIdx.x; i < n_pow2; i += RT) {
    float k = -INFINITY;
    if (i < NT) {
      float q;
      if (noise != nullptr) {
        q = noise[(long long)b * NT + i];
      } else {
        const uint4 r = philox4x32_r(make_uint4(uint32_t(b), uint32_t(i), uint32_t(offset), uint32_t(offset >> 32)),
                                     make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
        q = -__logf((float(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f));
      }
      const float p = __fdiv_rn(s[i], total);
      // q ** ctemp (mask_sampler.py:183); ctemp == 0 -> 1 exactly, as torch.pow
      const float d = ctemp == 0.f ? 1.0f : powf(q, ctemp);
      k = __fdiv_rn(p, d);
      if (k != k) k = -INFINITY;
    }
    key[i] = k;
    idx[i] = i < NT ? i : 0x7fffffff;
  }
  __syncthreads();

  for (int size = 2; size <= n_pow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n_pow2 >> 1); t += RT) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;   // this sub-sequence ends up in "comes_before" order
        const float ka = key[lo], kb = key[hi];
        const int ia = idx[lo], ib = idx[hi];
        const bool in_order = comes_before(ka, ia, kb, ib);
        if (in_order != up) {
          key[lo] = kb; key[hi] = ka;
          idx[lo] = ib; idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }

  const int64_t* tg = tgt_idx + (long long)b * tgt_stride;
  if (next_ctx != nullptr) {
    int64_t* nc = next_ctx + (long long)b * (NC + n_new);
    const int64_t* cx = ctx_idx + (long long)b * ctx_stride;
    for (int i = threadIdx.x; i < NC; i += RT) nc[i] = cx[i];
    for (int i = threadIdx.x; i < n_new; i += RT) nc[NC + i] = tg[idx[i]];
  }
  if (next_tgt != nullptr) {
    int64_t* nt = next_tgt + (long long)b * (NT - n_new);
    for (int i = threadIdx.x; i < NT - n_new; i += RT) nt[i] = tg[idx[n_new + i]];
  }
  if (order_out != nullptr)
    for (int i = threadIdx.x; i < NT; i += RT) order_out[(long long)b * NT + i] = idx[i];
}

}  // namespace
}  // namespace mebt

extern "C" int mebt_remask_sort(const float* score, const float* noise, float ctemp, const int64_t* ctx_idx,
                                int ctx_stride, const int64_t* tgt_idx, int tgt_stride, int B, int NC, int NT, int n_new,
                                unsigned long long seed, unsigned long long offset, int64_t* next_ctx,
                                int64_t* next_tgt, int64_t* order_out, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(B >= 0 && NC >= 0 && NT > 0 && n_new >= 0 && n_new <= NT, MEBT_ERR_SHAPE,
               "remask_sort: bad shape B=%d NC=%d NT=%d n_new=%d", B, NC, NT, n_new);
  MEBT_REQUIRE(NT <= 16384, MEBT_ERR_UNSUPPORTED, "remask_sort: NT=%d exceeds the 16384-key shared-memory sort", NT);
  if (B == 0) return MEBT_OK;
  int n_pow2 = 2;
  while (n_pow2 < NT) n_pow2 <<= 1;
  const size_t smem = size_t(n_pow2) * 8;
  static bool configured[64] = {};
  if (smem > 48 * 1024 && first_use_on_device(configured)) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(remask_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(128 * 1024)));
  }
  LaunchScope ls(FAM_REMASK, double(B) * (double(NT) * (4.0 + (noise != nullptr ? 4.0 : 0.0) + 16.0) + double(NC) * 16.0),
                 static_cast<cudaStream_t>(stream));
  remask_sort_kernel<<<B, RT, smem, static_cast<cudaStream_t>(stream)>>>(score, noise, ctemp, ctx_idx, ctx_stride,
                                                                         tgt_idx, tgt_stride, NC, NT, n_new, n_pow2,
                                                                         seed, offset, next_ctx, next_tgt, order_out);
  MEBT_LAUNCH_OK("remask_sort_kernel");
  return MEBT_OK;
}
