// fp32-accurate compute mode (north_star: logits within 1e-4 of the fp32 reference) on the bf16 tensor cores.
//
//  * split_f32_bf16x3_kernel: x = hi + lo (+ r, |r| <= 2^-16 |x|) with hi = bf16(x), lo = bf16(x - hi).  An fp32 GEMM
//    A W^T becomes ONE bf16 GEMM with the reduction dimension tripled,
//        [A_hi | A_lo | A_hi] . [W_hi | W_hi | W_lo]^T = A_hi W_hi + A_lo W_hi + A_hi W_lo,
//    which drops only the lo.lo and residual terms (relative 2^-16 per product) and accumulates in fp32 in TMEM: the
//    tcgen05 kernel of gemm.cu is reused unchanged (nn.Linear of mebt/modules/gpt.py:126-128,140,150-155,248).
//  * attention_f32_kernel: bmm / softmax / bmm of CrossAttention.forward (gpt.py:131-137) in plain fp32 FFMA + expf,
//    one thread per query row, K/V tiles staged in shared memory, online softmax, two key sources like K3.
//    A parity mode, not a throughput mode: it is what configs[0] (MeBT tiny, fp32) is checked with.
#include "common.cuh"

namespace mebt {
namespace {

// out row: [seg0 | seg1 | seg2], each K wide.  pattern 0 (activation side): hi | lo | hi; 1 (weight side): hi | hi | lo
__global__ void split_f32_bf16x3_kernel(const float* __restrict__ x, int ld, int rows, int K, __nv_bfloat16* __restrict__ out,
                                        int pattern) {
  const long long total = (long long)rows * (K / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = int(i / (K / 4)), c = int(i - (long long)r * (K / 4)) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * ld + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    float hi[4], lo[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hi[k] = __bfloat162float(__float2bfloat16_rn(f[k]));
      lo[k] = f[k] - hi[k];
    }
    uint2 h, l;
    h.x = pack_bf16x2(hi[0], hi[1]); h.y = pack_bf16x2(hi[2], hi[3]);
    l.x = pack_bf16x2(lo[0], lo[1]); l.y = pack_bf16x2(lo[2], lo[3]);
    __nv_bfloat16* o = out + (size_t)r * 3 * K + c;
    *reinterpret_cast<uint2*>(o) = h;
    *reinterpret_cast<uint2*>(o + K) = pattern == 0 ? l : h;
    *reinterpret_cast<uint2*>(o + 2 * K) = pattern == 0 ? h : l;
  }
}

constexpr int AF_ROWS = 64;      // query rows (threads) per CTA
constexpr int AF_KT = 32;        // keys per shared-memory tile

__global__ void __launch_bounds__(AF_ROWS) attention_f32_kernel(
    const float* __restrict__ Q, int ldq, int q_col0, const float* __restrict__ KV1, int ld1, int k1_col0, int v1_col0,
    int NK1, const float* __restrict__ KV2, int ld2, int k2_col0, int v2_col0, int NK2, float* __restrict__ O, int ldo,
    int NQ, float scale) {
  __shared__ float sK[AF_KT][64];
  __shared__ float sV[AF_KT][64];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qrow = blockIdx.x * AF_ROWS + threadIdx.x;
  const bool ok = qrow < NQ;
  float q[64], acc[64];
  if (ok) {
    const float4* qp = reinterpret_cast<const float4*>(Q + (size_t(b) * NQ + qrow) * ldq + q_col0 + h * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 v = qp[i];
      q[4 * i] = v.x * scale; q[4 * i + 1] = v.y * scale; q[4 * i + 2] = v.z * scale; q[4 * i + 3] = v.w * scale;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) q[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  float m = -INFINITY, l = 0.f;
  for (int src = 0; src < 2; ++src) {
    const float* KV = src == 0 ? KV1 : KV2;
    const int NK = src == 0 ? NK1 : NK2, ld = src == 0 ? ld1 : ld2;
    const int kc = (src == 0 ? k1_col0 : k2_col0) + h * 64, vc = (src == 0 ? v1_col0 : v2_col0) + h * 64;
    for (int k0 = 0; k0 < NK; k0 += AF_KT) {
      const int nk = min(AF_KT, NK - k0);
      __syncthreads();
      for (int i = threadIdx.x; i < AF_KT * 16; i += AF_ROWS) {      // float4 granules: key = i / 16, dims 4*(i % 16)
        const int kk = i >> 4, d4 = (i & 15) * 4;
        float4 kv4 = make_float4(0.f, 0.f, 0.f, 0.f), vv4 = kv4;
        if (kk < nk) {
          const float* rowp = KV + (size_t(b) * NK + k0 + kk) * ld;
          kv4 = *reinterpret_cast<const float4*>(rowp + kc + d4);
          vv4 = *reinterpret_cast<const float4*>(rowp + vc + d4);
        }
        *reinterpret_cast<float4*>(&sK[kk][d4]) = kv4;
        *reinterpret_cast<float4*>(&sV[kk][d4]) = vv4;
      }
      __syncthreads();
      float s[AF_KT];
      float mt = m;
#pragma unroll 4
      for (int kk = 0; kk < AF_KT; ++kk) {
        float d = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) d = fmaf(q[i], sK[kk][i], d);
        s[kk] = kk < nk ? d : -INFINITY;
        mt = fmaxf(mt, s[kk]);
      }
      const float alpha = m == -INFINITY ? 0.f : expf(m - mt);
      l *= alpha;
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] *= alpha;
      m = mt;
#pragma unroll 4
      for (int kk = 0; kk < AF_KT; ++kk) {
        const float pv = kk < nk ? expf(s[kk] - m) : 0.f;
        l += pv;
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = fmaf(pv, sV[kk][i], acc[i]);
      }
    }
  }
  if (ok) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    float4* op = reinterpret_cast<float4*>(O + (size_t(b) * NQ + qrow) * ldo + h * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      op[i] = make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv);
  }
}

}  // namespace
}  // namespace mebt

extern "C" {

int mebt_split_f32_bf16x3(const float* x, int ld, int rows, int K, void* out, int weight_side, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(rows >= 0 && K > 0 && K % 4 == 0 && ld % 4 == 0 && ld >= K, MEBT_ERR_SHAPE, "split_f32_bf16x3: bad shape");
  if (rows == 0) return MEBT_OK;
  const long long total = (long long)rows * (K / 4);
  const int blocks = int(std::min<long long>((total + 255) / 256, 148 * 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls(FAM_OTHER, double(rows) * K * 10.0, st);
    split_f32_bf16x3_kernel<<<blocks, 256, 0, st>>>(x, ld, rows, K, static_cast<__nv_bfloat16*>(out), weight_side ? 1 : 0);
  }
  MEBT_LAUNCH_OK("split_f32_bf16x3_kernel");
  return MEBT_OK;
}

int mebt_latent_attention_fwd_f32(const float* Q, int ldq, int q_col0, const float* KV1, int ld1, int k1_col0, int v1_col0,
                                  int NK1, const float* KV2, int ld2, int k2_col0, int v2_col0, int NK2, float* O, int ldo,
                                  int B, int H, int NQ, int head_dim, void* stream) {
  using namespace mebt;
  MEBT_REQUIRE(head_dim == 64, MEBT_ERR_UNSUPPORTED, "attention_f32: head_dim %d unsupported", head_dim);
  MEBT_REQUIRE(B > 0 && H > 0 && NQ > 0 && NK1 >= 0 && NK2 >= 0, MEBT_ERR_SHAPE, "attention_f32: bad shape");
  MEBT_REQUIRE(ldq % 4 == 0 && ldo % 4 == 0 && q_col0 % 4 == 0, MEBT_ERR_SHAPE, "attention_f32: Q/O must be 16B aligned");
  MEBT_REQUIRE(NK1 == 0 || (KV1 != nullptr && ld1 % 4 == 0 && k1_col0 % 4 == 0 && v1_col0 % 4 == 0), MEBT_ERR_SHAPE, "attention_f32: bad KV1");
  MEBT_REQUIRE(NK2 == 0 || (KV2 != nullptr && ld2 % 4 == 0 && k2_col0 % 4 == 0 && v2_col0 % 4 == 0), MEBT_ERR_SHAPE, "attention_f32: bad KV2");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    LaunchScope ls(FAM_ATTENTION, 4.0 * double(B) * H * double(NQ) * double(NK1 + NK2) * 64, st);
    attention_f32_kernel<<<dim3((NQ + AF_ROWS - 1) / AF_ROWS, H, B), AF_ROWS, 0, st>>>(
        Q, ldq, q_col0, KV1, ld1, k1_col0, v1_col0, NK1, KV2, ld2, k2_col0, v2_col0, NK2, O, ldo, NQ, 0.125f);
  }
  MEBT_LAUNCH_OK("attention_f32_kernel");
  return MEBT_OK;
}

}  // extern "C"
