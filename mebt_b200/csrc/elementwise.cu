// Memory-bound kernels of the MeBT hot path: K1 embedding stem, LayerNorm, K8 id scatter, K10 codebook
// row gather, dtype casts.  All are HBM-bound: one warp per row, 16-byte vectorised coalesced access,
// warp-shuffle reductions, no shared-memory staging needed (every byte is touched once).
#include "common.cuh"

namespace mebt {
namespace {

constexpr int ROWS_PER_BLOCK = 8;   // 8 warps, one row each

template <typename T> struct Vec4 {};   // 4 elements
template <> struct Vec4<float> {
  using type = float4;
  __device__ static float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  __device__ static void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<__nv_bfloat16> {
  using type = uint2;
  __device__ static float4 load(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  __device__ static void store(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// ---------------------------------------------------------------------------------------------
// K1: contexts = tok_emb[x[ctx_idx]] + pos_emb[ctx_idx]; targets = mask_emb + pos_emb[tgt_idx];
//     latents = sos_emb (broadcast over the batch).   reference: mebt/transformer.py:298-317
// ---------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void embed_gather_kernel(const int64_t* __restrict__ x, int x_stride, const int64_t* __restrict__ ctx_idx,
                                    int ctx_stride, const int64_t* __restrict__ tgt_idx, int tgt_stride,
                                    const float* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                    const float* __restrict__ mask_emb, const float* __restrict__ sos_emb,
                                    OutT* __restrict__ contexts, OutT* __restrict__ targets, OutT* __restrict__ latents,
                                    int B, int NC, int NT, int L, int D, int V, int n_pos, int* __restrict__ err_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + warp;
  const long long n_ctx = (long long)B * NC, n_tgt = (long long)B * NT, n_lat = (long long)B * L;
  if (row >= n_ctx + n_tgt + n_lat) return;
  const float* src_a;
  const float* src_b = nullptr;
  OutT* dst;
  if (row < n_ctx) {
    const int b = int(row / NC), i = int(row % NC);
    const long long pos = ctx_idx[(long long)b * ctx_stride + i];
    if (pos < 0 || pos >= n_pos) { if (lane == 0) atomicExch(err_flag, 1); return; }
    const long long tok = x[(long long)b * x_stride + pos];
    if (tok < 0 || tok >= V) { if (lane == 0) atomicExch(err_flag, 2); return; }
    src_a = tok_emb + tok * D;
    src_b = pos_emb + pos * D;
    dst = contexts + row * D;
  } else if (row < n_ctx + n_tgt) {
    const long long r = row - n_ctx;
    const int b = int(r / NT), i = int(r % NT);
    const long long pos = tgt_idx[(long long)b * tgt_stride + i];
    if (pos < 0 || pos >= n_pos) { if (lane == 0) atomicExch(err_flag, 1); return; }
    src_a = mask_emb;
    src_b = pos_emb + pos * D;
    dst = targets + r * D;
  } else {
    const long long r = row - n_ctx - n_tgt;
    src_a = sos_emb + (r % L) * D;
    dst = latents + r * D;
  }
  for (int c = lane * 4; c < D; c += 128) {
    float4 a = __ldg(reinterpret_cast<const float4*>(src_a + c));
    if (src_b != nullptr) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(src_b + c));
      // reference order: pos_emb row + token row (transformer.py:311-312); fp32 add is commutative
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    Vec4<OutT>::store(dst + c, a);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (eps inside the sqrt, biased variance) — nn.LayerNorm, gpt.py:147-148,216.
// Two-pass in registers (mean, then centred sum of squares) in fp32.
// ---------------------------------------------------------------------------------------------
template <typename InT, typename OutT, int MAX_VEC>   // MAX_VEC float4 groups per lane: D <= 128 * MAX_VEC
__global__ void layernorm_kernel(const InT* __restrict__ x, int ldx, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, OutT* __restrict__ y, int ldy, int rows, int D,
                                 float eps, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + warp;
  griddep_wait();
  if (row >= rows) return;
  const InT* xr = x + row * ldx;
  float4 v[MAX_VEC];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) {
      v[i] = Vec4<InT>::load(xr + c);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(sum) / float(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) {
      const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c2 * c2 + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / float(D) + eps);
  if (lane == 0) {
    if (mean_out != nullptr) mean_out[row] = mean;
    if (rstd_out != nullptr) rstd_out[row] = rstd;
  }
  OutT* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      Vec4<OutT>::store(yr + c, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K8: x[b, tgt_idx[b,i]] = ids[b,i]  (the sparse-COO write-back of transformer.py:413-439 ≡ scatter)
// ---------------------------------------------------------------------------------------------
__global__ void scatter_ids_kernel(int64_t* __restrict__ x, int x_stride, const int64_t* __restrict__ tgt_idx,
                                   int tgt_stride, const int64_t* __restrict__ ids, int B, int NT, int N,
                                   int* __restrict__ err_flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * NT) return;
  const int b = int(i / NT), j = int(i % NT);
  const long long pos = tgt_idx[(long long)b * tgt_stride + j];
  if (pos < 0 || pos >= N) { atomicExch(err_flag, 1); return; }
  x[(long long)b * x_stride + pos] = ids[i];
}

// ---------------------------------------------------------------------------------------------
// K10: codebook row gather, F.embedding (codebook.py:61, vqgan.py:91), optional channel-first output
//      (fusing shift_dim(h, -1, 1) of codebook.py:62 / vqgan.py:92).
// ---------------------------------------------------------------------------------------------
__global__ void row_gather_kernel(const int64_t* __restrict__ enc, const float* __restrict__ E, float* __restrict__ out,
                                  long long M, int C, int K, int* __restrict__ err_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * ROWS_PER_BLOCK + warp;
  if (row >= M) return;
  const long long code = enc[row];
  if (code < 0 || code >= K) { if (lane == 0) atomicExch(err_flag, 1); return; }
  const float4* src = reinterpret_cast<const float4*>(E + code * C);
  float4* dst = reinterpret_cast<float4*>(out + row * C);
  for (int c = lane; c < C / 4; c += 32) dst[c] = __ldg(src + c);
}

// out[b, c, s] = E[enc[b, s], c];  32x32 tile transposed through shared memory so both sides coalesce
__global__ void row_gather_cf_kernel(const int64_t* __restrict__ enc, const float* __restrict__ E,
                                     float* __restrict__ out, int S, int C, int K, int* __restrict__ err_flag) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int s = s0 + r;
    float v = 0.f;
    if (s < S && c0 + tx < C) {
      const long long code = enc[(long long)b * S + s];
      if (code < 0 || code >= K) atomicExch(err_flag, 1);
      else v = __ldg(E + code * C + c0 + tx);
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, s = s0 + tx;
    if (c < C && s < S) out[((long long)b * C + c) * S + s] = tile[tx][r];
  }
}

// fp32 -> bf16 cast (weights master copy -> tensor-core operand)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    Vec4<__nv_bfloat16>::store(out + i * 4, v);
  }
}

int* device_err_flag() {
  static int* flag = nullptr;
  if (flag == nullptr) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

// bf16 -> bf16 fast path for D = 256 * NV: 16-byte loads (8 elements per lane per load) and TWO rows per warp with
// all loads of both rows issued before the first reduction, which doubles the bytes each warp keeps in flight.
template <int NV>
__global__ void __launch_bounds__(256, 3) layernorm_bf16x2_kernel(const __nv_bfloat16* __restrict__ x, int ldx,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ y, int ldy, int rows,
                                                               float eps, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out) {
  constexpr int D = 256 * NV;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * ROWS_PER_BLOCK + warp) * 2;
  griddep_wait();
  if (row0 >= rows) return;
  const bool two = row0 + 1 < rows;
  uint4 raw[2][NV];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < NV; ++i)
      raw[r][i] = (r == 0 || two) ? *reinterpret_cast<const uint4*>(x + (row0 + r) * ldx + (lane + 32 * i) * 8)
                                  : make_uint4(0, 0, 0, 0);
  // the rows stay PACKED in registers (16 per row at D = 1024) and are unpacked on the fly in each of the three passes:
  // half the registers of an fp32 copy, so that twice as many rows are in flight per SM
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        sum[r] += f.x + f.y;
      }
    }
  float mu[2], rs[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) mu[r] = warp_sum(sum[r]) * (1.0f / D);
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        const float d0 = f.x - mu[r], d1 = f.y - mu[r];
        sq += d0 * d0;
        sq += d1 * d1;
      }
    }
    rs[r] = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
  }
  if (lane == 0) {
    if (mean_out != nullptr) { mean_out[row0] = mu[0]; if (two) mean_out[row0 + 1] = mu[1]; }
    if (rstd_out != nullptr) { rstd_out[row0] = rs[0]; if (two) rstd_out[row0 + 1] = rs[1]; }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !two) break;
      const uint32_t w[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(w[k]);
        o[2 * k] = (f.x - mu[r]) * rs[r] * gg[2 * k] + bb[2 * k];
        o[2 * k + 1] = (f.y - mu[r]) * rs[r] * gg[2 * k + 1] + bb[2 * k + 1];
      }
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(y + (row0 + r) * ldy + c) = u;
    }
  }
}

template <typename InT, typename OutT>
int launch_ln(const void* x, int ldx, const float* g, const float* b, void* y, int ldy, int rows, int D, float eps,
              float* mean, float* rstd, cudaStream_t st) {
  const int grid = (rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK;
  LaunchScope ls(FAM_LAYERNORM, double(rows) * D * double(sizeof(InT) + sizeof(OutT)), st);
  if constexpr (sizeof(InT) == 2 && sizeof(OutT) == 2) {
    if (D % 256 == 0 && D <= 1024 && ldx % 8 == 0 && ldy % 8 == 0) {
      const int g2 = (rows + 2 * ROWS_PER_BLOCK - 1) / (2 * ROWS_PER_BLOCK);
      const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
      __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
      switch (D / 256) {
        case 1: MEBT_CUDA_OK(launch_pdl(layernorm_bf16x2_kernel<1>, dim3(g2), dim3(256), 0, st, xb, ldx, g, b, yb, ldy, rows, eps, mean, rstd)); break;
        case 2: MEBT_CUDA_OK(launch_pdl(layernorm_bf16x2_kernel<2>, dim3(g2), dim3(256), 0, st, xb, ldx, g, b, yb, ldy, rows, eps, mean, rstd)); break;
        case 3: MEBT_CUDA_OK(launch_pdl(layernorm_bf16x2_kernel<3>, dim3(g2), dim3(256), 0, st, xb, ldx, g, b, yb, ldy, rows, eps, mean, rstd)); break;
        default: MEBT_CUDA_OK(launch_pdl(layernorm_bf16x2_kernel<4>, dim3(g2), dim3(256), 0, st, xb, ldx, g, b, yb, ldy, rows, eps, mean, rstd)); break;
      }
      MEBT_LAUNCH_OK("layernorm_bf16x2_kernel");
      return MEBT_OK;
    }
  }
  const InT* xi = static_cast<const InT*>(x);
  OutT* yo = static_cast<OutT*>(y);
  if (D <= 256) MEBT_CUDA_OK(launch_pdl(layernorm_kernel<InT, OutT, 2>, dim3(grid), dim3(256), 0, st, xi, ldx, g, b, yo, ldy, rows, D, eps, mean, rstd));
  else if (D <= 512) MEBT_CUDA_OK(launch_pdl(layernorm_kernel<InT, OutT, 4>, dim3(grid), dim3(256), 0, st, xi, ldx, g, b, yo, ldy, rows, D, eps, mean, rstd));
  else if (D <= 1024) MEBT_CUDA_OK(launch_pdl(layernorm_kernel<InT, OutT, 8>, dim3(grid), dim3(256), 0, st, xi, ldx, g, b, yo, ldy, rows, D, eps, mean, rstd));
  else MEBT_CUDA_OK(launch_pdl(layernorm_kernel<InT, OutT, 16>, dim3(grid), dim3(256), 0, st, xi, ldx, g, b, yo, ldy, rows, D, eps, mean, rstd));
  MEBT_LAUNCH_OK("layernorm_kernel");
  return MEBT_OK;
}

}  // namespace

int* err_flag_ptr() { return device_err_flag(); }

int embed_gather(const int64_t* x, int x_stride, const int64_t* ctx_idx, int ctx_stride, const int64_t* tgt_idx,
                 int tgt_stride, const float* tok_emb, const float* pos_emb, const float* mask_emb,
                 const float* sos_emb, void* contexts, void* targets, void* latents, int B, int NC, int NT, int L, int D,
                 int V, int n_pos, int out_dtype, cudaStream_t st) {
  MEBT_REQUIRE(B > 0 && NC >= 0 && NT >= 0 && L >= 0 && D > 0 && D % 4 == 0, MEBT_ERR_SHAPE,
               "embed_gather: bad shape B=%d NC=%d NT=%d L=%d D=%d", B, NC, NT, L, D);
  int* flag = device_err_flag();
  MEBT_REQUIRE(flag != nullptr, MEBT_ERR_CUDA, "embed_gather: cannot allocate error flag");
  const long long rows = (long long)B * (NC + NT + L);
  if (rows == 0) return MEBT_OK;
  const int grid = int((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK);
  const double e = out_dtype == MEBT_DTYPE_BF16 ? 2.0 : 4.0;
  LaunchScope ls(FAM_EMBED, double(B) * (NC * (16.0 + 8.0 * D + e * D) + NT * (8.0 + 4.0 * D + e * D) + L * (4.0 + e) * D), st);
  if (out_dtype == MEBT_DTYPE_BF16)
    embed_gather_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        x, x_stride, ctx_idx, ctx_stride, tgt_idx, tgt_stride, tok_emb, pos_emb, mask_emb, sos_emb,
        static_cast<__nv_bfloat16*>(contexts), static_cast<__nv_bfloat16*>(targets),
        static_cast<__nv_bfloat16*>(latents), B, NC, NT, L, D, V, n_pos, flag);
  else if (out_dtype == MEBT_DTYPE_FP32)
    embed_gather_kernel<float><<<grid, 256, 0, st>>>(x, x_stride, ctx_idx, ctx_stride, tgt_idx, tgt_stride, tok_emb,
                                                      pos_emb, mask_emb, sos_emb, static_cast<float*>(contexts),
                                                      static_cast<float*>(targets), static_cast<float*>(latents), B,
                                                      NC, NT, L, D, V, n_pos, flag);
  else
    MEBT_REQUIRE(false, MEBT_ERR_DTYPE, "embed_gather: unsupported out dtype %d", out_dtype);
  MEBT_LAUNCH_OK("embed_gather_kernel");
  return MEBT_OK;
}

int layernorm(const void* x, int ldx, int in_dtype, const float* gamma, const float* beta, void* y, int ldy,
              int out_dtype, int rows, int D, float eps, float* mean, float* rstd, cudaStream_t st) {
  MEBT_REQUIRE(rows >= 0 && D > 0 && D % 4 == 0 && D <= 2048, MEBT_ERR_SHAPE, "layernorm: bad shape rows=%d D=%d",
               rows, D);
  MEBT_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, MEBT_ERR_SHAPE, "layernorm: row strides must be multiples of 4");
  if (rows == 0) return MEBT_OK;
  if (in_dtype == MEBT_DTYPE_BF16 && out_dtype == MEBT_DTYPE_BF16)
    return launch_ln<__nv_bfloat16, __nv_bfloat16>(x, ldx, gamma, beta, y, ldy, rows, D, eps, mean, rstd, st);
  if (in_dtype == MEBT_DTYPE_FP32 && out_dtype == MEBT_DTYPE_BF16)
    return launch_ln<float, __nv_bfloat16>(x, ldx, gamma, beta, y, ldy, rows, D, eps, mean, rstd, st);
  if (in_dtype == MEBT_DTYPE_FP32 && out_dtype == MEBT_DTYPE_FP32)
    return launch_ln<float, float>(x, ldx, gamma, beta, y, ldy, rows, D, eps, mean, rstd, st);
  if (in_dtype == MEBT_DTYPE_BF16 && out_dtype == MEBT_DTYPE_FP32)
    return launch_ln<__nv_bfloat16, float>(x, ldx, gamma, beta, y, ldy, rows, D, eps, mean, rstd, st);
  MEBT_REQUIRE(false, MEBT_ERR_DTYPE, "layernorm: unsupported dtypes %d -> %d", in_dtype, out_dtype);
  return MEBT_OK;
}

// ---- AdamW over the flat parameter buffer, fused with the bf16 operand refresh ----
// torch.optim.AdamW(fused=True) arithmetic (decoupled decay, bias-corrected moments) on fp32 masters p, moments m / v and
// gradients g, all [n]; additionally writes bf16(p) for the tensor-core operands, which saves the separate cast pass.
// decay[i >> shift] != 0 marks the blocks of 2^shift elements that belong to a weight-decayed tensor
// (configure_optimizers, mebt/transformer.py:749-798: only the transformer's Linear weights decay).
__global__ void __launch_bounds__(256) adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, __nv_bfloat16* __restrict__ p16,
                                                         const unsigned char* __restrict__ decay, int shift, long long n4,
                                                         const AdamScalars a) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const unsigned char flag = decay[(i * 4) >> shift];
    if (flag & 2) continue;                  // updated by the fused weight-gradient epilogue (mebt_stack_backward_fused)
    float4 pv = reinterpret_cast<const float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<const float4*>(m)[i];
    float4 vv = reinterpret_cast<const float4*>(v)[i];
    const float keep = (flag & 1) ? 1.f - a.lr * a.wd : 1.f;
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gg = reinterpret_cast<const float*>(&gv);
    float* mm = reinterpret_cast<float*>(&mv);
    float* vq = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) adamw_element(pp[k], gg[k], mm[k], vq[k], keep, a);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    uint2 o;
    o.x = pack_bf16x2(pp[0], pp[1]);
    o.y = pack_bf16x2(pp[2], pp[3]);
    reinterpret_cast<uint2*>(p16)[i] = o;
  }
}

AdamScalars make_adam_scalars(float lr, float beta1, float beta2, float eps, float wd, int step) {
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  return AdamScalars{lr, beta1, beta2, eps, wd, float(double(lr) / bc1), float(1.0 / sqrt(bc2))};
}

int adamw_flat(float* p, const float* g, float* m, float* v, void* p16, const unsigned char* decay, int shift, long long n,
               float lr, float beta1, float beta2, float eps, float wd, int step, int max_ctas, cudaStream_t st) {
  MEBT_REQUIRE(n >= 0 && n % 4 == 0 && shift >= 2 && step >= 1, MEBT_ERR_SHAPE, "adamw: n %% 4 != 0, shift < 2 or step < 1");
  if (n == 0) return MEBT_OK;
  {
    LaunchScope ls(FAM_OTHER, double(n) * 30.0, st);
    // max_ctas > 0: a background update - few CTAs trickle through the buffers while latency-bound kernels of another
    // stream (the rest of backward) keep the SMs; the full grid saturates HBM and is for an update nothing overlaps
    // max_ctas < 0: a background update of SHORT-LIVED CTAs (-max_ctas 16-byte groups per thread): the grid is as large as
    // the range, so SM slots keep freeing up for the kernels of a higher-priority stream that arrive meanwhile
    int grid = max_ctas > 0 && max_ctas < 148 * 8 ? max_ctas : 148 * 8;
    if (max_ctas < 0) {
      const long long per_cta = 256LL * std::min(-max_ctas, 16);
      grid = int(std::min<long long>((n / 4 + per_cta - 1) / per_cta, 1 << 30));
    }
    adamw_flat_kernel<<<grid, 256, 0, st>>>(p, g, m, v, static_cast<__nv_bfloat16*>(p16), decay, shift, n / 4,
                                               make_adam_scalars(lr, beta1, beta2, eps, wd, step));
  }
  MEBT_LAUNCH_OK("adamw_flat_kernel");
  return MEBT_OK;
}

// ---- dropout on [rows, D] bf16 activations (training) ----
// forward : y = resid + x * keep / (1-p)   (resid optional; in place when y == x)
// backward: dx = dy * keep / (1-p)
// 8 elements (16 bytes) per thread; the keep decisions are regenerated from (seed, site, row, column).
__global__ void dropout_rows_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const __nv_bfloat16* __restrict__ resid,
                                    int ldres, __nv_bfloat16* __restrict__ y, int ldy, int rows, int D, DropKey key) {
  const int per_row = D >> 3;
  const long long total = (long long)rows * per_row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = int(i / per_row), c8 = int(i - (long long)row * per_row) * 8;
    const uint32_t rk = drop_row_key(key, uint32_t(row));
    uint4 v = *reinterpret_cast<const uint4*>(x + (size_t)row * ldx + c8);
    uint4 r = make_uint4(0, 0, 0, 0);
    if (resid != nullptr) r = *reinterpret_cast<const uint4*>(resid + (size_t)row * ldres + c8);
    uint32_t* vv = reinterpret_cast<uint32_t*>(&v);
    const uint32_t* rr = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f0, f1;
      drop_pair(key, rk, uint32_t(c8 / 2 + j), f0, f1);
      const float2 a = unpack_bf16x2(vv[j]);
      const float2 b = unpack_bf16x2(rr[j]);
      vv[j] = pack_bf16x2(fmaf(a.x, f0, b.x), fmaf(a.y, f1, b.y));
    }
    *reinterpret_cast<uint4*>(y + (size_t)row * ldy + c8) = v;
  }
}

int dropout_rows(const void* x, int ldx, const void* resid, int ldres, void* y, int ldy, int rows, int D, float p,
                 unsigned long long seed, unsigned long long site, cudaStream_t st) {
  MEBT_REQUIRE(rows >= 0 && D > 0 && D % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (resid == nullptr || ldres % 8 == 0),
               MEBT_ERR_SHAPE, "dropout: D and the row strides must be multiples of 8");
  MEBT_REQUIRE(p >= 0.f && p < 1.f, MEBT_ERR_SHAPE, "dropout: p = %f outside [0, 1)", p);
  if (rows == 0) return MEBT_OK;
  const DropKey key = make_drop_key(p, seed, site);
  const long long total = (long long)rows * (D / 8);
  const int blocks = int(std::min<long long>((total + 255) / 256, 148 * 8));
  {
    LaunchScope ls(FAM_OTHER, 0.0, st);
    dropout_rows_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx,
                                                static_cast<const __nv_bfloat16*>(resid), ldres,
                                                static_cast<__nv_bfloat16*>(y), ldy, rows, D, key);
  }
  MEBT_LAUNCH_OK("dropout_rows_kernel");
  return MEBT_OK;
}

}  // namespace mebt

extern "C" {

int mebt_adamw_flat(float* p, const float* g, float* m, float* v, void* p_bf16, const unsigned char* decay_blocks,
                    int block_shift, long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    void* stream) {
  return mebt::adamw_flat(p, g, m, v, p_bf16, decay_blocks, block_shift, n, lr, beta1, beta2, eps, weight_decay, step, 0,
                          static_cast<cudaStream_t>(stream));
}

int mebt_adamw_flat_bg(float* p, const float* g, float* m, float* v, void* p_bf16, const unsigned char* decay_blocks,
                       int block_shift, long long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                       int step, int max_ctas, void* stream) {
  return mebt::adamw_flat(p, g, m, v, p_bf16, decay_blocks, block_shift, n, lr, beta1, beta2, eps, weight_decay, step,
                          max_ctas, static_cast<cudaStream_t>(stream));
}

int mebt_dropout_rows(const void* x, int ldx, const void* resid, int ldres, void* y, int ldy, int rows, int D, float p,
                      unsigned long long seed, unsigned long long site, void* stream) {
  return mebt::dropout_rows(x, ldx, resid, ldres, y, ldy, rows, D, p, seed, site, static_cast<cudaStream_t>(stream));
}

int mebt_embed_gather(const int64_t* x_indices, int x_stride, const int64_t* ctx_idx, int ctx_stride,
                      const int64_t* tgt_idx, int tgt_stride, const float* tok_emb, const float* pos_emb,
                      const float* mask_emb, const float* sos_emb, void* contexts, void* targets, void* latents, int B,
                      int NC, int NT, int L, int D, int V, int n_pos, int out_dtype, void* stream) {
  return mebt::embed_gather(x_indices, x_stride, ctx_idx, ctx_stride, tgt_idx, tgt_stride, tok_emb, pos_emb, mask_emb,
                            sos_emb, contexts, targets, latents, B, NC, NT, L, D, V, n_pos, out_dtype,
                            static_cast<cudaStream_t>(stream));
}

int mebt_layernorm(const void* x, int ldx, int in_dtype, const float* gamma, const float* beta, void* y, int ldy,
                   int out_dtype, int rows, int D, float eps, float* mean_out, float* rstd_out, void* stream) {
  return mebt::layernorm(x, ldx, in_dtype, gamma, beta, y, ldy, out_dtype, rows, D, eps, mean_out, rstd_out,
                         static_cast<cudaStream_t>(stream));
}

int mebt_scatter_ids(int64_t* x, int x_stride, const int64_t* tgt_idx, int tgt_stride, const int64_t* ids, int B, int NT,
                     int N, void* stream) {
  MEBT_REQUIRE(B >= 0 && NT >= 0 && N > 0, MEBT_ERR_SHAPE, "scatter_ids: bad shape");
  const long long n = (long long)B * NT;
  if (n == 0) return MEBT_OK;
  int* flag = mebt::err_flag_ptr();
  MEBT_REQUIRE(flag != nullptr, MEBT_ERR_CUDA, "scatter_ids: cannot allocate error flag");
  mebt::LaunchScope ls(mebt::FAM_SCATTER, double(n) * 24.0, static_cast<cudaStream_t>(stream));
  mebt::scatter_ids_kernel<<<int((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, x_stride, tgt_idx, tgt_stride, ids, B, NT, N, flag);
  MEBT_LAUNCH_OK("scatter_ids_kernel");
  return MEBT_OK;
}

int mebt_row_gather(const int64_t* enc, const float* E, float* out, int batch, int S, int C, int K, int channel_first,
                    void* stream) {
  MEBT_REQUIRE(batch >= 0 && S >= 0 && C > 0 && K > 0, MEBT_ERR_SHAPE, "row_gather: bad shape");
  const long long M = (long long)batch * S;
  if (M == 0) return MEBT_OK;
  int* flag = mebt::err_flag_ptr();
  MEBT_REQUIRE(flag != nullptr, MEBT_ERR_CUDA, "row_gather: cannot allocate error flag");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mebt::LaunchScope ls(mebt::FAM_VQ, double(M) * (8.0 + 8.0 * C), st);
  if (channel_first) {
    dim3 grid((S + 31) / 32, (C + 31) / 32, batch), block(32, 8);
    mebt::row_gather_cf_kernel<<<grid, block, 0, st>>>(enc, E, out, S, C, K, flag);
  } else {
    MEBT_REQUIRE(C % 4 == 0, MEBT_ERR_SHAPE, "row_gather: C must be a multiple of 4");
    mebt::row_gather_kernel<<<int((M + mebt::ROWS_PER_BLOCK - 1) / mebt::ROWS_PER_BLOCK), 256, 0, st>>>(enc, E, out, M,
                                                                                                          C, K, flag);
  }
  MEBT_LAUNCH_OK("row_gather_kernel");
  return MEBT_OK;
}

int mebt_cast_f32_to_bf16(const float* in, void* out, long long n, void* stream) {
  MEBT_REQUIRE(n >= 0 && n % 4 == 0, MEBT_ERR_SHAPE, "cast: n must be a multiple of 4");
  if (n == 0) return MEBT_OK;
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mebt::LaunchScope ls(mebt::FAM_OTHER, double(n) * 6.0, static_cast<cudaStream_t>(stream));
  mebt::cast_f32_bf16_kernel<<<int(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, static_cast<__nv_bfloat16*>(out), n4);
  MEBT_LAUNCH_OK("cast_f32_bf16_kernel");
  return MEBT_OK;
}

// Index-range violations are recorded on the device (kernels never read out of bounds); this reads and
// clears the flag.  It synchronises the stream, so call it from tests / debug paths only.
int mebt_check_index_errors(void* stream) {
  int* flag = mebt::err_flag_ptr();
  MEBT_REQUIRE(flag != nullptr, MEBT_ERR_CUDA, "cannot allocate error flag");
  int h = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MEBT_CUDA_OK(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  MEBT_CUDA_OK(cudaStreamSynchronize(st));
  if (h != 0) {
    cudaMemsetAsync(flag, 0, sizeof(int), st);
    mebt::set_last_error("index out of range detected on device (kind %d)", h);
    return MEBT_ERR_SHAPE;
  }
  return MEBT_OK;
}

}  // extern "C"
