// Memory-bound backward kernels of the MeBT training step (what torch autograd runs for the reference's
// nn.LayerNorm / nn.Linear bias / nn.Embedding / gather on the path of mebt/transformer.py:216-286 and
// mebt/modules/gpt.py:159-253):
//   column sums (bias gradients), LayerNorm backward (dx + per-CTA dgamma/dbeta partials), the partial
//   reducer, and the stem's scatter-add into tok_emb / pos_emb / mask_emb / sos_emb gradients.
// Reductions over rows are two-stage with a fixed order (deterministic); only the embedding scatter-add uses
// fp32 atomics (several tokens of a batch may hit the same row), like torch's embedding backward.
#include <map>
#include <mutex>

#include "common.cuh"

namespace mebt {
namespace {

__device__ __forceinline__ float4 ld_bf16x4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ld_round_bf16x4(float4 v) {      // v rounded to bf16 and back
  const float2 a = unpack_bf16x2(pack_bf16x2(v.x, v.y)), b = unpack_bf16x2(pack_bf16x2(v.z, v.w));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_bf16x4(__nv_bfloat16* p, float4 v) {
  uint2 u;
  u.x = pack_bf16x2(v.x, v.y);
  u.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---- column sums: out[n] (+)= sum_r X[r, n] ----------------------------------------------------------------
// grid (N / 128, slabs); 256 threads = 32 column-quads x 8 row lanes.  Each CTA writes the partial sums of its row slab;
// the LAST CTA of a column block to finish (ticket counter) adds the slabs in slab order, so the result does not depend
// on which CTA that is (bitwise reproducible) and no second launch is needed.
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ X, int ld, int rows, int N, int rows_per_slab,
                              float* __restrict__ partial, float* __restrict__ out, int accumulate, int* __restrict__ counters) {
  __shared__ float4 red[8][32];
  __shared__ int s_last;
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  griddep_wait();
  const int col = blockIdx.x * 128 + cq * 4;
  const int r0 = blockIdx.y * rows_per_slab;
  const int r1 = min(rows, r0 + rows_per_slab);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N)
    for (int r = r0 + rl; r < r1; r += 8) {
      const float4 v = ld_bf16x4(X + size_t(r) * ld + col);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  red[rl][cq] = acc;
  __syncthreads();
  if (rl == 0 && col < N) {
    for (int i = 1; i < 8; ++i) {
      const float4 v = red[i][cq];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    __stcg(reinterpret_cast<float4*>(partial + size_t(blockIdx.y) * N + col), acc);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(counters + blockIdx.x, 1);
    s_last = ticket == int(gridDim.y) - 1;
    if (s_last) counters[blockIdx.x] = 0;            // ready for the next launch on this stream
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // slab sums: row lane rl adds slabs rl, rl+8, ... in order, then lane 0 adds the eight lane sums in order
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < N)
    for (int sl = rl; sl < int(gridDim.y); sl += 8) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + size_t(sl) * N + col));
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
  red[rl][cq] = t;
  __syncthreads();
  if (rl == 0 && col < N) {
    for (int i = 1; i < 8; ++i) {
      const float4 v = red[i][cq];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>(out + col);
    if (accumulate) { const float4 v = *o; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    *o = t;
  }
}

// Ticket counters of the fused two-stage reductions: one zero-initialised block per stream that uses them (the kernels
// of one stream are serialised; the training backward runs its bias-gradient sums on a second stream).
int* reduce_counters(cudaStream_t st) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, int*> table;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto it = table.find({dev, st});
  if (it != table.end()) return it->second;
  int* p = nullptr;
  if (cudaMalloc(&p, 1024 * sizeof(int)) != cudaSuccess) return nullptr;
  // zeroed on the stream that will use them: ordered before its first reduction, no device-wide synchronisation
  if (cudaMemsetAsync(p, 0, 1024 * sizeof(int), st) != cudaSuccess) { cudaFree(p); return nullptr; }
  table[{dev, st}] = p;
  return p;
}

// ---- LayerNorm backward --------------------------------------------------------------------------------------
// y = (x - mean) * rstd * gamma + beta
// dx = rstd * (g - mean_D(g) - xhat * mean_D(g * xhat)),  g = dy * gamma;   dgamma = sum_rows dy * xhat; dbeta = sum_rows dy
// Two kernels with opposite parallelism: dx is row-parallel (one warp per row), the parameter gradients are column sums
// (128 columns x a slab of rows per CTA, finished by the last CTA of a column block in slab order: reproducible).
// One row-parallel LayerNorm-backward problem: dx = (resid ? resid : 0) + ln'(dy [+ dy_add]) over `rows` rows.
struct LnDxProblem {
  const __nv_bfloat16* dy;
  const __nv_bfloat16* dy_add;     // optional second addend of the incoming gradient (a residual branch), or nullptr
  const __nv_bfloat16* x;
  const float* mean;
  const float* rstd;
  __nv_bfloat16* dx;
  const __nv_bfloat16* resid;      // may alias dx
  __nv_bfloat16* dx_drop;          // optional masked copy (see below)
  DropKey drop;
  int rows;
};

// Up to two problems that share gamma (ln1 applied to the query stream and to the key stream of one Block,
// gpt.py:180-181) in ONE launch: CTAs [0, blocks0) take problem 0, the rest problem 1.
template <int MAX_VEC>
__global__ void __launch_bounds__(256) layernorm_bwd_dx_kernel(const LnDxProblem p0, const LnDxProblem p1, int blocks0,
                                                               const float* __restrict__ gamma, int D) {
  // dx_drop (optional): a second copy of the result multiplied by the keep factors of `drop` - the gradient w.r.t. the
  // pre-dropout output of the Linear that produced this stream (proj / mlp.2), which its dgrad and wgrad GEMMs consume;
  // saves the separate dropout_rows launch between this kernel and those GEMMs.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool second = int(blockIdx.x) >= blocks0;
  const LnDxProblem& q = second ? p1 : p0;
  // gamma is a parameter, not a product of the preceding kernels: fetch it before the dependency wait
  float4 gm[MAX_VEC];
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = (lane + 32 * i) * 4;
    gm[i] = c < D ? __ldg(reinterpret_cast<const float4*>(gamma + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  griddep_wait();
  const float invD = 1.0f / float(D);
  const long long row = (long long)(int(blockIdx.x) - (second ? blocks0 : 0)) * 8 + warp;
  if (row >= q.rows) return;
  const __nv_bfloat16* dy = q.dy;
  const __nv_bfloat16* x = q.x;
  const __nv_bfloat16* resid = q.resid;
  const float mu = q.mean[row], rs = q.rstd[row];
  float4 xh[MAX_VEC], g[MAX_VEC], old[MAX_VEC];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) {
      const float4 xv = ld_bf16x4(x + row * D + c);
      float4 dv = ld_bf16x4(dy + row * D + c);
      if (q.dy_add != nullptr) {
        const float4 d2 = ld_bf16x4(q.dy_add + row * D + c);
        dv.x += d2.x; dv.y += d2.y; dv.z += d2.z; dv.w += d2.w;
      }
      if (resid != nullptr) old[i] = ld_bf16x4(resid + row * D + c);      // dx = resid + ln'(dy); resid may alias dx
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
  }
  s1 = warp_sum(s1) * invD;
  s2 = warp_sum(s2) * invD;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (c < D) {
      float4 o;
      o.x = rs * (g[i].x - s1 - xh[i].x * s2);
      o.y = rs * (g[i].y - s1 - xh[i].y * s2);
      o.z = rs * (g[i].z - s1 - xh[i].z * s2);
      o.w = rs * (g[i].w - s1 - xh[i].w * s2);
      if (resid != nullptr) { o.x += old[i].x; o.y += old[i].y; o.z += old[i].z; o.w += old[i].w; }
      st_bf16x4(q.dx + row * D + c, o);
      if (q.dx_drop != nullptr) {
        // the standalone kernel masks the bf16-rounded gradient: round first so that both forms agree bit for bit
        const float4 r = ld_round_bf16x4(o);
        const uint32_t rk = drop_row_key(q.drop, uint32_t(row));
        float f0, f1, f2, f3;
        drop_pair(q.drop, rk, uint32_t(c >> 1), f0, f1);
        drop_pair(q.drop, rk, uint32_t(c >> 1) + 1u, f2, f3);
        st_bf16x4(q.dx_drop + row * D + c, make_float4(r.x * f0, r.y * f1, r.z * f2, r.w * f3));
      }
    }
  }
}

// The production width (D = 1024) as its own kernel: the generic one holds xhat, g, the residual and gamma of a row in
// fp32 (123 registers, two 8-row CTAs per SM -> 3072 rows need a second wave); here a lane owns 4 x 8 consecutive columns,
// every access is 16 bytes, x and the residual stay packed, gamma is re-read (L1) in the second pass: four rows per
// 128-thread CTA, five CTAs per SM (740 of the 768 CTAs of 3072 rows resident at once).
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__global__ void __launch_bounds__(128, 5) layernorm_bwd_dx1024_kernel(const LnDxProblem p0, const LnDxProblem p1, int blocks0,
                                                                      const float* __restrict__ gamma) {
  constexpr int D = 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool second = int(blockIdx.x) >= blocks0;
  const LnDxProblem& q = second ? p1 : p0;
  griddep_wait();
  const long long row = (long long)(int(blockIdx.x) - (second ? blocks0 : 0)) * 4 + warp;
  if (row >= q.rows) return;
  const uint4* x4 = reinterpret_cast<const uint4*>(q.x + row * D);
  const uint4* dy4 = reinterpret_cast<const uint4*>(q.dy + row * D);
  const uint4* add4 = q.dy_add != nullptr ? reinterpret_cast<const uint4*>(q.dy_add + row * D) : nullptr;
  const uint4* res4 = q.resid != nullptr ? reinterpret_cast<const uint4*>(q.resid + row * D) : nullptr;
  const float4* gm4 = reinterpret_cast<const float4*>(gamma);
  uint4 xp[4], dp[4], ap[4], rp[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = lane + 32 * i;
    xp[i] = x4[v];
    dp[i] = dy4[v];
    if (add4 != nullptr) ap[i] = add4[v];
    if (res4 != nullptr) rp[i] = res4[v];       // dx = resid + ln'(dy); resid may alias dx
  }
  const float mu = q.mean[row], rs = q.rstd[row];
  float g[4][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = lane + 32 * i;
    float xv[8], dv[8];
    unpack8(xp[i], xv);
    unpack8(dp[i], dv);
    if (add4 != nullptr) {
      float av[8];
      unpack8(ap[i], av);
#pragma unroll
      for (int k = 0; k < 8; ++k) dv[k] += av[k];
    }
    const float4 ga = __ldg(gm4 + 2 * v), gb = __ldg(gm4 + 2 * v + 1);
    const float gmv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g[i][k] = dv[k] * gmv[k];
      s1 += g[i][k];
      s2 = fmaf(g[i][k], (xv[k] - mu) * rs, s2);
    }
  }
  s1 = warp_sum(s1) * (1.0f / D);
  s2 = warp_sum(s2) * (1.0f / D);
  const uint32_t rk = drop_row_key(q.drop, uint32_t(row));
  uint4* dx4 = reinterpret_cast<uint4*>(q.dx + row * D);
  uint4* dd4 = q.dx_drop != nullptr ? reinterpret_cast<uint4*>(q.dx_drop + row * D) : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = lane + 32 * i;
    float xv[8], o[8];
    unpack8(xp[i], xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = rs * (g[i][k] - s1 - (xv[k] - mu) * rs * s2);
    if (res4 != nullptr) {
      float rv[8];
      unpack8(rp[i], rv);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += rv[k];
    }
    const uint4 out = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    dx4[v] = out;
    if (dd4 != nullptr) {
      // the standalone kernel masks the bf16-rounded gradient: mask the rounded values so that both forms agree bit for bit
      float r[8];
      unpack8(out, r);
      float f[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) drop_pair(q.drop, rk, uint32_t(v * 4 + k), f[2 * k], f[2 * k + 1]);
      dd4[v] = make_uint4(pack_bf16x2(r[0] * f[0], r[1] * f[1]), pack_bf16x2(r[2] * f[2], r[3] * f[3]),
                          pack_bf16x2(r[4] * f[4], r[5] * f[5]), pack_bf16x2(r[6] * f[6], r[7] * f[7]));
    }
  }
}

// dgamma[c] (+)= sum_r dy[r,c] * xhat[r,c],  dbeta[c] (+)= sum_r dy[r,c].  grid (D / 128, slabs), 256 threads =
// 32 column-quads x 8 row lanes; partial: [slabs][2][D].
__global__ void __launch_bounds__(256) layernorm_bwd_param_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                       const float* __restrict__ mean, const float* __restrict__ rstd, int rows, int D,
                                       int rows_per_slab, float* __restrict__ partial, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int accumulate, int* __restrict__ counters) {
  __shared__ float4 red[2][8][32];
  __shared__ int s_last;
  const int cq = threadIdx.x & 31, rl = threadIdx.x >> 5;
  griddep_wait();
  const int col = blockIdx.x * 128 + cq * 4;
  const int r0 = blockIdx.y * rows_per_slab;
  const int r1 = min(rows, r0 + rows_per_slab);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < D)
    for (int r = r0 + rl; r < r1; r += 8) {
      const float4 dv = ld_bf16x4(dy + size_t(r) * D + col);
      const float4 xv = ld_bf16x4(x + size_t(r) * D + col);
      const float mu = mean[r], rs = rstd[r];
      ag.x += dv.x * ((xv.x - mu) * rs); ag.y += dv.y * ((xv.y - mu) * rs);
      ag.z += dv.z * ((xv.z - mu) * rs); ag.w += dv.w * ((xv.w - mu) * rs);
      ab.x += dv.x; ab.y += dv.y; ab.z += dv.z; ab.w += dv.w;
    }
  red[0][rl][cq] = ag;
  red[1][rl][cq] = ab;
  __syncthreads();
  if (rl < 2 && col < D) {                           // row lane 0 finishes dgamma, row lane 1 dbeta
    float4 t = red[rl][0][cq];
    for (int i = 1; i < 8; ++i) {
      const float4 v = red[rl][i][cq];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    __stcg(reinterpret_cast<float4*>(partial + (size_t(blockIdx.y) * 2 + rl) * D + col), t);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(counters + blockIdx.x, 1);
    s_last = ticket == int(gridDim.y) - 1;
    if (s_last) counters[blockIdx.x] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last CTA of the column block: row lanes 0-3 add the dgamma slabs (lane j: slabs j, j+4, ...), 4-7 the dbeta slabs
  const int which = rl >> 2, j = rl & 3;
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < D)
    for (int sl = j; sl < int(gridDim.y); sl += 4) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (size_t(sl) * 2 + which) * D + col));
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
  red[which][j][cq] = t;
  __syncthreads();
  if (j == 0 && col < D) {
    for (int i = 1; i < 4; ++i) {
      const float4 v = red[which][i][cq];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>((which == 0 ? dgamma : dbeta) + col);
    if (accumulate) { const float4 v = *o; t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w; }
    *o = t;
  }
}

// ---- stem backward ---------------------------------------------------------------------------------------------
// d_tok_emb[x[ctx_idx]] += d_ctx; d_pos_emb[ctx_idx] += d_ctx; d_pos_emb[tgt_idx] += d_tgt   (fp32 atomics)
__global__ void embed_scatter_add_kernel(const int64_t* __restrict__ x, int x_stride, const int64_t* __restrict__ ctx_idx,
                                         int ctx_stride, const int64_t* __restrict__ tgt_idx, int tgt_stride,
                                         const __nv_bfloat16* __restrict__ d_ctx, const __nv_bfloat16* __restrict__ d_tgt,
                                         float* __restrict__ d_tok, float* __restrict__ d_pos, int B, int NC, int NT,
                                         int D) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  const long long n_ctx = (long long)B * NC, n_tgt = (long long)B * NT;
  if (row >= n_ctx + n_tgt) return;
  const __nv_bfloat16* src;
  float* dst_tok = nullptr;
  float* dst_pos;
  if (row < n_ctx) {
    const int b = int(row / NC), i = int(row % NC);
    const long long pos = ctx_idx[(long long)b * ctx_stride + i];
    const long long tok = x[(long long)b * x_stride + pos];
    src = d_ctx + row * D;
    dst_tok = d_tok + tok * D;
    dst_pos = d_pos + pos * D;
  } else {
    const long long r = row - n_ctx;
    const int b = int(r / NT), i = int(r % NT);
    const long long pos = tgt_idx[(long long)b * tgt_stride + i];
    src = d_tgt + r * D;
    dst_pos = d_pos + pos * D;
  }
  for (int c = lane * 4; c < D; c += 128) {
    const float4 v = ld_bf16x4(src + c);
    atomicAdd(dst_pos + c + 0, v.x); atomicAdd(dst_pos + c + 1, v.y);
    atomicAdd(dst_pos + c + 2, v.z); atomicAdd(dst_pos + c + 3, v.w);
    if (dst_tok != nullptr) {
      atomicAdd(dst_tok + c + 0, v.x); atomicAdd(dst_tok + c + 1, v.y);
      atomicAdd(dst_tok + c + 2, v.z); atomicAdd(dst_tok + c + 3, v.w);
    }
  }
}

// d_sos[l, :] (+)= sum_b d_lat[b, l, :]   (fixed order over b)
__global__ void batch_sum_kernel(const __nv_bfloat16* __restrict__ d_lat, int B, long long per_batch,
                                 float* __restrict__ out, int accumulate) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= per_batch) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float4 v = ld_bf16x4(d_lat + b * per_batch + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float4* o = reinterpret_cast<float4*>(out + i);
  if (accumulate) { const float4 p = *o; acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w; }
  *o = acc;
}

// attention backward preprocess: delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ dO, int lddo, const __nv_bfloat16* __restrict__ O,
                                  int ldo, float* __restrict__ delta, int B, int H, int NQ) {
  // one warp per (b, q) row; lane pair handles one head (64 dims = 2 lanes x 32)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  griddep_wait();
  if (row >= (long long)B * NQ) return;
  const int b = int(row / NQ), q = int(row % NQ);
  for (int h0 = 0; h0 < H; h0 += 16) {
    const int h = h0 + (lane >> 1);
    float acc = 0.f;
    if (h < H) {
      const int c0 = h * 64 + (lane & 1) * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const float4 a = ld_bf16x4(dO + row * lddo + c0 + k);
        const float4 o = ld_bf16x4(O + row * ldo + c0 + k);
        acc += (a.x * o.x + a.y * o.y) + (a.z * o.z + a.w * o.w);
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (h < H && (lane & 1) == 0) delta[(size_t(b) * H + h) * NQ + q] = acc;
  }
}

}  // namespace

int colsum(const void* X, int ld, int rows, int N, float* out, int accumulate, float* workspace, size_t ws_bytes,
           cudaStream_t st) {
  MEBT_REQUIRE(rows >= 0 && N > 0 && N % 4 == 0 && ld % 4 == 0 && N <= 128 * 1000, MEBT_ERR_SHAPE,
               "colsum: bad shape rows=%d N=%d", rows, N);
  int slabs = (rows + 31) / 32;             // 4 rows per thread: enough CTAs in flight to hide the load latency
  if (slabs > 64) slabs = 64;
  if (slabs < 1) slabs = 1;
  const int rows_per_slab = (rows + slabs - 1) / slabs;
  MEBT_REQUIRE(workspace != nullptr && ws_bytes >= size_t(slabs) * N * 4, MEBT_ERR_WORKSPACE,
               "colsum: workspace too small (%zu < %zu)", ws_bytes, size_t(slabs) * N * 4);
  int* counters = reduce_counters(st);
  MEBT_REQUIRE(counters != nullptr, MEBT_ERR_CUDA, "colsum: cannot allocate the ticket counters");
  {
    LaunchScope ls(FAM_OTHER, double(rows) * N * 2.0, st);
    dim3 grid((N + 127) / 128, slabs);
    MEBT_CUDA_OK(launch_pdl(colsum_kernel, grid, dim3(256), 0, st, static_cast<const __nv_bfloat16*>(X), ld, rows, N,
                            rows_per_slab, workspace, out, accumulate, counters));
  }
  MEBT_LAUNCH_OK("colsum_kernel");
  return MEBT_OK;
}

constexpr int LNB_MAX_SLABS = 64;

size_t layernorm_bwd_workspace_bytes(int D) { return size_t(LNB_MAX_SLABS) * 2 * D * 4; }

// Parameter gradients of a LayerNorm: dgamma (+)= sum_r dy * xhat, dbeta (+)= sum_r dy.
int layernorm_bwd_params(const void* dy, const void* x, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                         int accumulate_params, int rows, int D, float* workspace, size_t ws_bytes, cudaStream_t st) {
  MEBT_REQUIRE(rows >= 0 && D > 0 && D % 4 == 0 && D <= 1024, MEBT_ERR_SHAPE, "layernorm_bwd: bad shape rows=%d D=%d", rows, D);
  MEBT_REQUIRE(workspace != nullptr && ws_bytes >= layernorm_bwd_workspace_bytes(D), MEBT_ERR_WORKSPACE,
               "layernorm_bwd: workspace too small");
  if (rows == 0) return MEBT_OK;
  int* counters = reduce_counters(st);
  MEBT_REQUIRE(counters != nullptr, MEBT_ERR_CUDA, "layernorm_bwd: cannot allocate the ticket counters");
  int slabs = (rows + 31) / 32;
  if (slabs > LNB_MAX_SLABS) slabs = LNB_MAX_SLABS;
  const int rows_per_slab = (rows + slabs - 1) / slabs;
  LaunchScope ls(FAM_LAYERNORM, double(rows) * D * 4.0, st);
  MEBT_CUDA_OK(launch_pdl(layernorm_bwd_param_kernel, dim3((D + 127) / 128, slabs), dim3(256), 0, st,
                          static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), mean, rstd, rows, D,
                          rows_per_slab, workspace, dgamma, dbeta, accumulate_params, counters + 1000));
  MEBT_LAUNCH_OK("layernorm_bwd_param_kernel");
  return MEBT_OK;
}

// dx = (resid != NULL ? resid : 0) + ln'(dy [+ dy_add]); resid may be dx itself.  dx_drop / drop: see the kernel.
// One or two problems (n = 1, 2) sharing gamma, in one launch.
struct LnDxDesc {
  const void* dy; const void* dy_add; const void* x; const float* mean; const float* rstd; void* dx; const void* resid;
  int rows; void* dx_drop; const DropKey* drop;
};
int layernorm_bwd_dx_multi(const LnDxDesc* d, int n, const float* gamma, int D, cudaStream_t st) {
  MEBT_REQUIRE(n >= 1 && n <= 2 && D > 0 && D % 4 == 0 && D <= 1024, MEBT_ERR_SHAPE, "layernorm_bwd: bad shape n=%d D=%d", n, D);
  LnDxProblem p[2];
  int blocks[2] = {0, 0};
  double bytes = 0.0;
  for (int i = 0; i < 2; ++i) {
    const LnDxDesc& s = d[i < n ? i : 0];
    p[i].dy = static_cast<const __nv_bfloat16*>(s.dy);
    p[i].dy_add = static_cast<const __nv_bfloat16*>(s.dy_add);
    p[i].x = static_cast<const __nv_bfloat16*>(s.x);
    p[i].mean = s.mean; p[i].rstd = s.rstd;
    p[i].dx = static_cast<__nv_bfloat16*>(s.dx);
    p[i].resid = static_cast<const __nv_bfloat16*>(s.resid);
    p[i].dx_drop = static_cast<__nv_bfloat16*>(s.dx_drop);          // with thr == 0 the second output is a plain copy
    p[i].drop = s.drop != nullptr ? *s.drop : DropKey{0u, 0u, 0u, 1.f};
    p[i].rows = i < n ? s.rows : 0;
    MEBT_REQUIRE(p[i].rows >= 0, MEBT_ERR_SHAPE, "layernorm_bwd: negative row count");
    blocks[i] = (p[i].rows + 7) / 8;
    bytes += double(p[i].rows) * D * (6.0 + (p[i].resid != nullptr ? 2.0 : 0.0) + (p[i].dy_add != nullptr ? 2.0 : 0.0) +
                                      (p[i].dx_drop != nullptr ? 2.0 : 0.0));
  }
  if (blocks[0] + blocks[1] == 0) return MEBT_OK;
  LaunchScope ls(FAM_LAYERNORM, bytes, st);
  if (D == 1024) {
    const int b0 = (p[0].rows + 3) / 4, b1 = (p[1].rows + 3) / 4;
    MEBT_CUDA_OK(launch_pdl(layernorm_bwd_dx1024_kernel, dim3(b0 + b1), dim3(128), 0, st, p[0], p[1], b0, gamma));
    MEBT_LAUNCH_OK("layernorm_bwd_dx1024_kernel");
    return MEBT_OK;
  }
  const dim3 grid(blocks[0] + blocks[1]);
  if (D <= 256) MEBT_CUDA_OK(launch_pdl(layernorm_bwd_dx_kernel<2>, grid, dim3(256), 0, st, p[0], p[1], blocks[0], gamma, D));
  else if (D <= 512) MEBT_CUDA_OK(launch_pdl(layernorm_bwd_dx_kernel<4>, grid, dim3(256), 0, st, p[0], p[1], blocks[0], gamma, D));
  else MEBT_CUDA_OK(launch_pdl(layernorm_bwd_dx_kernel<8>, grid, dim3(256), 0, st, p[0], p[1], blocks[0], gamma, D));
  MEBT_LAUNCH_OK("layernorm_bwd_dx_kernel");
  return MEBT_OK;
}

int layernorm_bwd_dx(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                     const void* resid, int rows, int D, void* dx_drop, const DropKey* drop, cudaStream_t st) {
  MEBT_REQUIRE(rows >= 0, MEBT_ERR_SHAPE, "layernorm_bwd: bad shape rows=%d D=%d", rows, D);
  if (rows == 0) return MEBT_OK;
  const LnDxDesc d{dy, nullptr, x, mean, rstd, dx, resid, rows, dx_drop, drop};
  return layernorm_bwd_dx_multi(&d, 1, gamma, D, st);
}

int layernorm_bwd_resid(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                        void* dx, const void* resid, float* dgamma, float* dbeta, int accumulate_params, int rows, int D,
                        float* workspace, size_t ws_bytes, cudaStream_t st) {
  // parameter gradients first: they read dy, which dx may overwrite when the caller accumulates in place
  int rc = layernorm_bwd_params(dy, x, mean, rstd, dgamma, dbeta, accumulate_params, rows, D, workspace, ws_bytes, st);
  if (rc) return rc;
  return layernorm_bwd_dx(dy, x, mean, rstd, gamma, dx, resid, rows, D, nullptr, nullptr, st);
}

int layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                  int accumulate_dx, float* dgamma, float* dbeta, int accumulate_params, int rows, int D,
                  float* workspace, size_t ws_bytes, cudaStream_t st) {
  return layernorm_bwd_resid(dy, x, mean, rstd, gamma, dx, accumulate_dx ? dx : nullptr, dgamma, dbeta, accumulate_params,
                             rows, D, workspace, ws_bytes, st);
}

int embed_backward(const int64_t* x, int x_stride, const int64_t* ctx_idx, int ctx_stride, const int64_t* tgt_idx,
                   int tgt_stride, const void* d_ctx, const void* d_tgt, const void* d_lat, float* d_tok, float* d_pos,
                   float* d_mask, float* d_sos, int B, int NC, int NT, int L, int D, float* workspace, size_t ws_bytes,
                   cudaStream_t st) {
  MEBT_REQUIRE(B > 0 && NC >= 0 && NT >= 0 && L >= 0 && D % 4 == 0, MEBT_ERR_SHAPE, "embed_backward: bad shape");
  const long long rows = (long long)B * (NC + NT);
  if (rows > 0) {
    LaunchScope ls(FAM_EMBED, double(B) * (NC * (16.0 + 2.0 * D + 16.0 * D) + NT * (8.0 + 2.0 * D + 8.0 * D)), st);
    embed_scatter_add_kernel<<<int((rows + 7) / 8), 256, 0, st>>>(
        x, x_stride, ctx_idx, ctx_stride, tgt_idx, tgt_stride, static_cast<const __nv_bfloat16*>(d_ctx),
        static_cast<const __nv_bfloat16*>(d_tgt), d_tok, d_pos, B, NC, NT, D);
    MEBT_LAUNCH_OK("embed_scatter_add_kernel");
  }
  if (NT > 0) {   // mask_emb is broadcast to every target row (transformer.py:263)
    int rc = colsum(d_tgt, D, B * NT, D, d_mask, 1, workspace, ws_bytes, st);
    if (rc) return rc;
  }
  if (L > 0) {    // sos_emb is broadcast over the batch (transformer.py:274)
    const long long per_batch = (long long)L * D;
    LaunchScope ls(FAM_OTHER, double(B) * per_batch * 2.0, st);
    batch_sum_kernel<<<int((per_batch / 4 + 255) / 256), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(d_lat), B,
                                                                        per_batch, d_sos, 1);
    MEBT_LAUNCH_OK("batch_sum_kernel");
  }
  return MEBT_OK;
}

int attn_delta(const void* dO, int lddo, const void* O, int ldo, float* delta, int B, int H, int NQ, cudaStream_t st) {
  const long long rows = (long long)B * NQ;
  if (rows == 0) return MEBT_OK;
  LaunchScope ls(FAM_ATTENTION, 0.0, st);
  MEBT_CUDA_OK(launch_pdl(attn_delta_kernel, dim3(int((rows + 7) / 8)), dim3(256), 0, st,
                          static_cast<const __nv_bfloat16*>(dO), lddo, static_cast<const __nv_bfloat16*>(O), ldo, delta, B, H,
                          NQ));
  MEBT_LAUNCH_OK("attn_delta_kernel");
  return MEBT_OK;
}

}  // namespace mebt

extern "C" {

size_t mebt_colsum_workspace_bytes(int N) { return size_t(64) * N * 4; }

int mebt_colsum(const void* X, int ld, int rows, int N, float* out, int accumulate, void* workspace,
                size_t workspace_bytes, void* stream) {
  return mebt::colsum(X, ld, rows, N, out, accumulate, static_cast<float*>(workspace), workspace_bytes,
                      static_cast<cudaStream_t>(stream));
}

size_t mebt_layernorm_bwd_workspace_bytes(int D) { return mebt::layernorm_bwd_workspace_bytes(D); }

int mebt_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                       int accumulate_dx, float* dgamma, float* dbeta, int accumulate_params, int rows, int D,
                       void* workspace, size_t workspace_bytes, void* stream) {
  return mebt::layernorm_bwd(dy, x, mean, rstd, gamma, dx, accumulate_dx, dgamma, dbeta, accumulate_params, rows, D,
                             static_cast<float*>(workspace), workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mebt_embed_backward(const int64_t* x_indices, int x_stride, const int64_t* ctx_idx, int ctx_stride,
                        const int64_t* tgt_idx, int tgt_stride, const void* d_contexts, const void* d_targets,
                        const void* d_latents, float* d_tok_emb, float* d_pos_emb, float* d_mask_emb, float* d_sos_emb,
                        int B, int NC, int NT, int L, int D, void* workspace, size_t workspace_bytes, void* stream) {
  return mebt::embed_backward(x_indices, x_stride, ctx_idx, ctx_stride, tgt_idx, tgt_stride, d_contexts, d_targets,
                              d_latents, d_tok_emb, d_pos_emb, d_mask_emb, d_sos_emb, B, NC, NT, L, D,
                              static_cast<float*>(workspace), workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
