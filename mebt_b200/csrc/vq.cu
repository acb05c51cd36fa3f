// K9: VQGAN codebook nearest-neighbour search, fused distance + argmin (no [M, n_codes] matrix in HBM).
// reference: Codebook.forward, mebt/modules/codebook.py:53-57
//     d = |z|^2 - 2 z.E^T + |E|^2 ;  encoding = argmin_k d
// Arithmetic is fp32 FFMA on purpose: with 16384 candidates the best/second-best gap on N(0,1) data goes
// down to ~5e-4 at |d| ~ 370 (SURVEY.md §7 hard part 5), which bf16/tf32 tensor-core products cannot
// resolve.  The same expression order as the reference is kept ((|z|^2 - 2 dot) + |e|^2) so rounding
// matches up to the summation order of the dot product; exact ties resolve to the lowest index, like
// torch.argmin.
//
// Tiling: CTA = 64 latent vectors x CODES_PER_CTA codes, 256 threads, 4x4 register micro-tiles, the
// z tile (64 x C) resident in shared memory, code tiles (64 x 16) streamed.  Partial winners are merged
// across CTAs with a 64-bit atomicMin on (orderable distance bits << 32 | index).
#include <cuda_fp16.h>

#include "common.cuh"

namespace mebt {

int gemm_f16_argmin(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* col_sq,
                    const float* row_sq, unsigned long long* packed, cudaStream_t stream);

namespace {

constexpr int VQ_BM = 64;        // latent vectors per CTA
constexpr int VQ_BN = 64;        // codes per inner tile
constexpr int VQ_BK = 16;        // channels per smem step
constexpr int VQ_THREADS = 256;

__device__ __forceinline__ uint32_t order_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void row_sqnorm_kernel(const float* __restrict__ E, int K, int C, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= K) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = E[(long long)row * C + c];
    s += v * v;
  }
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

__global__ void vq_init_kernel(unsigned long long* __restrict__ packed, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) packed[i] = ~0ull;
}

__global__ void vq_unpack_kernel(const unsigned long long* __restrict__ packed, int64_t* __restrict__ out, long long M) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) out[i] = int64_t(packed[i] & 0xFFFFFFFFull);
}

// z is channel-first: z[(b * C + c) * S + s]; logical row m = b * S + s  (shift_dim(z,1,-1).flatten, codebook.py:52)
__global__ void __launch_bounds__(VQ_THREADS) vq_argmin_kernel(const float* __restrict__ z, int S, int C,
                                                               const float* __restrict__ E,
                                                               const float* __restrict__ e_sq, int K, long long M,
                                                               int codes_per_cta,
                                                               unsigned long long* __restrict__ packed) {
  extern __shared__ float smf[];
  float* zs = smf;                         // [C][VQ_BM]   (k-major so a thread reads 4 consecutive rows as float4)
  float* es = zs + (size_t)C * VQ_BM;      // [VQ_BK][VQ_BN + 4]
  float* zsq = es + VQ_BK * (VQ_BN + 4);   // [VQ_BM]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4x4 outputs each
  const long long m0 = (long long)blockIdx.x * VQ_BM;
  const int code0 = blockIdx.y * codes_per_cta;

  // stage the z tile: for each channel the 64 rows are contiguous in s (coalesced) unless a batch boundary is crossed
  for (int i = threadIdx.x; i < C * VQ_BM; i += VQ_THREADS) {
    const int c = i / VQ_BM, r = i % VQ_BM;
    const long long m = m0 + r;
    float v = 0.f;
    if (m < M) {
      const long long b = m / S, s = m % S;
      v = z[(b * C + c) * S + s];
    }
    zs[c * VQ_BM + r] = v;
  }
  __syncthreads();
  if (threadIdx.x < VQ_BM) {               // |z|^2, sequential over channels
    float a = 0.f;
    for (int c = 0; c < C; ++c) { const float v = zs[c * VQ_BM + threadIdx.x]; a += v * v; }
    zsq[threadIdx.x] = a;
  }
  __syncthreads();

  float best[4];
  int besti[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { best[i] = INFINITY; besti[i] = 0x7fffffff; }

  for (int n0 = code0; n0 < code0 + codes_per_cta && n0 < K; n0 += VQ_BN) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += VQ_BK) {
      // E tile [64 codes][16 ch] -> es[ch][code]; each thread loads one float4 along channels
      {
        const int code = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + code < K) v = __ldg(reinterpret_cast<const float4*>(E + (long long)(n0 + code) * C + k0 + kq));
        es[(kq + 0) * (VQ_BN + 4) + code] = v.x;
        es[(kq + 1) * (VQ_BN + 4) + code] = v.y;
        es[(kq + 2) * (VQ_BN + 4) + code] = v.z;
        es[(kq + 3) * (VQ_BN + 4) + code] = v.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < VQ_BK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(zs + (k0 + k) * VQ_BM + ty * 4);
        const float4 b = *reinterpret_cast<const float4*>(es + k * (VQ_BN + 4) + tx * 4);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int code = n0 + tx * 4 + j;
      if (code < K) {
        const float esq = __ldg(e_sq + code);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d = (zsq[ty * 4 + i] - 2.0f * acc[i][j]) + esq;
          if (d < best[i]) { best[i] = d; besti[i] = code; }     // codes visited in increasing order
        }
      }
    }
  }
  // merge across the 16 threads (tx) that share rows, then across CTAs
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float bd = best[i];
    int bi = besti[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    const long long m = m0 + ty * 4 + i;
    if (tx == 0 && m < M && bi != 0x7fffffff)
      atomicMin(packed + m, (static_cast<unsigned long long>(order_bits(bd)) << 32) | uint32_t(bi));
  }
}


// ---- K9 on the tensor cores ------------------------------------------------------------------------------------
// fp32 x = hi + lo + r with hi = fp16(x), lo = fp16(x - hi), |r| <= 2^-22 |x|: the dot product z.e becomes ONE fp16
// GEMM with the reduction dimension tripled,
//     [z_hi | z_lo | z_hi] . [e_hi | e_hi | e_lo]^T = z_hi e_hi + z_lo e_hi + z_hi e_lo        (fp32 accumulation in TMEM)
// which drops lo.lo and the residuals: relative 2^-21 per product, 8 ulp of an fp32 product - far below the 1e-3 gap
// under which tests accept a flipped near-tie (bf16 halves would give 2^-15 and do flip them).  The distance and the
// running argmin are the GEMM's epilogue (csrc/gemm.cu, argmin_out), in the reference's expression order.

// z [b, C, S] fp32 channel-first -> A [b*S, 3C] fp16 rows (hi | lo | hi) and |z|^2 per row (fp32, channels in order).
// One CTA transposes a [C x 32] slab through shared memory: coalesced along s on the way in, along c on the way out.
__global__ void __launch_bounds__(256) vq_split_z_kernel(const float* __restrict__ z, int C, int S, __half* __restrict__ A,
                                                         float* __restrict__ zsq) {
  extern __shared__ float tile[];                    // [C][33]
  const int b = blockIdx.y, s0 = blockIdx.x * 32;
  const float* zb = z + (size_t)b * C * S;
  for (int i = threadIdx.x; i < C * 32; i += 256) {
    const int c = i >> 5, s = i & 31;
    tile[c * 33 + s] = (s0 + s < S) ? zb[(size_t)c * S + s0 + s] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < 32 && s0 + s < S; s += 8) {
    __half* row = A + ((size_t)b * S + s0 + s) * 3 * C;
    for (int c = lane; c < C; c += 32) {
      const float x = tile[c * 33 + s];
      const __half hi = __float2half_rn(x);
      const __half lo = __float2half_rn(x - __half2float(hi));
      row[c] = hi;
      row[C + c] = lo;
      row[2 * C + c] = hi;
    }
  }
  if (threadIdx.x < 32 && s0 + threadIdx.x < S) {    // |z|^2: sequential over channels, like the FFMA kernel
    float a = 0.f;
    for (int c = 0; c < C; ++c) { const float v = tile[c * 33 + threadIdx.x]; a += v * v; }
    zsq[(size_t)b * S + s0 + threadIdx.x] = a;
  }
}

// E [K, C] fp32 -> B [K, 3C] fp16 rows (hi | hi | lo)
__global__ void vq_split_e_kernel(const float* __restrict__ E, long long n, int C, __half* __restrict__ B) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = int(i - r * C);
    const float x = E[i];
    const __half hi = __float2half_rn(x);
    const __half lo = __float2half_rn(x - __half2float(hi));
    __half* row = B + r * 3 * C;
    row[c] = hi;
    row[C + c] = hi;
    row[2 * C + c] = lo;
  }
}

}  // namespace
}  // namespace mebt

extern "C" {

int mebt_row_sqnorm(const float* E, int K, int C, float* out, void* stream) {
  MEBT_REQUIRE(K > 0 && C > 0, MEBT_ERR_SHAPE, "row_sqnorm: bad shape");
  mebt::row_sqnorm_kernel<<<(K + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(E, K, C, out);
  MEBT_LAUNCH_OK("row_sqnorm_kernel");
  return MEBT_OK;
}

size_t mebt_vq_argmin_workspace_bytes(long long M) { return size_t(M) * sizeof(unsigned long long); }

int mebt_vq_argmin(const float* z_channel_first, int batch, int C, int S, const float* E, const float* e_sqnorm, int K,
                   int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mebt;
  const long long M = (long long)batch * S;
  MEBT_REQUIRE(batch >= 0 && S >= 0 && C > 0 && C % VQ_BK == 0 && K > 0, MEBT_ERR_SHAPE,
               "vq_argmin: bad shape batch=%d C=%d S=%d K=%d (C must be a multiple of %d)", batch, C, S, K, VQ_BK);
  if (M == 0) return MEBT_OK;
  MEBT_REQUIRE(workspace != nullptr && workspace_bytes >= size_t(M) * 8, MEBT_ERR_WORKSPACE,
               "vq_argmin: workspace too small (%zu < %zu)", workspace_bytes, size_t(M) * 8);
  const size_t smem = (size_t(C) * VQ_BM + VQ_BK * (VQ_BN + 4) + VQ_BM) * sizeof(float);
  MEBT_REQUIRE(smem <= 200 * 1024, MEBT_ERR_UNSUPPORTED, "vq_argmin: embedding_dim %d too large for the smem tile", C);
  static bool attr[64] = {};
  if (first_use_on_device(attr)) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned long long* packed = static_cast<unsigned long long*>(workspace);
  vq_init_kernel<<<int((M + 255) / 256), 256, 0, st>>>(packed, M);
  const int m_blocks = int((M + VQ_BM - 1) / VQ_BM);
  // split the codebook so that the grid covers the SMs a few times over
  int splits = 1;
  while (m_blocks * splits < 2 * sm_count() && (K / (splits * 2)) >= VQ_BN * 4 && (K % (splits * 2 * VQ_BN)) == 0) splits *= 2;
  const int codes_per_cta = ((K + splits - 1) / splits + VQ_BN - 1) / VQ_BN * VQ_BN;
  dim3 grid(m_blocks, (K + codes_per_cta - 1) / codes_per_cta);
  LaunchScope ls(FAM_VQ, 2.0 * double(M) * double(K) * double(C), st);
  vq_argmin_kernel<<<grid, VQ_THREADS, smem, st>>>(z_channel_first, S, C, E, e_sqnorm, K, M, codes_per_cta, packed);
  MEBT_LAUNCH_OK("vq_argmin_kernel");
  vq_unpack_kernel<<<int((M + 255) / 256), 256, 0, st>>>(packed, out_idx, M);
  MEBT_LAUNCH_OK("vq_unpack_kernel");
  return MEBT_OK;
}


/* ---- tensor-core path ---- */
size_t mebt_vq_codebook_split_bytes(int K, int C) { return size_t(K) * 3 * C * 2; }

int mebt_vq_split_codebook(const float* E, int K, int C, void* e_split, void* stream) {
  MEBT_REQUIRE(K > 0 && C > 0, MEBT_ERR_SHAPE, "vq_split_codebook: bad shape");
  const long long n = (long long)K * C;
  mebt::vq_split_e_kernel<<<int((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      E, n, C, static_cast<__half*>(e_split));
  MEBT_LAUNCH_OK("vq_split_e_kernel");
  return MEBT_OK;
}

size_t mebt_vq_argmin_tc_workspace_bytes(long long M, int C) {
  return ((size_t(M) * 3 * C * 2 + 255) & ~size_t(255)) + ((size_t(M) * 4 + 255) & ~size_t(255)) + size_t(M) * 8;
}

int mebt_vq_argmin_tc(const float* z_channel_first, int batch, int C, int S, const void* e_split, const float* e_sqnorm,
                      int K, int64_t* out_idx, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mebt;
  const long long M = (long long)batch * S;
  MEBT_REQUIRE(batch >= 0 && S >= 0 && C > 0 && C % 64 == 0 && C <= 1024 && K > 0 && K % 64 == 0, MEBT_ERR_SHAPE,
               "vq_argmin_tc: bad shape batch=%d C=%d S=%d K=%d (C and K must be multiples of 64)", batch, C, S, K);
  if (M == 0) return MEBT_OK;
  MEBT_REQUIRE(M < (1ll << 31), MEBT_ERR_SHAPE, "vq_argmin_tc: too many vectors");
  MEBT_REQUIRE(workspace != nullptr && workspace_bytes >= mebt_vq_argmin_tc_workspace_bytes(M, C), MEBT_ERR_WORKSPACE,
               "vq_argmin_tc: workspace too small (%zu < %zu)", workspace_bytes, mebt_vq_argmin_tc_workspace_bytes(M, C));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* W = static_cast<char*>(workspace);
  __half* A = reinterpret_cast<__half*>(W);
  float* zsq = reinterpret_cast<float*>(W + ((size_t(M) * 3 * C * 2 + 255) & ~size_t(255)));
  unsigned long long* packed = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(zsq) + ((size_t(M) * 4 + 255) & ~size_t(255)));
  const size_t smem = size_t(C) * 33 * sizeof(float);
  static bool attr[64] = {};
  if (first_use_on_device(attr)) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(vq_split_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * 33 * 4));
  }
  {
    LaunchScope ls(FAM_VQ, 0.0, st);
    vq_split_z_kernel<<<dim3((S + 31) / 32, batch), 256, smem, st>>>(z_channel_first, C, S, A, zsq);
    MEBT_LAUNCH_OK("vq_split_z_kernel");
    vq_init_kernel<<<int((M + 255) / 256), 256, 0, st>>>(packed, M);
  }
  int rc = gemm_f16_argmin(A, 3 * C, e_split, 3 * C, int(M), K, 3 * C, e_sqnorm, zsq, packed, st);
  if (rc) return rc;
  vq_unpack_kernel<<<int((M + 255) / 256), 256, 0, st>>>(packed, out_idx, M);
  MEBT_LAUNCH_OK("vq_unpack_kernel");
  return MEBT_OK;
}

}  // extern "C"
