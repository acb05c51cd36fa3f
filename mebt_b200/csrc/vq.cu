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
#include "common.cuh"

namespace mebt {
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
  static bool attr = false;
  if (!attr) {
    MEBT_CUDA_OK(cudaFuncSetAttribute(vq_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
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

}  // extern "C"
