// tcgen05.mma issue / execution rate microbenchmark (development tool, not part of the library): one thread issues R
// back-to-back MMAs on fixed shared-memory / tensor-memory operands (no TMA traffic, contents irrelevant), commits, and
// waits for completion; reports clk per instruction for
//   SS  : A and B from shared memory (K-major, 128B swizzle), M = 128 (1 CTA) or 256 (cta_group::2), N = 64 .. 256
//   TS  : A from tensor memory, B from shared memory (MN-major), the PV product of attention
//   alt : the same instructions alternating between two accumulators (independent dependency chains)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Iinclude tools/mma_probe.cu \
//        mebt_b200/csrc/runtime.cu -lcuda -o tools/mma_probe
#include <cstdio>
#include <cstdlib>

#include "../mebt_b200/csrc/common.cuh"

using namespace mebt;

__device__ __forceinline__ uint32_t cta_rank_in_cluster() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// mode: 0 SS, 1 SS alternating accumulators, 2 TS, 3 TS alternating accumulators
template <bool TWO_SM>
__global__ void __launch_bounds__(288, 1) mma_probe_kernel(int mode, int N, int reps, long long* out, int hammer) {
  __shared__ volatile int stop_flag;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ __align__(8) uint64_t ready_bar, sink_bar[8];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = TWO_SM ? cta_rank_in_cluster() : 0;
  if (threadIdx.x == 0) {
    stop_flag = 0;
    mbar_init(&done_bar, 1);
    mbar_init(&ready_bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&sink_bar[i], 1);
    fence_barrier_init();
    mbar_arrive(&ready_bar);          // phase 0 of ready_bar is complete from the start
  }
  if (warp == 8) {
    if constexpr (TWO_SM) { tmem_alloc_2sm(&tmem_slot, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (TWO_SM) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (hammer == 3 && warp < 8) {
    // ALU / MUFU-heavy warps, two per scheduler, like the softmax warpgroups of the attention kernel
    float a = float(threadIdx.x), b = 1.0001f, c = 0.f;
    while (!stop_flag) {
#pragma unroll
      for (int i = 0; i < 64; ++i) { a = fmaf(a, b, 0.5f); c += ex2_approx(-a * 1e-3f); b = fmaf(b, 0.999f, 1e-3f); }
    }
    if (c == 123.456f) out[1] = 1;
  } else if (hammer && hammer < 3 && warp < 4) {
    // the softmax warps of the attention kernel in miniature: tcgen05.ld of 32 columns, (hammer 2: + tcgen05.st of 16), forever
    const uint32_t lane_addr = uint32_t((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    while (!stop_flag) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + lane_addr + 128, r);
      tmem_ld_wait_regs(r);
      acc += r[0] ^ r[31];
      if (hammer == 2) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = r[2 * i] + acc;
        tmem_st_32x16(tmem_base + lane_addr + 300, pk);
        tmem_st_wait();
      }
    }
    if (acc == 0x12345678u) out[1] = acc;
  }
  if (threadIdx.x == 256 && rank == 0) {
    const uint32_t M = TWO_SM ? 256 : 128;
    const bool ts = mode == 2 || mode == 3;
    const bool alt = mode & 1;
    // SS: A [128 x 64] K-major at smem + 0, B [N(/2) x 64] K-major at smem + 16 KiB.  TS: B = V [16 keys x 64 dims] MN-major.
    const uint32_t idesc = ts ? make_idesc_bf16(M, N, 0, 1) : make_idesc_bf16(M, N, 0, 0);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 16384);
    const long long t0 = clock64();
    if (mode == 2 && !TWO_SM) {
      // TS, tight form: eight instructions per asm block (the attention kernel's PV issue path)
      const uint32_t v_lo = smem_desc_lo(sB, 64 * 128);
      for (int r = 0; r < reps; r += 8)
        umma_bf16_ts_x8(tmem_base, tmem_base + 448, 8u, v_lo, 2048u >> 4, smem_desc_hi_sw128(1024), idesc, 1u);
    } else if (mode == 8 || mode == 9) {
      // mode 8: test_wait poll (non-suspending) + four MMAs + commit; mode 9: wait only (no MMA, no commit): the bare cost
      const uint32_t a_lo = smem_desc_lo(sA, 16), b_lo = smem_desc_lo(sB, 16);
      for (int r = 0; r < reps; r += 4) {
        if (mode == 8) {
          while (!mbar_test_wait(&ready_bar, 0)) {}
          umma_bf16_ss_x4<TWO_SM>(tmem_base, a_lo, b_lo, 2, 2, smem_desc_hi_sw128(1024), smem_desc_hi_sw128(1024), idesc, 1u);
          if constexpr (TWO_SM) umma_commit_2sm_mc(&sink_bar[(r >> 2) & 7], 1);
          else umma_commit(&sink_bar[(r >> 2) & 7]);
        } else {
          mbar_wait(&ready_bar, 0);
        }
      }
    } else if (mode == 6 || mode == 7) {
      // mode 6: wait (no fence) + four MMAs + commit; mode 7: wait + EIGHT MMAs (two k-blocks) + one commit
      const uint32_t a_lo = smem_desc_lo(sA, 16), b_lo = smem_desc_lo(sB, 16);
      for (int r = 0; r < reps; r += (mode == 7 ? 8 : 4)) {
        mbar_wait(&ready_bar, 0);
        umma_bf16_ss_x4<TWO_SM>(tmem_base, a_lo, b_lo, 2, 2, smem_desc_hi_sw128(1024), smem_desc_hi_sw128(1024), idesc, 1u);
        if (mode == 7)
          umma_bf16_ss_x4<TWO_SM>(tmem_base, a_lo + 1024, b_lo + 1024, 2, 2, smem_desc_hi_sw128(1024), smem_desc_hi_sw128(1024), idesc, 1u);
        if constexpr (TWO_SM) umma_commit_2sm_mc(&sink_bar[(r >> 2) & 7], 1);
        else umma_commit(&sink_bar[(r >> 2) & 7]);
      }
    } else if (mode == 4 || mode == 5) {
      // the GEMM main loop's per-k-block sequence without any data movement: [wait on a barrier that is already complete +
      // fence (mode 5)] + four MMAs + one tcgen05.commit to a barrier nobody waits for
      const uint32_t a_lo = smem_desc_lo(sA, 16), b_lo = smem_desc_lo(sB, 16);
      for (int r = 0; r < reps; r += 4) {
        if (mode == 5) { mbar_wait(&ready_bar, 0); tc_fence_after(); }
        umma_bf16_ss_x4<TWO_SM>(tmem_base, a_lo, b_lo, 2, 2, smem_desc_hi_sw128(1024), smem_desc_hi_sw128(1024), idesc, 1u);
        if constexpr (TWO_SM) umma_commit_2sm_mc(&sink_bar[(r >> 2) & 7], 1);
        else umma_commit(&sink_bar[(r >> 2) & 7]);
      }
    } else if (!ts && !alt) {
      // tight form: four instructions per asm block over precomputed descriptor words (the library's issue path)
      const uint32_t a_lo = smem_desc_lo(sA, 16), b_lo = smem_desc_lo(sB, 16);
      for (int r = 0; r < reps; r += 4)
        umma_bf16_ss_x4<TWO_SM>(tmem_base, a_lo, b_lo, 2, 2, smem_desc_hi_sw128(1024), smem_desc_hi_sw128(1024), idesc, 1u);
    } else
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem_base + ((alt && (r & 1)) ? 256u : 0u);
      const int k = r & 3;
      if (!ts) {
        const uint64_t da = make_smem_desc_sw128(sA + k * 32, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(sB + k * 32, 16, 1024);
        if constexpr (TWO_SM) umma_bf16_ss_2sm(d, da, db, idesc, r > 1);
        else umma_bf16_ss(d, da, db, idesc, r > 1);
      } else {
        const uint64_t db = make_smem_desc_sw128(sB + k * 2048, 64 * 128, 1024);
        const uint32_t a = tmem_base + 448 + k * 8;          // packed bf16 P: 8 columns per K = 16 step
        if constexpr (TWO_SM) umma_bf16_ts_2sm(d, a, db, idesc, r > 1);
        else umma_bf16_ts(d, a, db, idesc, r > 1);
      }
    }
    const long long t1 = clock64();
    if constexpr (TWO_SM) umma_commit_2sm_mc(&done_bar, 1);
    else umma_commit(&done_bar);
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    stop_flag = 1;
  }
  if (threadIdx.x == 256 && rank != 0) stop_flag = 1;
  tc_fence_before();
  __syncthreads();
  if (TWO_SM) cluster_sync_all();
  if (warp == 8) {
    tc_fence_after();
    if constexpr (TWO_SM) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int smem_bytes = 96 * 1024;
  cudaFuncSetAttribute(mma_probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(mma_probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int reps = 2048;
  const char* names[10] = {"SS (4 per asm block)", "SS alternating accumulators", "TS (8 per asm block, 1-CTA)", "TS alternating accumulators",
                          "SS x4 + commit", "wait + fence + SS x4 + commit", "wait + SS x4 + commit", "wait + SS x8 + commit", "test_wait + SS x4 + commit", "try_wait alone (per 4)"};
  for (int hammer = 0; hammer < 4; ++hammer)
  for (int two = 0; two < (hammer ? 1 : 2); ++two)
    for (int mode = 0; mode < 10; ++mode)
      for (int N : {64, 128, 192, 256}) {
        if (hammer && !(mode == 0 || mode == 2)) continue;
        if ((mode == 2 || mode == 3) && N > 128) continue;
        if (mode & 1 && N > 128 && two == 0 && false) continue;
        if ((mode & 1) && N > 256) continue;
        cudaMemset(out, 0, 16);
        if (two) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(2); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = smem_bytes;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr; cfg.numAttrs = 1;
          cudaLaunchKernelEx(&cfg, mma_probe_kernel<true>, mode, N, reps, out, hammer);
        } else {
          mma_probe_kernel<false><<<1, 288, smem_bytes>>>(mode, N, reps, out, hammer);
        }
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        const int M = two ? 256 : 128;
        const double nominal = double(M) * N * 16 / 4096.0 / (two ? 2 : 1);    // clk per SM at 4096 MAC/clk/SM
        printf("%s%s M=%3d N=%3d K=16 %-28s: issue %7.1f clk/MMA, complete %7.1f clk/MMA (nominal %5.1f)%s\n",
               hammer == 0 ? "" : (hammer == 1 ? "[4 warps tcgen05.ld] " : (hammer == 2 ? "[4 warps tcgen05.ld+st] " : "[8 warps FMA+MUFU] ")), two ? "cta_group::2" : "cta_group::1", M, N, names[mode], double(h[0]) / reps, double(h[1]) / reps, nominal,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
