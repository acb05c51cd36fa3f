// TMEM -> register read throughput microbenchmark (development tool): 4 warps issue tcgen05.ld over their lane
// quarters, optionally while one thread keeps the tensor pipe busy with tcgen05.mma into another TMEM region.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Iinclude tools/tmem_probe.cu \
//        mebt_b200/csrc/runtime.cu -lcuda -o tools/tmem_probe
#include <cstdio>
#include "../mebt_b200/csrc/common.cuh"
using namespace mebt;

__device__ __forceinline__ void tmem_ld_32x32_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr) : "memory");
}

// mode bit 0: MMA running concurrently; bits 1..: 0 = x32 ld+wait each, 1 = two x32 in flight, 2 = x64
__global__ void __launch_bounds__(192, 1) tmem_kernel(int mode, int iters, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 192) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 2) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 1) {
    if (lane == 0 && (mode & 1)) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 256, 0, 0);
      const uint32_t sA = smem_u32(smem), sB = sA + 16384;
      uint32_t ph = 0;
      int n = 0;
      while (!stop) {
        for (int k = 0; k < 16; ++k)
          umma_bf16_ss(tb + 256, make_smem_desc_sw128(sA + (k & 3) * 32, 16, 1024), make_smem_desc_sw128(sB + (k & 3) * 32, 16, 1024), idesc, 1u);
        umma_commit(&bar);
        mbar_wait(&bar, ph); ph ^= 1; ++n;
      }
      out[8] = n;
    }
  } else if (warp >= 2) {
    const int q = warp & 3;
    const uint32_t t0 = tb + (uint32_t(q * 32) << 16);
    float acc = 0.f;
    __syncwarp();
    const long long c0 = clock64();
    const int kind = mode >> 1;
    for (int i = 0; i < iters; ++i) {
      if (kind == 0) {
        for (int c = 0; c < 8; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t0 + c * 32, r);
          tmem_ld_wait();
          acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
        }
      } else if (kind == 1) {
        for (int c = 0; c < 8; c += 2) {
          uint32_t r[32], r2[32];
          tmem_ld_32x32(t0 + c * 32, r);
          tmem_ld_32x32(t0 + c * 32 + 32, r2);
          tmem_ld_wait();
          acc += __uint_as_float(r[0]) + __uint_as_float(r2[31]);
        }
      } else {
        for (int c = 0; c < 8; c += 2) {
          uint32_t r[64];
          tmem_ld_32x32_x64(t0 + c * 32, r);
          tmem_ld_wait();
          acc += __uint_as_float(r[0]) + __uint_as_float(r[63]);
        }
      }
    }
    const long long c1 = clock64();
    if (lane == 0) out[q] = c1 - c0;
    sink[threadIdx.x] = acc;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) stop = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 16 * sizeof(long long)); cudaMalloc(&sink, 192 * 4);
  cudaFuncSetAttribute(tmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  for (int mode = 0; mode < 6; ++mode) {
    cudaMemset(out, 0, 16 * sizeof(long long));
    tmem_kernel<<<1, 192, 64 * 1024>>>(mode, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[16];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const char* kinds[3] = {"x32 ld+wait", "2 x32 in flight", "x64 ld+wait"};
    printf("mma %d, %-16s: %.0f clk per 128x256 fp32 accumulator sweep per warp (%.1f clk per 32 columns); mma batches %lld (%.0f clk per 128x256x16 mma)\n", mode & 1, kinds[mode >> 1],
           double(h[0]) / iters, double(h[0]) / iters / 8, h[8], h[8] ? double(h[0]) / (16.0 * h[8]) : 0.0);
  }
  return 0;
}
