// TMA per-box cost vs box height / swizzle / L2 promotion (development tool).  One producer thread streams boxes of a
// [rows, 64] bf16 matrix (L2 resident) through a 6-slot ring; prints clk per box.
#include <cstdio>
#include <cuda.h>
#include "../mebt_b200/csrc/common.cuh"
using namespace mebt;
constexpr int STAGES = 6;
__device__ __forceinline__ void wait_v(uint64_t* bar, uint32_t parity, int variant) {
  if (variant == 0) { mbar_wait(bar, parity); return; }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap tm, int box_rows, int iters, int total_rows, long long* cycles, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * 32768);
  uint64_t* empty = full + STAGES;
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); } fence_barrier_init(); }
  __syncthreads();
  const long long t0 = clock64();
  const int nbox = total_rows / box_rows;
  if (threadIdx.x == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      wait_v(&empty[stage], phase ^ 1, variant);
      mbar_arrive_expect_tx(&full[stage], box_rows * 128);
      tma_load_2d(smem + stage * 32768, &tm, &full[stage], 0, ((blockIdx.x * 977 + i * 31) % nbox) * box_rows);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      wait_v(&full[stage], phase, variant);
      mbar_arrive(&empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fnp);
  const size_t rows = 512 * 1024;
  void* buf; cudaMalloc(&buf, rows * 128); cudaMemset(buf, 1, rows * 128);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int smem_bytes = STAGES * 32768 + 256;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 3000;
  struct V { const char* name; CUtensorMapSwizzle sw; CUtensorMapL2promotion pr; } vs[] = {
    {"sw128 promo256", CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
};
  for (int variant : {0, 1}) for (auto& v : vs) for (int box_rows : {32, 256}) {
    CUtensorMap m;
    cuuint64_t dims[2] = {64, rows}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw, v.pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
    for (int ctas : {1, 148}) {
      k<<<ctas, 64, smem_bytes>>>(m, box_rows, iters, int(rows), cyc, variant);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("err\n"); return 1; }
      long long h[148]; cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < ctas; ++i) avg += double(h[i]); avg /= ctas;
      printf("wait %s | %s box %3d rows, %3d ctas: %6.0f clk per box, %5.1f B/clk/SM\n", variant ? "test_wait" : "try_wait", v.name, box_rows, ctas, avg / iters, double(box_rows) * 128 * iters / avg);
    }
  }
  return 0;
}
