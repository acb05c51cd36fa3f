// MUFU.EX2 issue rate per SM sub-partition (development tool): W warps per SMSP, each a chain-free stream of ex2.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int iters, float* out, long long* cyc) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  const long long c1 = clock64();
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int threads : {32, 128, 256, 512}) {
    k<<<148, threads>>>(iters, out, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_instr = double(h) / (double(iters) * 16);
    printf("%3d threads/SM: %.2f clk per warp-level MUFU.EX2 per warp -> %.1f ex2 lanes / clk / SM\n", threads, per_instr, (threads / 32) * 32.0 / per_instr);
  }
  return 0;
}
