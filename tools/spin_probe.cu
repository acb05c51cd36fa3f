// Does a thread spinning on mbarrier.try_wait slow down shared-memory traffic of other warps? (development tool)
#include <cstdio>
#include "../mebt_b200/csrc/common.cuh"
using namespace mebt;

__global__ void __launch_bounds__(192, 1) spin_kernel(int spinners, int backoff, int iters, long long* out, float* sink) {
  __shared__ __align__(16) float buf[4][32 * 36];
  __shared__ uint64_t never, done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&never, 1); mbar_init(&done, 128); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 4 * 32 * 36; i += 192) (&buf[0][0])[i] = 1.0f;
  __syncthreads();
  if (warp < 2) {
    if (lane == 0 && warp < spinners) {
      // spin exactly like mbar_wait() on a barrier that completes only when the workers are done
      while (!mbar_try_wait(&done, 0)) { if (backoff) __nanosleep(backoff); }
    }
  } else {
    float* b = buf[warp - 2];
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    __syncwarp();
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(b + ((lane + j) & 31) * 36 + (j & 7) * 4);
        acc[j] += v.x + v.y + v.z + v.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(b + lane * 36 + j * 4) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    const long long c1 = clock64();
    if (lane == 0) out[warp - 2] = c1 - c0;
    sink[threadIdx.x] = acc[0] + acc[7];
    mbar_arrive(&done);
  }
}

int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 8 * sizeof(long long)); cudaMalloc(&sink, 192 * 4);
  const int iters = 4000;
  for (int backoff : {0, 32, 128}) for (int spinners = 0; spinners <= 2; ++spinners) {
    spin_kernel<<<1, 192>>>(spinners, backoff, iters, out, sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[8];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("spinners %d backoff %3d ns: %.1f / %.1f / %.1f / %.1f clk per iteration (8 LDS.128 + 32 FADD + 4 STS.128) for warps 2..5\n", spinners, backoff,
           double(h[0]) / iters, double(h[1]) / iters, double(h[2]) / iters, double(h[3]) / iters);
  }
  return 0;
}
