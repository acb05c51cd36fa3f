// Stand-alone GEMM micro-benchmark (development tool, not part of the library): launches mebt_gemm_bf16_aux from
// C++ so that small shapes are not bound by the Python/ctypes call overhead, and prints the role-level cycle
// accounting compiled in with -DMEBT_GEMM_TRACE.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DMEBT_GEMM_TRACE \
//        -Iinclude tools/gemm_bench.cu mebt_b200/csrc/gemm.cu mebt_b200/csrc/runtime.cu -lcuda -o tools/gemm_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/mebt_b200.h"

extern "C" void mebt_gemm_set_trace(long long* buf);

struct Shape { int M, N, K, flags; const char* name; };

int main(int argc, char** argv) {
  const bool train = argc > 1 && !strcmp(argv[1], "train");
  std::vector<Shape> shapes;
  if (train) {
    shapes = {{128, 256, 64, 0, "tiny"},          {1536, 1024, 1024, 0, "lat proj"},  {1536, 3072, 1024, 0, "lat qkv"},
              {1536, 4096, 1024, MEBT_GEMM_GELU, "lat fc1"}, {1536, 1024, 4096, 0, "lat fc2"}, {3072, 1024, 1024, 0, "tgt proj"},
              {3072, 2048, 1024, 0, "kv"},        {3072, 4096, 1024, MEBT_GEMM_GELU, "tgt fc1"}, {3072, 1024, 4096, 0, "tgt fc2"},
              {3072, 16384, 1024, 0, "head"}};
  } else {
    shapes = {{4096, 1024, 1024, 0, "lat proj"},  {4096, 3072, 1024, 0, "lat qkv"},  {4096, 4096, 1024, MEBT_GEMM_GELU, "lat fc1"},
              {4096, 1024, 4096, 0, "lat fc2"},   {65536, 1024, 1024, 0, "tgt proj"}, {65536, 2048, 1024, 0, "kv"},
              {65536, 4096, 1024, MEBT_GEMM_GELU, "tgt fc1"}, {65536, 4096, 1024, 0, "tgt fc1 nogelu"}, {65536, 1024, 4096, 0, "tgt fc2"},
              {32768, 16384, 1024, MEBT_GEMM_OUT_FP32, "head fp32"}, {32768, 16384, 1024, 0, "head bf16"}};
  }
  cudaStream_t st;
  cudaStreamCreate(&st);
  long long* trace;
  cudaMalloc(&trace, 148 * 29 * sizeof(long long));
  for (const Shape& s : shapes) {
    const size_t pool = 8;   // rotate weights so that they come from HBM, as inside the model
    __nv_bfloat16 *A, *W, *R;
    void* C;
    float* bias;
    const bool f32 = s.flags & MEBT_GEMM_OUT_FP32;
    cudaMalloc(&A, size_t(s.M) * s.K * 2);
    cudaMalloc(&R, size_t(s.M) * s.N * 2);
    cudaMalloc(&W, pool * size_t(s.N) * s.K * 2);
    cudaMalloc(&C, size_t(s.M) * s.N * (f32 ? 4 : 2));
    cudaMalloc(&bias, s.N * 4);
    cudaMemset(A, 0x3c, size_t(s.M) * s.K * 2);       // bf16 0x3c3c = 0.0115
    cudaMemset(W, 0x3c, pool * size_t(s.N) * s.K * 2);
    cudaMemset(R, 0, size_t(s.M) * s.N * 2);
    cudaMemset(bias, 0, s.N * 4);
    auto run = [&](int i) {
      return mebt_gemm_bf16_aux(A, s.K, 0, W + (i % pool) * size_t(s.N) * s.K, s.K, 0, C, s.N, s.M, s.N, s.K, bias,
                                (s.N == 1024 && !f32) ? R : nullptr, s.N, nullptr, 0, s.flags, st);
    };
    mebt_gemm_set_trace(nullptr);
    for (int i = 0; i < 5; ++i)
      if (run(i)) { printf("%s: error %s\n", s.name, mebt_last_error()); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps; ++i) run(i);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double us = ms * 1e3 / reps;
    cudaMemset(trace, 0, 148 * 29 * sizeof(long long));
    mebt_gemm_set_trace(trace);
    run(0);
    cudaStreamSynchronize(st);
    static long long h[148 * 29];
    cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    double avg[8] = {0};
    int n = 0;
    for (int b = 0; b < 148; ++b) {
      if (h[b * 8 + 4] == 0) continue;
      ++n;
      for (int i = 0; i < 8; ++i) avg[i] += double(h[b * 8 + i]);
    }
    for (int i = 0; i < 8; ++i) avg[i] /= (n ? n : 1);
    double ex[5] = {0, 0, 0, 0, 0};
    int n2 = 0;
    for (int b = 0; b < 148; ++b) n2 += h[b * 8 + 6] != 0;
    for (int b = 0; b < 148; ++b) for (int i = 0; i < 5; ++i) ex[i] += double(h[148 * 8 + b * 5 + i]) / (n2 ? n2 : 1);
    printf("%-16s M=%6d N=%6d K=%5d | %8.1f us %7.1f TF | ctas %3d | prod wait_empty %7.0f / %7.0f | mma wait_full %7.0f wait_acc %7.0f / %7.0f | epi wait_full %7.0f / %7.0f [slot %6.0f fence %6.0f store %6.0f tmem %6.0f bias %6.0f]\n",
           s.name, s.M, s.N, s.K, us, 2.0 * s.M * s.N * s.K / us / 1e6, n, avg[0], avg[1], avg[2], avg[3], avg[4], avg[5], avg[6], ex[0], ex[1], ex[2], ex[3], ex[4]);
    {   // milestones of the CTAs (SM clocks after kernel entry): prologue done, dependency wait passed, first operands landed,
        // last MMA issued, accumulator complete, last unit staged, CTA end
      double t[16] = {0};
      int m = 0;
      for (int b = 0; b < 148; ++b) {
        const long long* q = h + 148 * 13 + b * 16;
        if (q[0] == 0) continue;
        ++m;
        for (int i = 1; i < 16; ++i) t[i] += double(q[i] - q[0]);
      }
      for (int i = 1; i < 16; ++i) t[i] /= (m ? m : 1);
      printf("    milestones (clk): prologue %6.0f | dep %6.0f | first-full %6.0f | last-mma %6.0f | acc-ready %6.0f | staged %6.0f | end %6.0f\n",
             t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
      printf("    epilogue unit 0: enter %6.0f in-ready %6.0f computed %6.0f done %6.0f | unit 1: enter %6.0f in-ready %6.0f computed %6.0f done %6.0f\n",
             t[8], t[9], t[10], t[11], t[12], t[13], t[14], t[15]);
    }
    fflush(stdout);
    cudaFree(A); cudaFree(W); cudaFree(C); cudaFree(bias); cudaFree(R);
  }
  return 0;
}
