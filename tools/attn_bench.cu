// Stand-alone attention micro-benchmark with per-warpgroup cycle accounting (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DMEBT_ATTN_TRACE -Iinclude \
//        tools/attn_bench.cu mebt_b200/csrc/attention.cu mebt_b200/csrc/runtime.cu -lcuda -o tools/attn_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../include/mebt_b200.h"
#ifdef NO_TRACE
static void mebt_attn_set_trace(long long*) {}
#else
extern "C" void mebt_attn_set_trace(long long* buf);
#endif
int main() {
  struct S { const char* name; int B, NQ, NK; } shapes[] = {{"latent_enc", 16, 256, 8192}, {"latent_dec", 16, 8192, 256}, {"latent_self", 16, 256, 256}, {"enc train", 6, 256, 512}, {"dec train", 6, 512, 256}};
  const int H = 16, D = 1024;
  long long* trace; cudaMalloc(&trace, 148 * 16 * 8);
  cudaStream_t st; cudaStreamCreate(&st);
  for (auto& s : shapes) {
    __nv_bfloat16 *q, *kv, *o;
    cudaMalloc(&q, size_t(s.B) * s.NQ * D * 2); cudaMalloc(&kv, size_t(s.B) * s.NK * 2 * D * 2); cudaMalloc(&o, size_t(s.B) * s.NQ * D * 2);
    cudaMemset(q, 0x3c, size_t(s.B) * s.NQ * D * 2); cudaMemset(kv, 0x3c, size_t(s.B) * s.NK * 2 * D * 2);
    auto run = [&]() { return mebt_latent_attention_fwd(q, D, 0, kv, 2 * D, 0, D, s.NK, nullptr, 0, 0, 0, 0, o, D, nullptr, s.B, H, s.NQ, 64, st); };
    mebt_attn_set_trace(nullptr);
    for (int i = 0; i < 3; ++i) if (run()) { printf("error %s\n", mebt_last_error()); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < 20; ++i) run();
    cudaEventRecord(e1, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemset(trace, 0, 148 * 16 * 8);
    mebt_attn_set_trace(trace);
    run(); cudaStreamSynchronize(st);
    long long h[148 * 16]; cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
    double a[4] = {0, 0, 0, 0}; int n = 0;
    for (int i = 0; i < 296; ++i) if (h[i * 4 + 3]) { ++n; for (int k = 0; k < 4; ++k) a[k] += double(h[i * 4 + k]); }
    for (int k = 0; k < 4; ++k) a[k] /= (n ? n : 1);
    const double tiles = double(s.B) * H * ((s.NQ + 127) / 128) * ((s.NK + 127) / 128) / 148.0;
    printf("%-12s B=%2d NQ=%5d NK=%5d | %8.1f us | per softmax WG: wait S %8.0f  wait O %8.0f  row pass %8.0f  total %8.0f clk | ~%.0f tiles per CTA -> %.0f clk per 128x128 tile\n",
           s.name, s.B, s.NQ, s.NK, ms * 1e3 / 20, a[0], a[1], a[2], a[3], tiles, a[3] / tiles);
#ifndef NO_TRACE
    { double s0 = 0, s1 = 0, s2 = 0, s3 = 0; int c = 0; for (int i = 0; i < 148; ++i) if (h[148 * 8 + 4 * i + 3]) { ++c; s0 += h[148 * 8 + 4 * i]; s1 += h[148 * 8 + 4 * i + 1]; s2 += h[148 * 8 + 4 * i + 2]; s3 += h[148 * 8 + 4 * i + 3]; }
      if (c) printf("   mma thread: wait P %.0f  wait K/V,Q %.0f  PV issue (8 MMAs) %.0f  total %.0f clk (%.0f steps)\n", s0 / c, s1 / c, s2 / c, s3 / c, tiles); }
#endif
    cudaFree(q); cudaFree(kv); cudaFree(o);
  }
  return 0;
}
