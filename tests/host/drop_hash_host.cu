// Host-side driver for the dropout keep-decision helpers of mebt_b200/csrc/common.cuh (they are __host__ __device__):
// prints the keep factors of a small grid so that tests/test_dropout_hash_cpu.py can compare them with an independent
// Python restatement of the documented algorithm and check their statistics without a GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -I include -o <out> tests/host/drop_hash_host.cu
// (only host code runs: no GPU is needed)
#include <cstdio>
#include <cstdlib>
#include "../../mebt_b200/csrc/common.cuh"

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: p seed site rows pairs\n"); return 2; }
  const float p = float(atof(argv[1]));
  const unsigned long long seed = strtoull(argv[2], nullptr, 10), site = strtoull(argv[3], nullptr, 10);
  const int rows = atoi(argv[4]), pairs = atoi(argv[5]);
  const mebt::DropKey key = mebt::make_drop_key(p, seed, site);
  printf("%u %u %u %.9g\n", key.k0, key.k1, key.thr, double(key.inv_keep));
  for (int r = 0; r < rows; ++r) {
    const uint32_t rk = mebt::drop_row_key(key, uint32_t(r));
    for (int c = 0; c < pairs; ++c) {
      float f0, f1;
      mebt::drop_pair(key, rk, uint32_t(c), f0, f1);
      putchar(f0 != 0.f ? '1' : '0');
      putchar(f1 != 0.f ? '1' : '0');
    }
    putchar('\n');
  }
  return 0;
}
