// L2 -> shared-memory feed microbenchmark (development tool).  Persistent CTAs stream 16 KiB TMA boxes of a
// [rows, 64] bf16 matrix through a shared-memory ring without consuming them and report delivered bytes / clk / SM:
//   mode 0: every CTA streams its own rows (no sharing)
//   mode 1: the CTAs of a cluster stream the SAME rows, each with its own unicast loads
//   mode 2: the CTAs of a cluster stream the same rows; each loads 1/csz of every box and multicasts it to all
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Iinclude tools/mc_probe.cu \
//        mebt_b200/csrc/runtime.cu -lcuda -o tools/mc_probe
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mebt_b200/csrc/common.cuh"

using namespace mebt;

constexpr int STAGES = 6;
constexpr int BOX_ROWS = 256;                    // 256 rows x 64 bf16 = 32 KiB
constexpr int BOX_BYTES = BOX_ROWS * 128;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }

__global__ void __launch_bounds__(64, 1) feed_kernel(const __grid_constant__ CUtensorMap tm_full,
                                                     const __grid_constant__ CUtensorMap tm_part,
                                                     const __grid_constant__ CUtensorMap tm_str,
                                                     const __grid_constant__ CUtensorMap tm_3d, int mode, int csz,
                                                     int iters, int total_boxes, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * BOX_BYTES);
  uint64_t* empty = full + STAGES;
  const uint32_t rank = csz > 1 ? cluster_ctarank() : 0;
  const uint32_t cid = csz > 1 ? cluster_id_x() : blockIdx.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], mode == 2 ? csz : 1); }
    fence_barrier_init();
  }
  __syncthreads();
  if (csz > 1) cluster_sync_all();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {           // producer
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], BOX_BYTES);
      const int box = mode == 0 ? (blockIdx.x * 977 + i * 31) % total_boxes : (cid * 977 + i * 31) % total_boxes;
      uint8_t* dst = smem + stage * BOX_BYTES;
      if (mode == 2) {
        const int part = BOX_ROWS / csz;
        tma_load_2d_mc(dst + rank * part * 128, &tm_part, &full[stage], 0, box * BOX_ROWS + rank * part, uint16_t((1u << csz) - 1));
      } else if (mode == 5) {       // ONE 3-D box = the two k-adjacent [128 x 64] tiles of mode 4's row block
        const int rb = (blockIdx.x * 977 + (i >> 3) * 31) % (total_boxes / 16);
        tma_load_3d(dst, &tm_3d, &full[stage], 0, rb * 256, (i & 7) * 2);
      } else if (mode == 6) {       // mode 4's two boxes, issued by two different threads (second one below)
        const int rb = (blockIdx.x * 977 + (i >> 4) * 31) % (total_boxes / 16);
        tma_load_2d(dst, &tm_str, &full[stage], (i & 15) * 64, rb * 256);
      } else if (mode == 4) {       // GEMM-operand-like: [128 rows][64 k] boxes out of a [rows, 1024] row-major matrix
        const int rb = (blockIdx.x * 977 + (i >> 4) * 31) % (total_boxes / 16);
        tma_load_2d(dst, &tm_str, &full[stage], (i & 15) * 64, rb * 256);
        tma_load_2d(dst + BOX_BYTES / 2, &tm_str, &full[stage], (i & 15) * 64, rb * 256 + 128);
      } else if (mode == 3) {
        tma_load_2d(dst, &tm_part, &full[stage], 0, box * BOX_ROWS);
        tma_load_2d(dst + BOX_BYTES / 2, &tm_part, &full[stage], 0, ((box + 1024) % total_boxes) * BOX_ROWS);
      } else {
        tma_load_2d(dst, &tm_full, &full[stage], 0, box * BOX_ROWS);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 33 && mode == 6) {   // second producer thread (own warp lane; same warp as the consumer)
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[stage], phase ^ 1);
      const int rb = (blockIdx.x * 977 + (i >> 4) * 31) % (total_boxes / 16);
      tma_load_2d(smem + stage * BOX_BYTES + BOX_BYTES / 2, &tm_str, &full[stage], (i & 15) * 64, rb * 256 + 128);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {   // consumer: releases the stage as soon as it has landed
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[stage], phase);
      if (mode == 2) {
        for (int r = 0; r < csz; ++r)
          asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32(&empty[stage]), r)) : "memory");
      } else {
        mbar_arrive(&empty[stage]);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (csz > 1) cluster_sync_all();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  const int total_boxes = 2048;
  const size_t rows = size_t(total_boxes) * BOX_ROWS;
  void* buf;
  cudaMalloc(&buf, rows * 128);
  cudaMemset(buf, 1, rows * 128);
  long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int smem_bytes = STAGES * BOX_BYTES + 256;
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 4000;
  for (int csz : {1}) {
    for (int mode : {0, 1, 2}) {
      if (csz == 1 && mode != 0) continue;
      CUtensorMap tf, tp;
      if (get_tensor_map_2d(&tf, buf, 2, 64, rows, 128, 64, BOX_ROWS)) { printf("map: %s\n", mebt_last_error()); return 1; }
      if (get_tensor_map_2d(&tp, buf, 2, 64, rows, 128, 64, BOX_ROWS / csz)) { printf("map: %s\n", mebt_last_error()); return 1; }
      cudaLaunchConfig_t cfg = {};
      cfg.blockDim = dim3(64);
      cfg.dynamicSmemBytes = smem_bytes;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int max_clusters = 0;
      cfg.gridDim = dim3(csz);
      cudaOccupancyMaxActiveClusters(&max_clusters, feed_kernel, &cfg);
      int ctas = csz == 1 ? 148 : max_clusters * csz;
      if (ctas > 148) ctas = 148 / csz * csz;
      cfg.gridDim = dim3(ctas);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&cfg, feed_kernel, tf, tp, tf, tf, mode, csz, iters, total_boxes, cyc);
        if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("sync failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < ctas; ++i) avg += double(h[i]); avg /= ctas;
      const double per_sm = double(iters) * BOX_BYTES / avg;
      printf("csz %d mode %d (%s): ctas %3d (max clusters %d) | %.3f ms | delivered %.1f B/clk/SM, %.0f B/clk chip, %.2f TB/s\n", csz, mode,
             mode == 0 ? "distinct" : mode == 1 ? "same, unicast" : "same, multicast", ctas, max_clusters, best, per_sm, per_sm * ctas,
             double(iters) * BOX_BYTES * ctas / (best * 1e-3) / 1e12);
    }
  }
  // unicast, distinct rows: how does the per-SM rate depend on the number of active SMs / boxes per stage?
  for (int ctas : {37, 148}) {
    CUtensorMap tf, tp;
    get_tensor_map_2d(&tf, buf, 2, 64, rows, 128, 64, BOX_ROWS);
    get_tensor_map_2d(&tp, buf, 2, 64, rows, 128, 64, BOX_ROWS / 2);
    CUtensorMap ts;
    get_tensor_map_2d(&ts, buf, 2, 1024, rows / 16, 2048, 64, 128);
    CUtensorMap t3;
    if (get_tensor_map_slabs(&t3, buf, 2, 1024, rows / 16, 2048, 128, 2)) { printf("3d map: %s\n", mebt_last_error()); return 1; }
    for (int mode : {0, 3, 4, 5, 6}) {
      cudaLaunchConfig_t cfg = {};
      cfg.blockDim = dim3(64);
      cfg.dynamicSmemBytes = smem_bytes;
      cfg.gridDim = dim3(ctas);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, feed_kernel, tf, tp, ts, t3, mode, 1, iters, total_boxes, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), cyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < ctas; ++i) avg += double(h[i]); avg /= ctas;
      const double per_sm = double(iters) * BOX_BYTES / avg;
      printf("ctas %3d mode %d (%s): %.3f ms | %.1f B/clk/SM, %.0f B/clk chip\n", ctas, mode, mode == 0 ? "1 contiguous 32K box per stage" : mode == 3 ? "2 contiguous 16K boxes per stage" : mode == 4 ? "2 strided [128 x 64] 16K boxes per stage" : mode == 5 ? "1 3-D box {64,128 rows,2 slabs} 32K per stage" : "2 strided 16K boxes, two issuing threads",
             best, per_sm, per_sm * ctas);
    }
  }
  return 0;
}
