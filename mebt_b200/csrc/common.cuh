// Common device/host helpers for the mebt_b200 sm_100a kernels.
//
// Everything here is hand-written inline PTX for Blackwell (sm_100a): mbarrier,
// TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the shared-memory
// matrix descriptors they consume.  No CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>          // CUtensorMap (types only; the driver symbol is fetched at run time)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mebt_b200.h"   // MEBT_OK / MEBT_ERR_* codes

namespace mebt {

// ----------------------------------------------------------------------------------------------
// Host-side error plumbing (C ABI returns int codes; message kept per thread)
// ----------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MEBT_CUDA_OK(expr)                                         \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::mebt::cuda_fail(_e, #expr);    \
  } while (0)

#define MEBT_REQUIRE(cond, code, ...)                              \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::mebt::set_last_error(__VA_ARGS__);                         \
      return (code);                                               \
    }                                                              \
  } while (0)

#define MEBT_LAUNCH_OK(what)                                       \
  do {                                                             \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return ::mebt::cuda_fail(_e, what);     \
  } while (0)

// 2-D bf16/fp32 tensor map with a 128-byte swizzle.  `inner` is the contiguous dimension.
// Cached by (ptr, dims, stride, box, elem size).
int get_tensor_map_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer);

// 3-D slab view: one box = `slabs` adjacent 128-byte-wide column slabs of `box_outer` rows (see runtime.cu)
int get_tensor_map_slabs(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                         uint64_t row_stride_bytes, uint32_t box_outer, uint32_t slabs);

int sm_count();
// SMs a persistent kernel launched now may occupy: sm_count() unless a GridCapScope is open on this thread (the training
// backward leaves a share of the SMs to its side-stream weight-gradient launch)
int grid_cap();
struct GridCapScope {
  int prev;
  explicit GridCapScope(int cap);
  ~GridCapScope();
};
// cudaFuncSetAttribute is per device: `seen` is the per-call-site record of the devices a kernel was configured on
inline bool first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (seen[dev]) return false;
  seen[dev] = true;
  return true;
}

// ---- lightweight in-library accounting (bench.py's gpu_launches and per-kernel roofline numbers) ----------
enum KernelFamily : int {
  FAM_GEMM = 0, FAM_ATTENTION = 1, FAM_LAYERNORM = 2, FAM_EMBED = 3, FAM_SAMPLE = 4, FAM_CE = 5, FAM_REMASK = 6,
  FAM_SCATTER = 7, FAM_VQ = 8, FAM_OTHER = 9, FAM_COUNT = 10
};
void note_launch(int family, double work, cudaStream_t st, bool begin);
// RAII: counts the launch and, when profiling is enabled, brackets it with CUDA events on its stream.
struct LaunchScope {
  int family; cudaStream_t st;
  LaunchScope(int fam, double work, cudaStream_t s) : family(fam), st(s) { note_launch(fam, work, s, true); }
  ~LaunchScope() { note_launch(family, 0.0, st, false); }
};

#ifdef __CUDACC__
// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining,
// so its launch latency and prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the predecessor's
// tail.  Every kernel launched this way calls griddep_wait() before it touches global memory.
// An event recorded between a kernel and its programmatic dependent is not signalled until the dependent has been
// scheduled behind it; a stream that forks work to other streams through such an event therefore asks for its NEXT
// launch to be an ordinary (fully serialised) one: pdl_fence(stream).
struct PdlFence { cudaStream_t stream; bool armed; };
inline PdlFence& pdl_fence_state() { static thread_local PdlFence f{nullptr, false}; return f; }
inline void pdl_fence(cudaStream_t st) { pdl_fence_state() = PdlFence{st, true}; }
inline int pdl_allowed(cudaStream_t st) {
  PdlFence& f = pdl_fence_state();
  if (f.armed && f.stream == st) { f.armed = false; return 0; }
  return 1;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_allowed(st);
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Same, as 2-CTA thread-block clusters (grid.x must be even; CTAs 2c and 2c+1 form cluster c).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster2(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_allowed(st);
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------------------------
// Device helpers
// ----------------------------------------------------------------------------------------------
#ifndef MEBT_WAIT_BACKOFF_NS
#define MEBT_WAIT_BACKOFF_NS 40
#endif
#ifndef MEBT_SPIN_LIMIT
#define MEBT_SPIN_LIMIT (1u << 26)   // a stuck pipeline traps instead of hanging the GPU
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte shared-memory accesses by 32-bit shared-window address (LDS/STS instead of generic LD/ST)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// Wait until every grid this launch depends on has completed and its writes are visible, then let the next
// kernel in the stream start launching (it will block in its own griddep_wait until this grid completes).
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > MEBT_SPIN_LIMIT) __trap();
  }
}

// Same for the single-thread producer / MMA-issuer roles: back off between polls so that the spinning warp does not
// take issue slots from the compute warps that share its SM sub-partition.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (MEBT_WAIT_BACKOFF_NS > 0) __nanosleep(MEBT_WAIT_BACKOFF_NS);
    if (++spins > MEBT_SPIN_LIMIT) __trap();
  }
}

// ---- TMA ----
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// tile load global -> shared, completion on an mbarrier (bytes). c0 = inner coordinate, c1 = outer.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 3-D (slab view) tile load: c0 = element inside the 128-byte slab (0), c1 = row, c2 = slab index
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// multicast tile load: the box lands at the same shared-memory offset in every CTA of `cta_mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx for the bytes it got
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// ---- SM-pair (cta_group::2) variants ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {   // same offset in CTA `cta_rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {       // arrive on a (possibly remote) barrier
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// tile load into THIS CTA's shared memory whose completion is signalled on a barrier given as a shared::cluster
// address (the leader CTA's barrier in pair mode)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t cluster_bar_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(cluster_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// D (256 x N, 128 rows in each CTA's TMEM) (+)= A (128 rows from each CTA) * B (N/2 rows from each CTA)
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// C (global, through the tensor map's dtype) += tile in shared memory
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// at most N of this thread's most recent bulk groups may still be reading their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// 1-D bulk copy global -> shared (no tensor map; 16-byte aligned, size a multiple of 16), completing on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// asynchronous request of `bytes` (multiple of 16, 16-byte aligned) of global memory into L2
__device__ __forceinline__ void l2_prefetch_bulk(const void* gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.commit: the mbarrier receives one arrival once all previously issued MMAs of this thread retire.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// same, arriving on the mbarrier at this offset in every CTA of `cta_mask` (cluster multicast)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// The four K = 16 steps of one 64-wide k-block as ONE instruction sequence.  The single issuing thread is what bounds small
// tiles (measured: ~157 clk per tcgen05.mma when every descriptor is rebuilt from a byte address - shift, mask, or, two
// 32-bit moves per operand - and every instruction sits in its own elect / branch wrapper, against 64 clk of tensor
// time for M128 N128 K16), so the descriptors are split into a loop-invariant high word and a low word that advances
// by a constant per step: (address >> 4) | (LBO >> 4) << 16, + a_step / b_step.
__host__ __device__ constexpr uint32_t smem_desc_hi_sw128(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
template <bool TWO_SM>
__device__ __forceinline__ void umma_bf16_ss_x4(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_step,
                                                uint32_t b_step, uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate_first) {
#define MEBT_MMA4(CG)                                                                     \
  asm volatile(                                                                           \
      "{\n"                                                                               \
      ".reg .pred p, t;\n"                                                                \
      ".reg .b64 da, db;\n"                                                               \
      ".reg .b32 al, bl;\n"                                                               \
      "setp.ne.b32 p, %8, 0;\n"                                                           \
      "setp.eq.b32 t, 0, 0;\n"                                                            \
      "mov.b64 da, {%1, %5};\n"                                                           \
      "mov.b64 db, {%2, %6};\n"                                                           \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, p;\n"                    \
      "add.u32 al, %1, %3;\n"                                                             \
      "add.u32 bl, %2, %4;\n"                                                             \
      "mov.b64 da, {al, %5};\n"                                                           \
      "mov.b64 db, {bl, %6};\n"                                                           \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, t;\n"                    \
      "add.u32 al, al, %3;\n"                                                             \
      "add.u32 bl, bl, %4;\n"                                                             \
      "mov.b64 da, {al, %5};\n"                                                           \
      "mov.b64 db, {bl, %6};\n"                                                           \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, t;\n"                    \
      "add.u32 al, al, %3;\n"                                                             \
      "add.u32 bl, bl, %4;\n"                                                             \
      "mov.b64 da, {al, %5};\n"                                                           \
      "mov.b64 db, {bl, %6};\n"                                                           \
      "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %7, t;\n"                    \
      "}\n" ::"r"(tmem_d),                                                                \
      "r"(a_lo), "r"(b_lo), "r"(a_step), "r"(b_step), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate_first) \
      : "memory")
  if constexpr (TWO_SM) MEBT_MMA4("2"); else MEBT_MMA4("1");
#undef MEBT_MMA4
}

// Eight K = 16 steps with the A operand in tensor memory (attention's O += P V over one 128-key tile): A advances by
// a_step TMEM columns per step, B's descriptor low word by b_step.
__device__ __forceinline__ void umma_bf16_ts_x8(uint32_t tmem_d, uint32_t tmem_a, uint32_t a_step, uint32_t b_lo,
                                                uint32_t b_step, uint32_t b_hi, uint32_t idesc, uint32_t accumulate_first) {
  asm volatile(
      "{\n"
      ".reg .pred p, t;\n"
      ".reg .b64 db;\n"
      ".reg .b32 ta, bl;\n"
      "setp.ne.b32 p, %7, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "mov.b64 db, {%3, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %6, p;\n"
      "add.u32 ta, %1, %2;\n add.u32 bl, %3, %4;\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %6, t;\n"
      "add.u32 ta, ta, %2;\n add.u32 bl, bl, %4;\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %6, t;\n"
      "add.u32 ta, ta, %2;\n add.u32 bl, bl, %4;\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %6, t;\n"
      "add.u32 ta, ta, %2;\n add.u32 bl, bl, %4;\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %6, t;\n"
      "add.u32 ta, ta, %2;\n add.u32 bl, bl, %4;\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %6, t;\n"
      "add.u32 ta, ta, %2;\n add.u32 bl, bl, %4;\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %6, t;\n"
      "add.u32 ta, ta, %2;\n add.u32 bl, bl, %4;\n mov.b64 db, {bl, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %6, t;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(a_step), "r"(b_lo), "r"(b_step), "r"(b_hi), "r"(idesc), "r"(accumulate_first)
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]: the A operand (M x K, K-major) is read from tensor memory - row m in lane m,
// two bf16 per 32-bit column, element k in column k / 2 (low half = even k).  A K = 16 step consumes 8 columns.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D (cute::UMMA::InstrDescriptor layout):
//  [4,6) D fmt (1 = f32) | [7,10) A fmt (1 = bf16) | [10,13) B fmt | bit 15 A major | bit 16 B major
//  [17,23) N >> 3 | [24,29) M >> 4.   major: 0 = K-major (reduction dim contiguous), 1 = MN-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// Shared-memory matrix descriptor, 128-byte swizzle (layout type 2), sm_100 version field = 1.
//  [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version | [61,64) layout type
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

// TMEM -> registers: 32 lanes (this warp's quarter) x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: the mirror image of tmem_ld_32x32
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {   // 32 lanes x 16 columns
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but naming the registers an earlier (still in flight) tcgen05.ld targets as read-write operands, so the
// compiler cannot schedule a use of them above the wait when other work sits between the load and the wait.
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// ---- dropout (training): counter-based keep decisions, identical in forward and backward ----
// One 32-bit hash per PAIR of adjacent elements (16 bits each); keep iff bits >= thr.  The hash is the "lowbias32"
// integer finaliser applied to (row key + pair index * golden ratio); the row key mixes the 64-bit site seed with the
// row id.  Reference: nn.Dropout / F.dropout (mebt/modules/gpt.py:112-113,140,150-155,239-242): y = x * keep / (1-p).
struct DropKey {
  uint32_t k0, k1;     // site seed
  uint32_t thr;        // round(p * 65536); 0 = dropout off
  float inv_keep;      // 65536 / (65536 - thr)
};
inline DropKey make_drop_key(float p, unsigned long long seed, unsigned long long site) {
  DropKey k;
  unsigned long long s = seed + (site + 1) * 0x9E3779B97F4A7C15ull;
  s ^= s >> 30; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 27; s *= 0x94D049BB133111EBull; s ^= s >> 31;   // splitmix64
  k.k0 = uint32_t(s);
  k.k1 = uint32_t(s >> 32);
  long t = lrintf(p * 65536.f);
  k.thr = uint32_t(t < 0 ? 0 : (t > 65535 ? 65535 : t));
  k.inv_keep = 65536.f / float(65536u - k.thr);
  return k;
}
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_row_key(const DropKey& k, uint32_t row_id) {
  return mix32(k.k0 ^ mix32(row_id + k.k1));
}
// keep factors (inv_keep or 0) of elements 2*pair and 2*pair+1 of the row
__host__ __device__ __forceinline__ void drop_pair(const DropKey& k, uint32_t row_key, uint32_t pair, float& f0, float& f1) {
  const uint32_t h = mix32(row_key + pair * 0x9E3779B9u);
  f0 = (h & 0xffffu) >= k.thr ? k.inv_keep : 0.f;
  f1 = (h >> 16) >= k.thr ? k.inv_keep : 0.f;
}

// ---- AdamW element update: torch.optim.AdamW(fused=True) arithmetic (decoupled decay, bias-corrected moments), shared by
// adamw_flat_kernel and the fused epilogue of the grouped weight-gradient GEMM so that both produce the same bits ----
struct AdamScalars {
  float lr, beta1, beta2, eps, wd;
  float step_size;        // lr / (1 - beta1^step)
  float inv_bc2_sqrt;     // 1 / sqrt(1 - beta2^step)
};
__device__ __forceinline__ void adamw_element(float& p, const float g, float& m, float& v, const float keep, const AdamScalars& a) {
  p *= keep;
  m = m + (1.f - a.beta1) * (g - m);
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  const float denom = sqrtf(v) * a.inv_bc2_sqrt + a.eps;
  p -= a.step_size * (m / denom);
}

// ---- small numeric helpers ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// 2^x on the SFU (MUFU.EX2), flush-to-zero, no range fix-up code: x <= 0 on the softmax paths that use it
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(x) on the SFU (MUFU.LG2)
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exact (erf) GELU, as torch.nn.GELU() default
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx of the erf GELU: Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * __expf(-0.5f * x * x);
}
// ---- packed (f32x2) GELU for the GEMM epilogues ----
// erf(x / sqrt(2)) = x * P(x^2) / Q(x^2) on |x| <= 4 sqrt(2) (the rational minimax form used by Eigen/XLA for float
// erf, rescaled so that it takes x rather than x / sqrt(2)); |erf error| < 5e-7, |gelu error| < 2e-6 over all x — the
// result is rounded to bf16 (2^-9 relative) right after.  Two elements per FFMA2: 11 issue slots per element against
// ~28 for the erff() path, which made the fc1 epilogue as long as its main loop.
namespace gelu_detail {
constexpr double kS = 0.70710678118654752440;
constexpr float kA6 = float(-2.72614225801306e-10 * kS / 64.0), kA5 = float(2.77068142495902e-08 * kS / 32.0),
                kA4 = float(-2.10102402082508e-06 * kS / 16.0), kA3 = float(-5.69250639462346e-05 * kS / 8.0),
                kA2 = float(-7.34990630326855e-04 * kS / 4.0), kA1 = float(-2.95459980854025e-03 * kS / 2.0),
                kA0 = float(-1.60960333262415e-02 * kS);
constexpr float kB4 = float(-1.45660718464996e-05 / 16.0), kB3 = float(-2.13374055278905e-04 / 8.0),
                kB2 = float(-1.68282697438203e-03 / 4.0), kB1 = float(-7.37332916720468e-03 / 2.0),
                kB0 = float(-1.42647390514189e-02);
constexpr float kClamp = 5.65685424949238f;   // 4 sqrt(2)
}  // namespace gelu_detail
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
// erf(x / sqrt(2)) for two values
__device__ __forceinline__ float2 erf_rsqrt2_x2(float2 x) {
  using namespace gelu_detail;
  x.x = fminf(fmaxf(x.x, -kClamp), kClamp);
  x.y = fminf(fmaxf(x.y, -kClamp), kClamp);
  const float2 x2 = __fmul2_rn(x, x);
  float2 p = __ffma2_rn(splat2(kA6), x2, splat2(kA5));
  p = __ffma2_rn(p, x2, splat2(kA4));
  p = __ffma2_rn(p, x2, splat2(kA3));
  p = __ffma2_rn(p, x2, splat2(kA2));
  p = __ffma2_rn(p, x2, splat2(kA1));
  p = __ffma2_rn(p, x2, splat2(kA0));
  p = __fmul2_rn(p, x);
  float2 q = __ffma2_rn(splat2(kB4), x2, splat2(kB3));
  q = __ffma2_rn(q, x2, splat2(kB2));
  q = __ffma2_rn(q, x2, splat2(kB1));
  q = __ffma2_rn(q, x2, splat2(kB0));
  float2 rq;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq.x) : "f"(q.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rq.y) : "f"(q.y));
  return __fmul2_rn(p, rq);
}
__device__ __forceinline__ float2 gelu_erf_x2(float2 x) {
  const float2 h = __fmul2_rn(x, splat2(0.5f));
  return __ffma2_rn(h, erf_rsqrt2_x2(x), h);              // 0.5 x (1 + erf)
}
// d/dx of the erf GELU for two values: Phi(x) + x phi(x)
__device__ __forceinline__ float2 gelu_erf_grad_x2(float2 x) {
  const float2 e = erf_rsqrt2_x2(x);
  const float2 x2 = __fmul2_rn(x, x);
  float2 g;                                                // exp(-x^2 / 2) = 2^(-x^2 * log2(e) / 2)
  g.x = ex2_approx(x2.x * -0.72134752044448170368f);
  g.y = ex2_approx(x2.y * -0.72134752044448170368f);
  const float2 phi = __fmul2_rn(__fmul2_rn(x, splat2(0.39894228040143267794f)), g);
  return __fadd2_rn(__ffma2_rn(e, splat2(0.5f), splat2(0.5f)), phi);
}
#endif  // __CUDACC__

}  // namespace mebt
