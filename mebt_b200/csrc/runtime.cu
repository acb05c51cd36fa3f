// Host-side runtime shared by every op: error strings, device check, TMA tensor-map cache.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace mebt {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_last_error("CUDA error %d (%s) at %s", int(e), cudaGetErrorString(e), what);
  return MEBT_ERR_CUDA;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

static thread_local int g_grid_cap = 0;
int grid_cap() {
  const int n = sm_count();
  return g_grid_cap > 0 && g_grid_cap < n ? g_grid_cap : n;
}
GridCapScope::GridCapScope(int cap) : prev(g_grid_cap) { if (cap > 0) g_grid_cap = cap; }
GridCapScope::~GridCapScope() { g_grid_cap = prev; }

// ---- launch accounting / per-family event timing ------------------------------------------------
struct ProfRec { cudaEvent_t e0, e1; int family; double work; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static unsigned long long g_launches[FAM_COUNT] = {0};

static cudaEvent_t take_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void note_launch(int family, double work, cudaStream_t st, bool begin) {
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (begin) {
    ++g_launches[family];
    if (g_prof_on) {
      ProfRec r{take_event(), take_event(), family, work};
      cudaEventRecord(r.e0, st);
      g_prof.push_back(r);
    }
  } else if (g_prof_on && !g_prof.empty()) {
    cudaEventRecord(g_prof.back().e1, st);
  }
}

// ---- tensor map cache ------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t inner, outer, stride;
  uint32_t box_inner, box_outer, elem, slabs;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return size_t(h);
  }
};

// 3-D view of a row-major [outer, inner] matrix as [inner / w][outer][w] with w = 128 bytes of elements, so that ONE
// box {w, box_outer, slabs} brings `slabs` adjacent 128-byte-wide column slabs of `box_outer` rows, laid out in shared
// memory slab after slab (each slab = box_outer rows of 128 B, 128B-swizzled).  `inner` must be a multiple of w.
int get_tensor_map_slabs(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                         uint64_t row_stride_bytes, uint32_t box_outer, uint32_t slabs) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const uint32_t w = 128u / uint32_t(elem_bytes);
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.inner = inner; key.outer = outer; key.stride = row_stride_bytes;
  key.box_inner = w; key.box_outer = box_outer; key.elem = uint32_t(elem_bytes); key.slabs = slabs;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return MEBT_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  MEBT_REQUIRE(fn != nullptr, MEBT_ERR_DEVICE, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  MEBT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (row_stride_bytes & 15) == 0, MEBT_ERR_SHAPE,
               "TMA operand must be 16-byte aligned (ptr=%p stride=%llu)", ptr, (unsigned long long)row_stride_bytes);
  MEBT_REQUIRE(inner % w == 0 && box_outer >= 1 && box_outer <= 256 && slabs >= 1 && slabs <= 256, MEBT_ERR_SHAPE,
               "TMA slab view needs inner %% %u == 0 (inner=%llu box_outer=%u slabs=%u)", w, (unsigned long long)inner,
               box_outer, slabs);
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[3] = {w, outer, inner / w};
  cuuint64_t strides[2] = {row_stride_bytes, 128};
  cuuint32_t box[3] = {w, box_outer, slabs};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MEBT_REQUIRE(r == CUDA_SUCCESS, MEBT_ERR_CUDA,
               "cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%ux%u", int(r), ptr,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, w, box_outer, slabs);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return MEBT_OK;
}

int get_tensor_map_2d(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t inner, uint64_t outer,
                      uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.inner = inner; key.outer = outer; key.stride = row_stride_bytes;
  key.box_inner = box_inner; key.box_outer = box_outer; key.elem = uint32_t(elem_bytes);
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return MEBT_OK; }
  }
  EncodeTiledFn fn = encode_fn();
  MEBT_REQUIRE(fn != nullptr, MEBT_ERR_DEVICE, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  MEBT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (row_stride_bytes & 15) == 0, MEBT_ERR_SHAPE,
               "TMA operand must be 16-byte aligned (ptr=%p stride=%llu)", ptr, (unsigned long long)row_stride_bytes);
  MEBT_REQUIRE(box_inner * uint32_t(elem_bytes) == 128 && box_outer >= 1 && box_outer <= 256, MEBT_ERR_SHAPE,
               "TMA box must be 128 bytes wide and <= 256 rows (got %u x %u)", box_inner, box_outer);
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MEBT_REQUIRE(r == CUDA_SUCCESS, MEBT_ERR_CUDA,
               "cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu stride=%llu box=%ux%u", int(r), ptr,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
               box_outer);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return MEBT_OK;
}

// 5-D tiled map over a channels-last activation tensor (csrc/conv3d.cu): dims / box innermost first, 128-byte swizzle,
// element strides > 1 = one element every `stride` along that dimension (the box then holds ceil(box / stride) of them).
// Not cached: a convolution launch is tens of microseconds and the encode is host arithmetic.
int get_tensor_map_5d(CUtensorMap* out, const void* ptr, const uint64_t dims[5], const uint64_t strides_bytes[4],
                      const uint32_t box[5], const uint32_t elem_strides[5]) {
  EncodeTiledFn fn = encode_fn();
  MEBT_REQUIRE(fn != nullptr, MEBT_ERR_DEVICE, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  MEBT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, MEBT_ERR_SHAPE, "TMA operand must be 16-byte aligned (ptr=%p)", ptr);
  cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5]; cuuint32_t e[5];
  for (int i = 0; i < 5; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = elem_strides[i]; }
  for (int i = 0; i < 4; ++i) {
    MEBT_REQUIRE((strides_bytes[i] & 15) == 0, MEBT_ERR_SHAPE, "TMA strides must be multiples of 16 bytes (dim %d: %llu)", i + 1,
                 (unsigned long long)strides_bytes[i]);
    s[i] = strides_bytes[i];
  }
  MEBT_REQUIRE(box[0] * 2 == 128, MEBT_ERR_SHAPE, "TMA box must be 128 bytes wide");
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MEBT_REQUIRE(r == CUDA_SUCCESS, MEBT_ERR_CUDA,
               "cuTensorMapEncodeTiled(5d) failed (%d) dims=%llu,%llu,%llu,%llu,%llu box=%u,%u,%u,%u,%u estr=%u,%u,%u,%u,%u", int(r),
               (unsigned long long)d[0], (unsigned long long)d[1], (unsigned long long)d[2], (unsigned long long)d[3],
               (unsigned long long)d[4], b[0], b[1], b[2], b[3], b[4], e[0], e[1], e[2], e[3], e[4]);
  *out = m;
  return MEBT_OK;
}

}  // namespace mebt

extern "C" {

const char* mebt_last_error(void) { return mebt::g_last_error; }

const char* mebt_version(void) { return "mebt_b200 0.1 (sm_100a)"; }

unsigned long long mebt_launch_count(void) {
  std::lock_guard<std::mutex> g(mebt::g_prof_mu);
  unsigned long long n = 0;
  for (int i = 0; i < mebt::FAM_COUNT; ++i) n += mebt::g_launches[i];
  return n;
}

void mebt_profile_enable(int on) {
  std::lock_guard<std::mutex> g(mebt::g_prof_mu);
  mebt::g_prof_on = on != 0;
}

int mebt_profile_report(double* time_ms, double* work, long long* launches) {
  using namespace mebt;
  MEBT_CUDA_OK(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> g(g_prof_mu);
  for (int i = 0; i < FAM_COUNT; ++i) { time_ms[i] = 0.0; work[i] = 0.0; launches[i] = 0; }
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      time_ms[r.family] += ms;
      work[r.family] += r.work;
      launches[r.family] += 1;
    }
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_prof.clear();
  return MEBT_OK;
}

int mebt_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return mebt::cuda_fail(e, "cudaGetDevice");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    mebt::set_last_error("mebt_b200 needs an sm_100-class GPU (Blackwell B200); found sm_%d%d", major, minor);
    return MEBT_ERR_DEVICE;
  }
  return MEBT_OK;
}

}  // extern "C"
