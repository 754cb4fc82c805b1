// C-ABI housekeeping: thread-local error string, version, device info.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <unordered_set>

#include "pmu_common.cuh"
#include "ctx.cuh"

// ---- the launch context (ctx.cuh) ------------------------------------------------------------------------------------
struct MapKey {
  pmu::TensorMapSpec s;
  bool operator==(const MapKey& o) const { return memcmp(&s, &o.s, sizeof(s)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k.s);
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < sizeof(k.s) / 8; ++i) { h ^= w[i]; h *= 0x100000001b3ull; h ^= h >> 29; }
    return (size_t)h;
  }
};
static_assert(sizeof(pmu::TensorMapSpec) % 8 == 0, "TensorMapSpec is hashed as 64-bit words");

struct pmu_ctx {
  int device = 0, sm_count = 0, cc_major = 0;
  std::mutex mu;
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
  std::unordered_set<const void*> attr_done;
  int64_t hits = 0, misses = 0;
};
static constexpr size_t CTX_MAX_MAPS = 8192;      // ~1.4 MB of descriptors; cleared when full

namespace pmu {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static thread_local pmu_ctx* g_ctx = nullptr;

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int tensor_map(CUtensorMap* out, const TensorMapSpec& spec, const char* what) {
  pmu_ctx* c = g_ctx;
  MapKey key;
  if (c) {
    memset(&key, 0, sizeof(key));                 // padding bytes take part in the comparison
    key.s.ptr = spec.ptr; key.s.rank = spec.rank; key.s.dtype = spec.dtype; key.s.swizzle = spec.swizzle;
    key.s.l2promo = spec.l2promo; key.s.oob = spec.oob;
    for (int i = 0; i < spec.rank; ++i) { key.s.dims[i] = spec.dims[i]; key.s.box[i] = spec.box[i]; }
    for (int i = 0; i + 1 < spec.rank; ++i) key.s.strides[i] = spec.strides[i];
    std::lock_guard<std::mutex> lk(c->mu);
    auto it = c->maps.find(key);
    if (it != c->maps.end()) { *out = it->second; ++c->hits; return PMU_OK; }
  }
  auto fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PMU_ERR_CUDA; }
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < spec.rank; ++i) { dims[i] = spec.dims[i]; box[i] = spec.box[i]; }
  for (int i = 0; i + 1 < spec.rank; ++i) strides[i] = spec.strides[i];
  CUresult r = fn(out, (CUtensorMapDataType)spec.dtype, (cuuint32_t)spec.rank, const_cast<void*>(spec.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)spec.swizzle, (CUtensorMapL2promotion)spec.l2promo,
                  (CUtensorMapFloatOOBfill)spec.oob);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s: rank %d, dims %llu x %llu x %llu x %llu) failed: %d", what, spec.rank,
              (unsigned long long)spec.dims[0], (unsigned long long)(spec.rank > 1 ? spec.dims[1] : 1),
              (unsigned long long)(spec.rank > 2 ? spec.dims[2] : 1), (unsigned long long)(spec.rank > 3 ? spec.dims[3] : 1), (int)r);
    return PMU_ERR_CUDA;
  }
  if (c) {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->maps.size() >= CTX_MAX_MAPS) c->maps.clear();
    c->maps.emplace(key, *out);
    ++c->misses;
  }
  return PMU_OK;
}

int set_max_dyn_smem(const void* kernel, int bytes) {
  pmu_ctx* c = g_ctx;
  if (c) {
    std::lock_guard<std::mutex> lk(c->mu);
    if (c->attr_done.count(kernel)) return PMU_OK;
  }
  PMU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (c) {
    std::lock_guard<std::mutex> lk(c->mu);
    c->attr_done.insert(kernel);
  }
  return PMU_OK;
}

int device_cc_major(int* cc_major) {
  if (g_ctx) { *cc_major = g_ctx->cc_major; return PMU_OK; }
  int dev = 0;
  PMU_CUDA(cudaGetDevice(&dev));
  PMU_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  return PMU_OK;
}

int sm_count() {
  if (g_ctx) return g_ctx->sm_count;
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace pmu

extern "C" const char* pmu_last_error(void) { return pmu::g_err; }

extern "C" int pmu_version(void) { return 100; }  // 0.1.0

extern "C" int pmu_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  PMU_CUDA(cudaGetDevice(&dev));
  if (sm_count) PMU_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) PMU_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) PMU_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return PMU_OK;
}

extern "C" int pmu_set_device(int device) {
  PMU_CHECK_ARG(device >= 0, "pmu_set_device: negative device index");
  pmu_ctx_bind(nullptr);              // the context-free path: every launch queries / encodes what it needs
  PMU_CUDA(cudaSetDevice(device));
  return PMU_OK;
}

extern "C" int pmu_ctx_create(int device, pmu_ctx** out) {
  PMU_CHECK_ARG(out != nullptr && device >= 0, "pmu_ctx_create: bad arguments");
  int n = 0, sm = 0, cc = 0;
  PMU_CUDA(cudaGetDeviceCount(&n));
  PMU_CHECK_ARG(device < n, "pmu_ctx_create: device %d of %d", device, n);
  PMU_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
  PMU_CUDA(cudaDeviceGetAttribute(&cc, cudaDevAttrComputeCapabilityMajor, device));
  pmu_ctx* c = new (std::nothrow) pmu_ctx();
  PMU_CHECK_ARG(c != nullptr, "pmu_ctx_create: out of host memory");
  c->device = device; c->sm_count = sm > 0 ? sm : 148; c->cc_major = cc;
  *out = c;
  return PMU_OK;
}

extern "C" int pmu_ctx_destroy(pmu_ctx* ctx) {
  if (!ctx) return PMU_OK;
  if (pmu::g_ctx == ctx) pmu::g_ctx = nullptr;
  delete ctx;
  return PMU_OK;
}

extern "C" int pmu_ctx_bind(pmu_ctx* ctx) {
  static thread_local int cur_dev = -1;
  if (ctx) {
    if (cur_dev != ctx->device) {
      PMU_CUDA(cudaSetDevice(ctx->device));
      cur_dev = ctx->device;
    }
  } else {
    cur_dev = -1;                     // the caller may switch devices behind our back from here on
  }
  pmu::g_ctx = ctx;
  return PMU_OK;
}

extern "C" int pmu_ctx_stats(pmu_ctx* ctx, int64_t* tensor_maps, int64_t* hits, int64_t* misses) {
  PMU_CHECK_ARG(ctx != nullptr, "pmu_ctx_stats: null context");
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (tensor_maps) *tensor_maps = (int64_t)ctx->maps.size();
  if (hits) *hits = ctx->hits;
  if (misses) *misses = ctx->misses;
  return PMU_OK;
}
