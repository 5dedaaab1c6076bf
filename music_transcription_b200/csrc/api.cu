// Library plumbing: error strings, device checks, tensor-map encoding.
#include <map>
#include <mutex>

#include "common.cuh"

namespace amt {

unsigned long long launch_count();

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// Per-DEVICE caches: a process may drive several GPUs (one model / frontend handle per device), so nothing about
// "the" device may be cached in a process-wide static.
constexpr int kMaxDevices = 64;
static int g_sms[kMaxDevices] = {0};
static int g_cc_major[kMaxDevices] = {0};       // 0 = not queried yet

static int current_device(int* dev) {
  cudaError_t e = cudaGetDevice(dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(AMT_ERR_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  }
  if (*dev < 0 || *dev >= kMaxDevices) return set_error(AMT_ERR_DEVICE, "device ordinal %d out of range", *dev);
  return 0;
}

static int query_device(int dev) {
  int sms = 0, major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_error(AMT_ERR_DEVICE, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  }
  __atomic_store_n(&g_sms[dev], sms, __ATOMIC_RELAXED);
  __atomic_store_n(&g_cc_major[dev], major, __ATOMIC_RELEASE);
  return 0;
}

int ensure_device() {
  int dev = 0;
  AMT_TRY(current_device(&dev));
  if (__atomic_load_n(&g_cc_major[dev], __ATOMIC_ACQUIRE) == 0) AMT_TRY(query_device(dev));
  if (g_cc_major[dev] != 10)
    return set_error(AMT_ERR_DEVICE, "libamt_sm100 needs a compute-capability 10.x GPU (B200); device %d is %d.x", dev,
                     g_cc_major[dev]);
  return 0;
}

int num_sms() {
  int dev = 0;
  if (current_device(&dev) != 0) return 148;
  if (__atomic_load_n(&g_cc_major[dev], __ATOMIC_ACQUIRE) == 0 && query_device(dev) != 0) return 148;
  return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

struct AttrKey {
  int dev;
  const void* func;
  int attr;
  bool operator<(const AttrKey& o) const {
    if (dev != o.dev) return dev < o.dev;
    if (func != o.func) return func < o.func;
    return attr < o.attr;
  }
};
static std::mutex g_attr_mutex;
static std::map<AttrKey, int> g_attr_values;

int ensure_func_attr(const void* func, cudaFuncAttribute attr, int value) {
  int dev = 0;
  AMT_TRY(current_device(&dev));
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  const AttrKey key{dev, func, static_cast<int>(attr)};
  auto it = g_attr_values.find(key);
  if (it != g_attr_values.end() && it->second >= value) return 0;
  AMT_CUDA(cudaFuncSetAttribute(func, attr, value));
  g_attr_values[key] = value;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int encode_tmap(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      cudaGetLastError();
      return set_error(AMT_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(map, dtype, static_cast<cuuint32_t>(rank),
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(AMT_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]",
                     static_cast<int>(r), rank, (unsigned long long)dims[0],
                     (unsigned long long)(rank > 1 ? dims[1] : 0), (unsigned long long)(rank > 2 ? dims[2] : 0),
                     (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], rank > 1 ? box[1] : 0,
                     rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return 0;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  return encode_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle);
}

int encode_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  return encode_tmap(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle);
}

}  // namespace amt

extern "C" {

const char* amt_version(void) { return "amt-sm100 0.1.0"; }
const char* amt_last_error(void) { return amt::last_error_buf(); }
int amt_device_check(void) { return amt::ensure_device(); }
uint64_t amt_launch_count(void) { return amt::launch_count(); }

}  // extern "C"
