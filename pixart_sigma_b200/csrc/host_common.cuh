// Host-side helpers shared by the entry points: error reporting, device check, TMA tensor-map encoding
// (driver entry point fetched at run time so the library does not link libcuda), launch counter.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/pixart_sm100.h"

namespace pxa {

inline char* last_error_buf() {
  static thread_local char buf[512] = "";
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
inline std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

struct DeviceInfo {
  int ok = 0;      // 1 when the current device is sm_100
  int sms = 0;
  int major = 0, minor = 0;
};
inline const DeviceInfo& device_info() {
  static thread_local DeviceInfo info;
  static thread_local int cached_dev = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    info = DeviceInfo{};
    return info;
  }
  if (dev != cached_dev) {
    cudaDeviceProp p;
    info = DeviceInfo{};
    if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) {
      info.major = p.major;
      info.minor = p.minor;
      info.sms = p.multiProcessorCount;
      info.ok = (p.major == 10);
    }
    cached_dev = dev;
  }
  return info;
}
#define PXA_REQUIRE_SM100()                                                                           \
  do {                                                                                                \
    const ::pxa::DeviceInfo& _di = ::pxa::device_info();                                              \
    if (!_di.ok)                                                                                      \
      return ::pxa::fail(PXA_ERR_ARCH, "libpixart_sm100 needs an sm_100 (B200) device, found sm_%d%d", \
                         _di.major, _di.minor);                                                       \
  } while (0)

#define PXA_CHECK_CUDA(expr)                                                                        \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return ::pxa::fail(PXA_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 tensor map of rank `rank` (<= 4). dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of
// dims[i+1]. OOB elements read as zero.
inline int make_tmap(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PXA_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (reinterpret_cast<uintptr_t>(base) & 15) return fail(PXA_ERR_ALIGN, "tensor base %p not 16-byte aligned", base);
  cuuint64_t d[4];
  cuuint64_t s[3];
  cuuint32_t b[4], e[4];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = 1;
    if (i + 1 < rank) {
      s[i] = strides_bytes[i];
      if (s[i] & 15) return fail(PXA_ERR_ALIGN, "tensor stride %llu bytes not a multiple of 16", (unsigned long long)s[i]);
    }
  }
  CUresult r = fn(map, dtype, rank, const_cast<void*>(base), d, s, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PXA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return PXA_OK;
}

inline int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  return make_tmap(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swz);
}

}  // namespace pxa
