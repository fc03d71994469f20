// ltr_host.cuh -- host-side helpers shared by the translation units of libltr_sm100.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ltr_sm100.h"

namespace ltr {

extern thread_local int tls_cuda_error;   // defined in ltr_kernels.cu

inline int cuda_fail(cudaError_t e) {
  tls_cuda_error = static_cast<int>(e);
  return LTR_ECUDA;
}
#define LTR_CUDA(call)                                 \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) return cuda_fail(e__);     \
  } while (0)

struct DeviceInfo { int sms; int major; bool ok; };
inline int device_info(DeviceInfo* out) {
  static DeviceInfo cache[64];
  int dev = 0;
  LTR_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return LTR_EUNSUPPORTED;
  if (!cache[dev].ok) {
    int sms = 0, major = 0;
    LTR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LTR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    cache[dev].sms = sms;
    cache[dev].major = major;
    cache[dev].ok = true;
  }
  *out = cache[dev];
  return out->major == 10 ? LTR_OK : LTR_EUNSUPPORTED;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace ltr
