// runtime.cu — error plumbing, device queries and small host utilities of libb200decode.
// get_device_attribute / get_max_shared_memory_per_block_device_attribute replace
// kernels/cuda_utils_kernels.cu of the reference (schema kernels/torch_bindings.cpp:497-504).
#include "common.cuh"

#include <string.h>

namespace b200 {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }
int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

int num_sms() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace b200

extern "C" const char* b200_last_error(void) { return b200::g_err.c_str(); }

extern "C" int b200_abi_version(void) { return 1; }

extern "C" int b200_parse_kv_cache_dtype(const char* s) {
  if (s == nullptr) return -1;
  if (strcmp(s, "auto") == 0) return B200_KV_AUTO;
  if (strcmp(s, "fp8") == 0 || strcmp(s, "fp8_e4m3") == 0) return B200_KV_FP8_E4M3;
  if (strcmp(s, "fp8_e5m2") == 0) return B200_KV_FP8_E5M2;
  return -1;
}

extern "C" int64_t b200_get_device_attribute(int64_t attribute, int64_t device_id) {
  int device = (int)device_id, value = 0;
  if (device < 0) cudaGetDevice(&device);
  cudaDeviceGetAttribute(&value, static_cast<cudaDeviceAttr>(attribute), device);
  return value;
}

extern "C" int64_t b200_get_max_shared_memory_per_block_device_attribute(int64_t device_id) {
  return b200_get_device_attribute((int64_t)cudaDevAttrMaxSharedMemoryPerBlockOptin, device_id);
}
