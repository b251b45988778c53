// Library-level entry points of libpcops.so: version, error strings, device queries.
#include <mutex>
#include "common.cuh"

namespace pc {
int num_sms() {
  static std::mutex mu;
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}
}  // namespace pc

extern "C" int pc_version(void) { return 100; }

extern "C" int pc_num_sms(void) {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return v;
}

extern "C" const char *pc_error_string(int code) {
  switch (code) {
    case PC_OK: return "PC_OK";
    case PC_ERR_INVALID_ARGUMENT: return "PC_ERR_INVALID_ARGUMENT";
    case PC_ERR_UNSUPPORTED: return "PC_ERR_UNSUPPORTED";
    case PC_ERR_WORKSPACE: return "PC_ERR_WORKSPACE";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "PC_ERR_UNKNOWN";
}
