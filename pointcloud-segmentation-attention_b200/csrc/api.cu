// Library-level entry points of libpcops.so: version, error strings, device queries.
#include <atomic>
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

cudaError_t allow_smem_cached(const void *kernel, size_t bytes) {
  struct Entry { const void *k; int dev; size_t bytes; };
  static std::mutex mu;
  static Entry table[256];
  static int used = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  Entry *slot = nullptr;
  for (int i = 0; i < used; ++i)
    if (table[i].k == kernel && table[i].dev == dev) { slot = &table[i]; break; }
  if (slot && slot->bytes >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  if (slot) slot->bytes = bytes;
  else if (used < 256) table[used++] = Entry{kernel, dev, bytes};
  return cudaSuccess;
}

static std::atomic<int> g_concurrency_hint{1};
int concurrency_hint() { return g_concurrency_hint.load(std::memory_order_relaxed); }

int resident_grid(const void *kernel, int threads, size_t smem, size_t needed) {
  // occupancy per (kernel, threads, smem), cached: cudaOccupancyMaxActiveBlocksPerMultiprocessor costs microseconds
  struct Entry { const void *k; int threads; size_t smem; int occ; };
  static std::mutex mu;
  static Entry table[64];
  static int used = 0;
  int occ = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < used; ++i)
      if (table[i].k == kernel && table[i].threads == threads && table[i].smem == smem) { occ = table[i].occ; break; }
    if (occ == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ <= 0) occ = 1;
      if (used < 64) table[used++] = Entry{kernel, threads, smem, occ};
    }
  }
  const int h = concurrency_hint();
  const int per_sm = occ / h > 1 ? occ / h : 1;
  size_t g = (size_t)num_sms() * (size_t)per_sm;
  if (g > needed) g = needed;
  return g < 1 ? 1 : (int)g;
}
}  // namespace pc

extern "C" int pc_version(void) { return 100; }

extern "C" int pc_set_concurrency_hint(int n) {
  if (n < 1) return PC_ERR_INVALID_ARGUMENT;
  pc::g_concurrency_hint.store(n, std::memory_order_relaxed);
  return PC_OK;
}

extern "C" int pc_get_concurrency_hint(void) { return pc::concurrency_hint(); }

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
