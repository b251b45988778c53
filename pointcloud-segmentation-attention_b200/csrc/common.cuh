// Shared device/host helpers for libpcops.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pcops.h"

#define PC_FULL_MASK 0xffffffffu

// Launch epilogue: report launch-configuration errors of THIS call without synchronising.
#define PC_RETURN_LAUNCH_STATUS()            \
  do {                                       \
    cudaError_t e__ = cudaPeekAtLastError(); \
    if (e__ != cudaSuccess) {                \
      cudaGetLastError();                    \
      return (int)e__;                       \
    }                                        \
    return PC_OK;                            \
  } while (0)

#define PC_CUDA_TRY(expr)                    \
  do {                                       \
    cudaError_t e__ = (expr);                \
    if (e__ != cudaSuccess) return (int)e__; \
  } while (0)

namespace pc {

// Squared distance in the reference's written order, (dx*dx + dy*dy) + dz*dz, with every
// operation individually rounded (the __f*_rn intrinsics are never contracted into FMA).
// Follows tf_sampling_g.cu:142, tf_grouping_g.cu:24, tf_interpolate.cpp:73.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// SM count of the current device, cached per device id.
int num_sms();
// Opt a kernel into > 48 KB dynamic shared memory (idempotent; cheap).
template <class K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// ---- CSR inverse of an index tensor + deterministic segmented reductions (segreduce.cu) ----
// Workspace per call: b * (nkeys + 1 + npos) int32.
size_t csr_workspace_bytes(int b, int nkeys, int npos);
// Builds, per scene, row_ptr (nkeys+1) and list (npos): list[row_ptr[i]..row_ptr[i+1]) = positions p with
// idx[p] == i, ascending.  Returns PC_OK or an error code.
int csr_build(int b, int nkeys, int npos, const int *idx, int *workspace, cudaStream_t stream);
// out[scene, i, :] = sum over list entries p (ascending) of src[scene, p / div, :] * (w ? w[scene, p] : 1)
int csr_reduce(int b, int nkeys, int npos, int c, int div, const float *src, const float *w, const int *workspace,
               float *out, cudaStream_t stream);

}  // namespace pc
