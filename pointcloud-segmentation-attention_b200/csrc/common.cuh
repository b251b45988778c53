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

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE fp32 results per lane per instruction) ----
// Used by the pair-test-heavy kernels (FPS, ball query, three_nn, kNN) to halve their FP-pipe instruction count while
// keeping every operation individually rounded.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even
// under --fmad=false, so the un-fused add is spelled fma(a, 1, b) with the 1 held in a register ptxas cannot see
// through (a kernel argument): round(a*1 + b) == round(a + b) exactly, and a product can never be folded into the
// multiplicand or the addend of an existing fma.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// Two squared distances at once, (dx*dx + dy*dy) + dz*dz per half, every step rounded on its own.
// `one2` must be pack2(one, one) with `one` == 1.0f coming from a kernel argument.
__device__ __forceinline__ f32x2 sqdist3_x2(f32x2 ax, f32x2 ay, f32x2 az, f32x2 bx, f32x2 by, f32x2 bz, f32x2 one2) {
  const f32x2 dx = sub2(ax, bx), dy = sub2(ay, by), dz = sub2(az, bz);
  return fma2(fma2(mul2(dx, dx), one2, mul2(dy, dy)), one2, mul2(dz, dz));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Exact floor(x / d) for 0 <= x < 2^31 as one widening multiply and a shift (s = 31 + ceil(log2 d),
// m = ceil(2^s / d) < 2^32).  The flat-index kernels (gathers, interpolation) decode (row, channel, scene) from a
// linear element index; a 64-bit hardware-less division there costs more issue slots than the gather itself.
struct FastDiv {
  uint32_t d, m, s;
  FastDiv() : d(1), m(0x80000000u), s(31) {}
  explicit FastDiv(uint32_t dd) : d(dd ? dd : 1) {
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    s = 31 + l;
    m = (uint32_t)(((1ull << s) + d - 1) / d);
  }
  __device__ __forceinline__ uint32_t div(uint32_t x) const { return (uint32_t)(((unsigned long long)x * m) >> s); }
};
// Grid size for a grid-stride kernel: every SM filled to the kernel's real occupancy, never more CTAs than `needed`
// (a second, partially filled wave costs a short HBM-bound kernel up to half its run time); with a concurrency hint h
// the occupancy is divided by h (at least one CTA per SM): co-resident kernels of other batches fill the rest.
int resident_grid(const void *kernel, int threads, size_t smem, size_t needed);

// SM count of the current device, cached per device id.
int num_sms();
// pc_set_concurrency_hint(): independent launches the caller keeps in flight (>= 1).
int concurrency_hint();
// Opt a kernel into > 48 KB dynamic shared memory (idempotent; cheap).
// The attribute is set once per (kernel, device) and raised only when a larger size is asked for, so steady-state
// calls (and calls made while a stream is being captured into a CUDA graph) touch no driver state.
cudaError_t allow_smem_cached(const void *kernel, size_t bytes);
template <class K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
  return allow_smem_cached((const void *)kernel, bytes);
}

// Largest float s with max(sqrtf(s), 1e-20f) < radius, or -1 if there is none (ball_query.cu): the exact form of the
// reference's hit test max(sqrtf(d2),1e-20f) < radius (tf_grouping_g.cu:24-25) as a compare on d2.
float ball_threshold(float radius);

// ---- CSR inverse of an index tensor + deterministic segmented reductions (segreduce.cu) ----
// Workspace per call: b * (nkeys + 1 + npos) int32.
size_t csr_workspace_bytes(int b, int nkeys, int npos);
// Builds, per scene, row_ptr (nkeys+1) and list (npos): list[row_ptr[i]..row_ptr[i+1]) = positions p with
// idx[p] == i, ascending.  Returns PC_OK or an error code.
int csr_build(int b, int nkeys, int npos, const int *idx, int *workspace, cudaStream_t stream);
// out[scene, i, :] = sum over list entries p (ascending) of src[scene, p / div, :] * (w ? w[scene, p] : 1)
int csr_reduce(int b, int nkeys, int npos, int c, int div, const float *src, const float *w, const int *workspace,
               float *out, cudaStream_t stream);


// ---- tcgen05 Dense engine (gemm_tf32.cu), for kernels that need a projection of their own ----
size_t dense_image_bytes(int K, int N);
// image of W: element (k, n) = w[k * sk + n * sn]
int dense_prepare(int K, int N, size_t sk, size_t sn, const float *w, void *image, cudaStream_t st);
// out (rows, N; stride ldo) = act(x (rows, K; stride ldx) . W + bias); four_products adds the lo x lo term of the TF32 split
// (fp32-grade products: for small projections whose result feeds an exponent, e.g. the attention query)
int dense_forward(size_t rows, int K, size_t ldx, int N, size_t ldo, int relu, const float *x, const void *image,
                  const float *bias, float *out, cudaStream_t st, bool four_products = false);

}  // namespace pc
