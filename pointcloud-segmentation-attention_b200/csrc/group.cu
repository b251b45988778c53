// gather_point and group_point for sm_100a: pure index gathers, HBM-write bound.
//
// Replaces gatherpointKernel (reference tf_ops/sampling/tf_sampling_g.cu:172-181) and group_point_gpu
// (tf_ops/grouping/tf_grouping_g.cu:40-57).  The reference walks one query per thread with serial nsample x c inner
// loops (warp accesses strided by nsample*c*4 bytes); here consecutive threads write consecutive 16-byte pieces of
// the output, so stores are fully coalesced 128-bit transactions and the source rows (a few MB per batch) are
// served from L2.
#include "common.cuh"

namespace pc {
namespace {

// c % 4 == 0 and 16-byte aligned bases: one float4 per element, four independent gathers in flight per thread
// (index loads first, then the four row loads, then the four coalesced 128-bit stores).  I = uint32_t with
// multiply-shift index decoding when the output has < 2^31 vectors, size_t with real divisions beyond.
constexpr int kGroupUnroll = 4;
template <class I> struct Decode;
template <> struct Decode<uint32_t> {
  FastDiv a, b;
  Decode(size_t da, size_t db) : a((uint32_t)da), b((uint32_t)db) {}
  __device__ __forceinline__ uint32_t by_a(uint32_t x) const { return a.div(x); }
  __device__ __forceinline__ uint32_t by_b(uint32_t x) const { return b.div(x); }
  __device__ __forceinline__ uint32_t da() const { return a.d; }
};
template <> struct Decode<size_t> {
  size_t a, b;
  Decode(size_t da, size_t db) : a(da ? da : 1), b(db ? db : 1) {}
  __device__ __forceinline__ size_t by_a(size_t x) const { return x / a; }
  __device__ __forceinline__ size_t by_b(size_t x) const { return x / b; }
  __device__ __forceinline__ size_t da() const { return a; }
};

template <class I>
__global__ void __launch_bounds__(256)
group_vec4_kernel(I total_vec, Decode<I> dec /* a = c4, b = rows per scene */, size_t src_rows_per_scene,
                  const float4 *__restrict__ points, const int *__restrict__ idx, float4 *__restrict__ out) {
  const I step = (I)gridDim.x * blockDim.x;
  for (I v0 = (I)blockIdx.x * blockDim.x + threadIdx.x; v0 < total_vec; v0 += step * kGroupUnroll) {
    const float4 *src[kGroupUnroll];
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) {
      const I v = v0 + u * step;
      src[u] = points;
      if (v < total_vec) {
        const I row = dec.by_a(v);
        const I q = v - row * dec.da();
        const I scene = dec.by_b(row);
        src[u] = points + ((size_t)scene * src_rows_per_scene + __ldg(idx + row)) * dec.da() + q;
      }
    }
    float4 val[kGroupUnroll];
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) val[u] = __ldg(src[u]);
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) {
      const I v = v0 + u * step;
      if (v < total_vec) __stcs(out + v, val[u]);  // streaming store: the grouped tensor is consumed by a later kernel
    }
  }
}

// Any c / alignment: each thread produces 4 consecutive output floats (possibly from 2-4 different rows) and
// stores them as one 128-bit word when the output base allows it.
template <class I>
__global__ void __launch_bounds__(256)
group_any_kernel(I total, Decode<I> dec /* a = c, b = rows per scene */, size_t src_rows_per_scene, bool vec_store,
                 const float *__restrict__ points, const int *__restrict__ idx, float *__restrict__ out) {
  const I nchunks = (total + 3) / 4;
  const I c = dec.da();
  for (I ch = (I)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += (I)gridDim.x * blockDim.x) {
    const I e0 = ch * 4;
    float v[4];
    I row = dec.by_a(e0);
    I l = e0 - row * c;
    int src = (row * c < total) ? __ldg(idx + row) : 0;
    I scene = dec.by_b(row);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (e0 + t < total) v[t] = __ldg(points + ((size_t)scene * src_rows_per_scene + src) * c + l);
      else v[t] = 0.0f;
      if (++l == c) {
        l = 0;
        ++row;
        if (row * c < total) { src = __ldg(idx + row); scene = dec.by_b(row); }
      }
    }
    if (vec_store && e0 + 3 < total) {
      *reinterpret_cast<float4 *>(out + e0) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (e0 + t < total) out[e0 + t] = v[t];
    }
  }
}

template <class I>
void launch_gather(size_t total, size_t rows, size_t src_rows, int c, const float *points, const int *idx, float *out,
                   cudaStream_t st) {
  if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
    const size_t nv = total / 4;
    const int blocks = resident_grid((const void *)group_vec4_kernel<I>, 256, 0, (nv + 256 * kGroupUnroll - 1) / (256 * kGroupUnroll));
    group_vec4_kernel<I><<<blocks, 256, 0, st>>>((I)nv, Decode<I>((size_t)c / 4, rows), src_rows, (const float4 *)points, idx,
                                                 (float4 *)out);
  } else {
    const size_t nch = (total + 3) / 4;
    const int blocks = resident_grid((const void *)group_any_kernel<I>, 256, 0, (nch + 255) / 256);
    group_any_kernel<I><<<blocks, 256, 0, st>>>((I)total, Decode<I>((size_t)c, rows), src_rows, aligned16(out), points, idx, out);
  }
}

int gather_rows(size_t scenes, size_t src_rows, size_t rows, int c, const float *points, const int *idx, float *out,
                cudaStream_t st) {
  const size_t total = scenes * rows * (size_t)c;
  if (total == 0) return PC_OK;
  if (total < (1ull << 31)) launch_gather<uint32_t>(total, rows, src_rows, c, points, idx, out, st);
  else launch_gather<size_t>(total, rows, src_rows, c, points, idx, out, st);
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace
}  // namespace pc

extern "C" int pc_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out,
                               pc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0) return PC_OK;
  if (n == 0 || !inp || !idx || !out) return PC_ERR_INVALID_ARGUMENT;
  return pc::gather_rows((size_t)b, (size_t)n, (size_t)m, 3, inp, idx, out, (cudaStream_t)stream);
}

extern "C" int pc_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                              float *out, pc_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || m < 0 || nsample < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0 || nsample == 0 || c == 0) return PC_OK;
  if (n == 0 || !points || !idx || !out) return PC_ERR_INVALID_ARGUMENT;
  return pc::gather_rows((size_t)b, (size_t)n, (size_t)m * nsample, c, points, idx, out, (cudaStream_t)stream);
}
