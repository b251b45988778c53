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
// (index loads first, then the four row loads, then the four coalesced 128-bit stores).
constexpr int kGroupUnroll = 4;
__global__ void __launch_bounds__(256)
group_vec4_kernel(size_t total_vec, int c4, size_t rows_per_scene, size_t src_rows_per_scene,
                  const float4 *__restrict__ points, const int *__restrict__ idx, float4 *__restrict__ out) {
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < total_vec; v0 += step * kGroupUnroll) {
    const float4 *src[kGroupUnroll];
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) {
      const size_t v = v0 + u * step;
      src[u] = points;
      if (v < total_vec) {
        const size_t row = v / c4;
        const int q = (int)(v - row * c4);
        const size_t scene = row / rows_per_scene;
        src[u] = points + (scene * src_rows_per_scene + __ldg(idx + row)) * c4 + q;
      }
    }
    float4 val[kGroupUnroll];
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) val[u] = __ldg(src[u]);
#pragma unroll
    for (int u = 0; u < kGroupUnroll; ++u) {
      const size_t v = v0 + u * step;
      if (v < total_vec) __stcs(out + v, val[u]);  // streaming store: the grouped tensor is consumed by a later kernel
    }
  }
}

// Any c / alignment: each thread produces 4 consecutive output floats (possibly from 2-4 different rows) and
// stores them as one 128-bit word when the output base allows it.
__global__ void __launch_bounds__(256)
group_any_kernel(size_t total, int c, size_t rows_per_scene, size_t src_rows_per_scene, bool vec_store,
                 const float *__restrict__ points, const int *__restrict__ idx, float *__restrict__ out) {
  const size_t nchunks = (total + 3) / 4;
  for (size_t ch = (size_t)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += (size_t)gridDim.x * blockDim.x) {
    const size_t e0 = ch * 4;
    float v[4];
    size_t row = e0 / c;
    int l = (int)(e0 - row * c);
    int src = (row * c < total) ? __ldg(idx + row) : 0;
    size_t scene = row / rows_per_scene;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (e0 + t < total) v[t] = __ldg(points + (scene * src_rows_per_scene + src) * c + l);
      else v[t] = 0.0f;
      if (++l == c) {
        l = 0;
        ++row;
        if (row * c < total) { src = __ldg(idx + row); scene = row / rows_per_scene; }
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

int gather_rows(size_t scenes, size_t src_rows, size_t rows, int c, const float *points, const int *idx, float *out,
                cudaStream_t st) {
  const size_t total = scenes * rows * (size_t)c;
  if (total == 0) return PC_OK;
  const int sms = num_sms();
  if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
    const size_t nv = total / 4;
    size_t blocks = (nv + 256 * kGroupUnroll - 1) / (256 * kGroupUnroll);
    if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
    group_vec4_kernel<<<(unsigned)blocks, 256, 0, st>>>(nv, c / 4, rows, src_rows, (const float4 *)points, idx,
                                                         (float4 *)out);
  } else {
    const size_t nch = (total + 3) / 4;
    size_t blocks = (nch + 255) / 256;
    if (blocks > (size_t)sms * 64) blocks = (size_t)sms * 64;
    group_any_kernel<<<(unsigned)blocks, 256, 0, st>>>(total, c, rows, src_rows, aligned16(out), points, idx, out);
  }
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
