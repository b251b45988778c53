// Fused front ends of the two PointNet++ layer types: what sample_and_group / pointnet_fp_module build around the
// geometry ops with stock TF elementwise ops and concats (reference pointnet2_tensorflow/utils/pointnet_util.py).
//
// pc_sa_group      = group_point(xyz, idx) - tile(new_xyz)  (+)  group_point(points, idx)  -> concat   (:39-52)
//                    one pass writes new_points (b,m,ns,3+c) and the centred grouped_xyz (b,m,ns,3); the reference
//                    runs two GroupPoint ops, a tile, a subtract and a concat over five intermediate tensors.
// pc_fp_interpolate = inverse-distance weights (:219-222) -> three_interpolate (:223) -> concat with the skip features
//                    (:226); one pass writes (b,n,c2+c1) and, optionally, the (b,n,3) weights the backward needs.
// Both are HBM-write bound gathers: consecutive threads produce consecutive output floats (4 per thread, one 128-bit
// store when the base allows), the gathered rows come out of L2.  Values are bit-identical to the op-by-op
// composition: the subtraction and the weight arithmetic are the same single un-fused fp32 operations.
#include "common.cuh"

namespace pc {
namespace {

// I = uint32_t with multiply-shift index decoding (FastDiv) when the output has < 2^31 floats, size_t with real divisions
// beyond: the three 64-bit divisions per 16 output bytes made this gather ALU-bound (ncu: issue slots 72 %, ALU 49 %).
template <class I> struct SaDecode;
template <> struct SaDecode<uint32_t> {
  FastDiv w, ns, m;
  SaDecode(int W, int NS, int M) : w((uint32_t)W), ns((uint32_t)NS), m((uint32_t)M) {}
  __device__ __forceinline__ uint32_t by_w(uint32_t x) const { return w.div(x); }
  __device__ __forceinline__ uint32_t by_ns(uint32_t x) const { return ns.div(x); }
  __device__ __forceinline__ uint32_t by_m(uint32_t x) const { return m.div(x); }
};
template <> struct SaDecode<size_t> {
  size_t w, ns, m;
  SaDecode(int W, int NS, int M) : w(W), ns(NS), m(M) {}
  __device__ __forceinline__ size_t by_w(size_t x) const { return x / w; }
  __device__ __forceinline__ size_t by_ns(size_t x) const { return x / ns; }
  __device__ __forceinline__ size_t by_m(size_t x) const { return x / m; }
};

template <class I>
__global__ void __launch_bounds__(256)
sa_group_kernel(I total, SaDecode<I> dec, int n, int c, bool vec_store, const float *__restrict__ xyz,
                const float *__restrict__ points, const int *__restrict__ idx, const float *__restrict__ new_xyz,
                float *__restrict__ out, float *__restrict__ gxyz) {
  const int W = 3 + c;
  const I nchunks = (total + 3) / 4;
  for (I ch = (I)blockIdx.x * blockDim.x + threadIdx.x; ch < nchunks; ch += (I)gridDim.x * blockDim.x) {
    const I e0 = ch * 4;
    I row = dec.by_w(e0);  // (scene*m + j)*ns + k
    int col = (int)(e0 - row * W);
    float v[4];
    int src = -1;
    size_t scene = 0, q = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      v[t] = 0.0f;
      if (e0 + t < total) {
        if (src < 0) {
          src = __ldg(idx + row);
          const I qq = dec.by_ns(row);        // scene*m + j
          q = qq;
          scene = dec.by_m(qq);
        }
        if (col < 3) {
          const float g = __fsub_rn(__ldg(xyz + (scene * n + src) * 3 + col), __ldg(new_xyz + q * 3 + col));
          v[t] = g;
          if (gxyz) gxyz[(size_t)row * 3 + col] = g;
        } else {
          v[t] = __ldg(points + (scene * n + src) * (size_t)c + (col - 3));
        }
      }
      if (++col == W) { col = 0; ++row; src = -1; }
    }
    if (vec_store && e0 + 3 < total) {
      __stcs(reinterpret_cast<float4 *>(out + e0), make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (e0 + t < total) out[e0 + t] = v[t];
    }
  }
}

__device__ __forceinline__ void fp_weights(const float *__restrict__ dist, size_t row, float &w1, float &w2, float &w3) {
  // pointnet_util.py:219-222, same operation order as three_weights_kernel (interpolate.cu)
  const float d0 = fmaxf(__ldg(dist + row * 3 + 0), 1e-10f), d1 = fmaxf(__ldg(dist + row * 3 + 1), 1e-10f),
              d2 = fmaxf(__ldg(dist + row * 3 + 2), 1e-10f);
  const float r0 = __fdiv_rn(1.0f, d0), r1 = __fdiv_rn(1.0f, d1), r2 = __fdiv_rn(1.0f, d2);
  const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
  w1 = __fdiv_rn(r0, norm); w2 = __fdiv_rn(r1, norm); w3 = __fdiv_rn(r2, norm);
}

// A warp owns RPW (<= 32) consecutive dense points: lane r computes the weights and fetches the indices of point r ONCE
// (they are shared by all c2 channels), then the warp walks its rows, broadcasting (i1,i2,i3,w1,w2,w3) by shuffle while
// its lanes cover the row's channels -- 128-bit loads / stores when the channel counts and bases allow.  RPW = 32 for
// the big levels; the host lowers it when there are few rows (FP1: 1024 rows x 768 channels ran on 4 CTAs, 69 us).
__global__ void __launch_bounds__(256)
fp_interp_kernel(size_t rows, int RPW, int n, int m, int c2, int c1, bool vec, const float *__restrict__ dist,
                 const int *__restrict__ idx, const float *__restrict__ points2, const float *__restrict__ points1,
                 float *__restrict__ out, float *__restrict__ weight) {
  const int lane = threadIdx.x & 31;
  const int W = c2 + c1;
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r0 = warp_global * RPW; r0 < rows; r0 += nwarps * RPW) {
    const size_t my = r0 + lane;
    float w1 = 0.f, w2 = 0.f, w3 = 0.f;
    int i1 = 0, i2 = 0, i3 = 0;
    if (lane < RPW && my < rows && c2 > 0) {
      fp_weights(dist, my, w1, w2, w3);
      i1 = __ldg(idx + my * 3 + 0); i2 = __ldg(idx + my * 3 + 1); i3 = __ldg(idx + my * 3 + 2);
      if (weight) { weight[my * 3 + 0] = w1; weight[my * 3 + 1] = w2; weight[my * 3 + 2] = w3; }
    }
    const int nr = (int)min((size_t)RPW, rows - r0);
    for (int rr = 0; rr < nr; ++rr) {
      const size_t row = r0 + rr;
      const float a1 = __shfl_sync(PC_FULL_MASK, w1, rr), a2 = __shfl_sync(PC_FULL_MASK, w2, rr),
                  a3 = __shfl_sync(PC_FULL_MASK, w3, rr);
      const int j1 = __shfl_sync(PC_FULL_MASK, i1, rr), j2 = __shfl_sync(PC_FULL_MASK, i2, rr),
                j3 = __shfl_sync(PC_FULL_MASK, i3, rr);
      const size_t scene = row / n;
      const float *base = points2 + scene * (size_t)m * c2;
      const float *p1 = base + (size_t)j1 * c2, *p2 = base + (size_t)j2 * c2, *p3 = base + (size_t)j3 * c2;
      float *o = out + row * (size_t)W;
      const float *s1 = points1 + row * (size_t)c1;
      if (vec) {  // c2 % 4 == 0, c1 % 4 == 0, 16-byte aligned bases
        for (int q = lane; q < c2 / 4; q += 32) {
          const float4 x = __ldg(reinterpret_cast<const float4 *>(p1) + q), y = __ldg(reinterpret_cast<const float4 *>(p2) + q),
                       z = __ldg(reinterpret_cast<const float4 *>(p3) + q);
          // tf_interpolate.cpp:119: p1*w1 + p2*w2 + p3*w3, left to right, un-fused
          float4 r;
          r.x = __fadd_rn(__fadd_rn(__fmul_rn(x.x, a1), __fmul_rn(y.x, a2)), __fmul_rn(z.x, a3));
          r.y = __fadd_rn(__fadd_rn(__fmul_rn(x.y, a1), __fmul_rn(y.y, a2)), __fmul_rn(z.y, a3));
          r.z = __fadd_rn(__fadd_rn(__fmul_rn(x.z, a1), __fmul_rn(y.z, a2)), __fmul_rn(z.z, a3));
          r.w = __fadd_rn(__fadd_rn(__fmul_rn(x.w, a1), __fmul_rn(y.w, a2)), __fmul_rn(z.w, a3));
          __stcs(reinterpret_cast<float4 *>(o) + q, r);
        }
        for (int q = lane; q < c1 / 4; q += 32)
          __stcs(reinterpret_cast<float4 *>(o + c2) + q, __ldg(reinterpret_cast<const float4 *>(s1) + q));
      } else {
        for (int col = lane; col < c2; col += 32)
          o[col] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(p1 + col), a1), __fmul_rn(__ldg(p2 + col), a2)),
                             __fmul_rn(__ldg(p3 + col), a3));
        for (int col = lane; col < c1; col += 32) o[c2 + col] = __ldg(s1 + col);
      }
    }
  }
}

}  // namespace
}  // namespace pc

extern "C" int pc_sa_group(int b, int n, int c, int m, int nsample, const float *xyz, const float *points,
                           const int *idx, const float *new_xyz, float *new_points, float *grouped_xyz,
                           pc_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || m < 0 || nsample < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0 || nsample == 0) return PC_OK;
  if (n == 0 || !xyz || !idx || !new_xyz || !new_points || (c > 0 && !points)) return PC_ERR_INVALID_ARGUMENT;
  const size_t total = (size_t)b * m * nsample * (3 + c);
  size_t blocks = ((total + 3) / 4 + 255) / 256;
  const size_t cap = (size_t)pc::num_sms() * 32;
  if (blocks > cap) blocks = cap;
  if (total < (1ull << 31))
    pc::sa_group_kernel<uint32_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (uint32_t)total, pc::SaDecode<uint32_t>(3 + c, nsample, m), n, c, pc::aligned16(new_points), xyz, points, idx, new_xyz,
        new_points, grouped_xyz);
  else
    pc::sa_group_kernel<size_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        total, pc::SaDecode<size_t>(3 + c, nsample, m), n, c, pc::aligned16(new_points), xyz, points, idx, new_xyz, new_points,
        grouped_xyz);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_fp_interpolate(int b, int n, int m, int c2, int c1, const float *dist, const int *idx,
                                 const float *points2, const float *points1, float *out, float *weight,
                                 pc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || c2 < 0 || c1 < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0 || c2 + c1 == 0) return PC_OK;
  if (!out || (c2 > 0 && (m == 0 || !dist || !idx || !points2)) || (c1 > 0 && !points1)) return PC_ERR_INVALID_ARGUMENT;
  const size_t rows = (size_t)b * n;
  int rpw = 32;  // rows per warp: fewer when the launch would not fill the chip
  while (rpw > 1 && (rows + rpw - 1) / rpw < (size_t)pc::num_sms() * 16) rpw >>= 1;
  size_t blocks = ((rows + rpw - 1) / rpw + 7) / 8;  // 8 warps per CTA, one group of rows per warp
  const size_t cap = (size_t)pc::num_sms() * 32;
  if (blocks > cap) blocks = cap;
  const bool vec = c2 % 4 == 0 && c1 % 4 == 0 && pc::aligned16(out) && pc::aligned16(points2) &&
                   (c1 == 0 || pc::aligned16(points1));
  pc::fp_interp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, rpw, n, m, c2, c1, vec, dist, idx, points2,
                                                                         points1, out, weight);
  PC_RETURN_LAUNCH_STATUS();
}
