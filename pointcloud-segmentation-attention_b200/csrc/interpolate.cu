// three_nn, inverse-distance weights and three_interpolate for sm_100a.
//
// The reference implements these three ops on the HOST only (tf_ops/interpolation_3d/tf_interpolate.cpp:60-127,
// registered for DEVICE_CPU at :187,:222), single-threaded, with device<->host copies around each call.  These are
// the first GPU kernels for them.
//
// three_nn: one thread per dense point; the sparse (known) cloud streams through shared memory as SoA rows so one
// broadcast LDS.128 feeds four candidates to packed fp32x2 arithmetic (two distances per FADD2/FMUL2/FFMA2, each
// operation still rounded on its own); the 3-slot insertion keeps the reference's strict '<' cascade and visits
// candidates in ascending index, so equal distances stay in ascending index order.  The reference compares in double after evaluating the distance in float;
// float values compare identically in either width, and its 1e40 initial slots (which print as +inf once stored to
// the float output) become +inf here.
// three_interpolate: one thread per (dense point, 4 channels), 128-bit loads of the three source rows (L2 resident),
// 128-bit coalesced stores, (p1*w1 + p2*w2) + p3*w3 un-fused as in tf_interpolate.cpp:119.
#include <math.h>
#include "common.cuh"

namespace pc {
namespace {

constexpr int kNNThreads = 256;
constexpr int kNNTile = 2048;  // known points per tile: 3 SoA rows of 2048 floats = 24 KB

// 3-slot insertion of tf_interpolate.cpp:74-89 (strict '<': equal distances keep ascending index order).
__device__ __forceinline__ void nn_insert(float d, int kk, float &b1, float &b2, float &b3, int &i1, int &i2, int &i3) {
  if (d < b3) {
    if (d < b1) {
      b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = kk;
    } else if (d < b2) {
      b3 = b2; i3 = i2; b2 = d; i2 = kk;
    } else {
      b3 = d; i3 = kk;
    }
  }
}

// One thread per dense point.  The known cloud is staged SoA in shared memory so one LDS.128 broadcast feeds four
// candidates as two packed fp32x2 operands; four distances cost 16 packed FP instructions, and only when the smallest
// of the four beats the current third-best does the thread enter the (in-order, exact) insertion cascade.
__global__ void __launch_bounds__(kNNThreads)
three_nn_kernel(int n, int m, float one, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                float *__restrict__ dist, int *__restrict__ idx) {
  __shared__ __align__(16) float tile[3][kNNTile];
  const int scene = blockIdx.y;
  const int j = blockIdx.x * kNNThreads + threadIdx.x;
  const bool live = j < n;
  const float *qp = xyz1 + ((size_t)scene * n + (live ? j : 0)) * 3;
  const float x1 = qp[0], y1 = qp[1], z1 = qp[2];
  const f32x2 qx2 = pack2(x1, x1), qy2 = pack2(y1, y1), qz2 = pack2(z1, z1), one2 = pack2(one, one);
  const float *known = xyz2 + (size_t)scene * m * 3;
  const float inf = __int_as_float(0x7f800000);
  float b1 = inf, b2 = inf, b3 = inf;
  int i1 = 0, i2 = 0, i3 = 0;
  constexpr int kLoads = kNNTile * 3 / kNNThreads;  // 24 coalesced loads per thread, all issued before the first store
  for (int t0 = 0; t0 < m; t0 += kNNTile) {
    const int tn = min(kNNTile, m - t0);
    const int tn4 = (tn + 3) & ~3;
    float pre[kLoads];
#pragma unroll
    for (int u = 0; u < kLoads; ++u) {
      const int i = threadIdx.x + kNNThreads * u;
      pre[u] = (i < tn * 3) ? __ldg(known + (size_t)t0 * 3 + i) : inf;  // slots past tn: +inf, never inserted
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kLoads; ++u) {  // AoS -> SoA
      const int i = threadIdx.x + kNNThreads * u, k = i / 3, c = i - k * 3;
      tile[c][k] = pre[u];
    }
    __syncthreads();
#pragma unroll 2
    for (int k = 0; k < tn4; k += 4) {
      const float4 xs = *reinterpret_cast<const float4 *>(&tile[0][k]);
      const float4 ys = *reinterpret_cast<const float4 *>(&tile[1][k]);
      const float4 zs = *reinterpret_cast<const float4 *>(&tile[2][k]);
      float d0, d1, d2, d3;
      // candidate minus query, as tf_interpolate.cpp:73 writes it
      unpack2(sqdist3_x2(pack2(xs.x, xs.y), pack2(ys.x, ys.y), pack2(zs.x, zs.y), qx2, qy2, qz2, one2), d0, d1);
      unpack2(sqdist3_x2(pack2(xs.z, xs.w), pack2(ys.z, ys.w), pack2(zs.z, zs.w), qx2, qy2, qz2, one2), d2, d3);
      if (fminf(fminf(d0, d1), fminf(d2, d3)) < b3) {
        const int kk = t0 + k;
        nn_insert(d0, kk, b1, b2, b3, i1, i2, i3);
        nn_insert(d1, kk + 1, b1, b2, b3, i1, i2, i3);
        nn_insert(d2, kk + 2, b1, b2, b3, i1, i2, i3);
        nn_insert(d3, kk + 3, b1, b2, b3, i1, i2, i3);
      }
    }
  }
  if (live) {
    float *dp = dist + ((size_t)scene * n + j) * 3;
    int *ip = idx + ((size_t)scene * n + j) * 3;
    dp[0] = b1; dp[1] = b2; dp[2] = b3;
    ip[0] = i1; ip[1] = i2; ip[2] = i3;
  }
}

__global__ void __launch_bounds__(256)
three_weights_kernel(size_t rows, const float *__restrict__ dist, float *__restrict__ weight) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    const float d0 = fmaxf(dist[r * 3 + 0], 1e-10f), d1 = fmaxf(dist[r * 3 + 1], 1e-10f),
                d2 = fmaxf(dist[r * 3 + 2], 1e-10f);
    const float r0 = __fdiv_rn(1.0f, d0), r1 = __fdiv_rn(1.0f, d1), r2 = __fdiv_rn(1.0f, d2);
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    weight[r * 3 + 0] = __fdiv_rn(r0, norm);
    weight[r * 3 + 1] = __fdiv_rn(r1, norm);
    weight[r * 3 + 2] = __fdiv_rn(r2, norm);
  }
}

__device__ __forceinline__ float blend3(float p1, float p2, float p3, float w1, float w2, float w3) {
  return __fadd_rn(__fadd_rn(__fmul_rn(p1, w1), __fmul_rn(p2, w2)), __fmul_rn(p3, w3));
}

constexpr int kInterpUnroll = 2;  // 2 x 3 independent 128-bit row loads in flight per thread
__global__ void __launch_bounds__(256)
interp_vec4_kernel(size_t total_vec, int c4, int n, int m, const float4 *__restrict__ points,
                   const int *__restrict__ idx, const float *__restrict__ weight, float4 *__restrict__ out) {
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < total_vec; v0 += step * kInterpUnroll) {
    float4 a[kInterpUnroll], b[kInterpUnroll], c[kInterpUnroll];
    float w1[kInterpUnroll], w2[kInterpUnroll], w3[kInterpUnroll];
#pragma unroll
    for (int u = 0; u < kInterpUnroll; ++u) {
      const size_t v = min(v0 + u * step, total_vec - 1);
      const size_t row = v / c4;  // scene*n + j
      const int q = (int)(v - row * c4);
      const size_t scene = row / n;
      const int i1 = __ldg(idx + row * 3), i2 = __ldg(idx + row * 3 + 1), i3 = __ldg(idx + row * 3 + 2);
      w1[u] = __ldg(weight + row * 3); w2[u] = __ldg(weight + row * 3 + 1); w3[u] = __ldg(weight + row * 3 + 2);
      const float4 *base = points + scene * (size_t)m * c4 + q;
      a[u] = __ldg(base + (size_t)i1 * c4); b[u] = __ldg(base + (size_t)i2 * c4); c[u] = __ldg(base + (size_t)i3 * c4);
    }
#pragma unroll
    for (int u = 0; u < kInterpUnroll; ++u) {
      const size_t v = v0 + u * step;
      if (v < total_vec)
        __stcs(out + v, make_float4(blend3(a[u].x, b[u].x, c[u].x, w1[u], w2[u], w3[u]),
                                    blend3(a[u].y, b[u].y, c[u].y, w1[u], w2[u], w3[u]),
                                    blend3(a[u].z, b[u].z, c[u].z, w1[u], w2[u], w3[u]),
                                    blend3(a[u].w, b[u].w, c[u].w, w1[u], w2[u], w3[u])));
    }
  }
}

__global__ void __launch_bounds__(256)
interp_scalar_kernel(size_t total, int c, int n, int m, const float *__restrict__ points,
                     const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / c;
    const int l = (int)(e - row * c);
    const size_t scene = row / n;
    const int i1 = __ldg(idx + row * 3), i2 = __ldg(idx + row * 3 + 1), i3 = __ldg(idx + row * 3 + 2);
    const float w1 = __ldg(weight + row * 3), w2 = __ldg(weight + row * 3 + 1), w3 = __ldg(weight + row * 3 + 2);
    const float *base = points + scene * (size_t)m * c + l;
    out[e] = blend3(__ldg(base + (size_t)i1 * c), __ldg(base + (size_t)i2 * c), __ldg(base + (size_t)i3 * c), w1, w2, w3);
  }
}

}  // namespace
}  // namespace pc

extern "C" int pc_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                           pc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return PC_OK;
  if (!xyz1 || !dist || !idx || (m > 0 && !xyz2)) return PC_ERR_INVALID_ARGUMENT;
  if (b > 65535) return PC_ERR_UNSUPPORTED;
  dim3 grid((n + pc::kNNThreads - 1) / pc::kNNThreads, b);
  pc::three_nn_kernel<<<grid, pc::kNNThreads, 0, (cudaStream_t)stream>>>(n, m, 1.0f, xyz1, xyz2, dist, idx);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_three_weights(size_t rows, const float *dist, float *weight, pc_stream_t stream) {
  if (rows == 0) return PC_OK;
  if (!dist || !weight) return PC_ERR_INVALID_ARGUMENT;
  size_t blocks = (rows + 255) / 256;
  const size_t cap = (size_t)pc::num_sms() * 32;
  if (blocks > cap) blocks = cap;
  pc::three_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, dist, weight);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx,
                                    const float *weight, float *out, pc_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0 || c == 0) return PC_OK;
  if (m == 0 || !points || !idx || !weight || !out) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)b * n * c;
  const size_t cap = (size_t)pc::num_sms() * 16;
  if (c % 4 == 0 && pc::aligned16(points) && pc::aligned16(out)) {
    const size_t nv = total / 4;
    size_t blocks = (nv + 256 * pc::kInterpUnroll - 1) / (256 * pc::kInterpUnroll);
    if (blocks > cap) blocks = cap;
    pc::interp_vec4_kernel<<<(unsigned)blocks, 256, 0, st>>>(nv, c / 4, n, m, (const float4 *)points, idx, weight,
                                                             (float4 *)out);
  } else {
    size_t blocks = (total + 255) / 256;
    if (blocks > cap) blocks = cap;
    pc::interp_scalar_kernel<<<(unsigned)blocks, 256, 0, st>>>(total, c, n, m, points, idx, weight, out);
  }
  PC_RETURN_LAUNCH_STATUS();
}
