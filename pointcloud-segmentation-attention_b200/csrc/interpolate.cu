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
// candidates as two packed fp32x2 operands.  The scan is a branch-free FILTER: for every word of 32 candidates a thread
// only records, as one bit per candidate, whether the distance beats its third-best AT THE START of the word (sign of
// an integer subtract funnel-shifted into the word -- non-negative floats order like their bit patterns).  The exact
// 3-slot insertion then runs only for the recorded candidates, in ascending index, recomputing the distance with the
// same un-fused arithmetic.  A candidate that fails the stale test can never enter the list (the third-best only
// shrinks), so the result is the reference's sequential scan exactly, while 32 lanes no longer serialise on a branch
// that some lane takes at almost every step (a top-3 list changes ~3 ln m times per lane).
// Q dense points per thread: one broadcast LDS.128 then feeds 4 candidates x Q queries (a broadcast LDS.128 still
// occupies the shared-memory crossbar for 4 cycles per warp, which bounds a Q = 1 scan before the fp32 pipe does).
template <int Q, int kNNThreads>
__global__ void __launch_bounds__(kNNThreads)
three_nn_kernel(int n, int m, float one, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                float *__restrict__ dist, int *__restrict__ idx) {
  __shared__ __align__(16) float tile[3][kNNTile];
  const int scene = blockIdx.y;
  const float *known = xyz2 + (size_t)scene * m * 3;
  const float inf = __int_as_float(0x7f800000);
  const f32x2 one2 = pack2(one, one);
  int j[Q];
  float x1[Q], y1[Q], z1[Q], b1[Q], b2[Q], b3[Q];
  int i1[Q], i2[Q], i3[Q];
  f32x2 qx2[Q], qy2[Q], qz2[Q];
#pragma unroll
  for (int u = 0; u < Q; ++u) {
    j[u] = (blockIdx.x * Q + u) * kNNThreads + threadIdx.x;
    const float *qp = xyz1 + ((size_t)scene * n + (j[u] < n ? j[u] : 0)) * 3;
    x1[u] = qp[0]; y1[u] = qp[1]; z1[u] = qp[2];
    qx2[u] = pack2(x1[u], x1[u]); qy2[u] = pack2(y1[u], y1[u]); qz2[u] = pack2(z1[u], z1[u]);
    b1[u] = b2[u] = b3[u] = inf;
    i1[u] = i2[u] = i3[u] = 0;
  }
  constexpr int kLoads = kNNTile * 3 / kNNThreads;  // 24 coalesced loads per thread, all issued before the first store
  for (int t0 = 0; t0 < m; t0 += kNNTile) {
    const int tn = min(kNNTile, m - t0);
    const int tn32 = (tn + 31) & ~31;
    float pre[kLoads];
#pragma unroll
    for (int u = 0; u < kLoads; ++u) {
      const int i = threadIdx.x + kNNThreads * u;
      pre[u] = (i < tn * 3) ? __ldg(known + (size_t)t0 * 3 + i) : inf;  // slots past tn: +inf, never below any b3
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kLoads; ++u) {  // AoS -> SoA
      const int i = threadIdx.x + kNNThreads * u, k = i / 3, c = i - k * 3;
      tile[c][k] = pre[u];
    }
    __syncthreads();
    for (int w0 = 0; w0 < tn32; w0 += 32) {
      int thr[Q];
      unsigned word[Q];
#pragma unroll
      for (int u = 0; u < Q; ++u) {
        thr[u] = __float_as_int(b3[u]);  // stale third-best for this word; +inf admits every finite distance
        word[u] = 0;
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int k = w0 + g * 4;
        const float4 xs = *reinterpret_cast<const float4 *>(&tile[0][k]);
        const float4 ys = *reinterpret_cast<const float4 *>(&tile[1][k]);
        const float4 zs = *reinterpret_cast<const float4 *>(&tile[2][k]);
        const f32x2 xa = pack2(xs.x, xs.y), ya = pack2(ys.x, ys.y), za = pack2(zs.x, zs.y);
        const f32x2 xb = pack2(xs.z, xs.w), yb = pack2(ys.z, ys.w), zb = pack2(zs.z, zs.w);
#pragma unroll
        for (int u = 0; u < Q; ++u) {
          float d0, d1, d2, d3;
          // candidate minus query, as tf_interpolate.cpp:73 writes it
          unpack2(sqdist3_x2(xa, ya, za, qx2[u], qy2[u], qz2[u], one2), d0, d1);
          unpack2(sqdist3_x2(xb, yb, zb, qx2[u], qy2[u], qz2[u], one2), d2, d3);
          word[u] = __funnelshift_l(__float_as_int(d0) - thr[u], word[u], 1);  // d < b3 <=> bits(d) - bits(b3) < 0
          word[u] = __funnelshift_l(__float_as_int(d1) - thr[u], word[u], 1);
          word[u] = __funnelshift_l(__float_as_int(d2) - thr[u], word[u], 1);
          word[u] = __funnelshift_l(__float_as_int(d3) - thr[u], word[u], 1);
        }
      }
#pragma unroll
      for (int u = 0; u < Q; ++u) {
        unsigned wd = __brev(word[u]);  // bit e <-> candidate w0 + e
        while (wd) {                    // exact insertion of the few admitted candidates, ascending index
          const int e = __ffs(wd) - 1;
          wd &= wd - 1;
          const int k = w0 + e;
          const float d = sqdist3(tile[0][k], tile[1][k], tile[2][k], x1[u], y1[u], z1[u]);
          nn_insert(d, t0 + k, b1[u], b2[u], b3[u], i1[u], i2[u], i3[u]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < Q; ++u) {
    if (j[u] < n) {
      float *dp = dist + ((size_t)scene * n + j[u]) * 3;
      int *ip = idx + ((size_t)scene * n + j[u]) * 3;
      dp[0] = b1[u]; dp[1] = b2[u]; dp[2] = b3[u];
      ip[0] = i1[u]; ip[1] = i2[u]; ip[2] = i3[u];
    }
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
// Index decoding: multiply-shift (FastDiv) while the output has < 2^31 elements, real 64-bit divisions beyond.
template <class I> struct RowDecode;
template <> struct RowDecode<uint32_t> {
  FastDiv a, b;
  RowDecode(size_t da, size_t db) : a((uint32_t)da), b((uint32_t)db) {}
  __device__ __forceinline__ uint32_t by_a(uint32_t x) const { return a.div(x); }
  __device__ __forceinline__ uint32_t by_b(uint32_t x) const { return b.div(x); }
  __device__ __forceinline__ uint32_t da() const { return a.d; }
};
template <> struct RowDecode<size_t> {
  size_t a, b;
  RowDecode(size_t da, size_t db) : a(da ? da : 1), b(db ? db : 1) {}
  __device__ __forceinline__ size_t by_a(size_t x) const { return x / a; }
  __device__ __forceinline__ size_t by_b(size_t x) const { return x / b; }
  __device__ __forceinline__ size_t da() const { return a; }
};

template <class I>
__global__ void __launch_bounds__(256)
interp_vec4_kernel(I total_vec, RowDecode<I> dec /* a = c4, b = n */, int m, const float4 *__restrict__ points,
                   const int *__restrict__ idx, const float *__restrict__ weight, float4 *__restrict__ out) {
  const I step = (I)gridDim.x * blockDim.x;
  const I c4 = dec.da();
  for (I v0 = (I)blockIdx.x * blockDim.x + threadIdx.x; v0 < total_vec; v0 += step * kInterpUnroll) {
    float4 a[kInterpUnroll], b[kInterpUnroll], c[kInterpUnroll];
    float w1[kInterpUnroll], w2[kInterpUnroll], w3[kInterpUnroll];
#pragma unroll
    for (int u = 0; u < kInterpUnroll; ++u) {
      const I v = min(v0 + u * step, total_vec - 1);
      const I row = dec.by_a(v);  // scene*n + j
      const I q = v - row * c4;
      const I scene = dec.by_b(row);
      const size_t r3 = (size_t)row * 3;
      const int i1 = __ldg(idx + r3), i2 = __ldg(idx + r3 + 1), i3 = __ldg(idx + r3 + 2);
      w1[u] = __ldg(weight + r3); w2[u] = __ldg(weight + r3 + 1); w3[u] = __ldg(weight + r3 + 2);
      const float4 *base = points + (size_t)scene * (size_t)m * c4 + q;
      a[u] = __ldg(base + (size_t)i1 * c4); b[u] = __ldg(base + (size_t)i2 * c4); c[u] = __ldg(base + (size_t)i3 * c4);
    }
#pragma unroll
    for (int u = 0; u < kInterpUnroll; ++u) {
      const I v = v0 + u * step;
      if (v < total_vec)
        __stcs(out + v, make_float4(blend3(a[u].x, b[u].x, c[u].x, w1[u], w2[u], w3[u]),
                                    blend3(a[u].y, b[u].y, c[u].y, w1[u], w2[u], w3[u]),
                                    blend3(a[u].z, b[u].z, c[u].z, w1[u], w2[u], w3[u]),
                                    blend3(a[u].w, b[u].w, c[u].w, w1[u], w2[u], w3[u])));
    }
  }
}


// c % 128 == 0 (every FP level of the ScanNet models: 128 / 256 / 512 channels): a warp owns rpw (<= 32) consecutive rows, lane r
// fetches (idx, weight) of row r ONCE, the warp then walks the rows two at a time -- (i, w) broadcast by shuffle, a lane
// covers the float4 columns lane, lane + 32, ... -- with six independent 128-bit row loads in flight per lane.  The
// thread-per-float4 kernel above re-reads the six (idx, weight) words and re-decodes the row for every float4: 45 warp
// instructions per output float4 against 22 here (it ran at 55 % of the issue slots for 0.45 of the HBM peak).
#ifndef PCOPS_INTERP_ROWS
#define PCOPS_INTERP_ROWS 1
#endif
__global__ void __launch_bounds__(256)
interp_rows_kernel(size_t rows, int rpw, int n, int m, int c4, const float4 *__restrict__ points, const int *__restrict__ idx,
                   const float *__restrict__ weight, float4 *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r0 = warp_global * rpw; r0 < rows; r0 += nwarps * rpw) {
    const size_t my = min(r0 + (lane & (rpw - 1)), rows - 1);   // rpw is a power of two: lanes >= rpw mirror the first ones
    const int i1 = __ldg(idx + my * 3 + 0), i2 = __ldg(idx + my * 3 + 1), i3 = __ldg(idx + my * 3 + 2);
    const float w1 = __ldg(weight + my * 3 + 0), w2 = __ldg(weight + my * 3 + 1), w3 = __ldg(weight + my * 3 + 2);
    const float4 *base = points + (my / n) * (size_t)m * c4;   // scene of this lane's row
    const float4 *p1 = base + (size_t)i1 * c4, *p2 = base + (size_t)i2 * c4, *p3 = base + (size_t)i3 * c4;
    const int nr = (int)min((size_t)rpw, rows - r0);
    for (int rr = 0; rr < nr; rr += 2) {
      const int rb = min(rr + 1, nr - 1);
      const float4 *a1 = (const float4 *)__shfl_sync(PC_FULL_MASK, (unsigned long long)p1, rr),
                   *a2 = (const float4 *)__shfl_sync(PC_FULL_MASK, (unsigned long long)p2, rr),
                   *a3 = (const float4 *)__shfl_sync(PC_FULL_MASK, (unsigned long long)p3, rr);
      const float4 *b1 = (const float4 *)__shfl_sync(PC_FULL_MASK, (unsigned long long)p1, rb),
                   *b2 = (const float4 *)__shfl_sync(PC_FULL_MASK, (unsigned long long)p2, rb),
                   *b3 = (const float4 *)__shfl_sync(PC_FULL_MASK, (unsigned long long)p3, rb);
      const float u1 = __shfl_sync(PC_FULL_MASK, w1, rr), u2 = __shfl_sync(PC_FULL_MASK, w2, rr), u3 = __shfl_sync(PC_FULL_MASK, w3, rr);
      const float v1 = __shfl_sync(PC_FULL_MASK, w1, rb), v2 = __shfl_sync(PC_FULL_MASK, w2, rb), v3 = __shfl_sync(PC_FULL_MASK, w3, rb);
      float4 *oa = out + (r0 + rr) * (size_t)c4, *ob = out + (r0 + rb) * (size_t)c4;
      for (int q = lane; q < c4; q += 32) {
        const float4 x = __ldg(a1 + q), y = __ldg(a2 + q), z = __ldg(a3 + q);
        const float4 X = __ldg(b1 + q), Y = __ldg(b2 + q), Z = __ldg(b3 + q);
        __stcs(oa + q, make_float4(blend3(x.x, y.x, z.x, u1, u2, u3), blend3(x.y, y.y, z.y, u1, u2, u3),
                                   blend3(x.z, y.z, z.z, u1, u2, u3), blend3(x.w, y.w, z.w, u1, u2, u3)));
        if (rb != rr)
          __stcs(ob + q, make_float4(blend3(X.x, Y.x, Z.x, v1, v2, v3), blend3(X.y, Y.y, Z.y, v1, v2, v3),
                                     blend3(X.z, Y.z, Z.z, v1, v2, v3), blend3(X.w, Y.w, Z.w, v1, v2, v3)));
      }
    }
  }
}

template <class I>
__global__ void __launch_bounds__(256)
interp_scalar_kernel(I total, RowDecode<I> dec /* a = c, b = n */, int m, const float *__restrict__ points,
                     const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out) {
  const I c = dec.da();
  for (I e = (I)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (I)gridDim.x * blockDim.x) {
    const I row = dec.by_a(e);
    const I l = e - row * c;
    const I scene = dec.by_b(row);
    const size_t r3 = (size_t)row * 3;
    const int i1 = __ldg(idx + r3), i2 = __ldg(idx + r3 + 1), i3 = __ldg(idx + r3 + 2);
    const float w1 = __ldg(weight + r3), w2 = __ldg(weight + r3 + 1), w3 = __ldg(weight + r3 + 2);
    const float *base = points + (size_t)scene * (size_t)m * c + l;
    out[e] = blend3(__ldg(base + (size_t)i1 * c), __ldg(base + (size_t)i2 * c), __ldg(base + (size_t)i3 * c), w1, w2, w3);
  }
}

template <class I>
void launch_interp(size_t total, int c, int n, int m, const float *points, const int *idx, const float *weight,
                   float *out, cudaStream_t st) {
  const size_t rows = c > 0 ? total / (size_t)c : 0;
  int rpw = 32;   // rows per warp: fewer when 32-row groups would leave the SMs short of warps (16 per SM wanted)
  while (rpw > 4 && (rows + rpw - 1) / rpw < (size_t)num_sms() * 48) rpw >>= 1;
  if (PCOPS_INTERP_ROWS && c % 128 == 0 && c > 0 && aligned16(points) && aligned16(out) && rpw >= 8) {   // small launches: thread-per-float4 kernel
    size_t blocks = ((rows + rpw - 1) / rpw + 7) / 8;
    const int h = concurrency_hint();
    const size_t cap = (size_t)num_sms() * (size_t)(8 / h > 1 ? 8 / h : 1);   // CTAs per SM: 8 alone, fewer among co-resident kernels
    if (blocks > cap) blocks = cap;
    interp_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(rows, rpw, n, m, c / 4, (const float4 *)points, idx, weight, (float4 *)out);
  } else if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
    const size_t nv = total / 4;
    const int blocks = resident_grid((const void *)interp_vec4_kernel<I>, 256, 0, (nv + 256 * kInterpUnroll - 1) / (256 * kInterpUnroll));
    interp_vec4_kernel<I><<<blocks, 256, 0, st>>>((I)nv, RowDecode<I>((size_t)c / 4, (size_t)n), m, (const float4 *)points, idx,
                                                  weight, (float4 *)out);
  } else {
    const int blocks = resident_grid((const void *)interp_scalar_kernel<I>, 256, 0, (total + 255) / 256);
    interp_scalar_kernel<I><<<blocks, 256, 0, st>>>((I)total, RowDecode<I>((size_t)c, (size_t)n), m, points, idx, weight, out);
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
  // a CTA owns 256 dense points: as 128 threads x 2 points when that still gives >= 2 CTAs per SM (halves the
  // shared-memory broadcasts per pair test), else as 256 threads x 1 point
  dim3 grid((n + 255) / 256, b);
  if ((long)grid.x * b >= 2L * pc::num_sms())
    pc::three_nn_kernel<2, 128><<<grid, 128, 0, (cudaStream_t)stream>>>(n, m, 1.0f, xyz1, xyz2, dist, idx);
  else
    pc::three_nn_kernel<1, 256><<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, 1.0f, xyz1, xyz2, dist, idx);
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
  if (total < (1ull << 31)) pc::launch_interp<uint32_t>(total, c, n, m, points, idx, weight, out, st);
  else pc::launch_interp<size_t>(total, c, n, m, points, idx, weight, out, st);
  PC_RETURN_LAUNCH_STATUS();
}
