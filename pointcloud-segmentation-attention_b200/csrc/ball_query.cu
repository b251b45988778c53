// Ball query for sm_100a.
//
// Replaces query_ball_point_gpu / queryBallPointLauncher (reference tf_ops/grouping/tf_grouping_g.cu:3-36,125-128).
// Semantics kept: per query, the FIRST nsample dataset indices (ascending) inside the ball; the hit test is the
// reference's max(sqrtf(d2),1e-20f) < radius evaluated exactly -- the host turns it into d2 <= s_star where s_star is
// the largest float whose correctly-rounded sqrt is < radius (sqrt is monotone, so the two tests agree on every
// input; d2 < radius*radius would NOT, radius*radius rounds up for 0.1/0.2/0.4/0.8).  Unfilled slots repeat the
// first hit, pts_cnt saturates at nsample, rows of empty balls are zero.
//
// Design: the dataset streams through shared memory in tiles shared by the whole CTA; a warp owns QW queries and
// tests 32 candidates per step (one per lane) against all of them, so each candidate load is amortised over QW
// distance evaluations; hits are appended in index order with ballot + popc prefix (order-preserving compaction),
// and a warp stops evaluating a query once it has nsample hits.  The reference runs one thread per query streaming
// global memory on b CTAs; this runs b * m / (8*QW) CTAs.
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace pc {
namespace {

constexpr int kWarps = 8;
constexpr int kTile = 2048;  // candidates per shared-memory tile (24 KB)

template <int QW>
__global__ void __launch_bounds__(kWarps * 32)
ball_query_kernel(int n, int m, float s_star, int nsample, const float *__restrict__ xyz1,
                  const float *__restrict__ xyz2, int *__restrict__ idx, int *__restrict__ pts_cnt) {
  __shared__ float tile[kTile * 3];
  const int scene = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float *data = xyz1 + (size_t)scene * n * 3;
  const int q0 = (blockIdx.x * kWarps + warp) * QW;
  const unsigned lt = lanemask_lt();

  float qx[QW], qy[QW], qz[QW];
  int cnt[QW], first[QW];
  int *row[QW];
#pragma unroll
  for (int i = 0; i < QW; ++i) {
    const int q = q0 + i;
    const bool live = q < m;
    const float *qp = xyz2 + ((size_t)scene * m + (live ? q : 0)) * 3;
    qx[i] = qp[0]; qy[i] = qp[1]; qz[i] = qp[2];
    cnt[i] = live ? 0 : nsample;  // dead slots behave as already full
    first[i] = 0;
    row[i] = idx + ((size_t)scene * m + (live ? q : 0)) * nsample;
  }

  for (int t0 = 0; t0 < n; t0 += kTile) {
    bool warp_full = true;
#pragma unroll
    for (int i = 0; i < QW; ++i) warp_full = warp_full && (cnt[i] >= nsample);
    if (__syncthreads_and(warp_full)) break;  // also fences the previous tile's readers
    const int tn = min(kTile, n - t0);
    for (int i = threadIdx.x; i < tn * 3; i += kWarps * 32) tile[i] = data[(size_t)t0 * 3 + i];
    __syncthreads();
    if (warp_full) continue;
    for (int k0 = 0; k0 < tn; k0 += 32) {
      const int k = k0 + lane;
      const bool valid = k < tn;
      const int kk = valid ? k : 0;
      const float x = tile[kk * 3 + 0], y = tile[kk * 3 + 1], z = tile[kk * 3 + 2];
#pragma unroll
      for (int i = 0; i < QW; ++i) {
        if (cnt[i] < nsample) {  // warp-uniform
          const float s = sqdist3(qx[i], qy[i], qz[i], x, y, z);
          const bool hit = valid && (s <= s_star);
          const unsigned mask = __ballot_sync(PC_FULL_MASK, hit);
          if (mask) {
            if (cnt[i] == 0) first[i] = t0 + k0 + __ffs(mask) - 1;
            const int pos = cnt[i] + __popc(mask & lt);
            if (hit && pos < nsample) row[i][pos] = t0 + k;
            cnt[i] = min(nsample, cnt[i] + __popc(mask));
          }
        }
      }
    }
  }

#pragma unroll
  for (int i = 0; i < QW; ++i) {
    if (q0 + i < m) {
      const int fill = cnt[i] ? first[i] : 0;  // tf_grouping_g.cu:26-29; empty ball -> zero row
      for (int l = cnt[i] + lane; l < nsample; l += 32) row[i][l] = fill;
      if (lane == 0) pts_cnt[(size_t)scene * m + q0 + i] = cnt[i];
    }
  }
}

// Largest float s with max(sqrtf(s), 1e-20f) < radius, or -1 if there is none.
float ball_threshold(float radius) {
  if (!(radius > 1e-20f)) return -1.0f;  // d is clamped to 1e-20f from below, so nothing can hit
  uint32_t lo = 0, hi = 0x7f7fffffu;     // bit patterns of +0 .. FLT_MAX, monotone in value
  auto ok = [&](uint32_t bits) {
    float s;
    memcpy(&s, &bits, 4);
    return sqrtf(s) < radius;
  };
  if (!ok(lo)) return -1.0f;
  while (lo < hi) {
    uint32_t mid = lo + (hi - lo + 1) / 2;
    if (ok(mid)) lo = mid; else hi = mid - 1;
  }
  float s;
  memcpy(&s, &lo, 4);
  return s;
}

template <int QW>
int launch(int b, int n, int m, float s_star, int nsample, const float *xyz1, const float *xyz2, int *idx,
           int *pts_cnt, cudaStream_t st) {
  dim3 grid((m + kWarps * QW - 1) / (kWarps * QW), b);
  ball_query_kernel<QW><<<grid, kWarps * 32, 0, st>>>(n, m, s_star, nsample, xyz1, xyz2, idx, pts_cnt);
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace
}  // namespace pc

extern "C" int pc_query_ball(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                             int *idx, int *pts_cnt, pc_stream_t stream) {
  if (!(radius > 0.0f) || nsample <= 0) return PC_ERR_INVALID_ARGUMENT;  // tf_grouping.cpp:70-74
  if (b < 0 || n < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0) return PC_OK;
  if (!xyz2 || !idx || !pts_cnt || (n > 0 && !xyz1)) return PC_ERR_INVALID_ARGUMENT;
  if (b > 65535) return PC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const float s_star = pc::ball_threshold(radius);
  // Enough warps to cover the chip first, then amortise candidate loads over more queries per warp.
  const long warps4 = (long)b * ((m + 3) / 4);
  const long target = (long)pc::num_sms() * 16;
  if (warps4 >= target) return pc::launch<4>(b, n, m, s_star, nsample, xyz1, xyz2, idx, pts_cnt, st);
  if (warps4 * 2 >= target) return pc::launch<2>(b, n, m, s_star, nsample, xyz1, xyz2, idx, pts_cnt, st);
  return pc::launch<1>(b, n, m, s_star, nsample, xyz1, xyz2, idx, pts_cnt, st);
}
