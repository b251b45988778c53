// Ball query for sm_100a.
//
// Replaces query_ball_point_gpu / queryBallPointLauncher (reference tf_ops/grouping/tf_grouping_g.cu:3-36,125-128).
// Semantics kept: per query, the FIRST nsample dataset indices (ascending) inside the ball; the hit test is the
// reference's max(sqrtf(d2),1e-20f) < radius evaluated exactly -- the host turns it into d2 <= s_star where s_star is
// the largest float whose correctly-rounded sqrt is < radius (sqrt is monotone, so the two tests agree on every
// input; d2 < radius*radius would NOT, radius*radius rounds up for 0.1/0.2/0.4/0.8).  Unfilled slots repeat the
// first hit, pts_cnt saturates at nsample, rows of empty balls are zero.
//
// Design (FP32-pipe bound; the pair tests are the whole cost):
//   * a CTA owns 64 queries of one scene (a lane holds two) and its 8 warps split the candidate range into 8
//     contiguous segments, so a level with few queries still spreads over the chip (B*m/64*8 warps);
//   * each warp streams its segment through a private shared-memory stage, transposed to SoA so that one LDS.128
//     broadcast feeds four candidates as two packed fp32x2 operands; a thread tests two candidates per FADD2/FMUL2/
//     FFMA2 (every operation still rounded on its own, common.cuh);
//   * the inner loop is branch-free: a hit sets one bit of a 32-candidate word (per query, per segment) kept in shared
//     memory; ordering is free because bit position == candidate index;
//   * after a chunk (<= 4096 candidates) the 8 segment counts of a query are exchanged through shared memory, each
//     thread turns its own words into indices written at (hits in earlier segments) + rank -- "first nsample in
//     ascending index" without any sort, ballot or atomics -- and the CTA stops as soon as all its queries are full.
// The reference runs one thread per query streaming global memory on b CTAs with a divergent early exit.
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace pc {
namespace {

constexpr int kSeg = 8;           // warps per CTA = candidate segments
constexpr int kQ = 2;             // queries per lane: one broadcast LDS.128 feeds 4 candidates x 2 queries
constexpr int kQPB = 32 * kQ;     // queries per CTA
constexpr int kStage = 256;       // candidates staged per warp per step (multiple of 32)
constexpr int kStagePad = 12;     // row stride 268 floats: 16-byte aligned rows, x / y / z rows 12 banks apart for the transpose
constexpr int kChunk = 4096;      // candidates per chunk (capacity of the hit bitmap)
constexpr int kStageRow = kStage + kStagePad;

struct BallSmem {
  float stage[kSeg][3][kStageRow];
  unsigned bits[kChunk / 32][kQPB];  // [word][query]: bank == lane for both queries of a lane, conflict-free
  int seg_cnt[kSeg][kQPB];
  int seg_first[kSeg][kQPB];
};

__global__ void __launch_bounds__(kSeg * 32, 2)
ball_query_kernel(int n, int m, float s_star, float radius, int nsample, float one, const float *__restrict__ xyz1,
                  const float *__restrict__ xyz2, int *__restrict__ idx, int *__restrict__ pts_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BallSmem &sm = *reinterpret_cast<BallSmem *>(smem_raw);
  const int scene = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float *data = xyz1 + (size_t)scene * n * 3;
  const f32x2 one2 = pack2(one, one);
  const float inf = __int_as_float(0x7f800000);
  const int thr = (s_star >= 0.0f) ? __float_as_int(s_star) + 1 : 0;  // s_star < 0: nothing can hit

  int q[kQ], total[kQ], first[kQ];
  bool live[kQ];
  f32x2 qx2[kQ], qy2[kQ], qz2[kQ];
  int *row[kQ];
  bool qbad = false;  // a non-finite query coordinate: distances may be NaN, which the reference COUNTS as a hit
#pragma unroll
  for (int u = 0; u < kQ; ++u) {
    q[u] = blockIdx.x * kQPB + u * 32 + lane;
    live[u] = q[u] < m;
    const float *qp = xyz2 + ((size_t)scene * m + (live[u] ? q[u] : 0)) * 3;
    qbad = qbad || !(fabsf(qp[0]) < inf) || !(fabsf(qp[1]) < inf) || !(fabsf(qp[2]) < inf);
    qx2[u] = pack2(qp[0], qp[0]); qy2[u] = pack2(qp[1], qp[1]); qz2[u] = pack2(qp[2], qp[2]);
    row[u] = idx + ((size_t)scene * m + (live[u] ? q[u] : 0)) * nsample;
    total[u] = live[u] ? 0 : nsample;  // hits in all earlier chunks (dead slots count as full)
    first[u] = INT_MAX;                // first hit so far (absolute index)
  }
  float(*st)[kStageRow] = sm.stage[warp];
  const bool warp_qbad = __any_sync(PC_FULL_MASK, qbad);

  for (int c0 = 0; c0 < n; c0 += kChunk) {
    const int cn = min(kChunk, n - c0);
    const int seg_len = ((cn + kSeg - 1) / kSeg + 31) / 32 * 32;  // multiple of 32: segments own whole words
    const int s_lo = min(cn, warp * seg_len), s_hi = min(cn, s_lo + seg_len);
    int cnt[kQ], sfirst[kQ];
#pragma unroll
    for (int u = 0; u < kQ; ++u) { cnt[u] = 0; sfirst[u] = INT_MAX; }
    // The stage is double-buffered through registers: the 24 coalesced loads of the NEXT stage are in flight while
    // the current one is being tested, so global latency never stalls the warp.
    constexpr int kLoads = kStage * 3 / 32;
    float pre[kLoads];
    bool pre_bad = false;  // a non-finite candidate coordinate among this lane's fetched values
    auto fetch = [&](int t0) {
      const int tn = min(kStage, s_hi - t0);
      const float *src = data + (size_t)(c0 + t0) * 3;
      pre_bad = false;
#pragma unroll
      for (int u = 0; u < kLoads; ++u) {
        const int i = lane + 32 * u;
        const bool in = i < tn * 3;
        pre[u] = in ? __ldg(src + i) : inf;  // slots past tn become +inf (never hit)
        pre_bad = pre_bad || (in && !(fabsf(pre[u]) < inf));
      }
    };
    if (s_lo < s_hi) fetch(s_lo);
    for (int t0 = s_lo; t0 < s_hi; t0 += kStage) {
      const int tn = min(kStage, s_hi - t0);
      __syncwarp();
      // AoS -> SoA transpose of the fetched candidates into this warp's stage
#pragma unroll
      for (int u = 0; u < kLoads; ++u) {
        const int i = lane + 32 * u, k = i / 3, c = i - k * 3;
        st[c][k] = pre[u];
      }
      // Non-finite coordinates (never in real clouds): the reference's test is max(sqrtf(d2),1e-20f) < radius with
      // CUDA's max = fmaxf (tf_grouping_g.cu:24), which turns a NaN distance into 1e-20 -- a HIT -- so the bit-pattern
      // compare below (NaN never hits) is replaced for this stage by the literal expression, candidate by candidate.
      const bool exact = warp_qbad || __any_sync(PC_FULL_MASK, pre_bad);
      __syncwarp();
      if (t0 + kStage < s_hi) fetch(t0 + kStage);
      const int nwords = (tn + 31) / 32;
      for (int w = 0; w < nwords; ++w) {
        unsigned word[kQ];
#pragma unroll
        for (int u = 0; u < kQ; ++u) word[u] = 0;
        if (exact) {
          for (int e = 0; e < 32; ++e) {
            const int k = w * 32 + e;
            if (k >= tn) break;
#pragma unroll
            for (int u = 0; u < kQ; ++u) {
              float qx, qy, qz, dup;
              unpack2(qx2[u], qx, dup); unpack2(qy2[u], qy, dup); unpack2(qz2[u], qz, dup);
              const float d2 = sqdist3(qx, qy, qz, st[0][k], st[1][k], st[2][k]);
              if (fmaxf(sqrtf(d2), 1e-20f) < radius) word[u] |= 0x80000000u >> e;  // bit 31-e, as the fast path
            }
          }
        } else
#pragma unroll
        for (int g = 0; g < 8; ++g) {  // 4 candidates per step
          const int k = w * 32 + g * 4;
          const float4 xs = *reinterpret_cast<const float4 *>(&st[0][k]);
          const float4 ys = *reinterpret_cast<const float4 *>(&st[1][k]);
          const float4 zs = *reinterpret_cast<const float4 *>(&st[2][k]);
          const f32x2 xa = pack2(xs.x, xs.y), ya = pack2(ys.x, ys.y), za = pack2(zs.x, zs.y);
          const f32x2 xb = pack2(xs.z, xs.w), yb = pack2(ys.z, ys.w), zb = pack2(zs.z, zs.w);
#pragma unroll
          for (int u = 0; u < kQ; ++u) {
            float d0, d1, d2, d3;
            unpack2(sqdist3_x2(qx2[u], qy2[u], qz2[u], xa, ya, za, one2), d0, d1);
            unpack2(sqdist3_x2(qx2[u], qy2[u], qz2[u], xb, yb, zb, one2), d2, d3);
            // d <= s_star  <=>  bits(d) - thr < 0 (non-negative floats order like ints; NaN bits exceed every thr):
            // the sign bit is shifted into the word, candidate j of the word lands in bit 31-j
            word[u] = __funnelshift_l(__float_as_int(d0) - thr, word[u], 1);
            word[u] = __funnelshift_l(__float_as_int(d1) - thr, word[u], 1);
            word[u] = __funnelshift_l(__float_as_int(d2) - thr, word[u], 1);
            word[u] = __funnelshift_l(__float_as_int(d3) - thr, word[u], 1);
          }
        }
        const int wi = (t0 >> 5) + w;  // word index inside the chunk
#pragma unroll
        for (int u = 0; u < kQ; ++u) {
          const unsigned wd = __brev(word[u]);  // bit j <-> candidate j
          sm.bits[wi][u * 32 + lane] = wd;
          if (wd && sfirst[u] == INT_MAX) sfirst[u] = c0 + wi * 32 + __ffs(wd) - 1;
          cnt[u] += __popc(wd);
        }
      }
    }
    bool full = true;
#pragma unroll
    for (int u = 0; u < kQ; ++u) {
      sm.seg_cnt[warp][u * 32 + lane] = cnt[u];
      sm.seg_first[warp][u * 32 + lane] = sfirst[u];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kQ; ++u) {
      int base = total[u], chunk_total = 0;
#pragma unroll
      for (int s = 0; s < kSeg; ++s) {
        const int c = sm.seg_cnt[s][u * 32 + lane];
        if (s < warp) base += c;
        chunk_total += c;
        first[u] = min(first[u], sm.seg_first[s][u * 32 + lane]);
      }
      if (live[u] && cnt[u] > 0 && base < nsample) {  // this segment contributes slots [base, base + cnt) of the row
        int pos = base;
        for (int wi = s_lo >> 5; wi < (s_hi + 31) >> 5 && pos < nsample; ++wi) {
          unsigned wd = sm.bits[wi][u * 32 + lane];
          while (wd && pos < nsample) {
            row[u][pos++] = c0 + wi * 32 + __ffs(wd) - 1;
            wd &= wd - 1;
          }
        }
      }
      total[u] = min(nsample, total[u] + chunk_total);
      full = full && (total[u] >= nsample);
    }
    if (__syncthreads_and(full)) break;  // also protects bits / seg_* against the next chunk's writers
  }

#pragma unroll
  for (int u = 0; u < kQ; ++u) {
    if (live[u]) {
      const int fill = (first[u] == INT_MAX) ? 0 : first[u];  // tf_grouping_g.cu:26-29; empty ball -> zero row
      for (int l = total[u] + warp; l < nsample; l += kSeg) row[u][l] = fill;
      if (warp == 0) pts_cnt[(size_t)scene * m + q[u]] = total[u];
    }
  }
}

}  // namespace

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
  const size_t smem = sizeof(pc::BallSmem);
  PC_CUDA_TRY(pc::allow_smem(pc::ball_query_kernel, smem));
  dim3 grid((m + pc::kQPB - 1) / pc::kQPB, b);
  pc::ball_query_kernel<<<grid, pc::kSeg * 32, smem, st>>>(n, m, s_star, radius, nsample, 1.0f, xyz1, xyz2, idx, pts_cnt);
  PC_RETURN_LAUNCH_STATUS();
}
