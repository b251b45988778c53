// SelectionSort and fused knn_point for sm_100a.
//
// SelectionSort replaces selection_sort_gpu / selectionSortLauncher (reference tf_ops/grouping/tf_grouping_g.cu:
// 83-123,129-132): full-length (b,m,n) outputs holding exactly what the reference's k swap steps leave behind.  One
// warp owns a row: each step is a parallel arg-min over [s,n) by (value, position) -- the same element the reference's
// strict-'<' scan finds -- followed by the same swap.
//
// pc_knn replaces the TF graph of tf_ops/grouping/tf_grouping.py:48-73 (tile, subtract, square, reduce_sum,
// SelectionSort, slice) without ever forming the (b,m,n) matrix, and still returns the swap-induced (unstable) tie
// order of the selection sort.  Why that is possible: after s swap steps the unsorted tail equals the original
// array except at <= s positions that received displaced elements, and every element ever displaced started at a
// position < k.  So the first k outputs depend only on (a) the k "head" elements at positions 0..k-1 and (b) the k
// best elements by (value, position) among positions >= k; replaying the k swap steps on that sparse 2k-entry
// array gives the reference's output exactly -- and so does any superset of (b), because a tail element outside
// the k best is never selected within k steps.  One warp per query: a first pass over the dataset finds a distance
// bound that provably covers (b), a second pass collects the points under it, then the swaps are replayed.
#include "common.cuh"

namespace pc {
namespace {

// ------------------------------------------------------------------------------------------- SelectionSort
__global__ void __launch_bounds__(256)
selection_sort_kernel(size_t rows, int n, int k, const float *__restrict__ dist, int *__restrict__ outi,
                      float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t warp_global = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t r = warp_global; r < rows; r += nwarps) {
    const float *d = dist + r * n;
    float *o = out + r * n;
    int *oi = outi + r * n;
    for (int s = lane; s < n; s += 32) { o[s] = d[s]; oi[s] = s; }  // tf_grouping_g.cu:96-101
    __syncwarp();
    const int kk = min(k, n);
    for (int s = 0; s < kk; ++s) {
      // lane-local first minimum over t = s+lane, s+lane+32, ...
      float bv = 0.f;
      int bt = -1;
      for (int t = s + lane; t < n; t += 32) {
        const float v = o[t];
        if (bt < 0 || v < bv) { bv = v; bt = t; }
      }
      // warp arg-min by (value, position); lanes without elements carry bt = -1
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_xor_sync(PC_FULL_MASK, bv, off);
        const int ot = __shfl_xor_sync(PC_FULL_MASK, bt, off);
        const bool take = (ot >= 0) && (bt < 0 || ov < bv || (ov == bv && ot < bt));
        if (take) { bv = ov; bt = ot; }
      }
      if (lane == 0 && bt != s) {  // tf_grouping_g.cu:116-121
        const float tv = o[bt]; o[bt] = o[s]; o[s] = tv;
        const int ti = oi[bt]; oi[bt] = oi[s]; oi[s] = ti;
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------- fused kNN
constexpr int kKnnWarps = 4;
constexpr int kKnnQW = 4;       // queries per warp: one staged dataset tile serves 16 queries
constexpr int kKnnTile = 1024;  // dataset points per shared-memory tile (stored channel-major: tile[l][point])
constexpr int kKnnStep = 128;   // points per warp step: lane L takes points 4L .. 4L+3 of the step

// k-th smallest (1-based) of the 32*E values spread over the warp's registers, built bit by bit from the top: squared
// distances are non-negative floats, so their bit patterns order like integers.  INT_MAX pads; fewer than k real
// values give INT_MAX.
template <int E>
__device__ __forceinline__ int kth_smallest(const int (&v)[E], int k) {
  int T = 0;
  for (int bit = 30; bit >= 0; --bit) {
    const int cand = T | (1 << bit);
    int c = 0;
#pragma unroll
    for (int r = 0; r < E; ++r) c += v[r] < cand;
    if (__reduce_add_sync(PC_FULL_MASK, c) < k) T = cand;  // fewer than k values below cand: the k-th is >= cand
  }
  return T;
}

// Overflow path: cuts a candidate list (cnt >= k entries, any order, cnt <= 32 E) back to exactly its k best by
// (value, position) and returns the k-th smallest value.  Ties at that value keep the smallest positions.
template <int E>
__device__ __noinline__ int knn_compact(int *bv, int *bp, int cnt, int k) {
  const int lane = lane_id();
  int v[E], p[E];
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int e = r * 32 + lane;
    v[r] = e < cnt ? bv[e] : INT_MAX;
    p[r] = e < cnt ? bp[e] : INT_MAX;
  }
  __syncwarp();
  const int T = kth_smallest<E>(v, k);
  int clt = 0, ceq = 0;
#pragma unroll
  for (int r = 0; r < E; ++r) { clt += v[r] < T; ceq += v[r] == T; }
  const int need = k - __reduce_add_sync(PC_FULL_MASK, clt);  // entries equal to T to keep
  int P = INT_MAX;
  if (__reduce_add_sync(PC_FULL_MASK, ceq) > need) {          // more ties than room: the need-th smallest position
    int t[E];
#pragma unroll
    for (int r = 0; r < E; ++r) t[r] = v[r] == T ? p[r] : INT_MAX;
    P = kth_smallest<E>(t, need);
  }
  const unsigned lt_mask = lanemask_lt();
  int off = 0;
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const bool keep = v[r] < T || (v[r] == T && p[r] <= P);
    const unsigned mk = __ballot_sync(PC_FULL_MASK, keep);
    if (keep) { const int o = off + __popc(mk & lt_mask); bv[o] = v[r]; bp[o] = p[r]; }
    off += __popc(mk);
  }
  __syncwarp();
  return T;
}

// Four squared distances (value bits) of one query to the points 4L .. 4L+3 of a step, full step, no head element.
__device__ __forceinline__ void knn_dist4(const float4 &X, const float4 &Y, const float4 &Z, f32x2 qx2, f32x2 qy2,
                                          f32x2 qz2, f32x2 one2, int (&db)[4]) {
  float d0, d1, d2, d3;
  unpack2(sqdist3_x2(pack2(X.x, X.y), pack2(Y.x, Y.y), pack2(Z.x, Z.y), qx2, qy2, qz2, one2), d0, d1);
  unpack2(sqdist3_x2(pack2(X.z, X.w), pack2(Y.z, Y.w), pack2(Z.z, Z.w), qx2, qy2, qz2, one2), d2, d3);
  db[0] = __float_as_int(d0); db[1] = __float_as_int(d1); db[2] = __float_as_int(d2); db[3] = __float_as_int(d3);
}

// The same with every check: partial steps, the k head elements (written to the replay arrays, never candidates),
// any channel count.  tf_grouping.py:64-66: sum_l (xyz1 - xyz2)^2, channels in order.  Out of line: it runs for the
// first and the last step of a scene only and must not bloat the streaming loops.
template <bool kC3>
__device__ __noinline__ int4 knn_dist4_checked(const float *tile, int kk, int tn, int t0, int k, int c,
                                               const float *__restrict__ qp, float *eV, int *eP) {
  int db[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int pos = t0 + kk + j;
    db[j] = INT_MAX;
    if (kk + j < tn) {
      float d;
      if (kC3) {
        d = sqdist3(tile[kk + j], tile[kKnnTile + kk + j], tile[2 * kKnnTile + kk + j], __ldg(qp), __ldg(qp + 1),
                    __ldg(qp + 2));
      } else {
        d = 0.f;
        for (int l = 0; l < c; ++l) {
          const float t = __fsub_rn(tile[l * kKnnTile + kk + j], __ldg(qp + l));
          const float sq = __fmul_rn(t, t);
          d = (l == 0) ? sq : __fadd_rn(d, sq);
        }
      }
      if (pos < k) { eV[pos] = d; eP[pos] = pos; }
      else db[j] = __float_as_int(d);
    }
  }
  return make_int4(db[0], db[1], db[2], db[3]);
}

// Pass 2, some lane holds a candidate: lanes claim list slots with a shared-memory atomic (list order is irrelevant),
// and a list that could overflow on the next step is cut back to its k best.  Returns the (possibly tightened)
// admission bound.  Out of line for the same reason.
template <int CB>
__device__ __noinline__ int knn_append(int4 d, int taux, int pos0, int *bv, int *bp, int *counter, int k) {
  const int db[4] = {d.x, d.y, d.z, d.w};
  const int nl = (db[0] < taux) + (db[1] < taux) + (db[2] < taux) + (db[3] < taux);
  if (nl) {
    int o = atomicAdd(counter, nl);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (db[j] < taux) { bv[o] = db[j]; bp[o] = pos0 + j; ++o; }
  }
  __syncwarp();
  const int cq = *reinterpret_cast<volatile int *>(counter);
  if (cq > CB - kKnnStep) {
    taux = knn_compact<CB / 32>(bv, bp, cq, k);
    if (lane_id() == 0) *counter = k;
    __syncwarp();
  }
  return taux;
}

// One warp per kKnnQW queries, two passes over the dataset, which streams through a double-buffered shared-memory
// tile (cp.async, transposed to channel-major on the way in) shared by the block's 16 queries.
//   Pass 1 keeps, per lane and query, the minimum distance of GL disjoint groups of points (32 GL >= 4k groups per
//          query).  The k-th smallest group minimum t has at least k points at or below it (one per group), so each of
//          the k best candidates has distance <= t; with 4k groups about 1.15 k points of the whole dataset do.
//   Pass 2 recomputes the distances and appends the points with distance <= t to the query's candidate list (ballot +
//          popc).  A superset of the k best is all the replay needs (an extra tail element is never selected within k
//          steps), so there is no selection step at all unless duplicates overflow the list (knn_compact).
//   Phase 3 replays the k swap steps on head + candidates, entries in registers.
template <bool kC3, int GL>
__global__ void __launch_bounds__(kKnnWarps * 32, GL == 4 ? 4 : 2)
knn_kernel(int n, int m, int k, int c, float one, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
           float *__restrict__ val, int *__restrict__ idx) {
  constexpr int QW = kKnnQW, U = GL / 4;
  constexpr int CB = 8 * GL + 160;  // candidate list capacity; compaction above CB - kKnnStep entries
  constexpr int ER = GL / 2 + 1;    // replay registers per lane: k + cnt <= 8 GL + (CB - kKnnStep) = 32 ER
  extern __shared__ __align__(16) float smem[];
  const int cc = kC3 ? 3 : c;
  const int cap = k + CB;
  float *lists = smem + 2 * (size_t)kKnnTile * cc;
  __shared__ int scnt[kKnnWarps * kKnnQW];  // candidates per query
  if (threadIdx.x < kKnnWarps * kKnnQW) scnt[threadIdx.x] = 0;

  const int scene = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float *data = xyz1 + (size_t)scene * n * cc;
  const unsigned lt_mask = lanemask_lt();
  const f32x2 one2 = pack2(one, one);

  float *eV[QW];  // replay arrays per query: [0,k) head, [k, k+cnt) candidates: value, position
  int *eP[QW];
  const float *qp[QW];
  f32x2 qx2[QW], qy2[QW], qz2[QW];
  int gmin[QW][GL], taux[QW];
#pragma unroll
  for (int qi = 0; qi < QW; ++qi) {
    const int q = min((blockIdx.x * kKnnWarps + warp) * QW + qi, m - 1);  // queries past m repeat the last one
    eV[qi] = lists + (size_t)(warp * QW + qi) * 2 * cap;
    eP[qi] = reinterpret_cast<int *>(eV[qi] + cap);
    qp[qi] = xyz2 + ((size_t)scene * m + q) * cc;
    if (kC3) {
      const float x = __ldg(qp[qi]), y = __ldg(qp[qi] + 1), z = __ldg(qp[qi] + 2);
      qx2[qi] = pack2(x, x); qy2[qi] = pack2(y, y); qz2[qi] = pack2(z, z);
    }
#pragma unroll
    for (int g = 0; g < GL; ++g) gmin[qi][g] = INT_MAX;
    taux[qi] = INT_MAX;  // pass 2 admits value bits < taux
  }

  const int ntiles = (n + kKnnTile - 1) / kKnnTile;
  auto prefetch = [&](int it) {  // tile (it mod ntiles) -> buffer (it & 1), 4-byte cp.async, AoS -> channel-major
    const int t0 = (it % ntiles) * kKnnTile, tn = min(kKnnTile, n - t0);
    float *dst = smem + (size_t)(it & 1) * kKnnTile * cc;
    for (int pt = threadIdx.x; pt < tn; pt += kKnnWarps * 32) {
      const float *src = data + (size_t)(t0 + pt) * cc;
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(dst + pt);
#pragma unroll 3
      for (int l = 0; l < cc; ++l)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + l * kKnnTile * 4), "l"(src + l) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0);
  for (int it = 0; it < 2 * ntiles; ++it) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile `it` has landed for every thread; everyone is done with the other buffer
    if (it + 1 < 2 * ntiles) prefetch(it + 1);
    const bool pass2 = it >= ntiles;
    const int t0 = (it % ntiles) * kKnnTile, tn = min(kKnnTile, n - t0);
    const float *tile = smem + (size_t)(it & 1) * kKnnTile * cc;
    if (it == ntiles) {
#pragma unroll
      for (int qi = 0; qi < QW; ++qi) {
        const int t = kth_smallest<GL>(gmin[qi], k);
        taux[qi] = t == INT_MAX ? INT_MAX : t + 1;
      }
    }
    // steps [lo, hi) of the tile are full and past the k head elements: streaming loops; the rest takes the checks
    const int hi = kC3 ? (tn & ~(kKnnStep - 1)) : 0;
    const int lo = min(hi, t0 >= k ? 0 : (k - t0 + kKnnStep - 1) & ~(kKnnStep - 1));
    // (in position order: the admission bound turns strict after a compaction, which is only right for later positions)
    auto checked_steps = [&](int from, int to) {
      for (int k0 = from; k0 < to; k0 += kKnnStep) {
        const int kk = k0 + lane * 4;
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
          const int4 d = knn_dist4_checked<kC3>(tile, kk, tn, t0, k, c, qp[qi], eV[qi], eP[qi]);
          if (!pass2) {
            gmin[qi][0] = min(gmin[qi][0], d.x); gmin[qi][1] = min(gmin[qi][1], d.y);
            gmin[qi][2] = min(gmin[qi][2], d.z); gmin[qi][3] = min(gmin[qi][3], d.w);
          } else if (__any_sync(PC_FULL_MASK, min(min(d.x, d.y), min(d.z, d.w)) < taux[qi])) {
            taux[qi] = knn_append<CB>(d, taux[qi], t0 + kk, reinterpret_cast<int *>(eV[qi]) + k, eP[qi] + k,
                                      scnt + warp * QW + qi, k);
          }
        }
      }
    };
    checked_steps(0, lo);
    if (!pass2) {
      for (int k0 = lo; k0 < hi; k0 += kKnnStep * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (u > 0 && k0 + u * kKnnStep >= hi) break;
          const int kk = k0 + u * kKnnStep + lane * 4;
          const float4 X = *reinterpret_cast<const float4 *>(tile + kk);
          const float4 Y = *reinterpret_cast<const float4 *>(tile + kKnnTile + kk);
          const float4 Z = *reinterpret_cast<const float4 *>(tile + 2 * kKnnTile + kk);
#pragma unroll
          for (int qi = 0; qi < QW; ++qi) {
            int db[4];
            knn_dist4(X, Y, Z, qx2[qi], qy2[qi], qz2[qi], one2, db);
#pragma unroll
            for (int j = 0; j < 4; ++j) gmin[qi][u * 4 + j] = min(gmin[qi][u * 4 + j], db[j]);
          }
        }
      }
    } else {
      for (int k0 = lo; k0 < hi; k0 += kKnnStep) {
        const int kk = k0 + lane * 4;
        const float4 X = *reinterpret_cast<const float4 *>(tile + kk);
        const float4 Y = *reinterpret_cast<const float4 *>(tile + kKnnTile + kk);
        const float4 Z = *reinterpret_cast<const float4 *>(tile + 2 * kKnnTile + kk);
#pragma unroll
        for (int qi = 0; qi < QW; ++qi) {
          int db[4];
          knn_dist4(X, Y, Z, qx2[qi], qy2[qi], qz2[qi], one2, db);
          if (__any_sync(PC_FULL_MASK, min(min(db[0], db[1]), min(db[2], db[3])) < taux[qi]))
            taux[qi] = knn_append<CB>(make_int4(db[0], db[1], db[2], db[3]), taux[qi], t0 + kk,
                                      reinterpret_cast<int *>(eV[qi]) + k, eP[qi] + k, scnt + warp * QW + qi, k);
        }
      }
    }
    checked_steps(hi, tn);
  }
  __syncwarp();

  // replay in registers, the warp's queries interleaved: entry e lives in register e / 32 of lane e % 32
  int rv[QW][ER], rp[QW][ER], ri[QW][ER];
#pragma unroll
  for (int qi = 0; qi < QW; ++qi) {
    const int E = k + scnt[warp * QW + qi];
#pragma unroll
    for (int r = 0; r < ER; ++r) {
      const int e = r * 32 + lane;
      rv[qi][r] = e < E ? __float_as_int(eV[qi][e]) : INT_MAX;
      rp[qi][r] = e < E ? eP[qi][e] : INT_MAX;
      ri[qi][r] = rp[qi][r];
    }
  }
  for (int s = 0; s < k; ++s) {
    const int sr = s >> 5, sl = s & 31;
#pragma unroll
    for (int qi = 0; qi < QW; ++qi) {
      int bvv = INT_MAX, bpp = INT_MAX, bi = 0, br = -1;
#pragma unroll
      for (int r = 0; r < ER; ++r) {
        const bool better = r * 32 + lane >= s && (rv[qi][r] < bvv || (rv[qi][r] == bvv && rp[qi][r] < bpp));
        if (better) { bvv = rv[qi][r]; bpp = rp[qi][r]; bi = ri[qi][r]; br = r; }
      }
      const int vmin = __reduce_min_sync(PC_FULL_MASK, bvv);
      const int pmin = __reduce_min_sync(PC_FULL_MASK, bvv == vmin ? bpp : INT_MAX);
      const int owner = __ffs(__ballot_sync(PC_FULL_MASK, bvv == vmin && bpp == pmin && br >= 0)) - 1;
      const int imin = __shfl_sync(PC_FULL_MASK, bi, owner);
      int mv = rv[qi][0], mi = ri[qi][0];
#pragma unroll
      for (int r = 1; r < ER; ++r)
        if (sr == r) { mv = rv[qi][r]; mi = ri[qi][r]; }
      const int vs = __shfl_sync(PC_FULL_MASK, mv, sl), is = __shfl_sync(PC_FULL_MASK, mi, sl);
#pragma unroll
      for (int r = 0; r < ER; ++r) {  // the swap of tf_grouping_g.cu:116-121: slot s <-> the minimum's slot
        if (lane == owner && r == br) { rv[qi][r] = vs; ri[qi][r] = is; }
        if (lane == sl && r == sr) { rv[qi][r] = vmin; ri[qi][r] = imin; }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < QW; ++qi) {
    const int q = (blockIdx.x * kKnnWarps + warp) * QW + qi;
    if (q < m) {
      float *vo = val + ((size_t)scene * m + q) * k;
      int *io = idx + ((size_t)scene * m + q) * k;
#pragma unroll
      for (int r = 0; r < ER; ++r) {
        const int e = r * 32 + lane;
        if (e < k) { vo[e] = __int_as_float(rv[qi][r]); io[e] = ri[qi][r]; }
      }
    }
  }
}

template <bool kC3, int GL>
int launch_knn(int b, int n, int m, int k, int c, const float *xyz1, const float *xyz2, float *val, int *idx,
               cudaStream_t st) {
  const size_t smem = (2 * (size_t)kKnnTile * c + (size_t)kKnnWarps * kKnnQW * 2 * (k + 8 * GL + 160)) * sizeof(float);
  if (smem > 48 * 1024) PC_CUDA_TRY(allow_smem(knn_kernel<kC3, GL>, smem));
  const int qb = kKnnWarps * kKnnQW;
  dim3 grid((m + qb - 1) / qb, b);
  knn_kernel<kC3, GL><<<grid, kKnnWarps * 32, smem, st>>>(n, m, k, c, 1.0f, xyz1, xyz2, val, idx);
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace
}  // namespace pc

extern "C" int pc_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out,
                                 pc_stream_t stream) {
  if (k <= 0) return PC_ERR_INVALID_ARGUMENT;  // tf_grouping.cpp:112-113
  if (b < 0 || n < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  const size_t rows = (size_t)b * m;
  if (rows == 0 || n == 0) return PC_OK;
  if (!dist || !outi || !out) return PC_ERR_INVALID_ARGUMENT;
  size_t blocks = (rows + 7) / 8;
  const size_t cap = (size_t)pc::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  pc::selection_sort_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rows, n, k, dist, outi, out);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_knn(int b, int n, int m, int k, int c, const float *xyz1, const float *xyz2, float *val, int *idx,
                      pc_stream_t stream) {
  if (k <= 0 || b < 0 || n < 0 || m < 0 || c < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0) return PC_OK;
  if (k > 128 || k > n || c < 1 || c > 16 || b > 65535) return PC_ERR_UNSUPPORTED;
  if (!xyz1 || !xyz2 || !val || !idx) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 32) {  // GL = k_max / 8 group minima per lane: 32 GL = 4 k_max groups
    if (c == 3) return pc::launch_knn<true, 4>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
    return pc::launch_knn<false, 4>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
  }
  if (k <= 64) {
    if (c == 3) return pc::launch_knn<true, 8>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
    return pc::launch_knn<false, 8>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
  }
  if (c == 3) return pc::launch_knn<true, 16>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
  return pc::launch_knn<false, 16>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
}
