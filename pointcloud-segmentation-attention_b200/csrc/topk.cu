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
// array gives the reference's output exactly.  Phase 1 streams the dataset through shared memory, one warp per
// query, keeping (b) as a sorted list spread over the warp's registers (insert = ballot + shuffle-up); phase 2
// replays the swaps in shared memory.
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
constexpr int kKnnWarps = 8;
constexpr int kKnnTile = 1024;  // dataset points per shared-memory tile
constexpr int kInfBits = 0x7f800000;

// R = registers per lane for the candidate list, capacity 32*R >= k.  Sorted index e lives in slot e/32, lane e%32.
template <int R>
__global__ void __launch_bounds__(kKnnWarps * 32)
knn_kernel(int n, int m, int k, int c, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
           float *__restrict__ val, int *__restrict__ idx) {
  extern __shared__ float smem[];
  float *tile = smem;                                            // kKnnTile * c
  const int cap = 32 * R;
  float *wbase = smem + (size_t)kKnnTile * c + (size_t)(threadIdx.x >> 5) * (3 * (size_t)(k + cap));
  float *eV = wbase;                                             // replay arrays, k + cap entries each
  int *eP = reinterpret_cast<int *>(wbase + (k + cap));
  int *eI = reinterpret_cast<int *>(wbase + 2 * (k + cap));

  const int scene = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * kKnnWarps + warp;
  const bool live = q < m;
  const float *data = xyz1 + (size_t)scene * n * c;
  const float *qp = xyz2 + ((size_t)scene * m + (live ? q : 0)) * c;

  float cv[R];
  int cp[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { cv[r] = __int_as_float(kInfBits); cp[r] = INT_MAX; }
  float tau = __int_as_float(kInfBits);  // value of the last (worst) list entry

  for (int t0 = 0; t0 < n; t0 += kKnnTile) {
    const int tn = min(kKnnTile, n - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * c; i += kKnnWarps * 32) tile[i] = data[(size_t)t0 * c + i];
    __syncthreads();
    if (!live) continue;
    for (int k0 = 0; k0 < tn; k0 += 32) {
      const int kk = k0 + lane;
      const int pos = t0 + kk;
      float d = __int_as_float(kInfBits);
      if (kk < tn) {
        const float *p = tile + (size_t)kk * c;
        for (int l = 0; l < c; ++l) {  // tf_grouping.py:64-66: sum_l (xyz1 - xyz2)^2, channels in order
          const float t = __fsub_rn(p[l], __ldg(qp + l));
          const float sq = __fmul_rn(t, t);
          d = (l == 0) ? sq : __fadd_rn(d, sq);
        }
        if (pos < k) { eV[pos] = d; eP[pos] = pos; eI[pos] = pos; }  // head element
      }
      unsigned mask = __ballot_sync(PC_FULL_MASK, kk < tn && pos >= k && d < tau);
      while (mask) {  // insert candidates in ascending position
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const float xv = __shfl_sync(PC_FULL_MASK, d, src);
        const int xp = t0 + k0 + src;
        if (!(xv < tau)) continue;  // tau may have tightened since the ballot
        // first sorted index whose value is > xv (strict: equal values keep the earlier position first)
        int ins = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) ins += __popc(__ballot_sync(PC_FULL_MASK, !(xv < cv[r])));
        // shift entries >= ins up by one, from the top slot down
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
          float upv = __shfl_up_sync(PC_FULL_MASK, cv[r], 1);
          int upp = __shfl_up_sync(PC_FULL_MASK, cp[r], 1);
          if (r > 0) {  // lane 0 of this slot continues from lane 31 of the slot below
            const float lastv = __shfl_sync(PC_FULL_MASK, cv[r > 0 ? r - 1 : 0], 31);
            const int lastp = __shfl_sync(PC_FULL_MASK, cp[r > 0 ? r - 1 : 0], 31);
            if (lane == 0) { upv = lastv; upp = lastp; }
          }
          const int e = r * 32 + lane;
          if (e > ins) { cv[r] = upv; cp[r] = upp; }
          else if (e == ins) { cv[r] = xv; cp[r] = xp; }
        }
        tau = __shfl_sync(PC_FULL_MASK, cv[R - 1], 31);
      }
    }
  }
  if (!live) return;
  __syncwarp();

  // replay entries: [0,k) head, [k, k+ncand) best candidates among positions >= k (already sorted)
  int ncand = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool real = cp[r] != INT_MAX;
    ncand += __popc(__ballot_sync(PC_FULL_MASK, real));
    if (real) { eV[k + r * 32 + lane] = cv[r]; eP[k + r * 32 + lane] = cp[r]; eI[k + r * 32 + lane] = cp[r]; }
  }
  const int E = k + ncand;
  __syncwarp();
  float *vo = val + ((size_t)scene * m + q) * k;
  int *io = idx + ((size_t)scene * m + q) * k;
  for (int s = 0; s < k; ++s) {
    int bv = INT_MAX, bp = INT_MAX, be = -1;  // value bits (non-negative floats order like ints), position, entry
    for (int e = s + lane; e < E; e += 32) {
      const int v = __float_as_int(eV[e]);
      const int p = eP[e];
      if (v < bv || (v == bv && p < bp)) { bv = v; bp = p; be = e; }
    }
    const int vmin = __reduce_min_sync(PC_FULL_MASK, bv);
    const int pmin = __reduce_min_sync(PC_FULL_MASK, bv == vmin ? bp : INT_MAX);
    const unsigned owner = __ballot_sync(PC_FULL_MASK, bv == vmin && bp == pmin);
    const int emin = __shfl_sync(PC_FULL_MASK, be, __ffs(owner) - 1);
    if (lane == 0) {
      if (emin != s) {
        const float tv = eV[emin]; eV[emin] = eV[s]; eV[s] = tv;
        const int ti = eI[emin]; eI[emin] = eI[s]; eI[s] = ti;
      }
      vo[s] = eV[s];
      io[s] = eI[s];
    }
    __syncwarp();
  }
}

template <int R>
int launch_knn(int b, int n, int m, int k, int c, const float *xyz1, const float *xyz2, float *val, int *idx,
               cudaStream_t st) {
  const size_t smem = ((size_t)kKnnTile * c + (size_t)kKnnWarps * 3 * (k + 32 * R)) * sizeof(float);
  if (smem > 48 * 1024) PC_CUDA_TRY(allow_smem(knn_kernel<R>, smem));
  dim3 grid((m + kKnnWarps - 1) / kKnnWarps, b);
  knn_kernel<R><<<grid, kKnnWarps * 32, smem, st>>>(n, m, k, c, xyz1, xyz2, val, idx);
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
  if (k <= 32) return pc::launch_knn<1>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
  if (k <= 64) return pc::launch_knn<2>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
  return pc::launch_knn<4>(b, n, m, k, c, xyz1, xyz2, val, idx, st);
}
