// Whole-scene chunker and map-back for sm_100a (BASELINE config 4).
//
// Replaces the numpy body of get_all_subsets_with_all_points_for_scene_features
// (attention_points/scannet_dataset/complete_scene_loader.py:4-117) and map_back
// (attention_points/benchmark/generate_predictions.py:19-37).  The reference makes one O(N) boolean pass over the
// whole scan per 1.5 m cell, fancy-indexes every feature array per cell, shuffles a Python list and concatenates
// chunk by chunk.  Here:
//   pc_scene_cells          one launch: membership of every point in every padded cell, stable compaction into
//                           per-cell index lists (ascending point index = numpy boolean-mask order) + inner-cell flags
//   pc_scene_chunk_masksum  per candidate chunk: how many of its points lie in the un-padded cell (the reference drops
//                           chunks with none, :63,:99)
//   pc_scene_chunk_assemble the kept chunks: source index, coordinates, mask and original index of each of the 8192
//                           rows, following the host-drawn shuffle order / fill-up indices (numpy's RNG stream stays on
//                           the host so results are identical to the reference under the same np.random state)
//   pc_gather_rows_bytes    feature rows of any dtype by source index (labels, colours, normals, predictions)
//   pc_scene_sample_weights label weights x mask (:66-70 full chunks, :100-103 the fill-up chunk -- not masked there)
//   pc_map_back_winner      last-write-wins owner of every original point (numpy fancy assignment semantics)
// All of it is byte / index work: HBM- and latency-bound, no floating-point arithmetic except compares.
#include <type_traits>

#include "common.cuh"

namespace pc {
namespace {

constexpr int kCellThreads = 256;
constexpr int kCellSlabs = 8;  // a CTA covers 8 slabs of 256 consecutive points of one cell

// boxes: per cell 12 floats = padded lo[3], padded hi[3], inner lo[3], inner hi[3]; thresholds already rounded on
// the host so that the fp32 compare equals the reference's float64 compare (complete_scene_loader.py:35,41).
__device__ __forceinline__ bool inside(float x, float y, float z, const float *b) {
  return x >= b[0] && y >= b[1] && z >= b[2] && x <= b[3] && y <= b[4] && z <= b[5];
}

__global__ void __launch_bounds__(kCellThreads)
cell_count_kernel(int n, int nblk, const float *__restrict__ points, const float *__restrict__ boxes,
                  int *__restrict__ counts) {
  __shared__ float box[12];
  __shared__ int total;
  const int cell = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x;
  if (tid < 12) box[tid] = boxes[cell * 12 + tid];
  if (tid == 0) total = 0;
  __syncthreads();
  int c = 0;
#pragma unroll
  for (int k = 0; k < kCellSlabs; ++k) {
    const int p = (blk * kCellSlabs + k) * kCellThreads + tid;
    if (p < n) c += inside(__ldg(points + 3 * (size_t)p), __ldg(points + 3 * (size_t)p + 1), __ldg(points + 3 * (size_t)p + 2), box);
  }
  c = __reduce_add_sync(PC_FULL_MASK, c);
  if ((tid & 31) == 0 && c) atomicAdd(&total, c);
  __syncthreads();
  if (tid == 0) counts[cell * nblk + blk] = total;
}

// exclusive scan of counts[0..total) -> offsets; cell_base[cell] = offsets[cell*nblk], cell_base[ncells] = grand total
__global__ void __launch_bounds__(1024)
cell_scan_kernel(int ncells, int nblk, const int *__restrict__ counts, int *__restrict__ offsets,
                 int *__restrict__ cell_base) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = ncells * nblk;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < total; base += 1024) {
    const int i = base + tid;
    const int v = i < total ? counts[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(PC_FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      int iw = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(PC_FULL_MASK, iw, o);
        if (lane >= o) iw += u;
      }
      s_warp[lane] = iw - w;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - v;
    if (i < total) {
      offsets[i] = excl;
      if (i % nblk == 0) cell_base[i / nblk] = excl;
    }
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) cell_base[ncells] = s_carry;
}

__global__ void __launch_bounds__(kCellThreads)
cell_scatter_kernel(int n, int nblk, const float *__restrict__ points, const float *__restrict__ boxes,
                    const int *__restrict__ offsets, int *__restrict__ list, unsigned char *__restrict__ inner) {
  __shared__ float box[12];
  __shared__ int s_cnt[kCellThreads / 32];
  const int cell = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 12) box[tid] = boxes[cell * 12 + tid];
  __syncthreads();
  int base = offsets[cell * nblk + blk];
  for (int k = 0; k < kCellSlabs; ++k) {
    const int p = (blk * kCellSlabs + k) * kCellThreads + tid;
    bool hit = false, in = false;
    if (p < n) {
      const float x = __ldg(points + 3 * (size_t)p), y = __ldg(points + 3 * (size_t)p + 1), z = __ldg(points + 3 * (size_t)p + 2);
      hit = inside(x, y, z, box);
      in = hit && inside(x, y, z, box + 6);
    }
    const unsigned bal = __ballot_sync(PC_FULL_MASK, hit);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kCellThreads / 32; ++w) {
      const int c = s_cnt[w];
      before += (w < warp) ? c : 0;
      all += c;
    }
    if (hit) {
      const int dst = base + before + __popc(bal & lanemask_lt());
      list[dst] = p;
      inner[dst] = in ? 1 : 0;
    }
    base += all;
    __syncthreads();
  }
}

// chunk descriptor: 5 ints {list_base, order_off, start, rest, fill_off}: row t of the chunk is
//   t <  rest : cell position order[order_off + start + t]
//   t >= rest : cell position order[order_off + fill[fill_off + t - rest]]   (fill-up rows, complete_scene_loader.py:87-90)
// and its source point is list[list_base + position].
struct ChunkDesc { int list_base, order_off, start, rest, fill_off; };
__device__ __forceinline__ ChunkDesc load_desc(const int *desc, int chunk) {
  const int *d = desc + 5 * (size_t)chunk;
  return ChunkDesc{d[0], d[1], d[2], d[3], d[4]};
}
__device__ __forceinline__ int chunk_pos(const ChunkDesc &d, int t, const int *order, const int *fill) {
  return t < d.rest ? __ldg(order + d.order_off + d.start + t) : __ldg(order + d.order_off + __ldg(fill + d.fill_off + (t - d.rest)));
}

__global__ void __launch_bounds__(256)
chunk_masksum_kernel(int npoints, const int *__restrict__ desc, const int *__restrict__ order,
                     const unsigned char *__restrict__ inner, int *__restrict__ masksum) {
  __shared__ int total;
  const ChunkDesc d = load_desc(desc, blockIdx.x);
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int c = 0;
  for (int t = threadIdx.x; t < d.rest && t < npoints; t += blockDim.x)
    c += inner[d.list_base + __ldg(order + d.order_off + d.start + t)];
  c = __reduce_add_sync(PC_FULL_MASK, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&total, c);
  __syncthreads();
  if (threadIdx.x == 0) masksum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256)
chunk_assemble_kernel(int npoints, const int *__restrict__ desc, const int *__restrict__ order,
                      const int *__restrict__ fill, const int *__restrict__ list, const unsigned char *__restrict__ inner,
                      const float *__restrict__ points, int *__restrict__ src_index, float *__restrict__ point_sets,
                      unsigned char *__restrict__ masks, long long *__restrict__ orig_idx) {
  const ChunkDesc d = load_desc(desc, blockIdx.y);
  const size_t row0 = (size_t)blockIdx.y * npoints;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < npoints; t += gridDim.x * blockDim.x) {
    const int pos = chunk_pos(d, t, order, fill);
    const int src = __ldg(list + d.list_base + pos);
    const bool own = t < d.rest;
    src_index[row0 + t] = src;
    masks[row0 + t] = own ? inner[d.list_base + pos] : 0;          // fill-up rows never count (:92)
    orig_idx[row0 + t] = own ? (long long)src : 0ll;               // and carry original index 0 (:93-94)
    float *o = point_sets + (row0 + t) * 3;
    o[0] = __ldg(points + 3 * (size_t)src);
    o[1] = __ldg(points + 3 * (size_t)src + 1);
    o[2] = __ldg(points + 3 * (size_t)src + 2);
  }
}

// out[r] = idx[r] >= 0 ? src[idx[r]] : zeros, rows of row_bytes bytes.  W = bytes moved per thread step.
template <int W>
__global__ void __launch_bounds__(256)
gather_rows_bytes_kernel(size_t rows, FastDiv per_row, bool small, const unsigned char *__restrict__ src,
                         const int *__restrict__ idx, unsigned char *__restrict__ out) {
  typedef typename std::conditional<W == 4, unsigned int, unsigned char>::type T;
  const size_t units = per_row.d;  // row_bytes / W
  const size_t total = rows * units;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = small ? (size_t)per_row.div((uint32_t)e) : e / units;
    const size_t u = e - r * units;
    const int s = __ldg(idx + r);
    reinterpret_cast<T *>(out)[e] = s >= 0 ? __ldg(reinterpret_cast<const T *>(src) + (size_t)s * units + u) : T(0);
  }
}

__global__ void __launch_bounds__(256)
sample_weights_kernel(size_t total, FastDiv per_chunk, const int *__restrict__ desc, const int *__restrict__ labels,
                      const unsigned char *__restrict__ masks, double *__restrict__ out) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const uint32_t chunk = per_chunk.div((uint32_t)e);
    const bool full = desc[5 * (size_t)chunk + 3] >= (int)per_chunk.d;
    double w = 1.0;
    if (labels) w = (labels[e] == 0) ? 0.0 : 1.0;   // label_weights = ones(21), [0] = 0 (:12-13)
    if (full) w *= masks[e] ? 1.0 : 0.0;            // only the full chunks are masked (:70 vs :100-103)
    out[e] = w;
  }
}

__global__ void __launch_bounds__(256)
winner_kernel(size_t rows, const long long *__restrict__ orig_idx, const unsigned char *__restrict__ mask, int nres,
              int *__restrict__ winner) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    if (!mask[r]) continue;
    const long long o = orig_idx[r];
    if (o >= 0 && o < nres) atomicMax(winner + o, (int)r);   // numpy fancy assignment: the last occurrence wins
  }
}

// coordmin / coordmax of a scan (complete_scene_loader.py:21-22): one CTA, no atomics.  out = min[3], max[3].
__global__ void __launch_bounds__(1024)
bbox_kernel(int n, const float *__restrict__ points, float *__restrict__ out) {
  __shared__ float s_lo[32][3], s_hi[32][3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int p = tid; p < n; p += 1024)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = __ldg(points + 3 * (size_t)p + a);
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(PC_FULL_MASK, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(PC_FULL_MASK, hi[a], o));
    }
    if (lane == 0) { s_lo[warp][a] = lo[a]; s_hi[warp][a] = hi[a]; }
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float l = s_lo[lane][a], h = s_hi[lane][a];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        l = fminf(l, __shfl_xor_sync(PC_FULL_MASK, l, o));
        h = fmaxf(h, __shfl_xor_sync(PC_FULL_MASK, h, o));
      }
      if (lane == 0) { out[a] = l; out[3 + a] = h; }
    }
  }
}

}  // namespace
}  // namespace pc

extern "C" int pc_scene_bbox(int n, const float *points, float *out6, pc_stream_t stream) {
  if (n <= 0 || !points || !out6) return PC_ERR_INVALID_ARGUMENT;
  pc::bbox_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, points, out6);
  PC_RETURN_LAUNCH_STATUS();
}

static inline int scene_nblk(int n) {
  const int per = pc::kCellThreads * pc::kCellSlabs;
  return (n + per - 1) / per;
}

extern "C" size_t pc_scene_cells_workspace_bytes(int n, int ncells) {
  if (n <= 0 || ncells <= 0) return 0;
  return 2 * (size_t)ncells * scene_nblk(n) * sizeof(int);
}

extern "C" int pc_scene_cells(int n, int ncells, const float *points, const float *boxes, int *cell_base, int *list,
                              unsigned char *inner, void *workspace, pc_stream_t stream) {
  if (n < 0 || ncells < 0) return PC_ERR_INVALID_ARGUMENT;
  if (ncells == 0) return PC_OK;
  if (!cell_base) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    PC_CUDA_TRY(cudaMemsetAsync(cell_base, 0, sizeof(int) * ((size_t)ncells + 1), st));
    return PC_OK;
  }
  if (!points || !boxes || !list || !inner) return PC_ERR_INVALID_ARGUMENT;
  if (!workspace) return PC_ERR_WORKSPACE;
  if (ncells > 65535 || (size_t)ncells * scene_nblk(n) > (size_t)INT32_MAX) return PC_ERR_UNSUPPORTED;
  const int nblk = scene_nblk(n);
  int *counts = (int *)workspace, *offsets = counts + (size_t)ncells * nblk;
  dim3 grid(nblk, ncells);
  pc::cell_count_kernel<<<grid, pc::kCellThreads, 0, st>>>(n, nblk, points, boxes, counts);
  pc::cell_scan_kernel<<<1, 1024, 0, st>>>(ncells, nblk, counts, offsets, cell_base);
  pc::cell_scatter_kernel<<<grid, pc::kCellThreads, 0, st>>>(n, nblk, points, boxes, offsets, list, inner);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_scene_chunk_masksum(int nchunks, int npoints, const int *desc, const int *order,
                                      const unsigned char *inner, int *masksum, pc_stream_t stream) {
  if (nchunks < 0 || npoints <= 0) return PC_ERR_INVALID_ARGUMENT;
  if (nchunks == 0) return PC_OK;
  if (!desc || !order || !inner || !masksum) return PC_ERR_INVALID_ARGUMENT;
  pc::chunk_masksum_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>(npoints, desc, order, inner, masksum);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_scene_chunk_assemble(int nchunks, int npoints, const int *desc, const int *order, const int *fill,
                                       const int *list, const unsigned char *inner, const float *points,
                                       int *src_index, float *point_sets, unsigned char *masks, long long *orig_idx,
                                       pc_stream_t stream) {
  if (nchunks < 0 || npoints <= 0) return PC_ERR_INVALID_ARGUMENT;
  if (nchunks == 0) return PC_OK;
  if (!desc || !order || !fill || !list || !inner || !points || !src_index || !point_sets || !masks || !orig_idx)
    return PC_ERR_INVALID_ARGUMENT;
  if (nchunks > 65535) return PC_ERR_UNSUPPORTED;
  int gx = (npoints + 255) / 256;
  if (gx > 32) gx = 32;
  pc::chunk_assemble_kernel<<<dim3(gx, nchunks), 256, 0, (cudaStream_t)stream>>>(
      npoints, desc, order, fill, list, inner, points, src_index, point_sets, masks, orig_idx);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_gather_rows_bytes(size_t rows, int row_bytes, const void *src, const int *src_index, void *out,
                                    pc_stream_t stream) {
  if (row_bytes < 0) return PC_ERR_INVALID_ARGUMENT;
  if (rows == 0 || row_bytes == 0) return PC_OK;
  if (!src || !src_index || !out) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const bool words = row_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0;
  const size_t units = words ? row_bytes / 4 : row_bytes, total = rows * units;
  const bool small = total < (1ull << 31);
  if (words) {
    const int blocks = pc::resident_grid((const void *)pc::gather_rows_bytes_kernel<4>, 256, 0, (total + 255) / 256);
    pc::gather_rows_bytes_kernel<4><<<blocks, 256, 0, st>>>(rows, pc::FastDiv((uint32_t)units), small, (const unsigned char *)src,
                                                            src_index, (unsigned char *)out);
  } else {
    const int blocks = pc::resident_grid((const void *)pc::gather_rows_bytes_kernel<1>, 256, 0, (total + 255) / 256);
    pc::gather_rows_bytes_kernel<1><<<blocks, 256, 0, st>>>(rows, pc::FastDiv((uint32_t)units), small, (const unsigned char *)src,
                                                            src_index, (unsigned char *)out);
  }
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_scene_sample_weights(int nchunks, int npoints, const int *desc, const int *labels,
                                       const unsigned char *masks, double *out, pc_stream_t stream) {
  if (nchunks < 0 || npoints <= 0) return PC_ERR_INVALID_ARGUMENT;
  if (nchunks == 0) return PC_OK;
  if (!desc || !masks || !out) return PC_ERR_INVALID_ARGUMENT;
  const size_t total = (size_t)nchunks * npoints;
  if (total >= (1ull << 31)) return PC_ERR_UNSUPPORTED;
  const int blocks = pc::resident_grid((const void *)pc::sample_weights_kernel, 256, 0, (total + 255) / 256);
  pc::sample_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(total, pc::FastDiv((uint32_t)npoints), desc, labels, masks, out);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_map_back_winner(size_t rows, int nres, const long long *orig_idx, const unsigned char *mask,
                                  int *winner, pc_stream_t stream) {
  if (nres < 0) return PC_ERR_INVALID_ARGUMENT;
  if (nres == 0) return PC_OK;
  if (!winner) return PC_ERR_INVALID_ARGUMENT;
  if (rows >= (1ull << 31)) return PC_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  PC_CUDA_TRY(cudaMemsetAsync(winner, 0xff, sizeof(int) * (size_t)nres, st));   // -1: nobody wrote this point
  if (rows == 0) return PC_OK;
  if (!orig_idx || !mask) return PC_ERR_INVALID_ARGUMENT;
  const int blocks = pc::resident_grid((const void *)pc::winner_kernel, 256, 0, (rows + 255) / 256);
  pc::winner_kernel<<<blocks, 256, 0, st>>>(rows, orig_idx, mask, nres, winner);
  PC_RETURN_LAUNCH_STATUS();
}
