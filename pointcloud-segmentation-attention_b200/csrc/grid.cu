// Cell-grid accelerated ball query and three_nn for sm_100a.
//
// pc_query_ball / pc_three_nn test every (query, candidate) pair like the reference kernels do
// (tf_ops/grouping/tf_grouping_g.cu:3-36, tf_ops/interpolation_3d/tf_interpolate.cpp:60-103).  The *_grid entry points
// return the SAME outputs, bit for bit, from far fewer pair tests:
//   1. grid_build_kernel bins the candidate cloud of every scene into a uniform cell grid (one CTA per scene:
//      bounding box -> shared-memory histogram -> scan -> scatter into a cell-sorted float4 array that carries the
//      original index), and bins the QUERY cloud by the same cells, so 32 consecutive sorted queries are neighbours;
//   2. a warp owns such a tile of 32 queries, stages the union of their cell neighbourhoods once, and runs the
//      branch-free packed all-pairs filter on that local candidate set only.
//   ball query   cell edge >= r * (1 + 1e-4): every point inside a ball lies in the 3x3x3 neighbourhood of the query's
//                cell.  The cell filter is only a conservative superset -- the hit test on the survivors is the
//                reference's exact un-fused arithmetic -- and hits are recorded as bits at their ORIGINAL index in a
//                per-query shared-memory bitmap, so "first nsample in ascending index", first-hit padding and pts_cnt
//                come out unchanged.
//   three_nn     after scanning (a superset of) the 3x3x3 neighbourhood every unvisited point is farther than one
//                cell edge h, so a lane whose third-best squared distance is below (0.999 h)^2 is certified; the few
//                that are not get a warp-cooperative scan of the whole cloud.  Visiting order is no longer ascending
//                index, so the 3-slot insertion orders by (distance, index) -- exactly what the reference's strict-'<'
//                scan over ascending indices keeps.
#include <math.h>
#include "common.cuh"

namespace pc {
namespace {

constexpr int kMaxCells = 16384;  // 64 KB shared-memory histogram in the build kernel
// Build CTA size follows the caller's concurrency hint.  Alone, 1024 threads finish a scene soonest (23 us at SA1).
// Among co-resident kernels a 1024-thread CTA with a 64 KB histogram (51 k registers) fits on no SM that holds a
// 512-thread FPS CTA (45 k of the SM's 64 k registers): with eight batches in flight every binned op of every batch
// queued for the ~20 FPS-free SMs, and 256-thread CTAs (<= 64 registers) with a histogram sized by the candidate count
// took the pipelined step from 69.6 k to 75.5 k scenes/s -- although the lone build takes 19-35 us that way.
constexpr int kHdrInts = 16;

struct GridHdr {  // 16 x 4 bytes, first thing in a scene's workspace
  float ox, oy, oz, inv_h, h;
  int nx, ny, nz, ncells;
  int nonfinite;  // the cloud holds a +/-inf or NaN coordinate (ball query then scans every cell with the literal test)
  int pad[6];
};
static_assert(sizeof(GridHdr) == kHdrInts * 4, "GridHdr layout");

// per-scene workspace: GridHdr | cell_start[kMaxCells + 1] | (16-byte aligned) sorted float4[n]
__host__ __device__ inline size_t grid_sorted_offset_ints() { return (size_t)((kHdrInts + kMaxCells + 1 + 3) / 4 * 4); }
__host__ __device__ inline size_t grid_scene_ints(int n) { return grid_sorted_offset_ints() + (size_t)n * 4; }

__device__ __forceinline__ int cell_coord(float x, float o, float inv_h, int dim) {
  // monotone in x (fp32 subtract of a constant, multiply by a positive constant, floor): the conservative cell range
  // of an interval is the range of its end points
  const int c = (int)floorf((x - o) * inv_h);
  return min(max(c, 0), dim - 1);
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(PC_FULL_MASK, v, o);
    v = is_max ? fmaxf(v, w) : fminf(v, w);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, s_red[i]) : fminf(r, s_red[i]);
  return r;
}

// mode 0: cell edge = min_edge (ball query, min_edge = r'); mode 1: cell edge from the point density (three_nn).
// ONE launch bins both clouds of an op, grid (scenes, 2): CTA (s, 0) bins the candidate cloud of scene s into its grid;
// CTA (s, 1) bins the QUERY cloud by the same cells (so that 32 consecutive sorted queries are spatial neighbours) --
// it derives the grid header from the candidates itself (the same reduction over the same data gives the same bits;
// one extra pass over 12 n bytes) instead of waiting for a second launch behind the first.
template <int kBuildThreads>
__global__ void __launch_bounds__(kBuildThreads, kBuildThreads == 1024 ? 1 : 4)
grid_build_kernel(int nc, int nq, float min_edge, int mode, int max_cells, const float *__restrict__ xyz_c,
                  const float *__restrict__ xyz_q, int *__restrict__ ws_c, int *__restrict__ ws_q) {
  extern __shared__ int s_cnt[];  // max_cells
  __shared__ float s_red[32];
  __shared__ GridHdr s_hdr;
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int scene = blockIdx.x, role = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *hp = xyz_c + (size_t)scene * nc * 3;                       // the cloud that defines the grid
  const int n = role ? nq : nc;                                           // the cloud this CTA bins
  const float *p = role ? xyz_q + (size_t)scene * nq * 3 : hp;
  int *base = role ? ws_q + (size_t)scene * grid_scene_ints(nq) : ws_c + (size_t)scene * grid_scene_ints(nc);
  int *cell_start = base + kHdrInts;
  float4 *sorted = reinterpret_cast<float4 *>(base + grid_sorted_offset_ints());

  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  float bad = 0.0f;
  for (int k = tid; k < nc; k += kBuildThreads) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __ldg(hp + k * 3 + c);
      if (fabsf(v) <= 3.0e38f) {  // non-finite coordinates must not stretch the box: an infinite extent has no cell
        lo[c] = fminf(lo[c], v);  // size.  three_nn never selects them (d < best is false); ball query does count a
        hi[c] = fmaxf(hi[c], v);  // NaN distance as a hit (fmaxf, tf_grouping_g.cu:24): flagged below
      } else {
        bad = 1.0f;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    lo[c] = block_reduce(lo[c], false, s_red);
    hi[c] = block_reduce(hi[c], true, s_red);
  }
  bad = block_reduce(bad, true, s_red);
  if (tid == 0) {
    // finite by construction (box over finite coordinates; a scene without any gets a 1-cell grid at the origin)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (!(lo[c] <= hi[c])) lo[c] = hi[c] = 0.f;
    const float ex = fminf(fmaxf(hi[0] - lo[0], 0.f), 3.0e38f), ey = fminf(fmaxf(hi[1] - lo[1], 0.f), 3.0e38f),
                ez = fminf(fmaxf(hi[2] - lo[2], 0.f), 3.0e38f);
    float h = min_edge;
    if (mode == 1) {  // ~4 points per cell in volume terms, never finer than the surface / line density suggests: the
      // third neighbour must lie within one cell edge for a lane to be certified without the whole-cloud scan
      const float vol = ex * ey * ez, area = fmaxf(ex * ey, fmaxf(ex * ez, ey * ez)), len = fmaxf(ex, fmaxf(ey, ez));
      h = fmaxf(fmaxf(1.6f * cbrtf(vol / nc), 1.1f * sqrtf(area / nc)), fmaxf(2.0f * len / nc, 1e-12f));
      h = fmaxf(h, min_edge);
    }
    if (!(h > 0.f) || !isfinite(h)) h = 1.0f;
    int nx = 1, ny = 1, nz = 1;
    for (int it = 0; it < 1024; ++it) {  // coarsen until the grid fits the histogram (1.25^1024 > FLT_MAX: bounded)
      const float fx = floorf(ex / h) + 1.f, fy = floorf(ey / h) + 1.f, fz = floorf(ez / h) + 1.f;
      if (fx * fy * fz <= (float)max_cells) { nx = (int)fx; ny = (int)fy; nz = (int)fz; break; }
      h *= 1.25f;
      if (!isfinite(h)) { h = 3.0e38f; break; }  // one cell
    }
    s_hdr.ox = lo[0]; s_hdr.oy = lo[1]; s_hdr.oz = lo[2];
    s_hdr.h = h; s_hdr.inv_h = 1.0f / h;
    s_hdr.nx = nx; s_hdr.ny = ny; s_hdr.nz = nz; s_hdr.ncells = nx * ny * nz;
    s_hdr.nonfinite = bad > 0.0f;
    *reinterpret_cast<GridHdr *>(base) = s_hdr;
    s_carry = 0;
  }
  __syncthreads();
  const GridHdr g = s_hdr;
  for (int i = tid; i < g.ncells; i += kBuildThreads) s_cnt[i] = 0;
  __syncthreads();
  for (int k = tid; k < n; k += kBuildThreads) {
    const int c = (cell_coord(__ldg(p + k * 3 + 2), g.oz, g.inv_h, g.nz) * g.ny +
                   cell_coord(__ldg(p + k * 3 + 1), g.oy, g.inv_h, g.ny)) * g.nx +
                  cell_coord(__ldg(p + k * 3 + 0), g.ox, g.inv_h, g.nx);
    atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  // exclusive scan of s_cnt[0..ncells) in chunks of kBuildThreads; s_cnt becomes the scatter cursor
  for (int b0 = 0; b0 < g.ncells; b0 += kBuildThreads) {
    const int i = b0 + tid;
    const int v = i < g.ncells ? s_cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(PC_FULL_MASK, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = lane < kBuildThreads / 32 ? s_warp[lane] : 0;
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
    if (i < g.ncells) { s_cnt[i] = excl; cell_start[i] = excl; }
    __syncthreads();
    if (tid == kBuildThreads - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) cell_start[g.ncells] = n;
  for (int k = tid; k < n; k += kBuildThreads) {
    const float x = __ldg(p + k * 3 + 0), y = __ldg(p + k * 3 + 1), z = __ldg(p + k * 3 + 2);
    const int c = (cell_coord(z, g.oz, g.inv_h, g.nz) * g.ny + cell_coord(y, g.oy, g.inv_h, g.ny)) * g.nx +
                  cell_coord(x, g.ox, g.inv_h, g.nx);
    const int pos = atomicAdd(&s_cnt[c], 1);  // order inside a cell is arbitrary; no result depends on it
    sorted[pos] = make_float4(x, y, z, __int_as_float(k));
  }
}

// ------------------------------------------------------------------------------------------------ three_nn
// (distance, index) ordered 3-slot insertion: the reference visits ascending indices with strict '<'
// (tf_interpolate.cpp:74-89), i.e. keeps the 3 smallest by (distance, index); here the visiting order is arbitrary.
__device__ __forceinline__ bool before(float d, int k, float bd, int bi) { return d < bd || (d == bd && k < bi); }
__device__ __forceinline__ void nn_insert_lex(float d, int k, float &b1, float &b2, float &b3, int &i1, int &i2, int &i3) {
  // d < inf: the reference starts from best = 1e40 in double (tf_interpolate.cpp:66) and inserts on strict '<', so an
  // infinite (or NaN) distance -- a non-finite coordinate on either side -- never enters the list, whatever its index
  if (d < __int_as_float(0x7f800000) && before(d, k, b3, i3)) {
    if (before(d, k, b1, i1)) {
      b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
    } else if (before(d, k, b2, i2)) {
      b3 = b2; i3 = i2; b2 = d; i2 = k;
    } else {
      b3 = d; i3 = k;
    }
  }
}

// ------------------------------------------------------------------------------------------------ tile kernels
// Thread-per-query traversal of the grid (kernels above) is latency- and divergence-bound: every lane chases its own
// cell runs.  The tile kernels sort the QUERIES by cell as well (grid_build_kernel mode 2), so a warp owns 32
// spatial neighbours, whose cell neighbourhoods overlap almost completely.  The warp stages the UNION of those
// neighbourhoods once (coalesced cell-run copies into an SoA shared-memory stage) and then runs the branch-free packed
// all-pairs filter of ball_query.cu / interpolate.cu on that local candidate set only -- a few hundred candidates per
// warp instead of the whole scene.
constexpr int kMaxRows = 256;  // (z, y) cell rows of the union box handled per pass
constexpr int kCap = 512;      // staged candidates per batch
constexpr int kTileWarpInts = 3 * kMaxRows + 4 * kCap;  // per-warp shared memory: rows (start, len, offset) + stage

struct TileSmem {
  int *rs, *rl, *ro;
  float *x, *y, *z;
  int *i;
  __device__ explicit TileSmem(int *base)
      : rs(base), rl(base + kMaxRows), ro(base + 2 * kMaxRows), x(reinterpret_cast<float *>(base + 3 * kMaxRows)),
        y(x + kCap), z(y + kCap), i(base + 3 * kMaxRows + 3 * kCap) {}
};

// Stages the candidates of the cell box [X0..X1] x [Y0..Y1] x [Z0..Z1] batch by batch and calls process(count, real)
// with `count` (a multiple of 32; the first `real` are candidates, the rest +inf padding) candidates in the stage.  Collective over NW warps that share
// `st` (NW == 1: one warp, __syncwarp; NW > 1: the whole CTA of NW warps, __syncthreads).
template <int NW, class F>
__device__ __forceinline__ void for_each_batch(const GridHdr &g, const int *__restrict__ cell_start,
                                               const float4 *__restrict__ sorted, int X0, int X1, int Y0, int Y1, int Z0,
                                               int Z1, TileSmem &st, F &&process) {
  const int lane = threadIdx.x & 31, t = (NW == 1) ? lane : (int)threadIdx.x;
  constexpr int NT = NW * 32;
  auto sync = [] { if (NW == 1) __syncwarp(); else __syncthreads(); };
  const float inf = __int_as_float(0x7f800000);
  const int nyu = Y1 - Y0 + 1, nrows = nyu * (Z1 - Z0 + 1);
  for (int r0 = 0; r0 < nrows; r0 += kMaxRows) {
    const int nr = min(kMaxRows, nrows - r0);
    for (int r = t; r < kMaxRows; r += NT) {
      int s = 0, len = 0;
      if (r < nr) {
        const int rr = r0 + r, z = Z0 + rr / nyu, y = Y0 + rr % nyu, row = (z * g.ny + y) * g.nx;
        s = __ldg(cell_start + row + X0);  // cells X0..X1 of a row are contiguous in the sorted array
        len = __ldg(cell_start + row + X1 + 1) - s;
      }
      st.rs[r] = s;
      st.rl[r] = len;
    }
    sync();
    // exclusive prefix of the row lengths (every warp computes it redundantly; warp 0 stores): lane owns 8 rows
    int mine = 0;
#pragma unroll
    for (int u = 0; u < kMaxRows / 32; ++u) mine += st.rl[lane * (kMaxRows / 32) + u];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(PC_FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(PC_FULL_MASK, incl, 31);
    if (NW == 1 || threadIdx.x < 32) {
      int run = incl - mine;
#pragma unroll
      for (int u = 0; u < kMaxRows / 32; ++u) {
        st.ro[lane * (kMaxRows / 32) + u] = run;
        run += st.rl[lane * (kMaxRows / 32) + u];
      }
    }
    sync();
    for (int b0 = 0; b0 < total; b0 += kCap) {
      const int bn = min(kCap, total - b0);
      // flat copy: candidate f of the concatenated runs lives in row r = last index with ro[r] <= f (binary search,
      // zero-length rows sort before the row that holds f); four independent loads per thread are issued together
      for (int f0 = b0 + t; f0 < b0 + bn; f0 += NT * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int f = f0 + u * NT;
          v[u] = make_float4(inf, inf, inf, 0.f);
          if (f < b0 + bn) {
            int lo = 0, hi = nr;  // invariant: ro[lo] <= f, answer in [lo, hi)
#pragma unroll
            for (int step = 0; step < 8; ++step) {  // kMaxRows = 256 = 2^8
              const int mid = (lo + hi) >> 1;
              if (mid > lo && st.ro[mid] <= f) lo = mid; else hi = max(mid, lo + 1);
            }
            v[u] = __ldg(sorted + st.rs[lo] + (f - st.ro[lo]));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int f = f0 + u * NT;
          if (f < b0 + bn) {
            st.x[f - b0] = v[u].x; st.y[f - b0] = v[u].y; st.z[f - b0] = v[u].z; st.i[f - b0] = __float_as_int(v[u].w);
          }
        }
      }
      const int bp = (bn + 31) & ~31;
      for (int f = bn + t; f < bp; f += NT) { st.x[f] = inf; st.y[f] = inf; st.z[f] = inf; st.i[f] = 0; }
      sync();
      process(bp, bn);
      sync();
    }
  }
}

// 32 hit bits (bit e <-> staged candidate w0 + e) of `d(query, candidate) - thr < 0` for one word of the stage.
__device__ __forceinline__ unsigned filter_word(const TileSmem &st, int w0, f32x2 qx2, f32x2 qy2, f32x2 qz2, f32x2 one2,
                                                int thr, bool query_minus_candidate) {
  unsigned word = 0;
#pragma unroll
  for (int gq = 0; gq < 8; ++gq) {
    const int k = w0 + gq * 4;
    const float4 xs = *reinterpret_cast<const float4 *>(st.x + k);
    const float4 ys = *reinterpret_cast<const float4 *>(st.y + k);
    const float4 zs = *reinterpret_cast<const float4 *>(st.z + k);
    const f32x2 xa = pack2(xs.x, xs.y), ya = pack2(ys.x, ys.y), za = pack2(zs.x, zs.y);
    const f32x2 xb = pack2(xs.z, xs.w), yb = pack2(ys.z, ys.w), zb = pack2(zs.z, zs.w);
    float d0, d1, d2, d3;
    if (query_minus_candidate) {
      unpack2(sqdist3_x2(qx2, qy2, qz2, xa, ya, za, one2), d0, d1);
      unpack2(sqdist3_x2(qx2, qy2, qz2, xb, yb, zb, one2), d2, d3);
    } else {
      unpack2(sqdist3_x2(xa, ya, za, qx2, qy2, qz2, one2), d0, d1);
      unpack2(sqdist3_x2(xb, yb, zb, qx2, qy2, qz2, one2), d2, d3);
    }
    word = __funnelshift_l(__float_as_int(d0) - thr, word, 1);
    word = __funnelshift_l(__float_as_int(d1) - thr, word, 1);
    word = __funnelshift_l(__float_as_int(d2) - thr, word, 1);
    word = __funnelshift_l(__float_as_int(d3) - thr, word, 1);
  }
  return __brev(word);
}

// Ball query: a tile = 32 consecutive cell-sorted queries (lane = query).  The kBallWarps warps of a CTA share the
// tile: they stage the union neighbourhood together and split its 32-candidate words among themselves, recording hits
// with shared-memory atomicOr into ONE per-query bitmap over original indices (a 1-warp CTA would pin 45 KB of shared
// memory per resident warp); warp 0 then extracts the rows.
constexpr int kBallWarps = 4;
__global__ void __launch_bounds__(kBallWarps * 32)
ball_query_tile_kernel(int n, int m, float s_star, float radius, float reach, int nsample, float one,
                       const int *__restrict__ ws_c, const int *__restrict__ ws_q, int *__restrict__ idx,
                       int *__restrict__ pts_cnt) {
  extern __shared__ __align__(16) int s_tile[];  // TileSmem | bitmap (nwords + nsumm) * 32
  TileSmem st(s_tile);
  const int nwords = (n + 31) >> 5, nsumm = (nwords + 31) >> 5;
  unsigned *s_bits = reinterpret_cast<unsigned *>(s_tile + kTileWarpInts);
  unsigned *s_summ = s_bits + (size_t)nwords * 32;
  const int scene = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int *cbase = ws_c + (size_t)scene * grid_scene_ints(n);
  const GridHdr g = *reinterpret_cast<const GridHdr *>(cbase);
  const int *cell_start = cbase + kHdrInts;
  const float4 *sorted = reinterpret_cast<const float4 *>(cbase + grid_sorted_offset_ints());
  const float4 *qsorted =
      reinterpret_cast<const float4 *>(ws_q + (size_t)scene * grid_scene_ints(m) + grid_sorted_offset_ints());
  const f32x2 one2 = pack2(one, one);
  const int thr = (s_star >= 0.0f) ? __float_as_int(s_star) + 1 : 0;
  for (int i = threadIdx.x; i < (nwords + nsumm) * 32; i += kBallWarps * 32) s_bits[i] = 0;
  __syncthreads();
  const int ntiles = (m + 31) >> 5;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int qs = tile * 32 + lane;
    const bool live = qs < m;
    const float4 qv = __ldg(qsorted + (live ? qs : tile * 32));
    const int qi = __float_as_int(qv.w);
    const f32x2 qx2 = pack2(qv.x, qv.x), qy2 = pack2(qv.y, qv.y), qz2 = pack2(qv.z, qv.z);
    // cell box of this lane's ball (dead lanes copy the tile's first query) and the tile's union box
    int X0 = __reduce_min_sync(PC_FULL_MASK, cell_coord(qv.x - reach, g.ox, g.inv_h, g.nx));
    int X1 = __reduce_max_sync(PC_FULL_MASK, cell_coord(qv.x + reach, g.ox, g.inv_h, g.nx));
    int Y0 = __reduce_min_sync(PC_FULL_MASK, cell_coord(qv.y - reach, g.oy, g.inv_h, g.ny));
    int Y1 = __reduce_max_sync(PC_FULL_MASK, cell_coord(qv.y + reach, g.oy, g.inv_h, g.ny));
    int Z0 = __reduce_min_sync(PC_FULL_MASK, cell_coord(qv.z - reach, g.oz, g.inv_h, g.nz));
    int Z1 = __reduce_max_sync(PC_FULL_MASK, cell_coord(qv.z + reach, g.oz, g.inv_h, g.nz));
    // Non-finite coordinates (never in real clouds) on either side: a NaN distance is a HIT in the reference
    // (max = fmaxf, tf_grouping_g.cu:24), wherever the point was binned -- the tile then meets EVERY cell and uses the
    // literal hit expression instead of the bit-pattern compare.
    const float inf = __int_as_float(0x7f800000);
    const bool exact = g.nonfinite ||
                       __any_sync(PC_FULL_MASK, !(fabsf(qv.x) < inf) || !(fabsf(qv.y) < inf) || !(fabsf(qv.z) < inf));
    if (exact) { X0 = Y0 = Z0 = 0; X1 = g.nx - 1; Y1 = g.ny - 1; Z1 = g.nz - 1; }
    for_each_batch<kBallWarps>(g, cell_start, sorted, X0, X1, Y0, Y1, Z0, Z1, st, [&](int count, int real) {
      for (int w0 = warp * 32; w0 < count; w0 += kBallWarps * 32) {  // the CTA's warps take alternate words
        unsigned word = 0;
        if (exact) {
          for (int e = 0; e < 32 && w0 + e < real; ++e) {  // the +inf padding would meet an infinite query as NaN
            const float d2 = sqdist3(qv.x, qv.y, qv.z, st.x[w0 + e], st.y[w0 + e], st.z[w0 + e]);
            if (fmaxf(sqrtf(d2), 1e-20f) < radius) word |= 1u << e;
          }
        } else {
          word = filter_word(st, w0, qx2, qy2, qz2, one2, thr, true);
        }
        while (word) {  // record each hit at its ORIGINAL index (other warps may touch the same bitmap word)
          const int k = st.i[w0 + __ffs(word) - 1], w = k >> 5;
          word &= word - 1;
          atomicOr(&s_bits[w * 32 + lane], 1u << (k & 31));
          atomicOr(&s_summ[(w >> 5) * 32 + lane], 1u << (w & 31));
        }
      }
    });
    // (for_each_batch ends with a CTA barrier: every warp's hits are in the bitmap)
    if (warp == 0) {
      // ascending extraction; every touched word is cleared on the way so the bitmap is clean for the next tile
      int *row_out = idx + ((size_t)scene * m + (live ? qi : 0)) * nsample;
      int cnt = 0, first = 0;
      for (int sidx = 0; sidx < nsumm; ++sidx) {
        unsigned sw = s_summ[sidx * 32 + lane];
        if (sw) s_summ[sidx * 32 + lane] = 0;
        while (sw) {
          const int w = sidx * 32 + __ffs(sw) - 1;
          sw &= sw - 1;
          unsigned word = s_bits[w * 32 + lane];
          s_bits[w * 32 + lane] = 0;
          while (word && cnt < nsample) {
            const int k = w * 32 + __ffs(word) - 1;
            word &= word - 1;
            if (cnt == 0) first = k;
            if (live) row_out[cnt] = k;
            ++cnt;
          }
        }
      }
      if (live) {
        for (int l = cnt; l < nsample; ++l) row_out[l] = first;  // tf_grouping_g.cu:26-29; empty ball -> zero row
        pts_cnt[(size_t)scene * m + qi] = cnt;
      }
    }
    __syncthreads();  // bitmap clean before the next tile's hits
  }
}

// three_nn over tiles of 32 cell-sorted dense points; kWarps independent warps per CTA.
constexpr int kNNTileWarps = 4;
__global__ void __launch_bounds__(kNNTileWarps * 32)
three_nn_tile_kernel(int n, int m, float one, const int *__restrict__ ws_c, const int *__restrict__ ws_q,
                     float *__restrict__ dist, int *__restrict__ idx) {
  extern __shared__ __align__(16) int s_tile[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TileSmem st(s_tile + warp * kTileWarpInts);
  const int scene = blockIdx.y;
  const int *cbase = ws_c + (size_t)scene * grid_scene_ints(m);
  const GridHdr g = *reinterpret_cast<const GridHdr *>(cbase);
  const int *cell_start = cbase + kHdrInts;
  const float4 *sorted = reinterpret_cast<const float4 *>(cbase + grid_sorted_offset_ints());
  const float4 *qsorted =
      reinterpret_cast<const float4 *>(ws_q + (size_t)scene * grid_scene_ints(n) + grid_sorted_offset_ints());
  const f32x2 one2 = pack2(one, one);
  const float inf = __int_as_float(0x7f800000);
  const int ntiles = (n + 31) >> 5;
  for (int tile = blockIdx.x * kNNTileWarps + warp; tile < ntiles; tile += gridDim.x * kNNTileWarps) {
    const int qs = tile * 32 + lane;
    const bool live = qs < n;
    const float4 qv = __ldg(qsorted + (live ? qs : tile * 32));
    const int qi = __float_as_int(qv.w);
    const f32x2 qx2 = pack2(qv.x, qv.x), qy2 = pack2(qv.y, qv.y), qz2 = pack2(qv.z, qv.z);
    const int cx = cell_coord(qv.x, g.ox, g.inv_h, g.nx), cy = cell_coord(qv.y, g.oy, g.inv_h, g.ny),
              cz = cell_coord(qv.z, g.oz, g.inv_h, g.nz);
    // union of the lanes' ring-1 neighbourhoods (a superset for every lane: more candidates never hurt)
    const int X0 = max(__reduce_min_sync(PC_FULL_MASK, cx) - 1, 0), X1 = min(__reduce_max_sync(PC_FULL_MASK, cx) + 1, g.nx - 1);
    const int Y0 = max(__reduce_min_sync(PC_FULL_MASK, cy) - 1, 0), Y1 = min(__reduce_max_sync(PC_FULL_MASK, cy) + 1, g.ny - 1);
    const int Z0 = max(__reduce_min_sync(PC_FULL_MASK, cz) - 1, 0), Z1 = min(__reduce_max_sync(PC_FULL_MASK, cz) + 1, g.nz - 1);
    float b1 = inf, b2 = inf, b3 = inf;
    int i1 = INT_MAX, i2 = INT_MAX, i3 = INT_MAX;
    bool seeded = false;
    for_each_batch<1>(g, cell_start, sorted, X0, X1, Y0, Y1, Z0, Z1, st, [&](int count, int real) {
      // Seeds: with b3 = inf the first 32 candidates all pass the filter and cost 32 exact insertions per lane -- 58 % of
      // the replays of a typical FP4 tile (55 per 135 candidates).  ANY three candidates bound the third-best distance
      // from above, so three neighbours from the MIDDLE of the staged union (the cells the tile's queries sit in) are
      // inserted first; the filter then only passes what lies inside that radius.  Their bits are masked out of their
      // word below, so every candidate is still inserted at most once and the result is unchanged.
      int seed_w0 = -1;
      unsigned seed_mask = 0;
      if (!seeded && real >= 3) {
        int e0 = min(real / 2, real - 3);
        e0 = min(e0, (e0 & ~31) + 29);
#pragma unroll
        for (int t = 0; t < 3; ++t)
          nn_insert_lex(sqdist3(st.x[e0 + t], st.y[e0 + t], st.z[e0 + t], qv.x, qv.y, qv.z), st.i[e0 + t], b1, b2, b3, i1, i2, i3);
        seed_w0 = e0 & ~31;
        seed_mask = 7u << (e0 & 31);
        seeded = true;
      }
      for (int w0 = 0; w0 < count; w0 += 32) {
        // stale filter d <= b3 (ties on the distance may still win on the index); exact replay of the survivors
        const int thr = (b3 == inf) ? 0x7f800000 : __float_as_int(b3) + 1;
        unsigned word = filter_word(st, w0, qx2, qy2, qz2, one2, thr, false);
        if (w0 == seed_w0) word &= ~seed_mask;
        while (word) {
          const int e = w0 + __ffs(word) - 1;
          word &= word - 1;
          nn_insert_lex(sqdist3(st.x[e], st.y[e], st.z[e], qv.x, qv.y, qv.z), st.i[e], b1, b2, b3, i1, i2, i3);
        }
      }
    });
    // certified when the third-best lies strictly inside the lane's own ring-1 box (see three_nn_grid_kernel)
    const float bound = 0.999f * g.h;
    unsigned todo = __ballot_sync(PC_FULL_MASK, live && !(b3 < bound * bound));
    while (todo) {  // rare: the whole warp scans the cloud for that lane, then merges its 32 partial lists
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const float fx = __shfl_sync(PC_FULL_MASK, qv.x, src), fy = __shfl_sync(PC_FULL_MASK, qv.y, src),
                  fz = __shfl_sync(PC_FULL_MASK, qv.z, src);
      float c1 = inf, c2 = inf, c3 = inf;
      int j1 = INT_MAX, j2 = INT_MAX, j3 = INT_MAX;
      for (int p = lane; p < m; p += 32) {
        const float4 v = __ldg(sorted + p);
        nn_insert_lex(sqdist3(v.x, v.y, v.z, fx, fy, fz), __float_as_int(v.w), c1, c2, c3, j1, j2, j3);
      }
      float r1 = inf, r2 = inf, r3 = inf;
      int k1 = INT_MAX, k2 = INT_MAX, k3 = INT_MAX;
#pragma unroll
      for (int round = 0; round < 3; ++round) {  // pop the warp-wide (distance, index) minimum three times
        float bd = c1;
        int bi = j1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(PC_FULL_MASK, bd, o);
          const int oi = __shfl_xor_sync(PC_FULL_MASK, bi, o);
          if (before(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        if (round == 0) { r1 = bd; k1 = bi; } else if (round == 1) { r2 = bd; k2 = bi; } else { r3 = bd; k3 = bi; }
        if (c1 == bd && j1 == bi) { c1 = c2; j1 = j2; c2 = c3; j2 = j3; c3 = inf; j3 = INT_MAX; }  // owner pops
      }
      if (lane == src) { b1 = r1; b2 = r2; b3 = r3; i1 = k1; i2 = k2; i3 = k3; }
    }
    if (live) {
      float *dp = dist + ((size_t)scene * n + qi) * 3;
      int *ip = idx + ((size_t)scene * n + qi) * 3;
      dp[0] = b1; dp[1] = b2; dp[2] = b3;
      ip[0] = (i1 == INT_MAX) ? 0 : i1; ip[1] = (i2 == INT_MAX) ? 0 : i2; ip[2] = (i3 == INT_MAX) ? 0 : i3;
    }
  }
}

int build_pair(int b, int nc, int nq, float min_edge, int mode, const float *xyz_c, const float *xyz_q, int *ws_c,
               int *ws_q, cudaStream_t st) {
  // The histogram (and with it the finest grid) is sized by the candidate count: two cells per candidate is finer than
  // either op wants (the build coarsens the cell edge until the grid fits -- results do not depend on the edge), and a
  // build CTA that reserves 8 KB instead of 64 KB finds room on an SM whose shared memory an FPS CTA half fills.
  const bool lone = concurrency_hint() == 1;
  int max_cells = lone ? kMaxCells : 2 * nc;
  max_cells = max_cells < 1024 ? 1024 : max_cells > kMaxCells ? kMaxCells : max_cells;
  const size_t smem = (size_t)max_cells * sizeof(int);
  if (lone) {
    PC_CUDA_TRY(allow_smem(grid_build_kernel<1024>, (size_t)kMaxCells * sizeof(int)));
    grid_build_kernel<1024><<<dim3(b, 2), 1024, smem, st>>>(nc, nq, min_edge, mode, max_cells, xyz_c, xyz_q, ws_c, ws_q);
  } else {
    PC_CUDA_TRY(allow_smem(grid_build_kernel<256>, (size_t)kMaxCells * sizeof(int)));
    grid_build_kernel<256><<<dim3(b, 2), 256, smem, st>>>(nc, nq, min_edge, mode, max_cells, xyz_c, xyz_q, ws_c, ws_q);
  }
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace

}  // namespace pc

extern "C" size_t pc_query_ball_grid_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 0;
  return (size_t)b * (pc::grid_scene_ints(n) + pc::grid_scene_ints(m)) * sizeof(int);
}

extern "C" int pc_query_ball_grid(int b, int n, int m, float radius, int nsample, const float *xyz1,
                                  const float *xyz2, int *idx, int *pts_cnt, void *workspace, pc_stream_t stream) {
  if (!(radius > 0.0f) || nsample <= 0) return PC_ERR_INVALID_ARGUMENT;  // tf_grouping.cpp:70-74
  if (b < 0 || n < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0) return PC_OK;
  if (!xyz2 || !idx || !pts_cnt || (n > 0 && !xyz1)) return PC_ERR_INVALID_ARGUMENT;
  // shapes the bitmap does not cover (or nothing to bin): the all-pairs kernel gives the same outputs
  const int nwords = (n + 31) / 32, nsumm = (nwords + 31) / 32;
  const size_t smem = ((size_t)pc::kTileWarpInts + (size_t)(nwords + nsumm) * 32) * sizeof(int);
  if (n == 0 || smem > 96 * 1024 || b > 65535 || !isfinite(radius))
    return pc_query_ball(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, stream);
  if (!workspace) return PC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const float s_star = pc::ball_threshold(radius);
  const float reach = radius * 1.0001f + 1e-30f;  // conservative: fp32 rounding of the distance is ~1e-7 relative
  int *ws_c = (int *)workspace, *ws_q = ws_c + (size_t)b * pc::grid_scene_ints(n);
  int rc = pc::build_pair(b, n, m, reach, 0, xyz1, xyz2, ws_c, ws_q, st);  // candidates + queries, one launch
  if (rc) return rc;
  PC_CUDA_TRY(pc::allow_smem(pc::ball_query_tile_kernel, smem));
  const int ntiles = (m + 31) / 32;
  dim3 grid(ntiles < 64 ? ntiles : 64, b);
  pc::ball_query_tile_kernel<<<grid, pc::kBallWarps * 32, smem, st>>>(n, m, s_star, radius, reach, nsample, 1.0f, ws_c, ws_q, idx, pts_cnt);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" size_t pc_three_nn_grid_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || m <= 0 || n <= 0) return 0;
  return (size_t)b * (pc::grid_scene_ints(m) + pc::grid_scene_ints(n)) * sizeof(int);
}

extern "C" int pc_three_nn_grid(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx,
                                void *workspace, pc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return PC_OK;
  if (!xyz1 || !dist || !idx || (m > 0 && !xyz2)) return PC_ERR_INVALID_ARGUMENT;
  if (m < 64 || b > 65535) return pc_three_nn(b, n, m, xyz1, xyz2, dist, idx, stream);  // nothing to gain from binning
  if (!workspace) return PC_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int *ws_c = (int *)workspace, *ws_q = ws_c + (size_t)b * pc::grid_scene_ints(m);
  int rc = pc::build_pair(b, m, n, 0.0f, 1, xyz2, xyz1, ws_c, ws_q, st);  // known cloud + dense points, one launch
  if (rc) return rc;
  const size_t smem = (size_t)pc::kNNTileWarps * pc::kTileWarpInts * sizeof(int);
  PC_CUDA_TRY(pc::allow_smem(pc::three_nn_tile_kernel, smem));
  const int ntiles = (n + 31) / 32, ctas = (ntiles + pc::kNNTileWarps - 1) / pc::kNNTileWarps;
  dim3 grid(ctas < 256 ? ctas : 256, b);
  pc::three_nn_tile_kernel<<<grid, pc::kNNTileWarps * 32, smem, st>>>(n, m, 1.0f, ws_c, ws_q, dist, idx);
  PC_RETURN_LAUNCH_STATUS();
}
