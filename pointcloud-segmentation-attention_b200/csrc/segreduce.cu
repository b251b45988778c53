// Deterministic backward passes: group_point_grad, gather_point_grad, three_interpolate_grad.
//
// The reference scatters with float atomicAdd (tf_ops/grouping/tf_grouping_g.cu:61-78, tf_ops/sampling/
// tf_sampling_g.cu:183-192: order undefined, run-to-run different bits) or with a serial host loop
// (tf_ops/interpolation_3d/tf_interpolate.cpp:131-153).  Here every op is a SEGMENTED REDUCTION:
//   1. csr_build: per scene, a stable counting sort of the index tensor's flat positions by the point they refer to
//      -> row_ptr (nkeys+1) and list (npos) with each point's contributors in ascending position;
//   2. csr_reduce: one thread per (point, 4 channels) adds its contributors in that fixed order with 128-bit loads.
// The sum order equals the reference's serial CPU loops (ascending (j,k) / (j,t)), so results are bit-identical to
// them, and identical from run to run.  No atomics on floats anywhere; outputs are fully overwritten (no memset).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace pc {
namespace {

// Lanes of `act` holding the same key, as a mask (what __match_any_sync returns).  MATCH.ANY measured at ~135 cycles per
// warp instruction and serialised per SM on B200 (it was 75 % of csr_build's run time); one ballot per key bit costs a
// few issue slots and runs on all four schedulers.  nbits = bits needed for nkeys-1.  Lanes outside `act` get garbage.
__device__ __forceinline__ unsigned same_key_lanes(unsigned act, int key, int nbits) {
  unsigned peers = act;
  for (int b = 0; b < nbits; ++b) {
    const bool bit = (key >> b) & 1;
    const unsigned bal = __ballot_sync(PC_FULL_MASK, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

constexpr int kCsrThreads = 1024;
constexpr size_t kCsrSmemBudget = 200 * 1024;

// A cluster of CL CTAs per scene (CL = 1: a plain CTA).  CTA r of the cluster owns the contiguous position slice
// [r * slice, (r + 1) * slice); inside it the slice is split into W contiguous parts, one per warp, with counters
// cnt[w][key] in shared memory.  The per-key totals of the CTAs meet through distributed shared memory (one
// cluster.sync), so every CTA knows how many positions with its key lie in earlier slices, and the fill -- warp w walks
// part w in order, 32 positions per step, ranking equal keys inside a step by ballot -- writes each position to its
// final, ascending place: a stable counting sort with CL * W independent walkers and no sort pass.
// (One CTA per scene left 132 of 148 SMs idle and took 38 us at the FP4 shape; measured after: see DESIGN.md.)
template <int CL>
__global__ void __launch_bounds__(kCsrThreads, 1)
csr_build_kernel(int nkeys, int npos, int slice, int W, int part, int nbits, const int *__restrict__ idx,
                 int *__restrict__ ws) {
  extern __shared__ int cnt[];  // W * nkeys counters | nkeys totals of this CTA
  __shared__ int s_carry;
  __shared__ int s_warp[32];
  int *s_tot = cnt + (size_t)W * nkeys;
  const int scene = blockIdx.x / CL;
  const int rank = (CL > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *keys = idx + (size_t)scene * npos;
  int *row_ptr = ws + (size_t)scene * (nkeys + 1 + npos);
  int *list = row_ptr + nkeys + 1;
  const int p_lo = min(npos, rank * slice), p_hi = min(npos, p_lo + slice);   // this CTA's positions

  for (int i = tid; i < W * nkeys; i += kCsrThreads) cnt[i] = 0;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  // histogram per part (integer atomics: the final counts do not depend on order)
  for (int p = p_lo + tid; p < p_hi; p += kCsrThreads) atomicAdd(&cnt[((p - p_lo) / part) * nkeys + keys[p]], 1);
  __syncthreads();
  // cnt[w][key] <- number of positions with this key in parts < w of this slice; s_tot[key] <- the slice's total
  for (int key = tid; key < nkeys; key += kCsrThreads) {
    int total = 0;
    for (int w = 0; w < W; ++w) {
      const int t = cnt[w * nkeys + key];
      cnt[w * nkeys + key] = total;
      total += t;
    }
    s_tot[key] = total;
  }
  if (CL > 1) cg::this_cluster().sync(); else __syncthreads();
  // row_ptr: exclusive scan over the keys of the cluster-wide totals; `before` = this key's positions in earlier slices
  for (int base = 0; base < nkeys; base += kCsrThreads) {
    const int key = base + tid;
    int total = 0, before = 0;
    if (key < nkeys) {
      if (CL > 1) {
#pragma unroll
        for (int c = 0; c < CL; ++c) {
          const int t = *cg::this_cluster().map_shared_rank(&s_tot[key], c);
          if (c < rank) before += t;
          total += t;
        }
      } else {
        total = s_tot[key];
      }
    }
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(PC_FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane], iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(PC_FULL_MASK, iv, o);
        if (lane >= o) iv += u;
      }
      s_warp[lane] = iv - v;  // exclusive prefix of warp sums
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - total;
    if (key < nkeys) {
      if (rank == 0) row_ptr[key] = excl;
      for (int w = 0; w < W; ++w) cnt[w * nkeys + key] += excl + before;
    }
    __syncthreads();
    if (tid == kCsrThreads - 1) s_carry = excl + total;
    __syncthreads();
  }
  if (rank == 0 && tid == 0) row_ptr[nkeys] = npos;
  // stable fill: warp w walks part w of the slice in order, 32 positions per step
  if (warp < W) {
    int *c = cnt + warp * nkeys;
    const int lo = min(p_hi, p_lo + warp * part), hi = min(p_hi, lo + part);
    const unsigned lt = lanemask_lt();
    // keys of kPre steps are fetched ahead of the serial counter walk (the walk itself is a dependent
    // shared-memory chain; a global load per step in front of it made every step an L2 / DRAM round trip)
    constexpr int kPre = 8;
    for (int p0 = lo; p0 < hi; p0 += 32 * kPre) {
      int kreg[kPre];
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int p = p0 + 32 * u + lane;
        kreg[u] = p < hi ? __ldg(keys + p) : -1;
      }
      int dst[kPre];
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int p = p0 + 32 * u + lane;
        const bool valid = p < hi;
        const unsigned act = __ballot_sync(PC_FULL_MASK, valid);
        dst[u] = -1;
        if (act == 0) continue;
        const unsigned peers = same_key_lanes(act, kreg[u], nbits);
        // full-mask shuffle executed by every lane: a *_sync with a run-time lane mask makes nvcc emit its own MATCH.ANY
        // convergence check (the very instruction same_key_lanes avoids)
        const int leader = valid ? __ffs(peers) - 1 : lane;
        int slot = 0;
        if (valid && lane == leader) {
          const int key = kreg[u];
          slot = c[key];
          c[key] = slot + __popc(peers);
        }
        slot = __shfl_sync(PC_FULL_MASK, slot, leader);
        if (valid) dst[u] = slot + __popc(peers & lt);
        __syncwarp();  // orders this step's counter stores before the next step's loads
      }
      // the scattered global stores of the batch go out together, AFTER its warp barriers: a barrier with stores in
      // flight waits for them (measured ~1.5 us per step when each step stored before its barrier)
#pragma unroll
      for (int u = 0; u < kPre; ++u)
        if (dst[u] >= 0) list[dst[u]] = p0 + 32 * u + lane;
    }
  }
  if (CL > 1) cg::this_cluster().sync();  // no CTA may exit while a peer can still read its totals
}

template <int CL>
int launch_csr_build(int b, int nkeys, int npos, int slice, int W, int part, int nbits, size_t smem, const int *idx,
                     int *workspace, cudaStream_t st) {
  auto kernel = csr_build_kernel<CL>;
  if (smem > 48 * 1024) PC_CUDA_TRY(allow_smem(kernel, smem));
  if (CL == 1) {
    kernel<<<b, kCsrThreads, smem, st>>>(nkeys, npos, slice, W, part, nbits, idx, workspace);
    PC_RETURN_LAUNCH_STATUS();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)b * CL);
  cfg.blockDim = dim3(kCsrThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, nkeys, npos, slice, W, part, nbits, idx, workspace));
  return PC_OK;
}

// Fallback for key ranges too large for shared memory: counters in the row_ptr area itself, one warp walks all
// positions in order.  Slow, correct, only reached for nkeys > ~50k.
__global__ void __launch_bounds__(1024, 1)
csr_build_big_kernel(int nkeys, int npos, const int *__restrict__ idx, int *__restrict__ ws) {
  __shared__ int s_carry;
  __shared__ int s_warp[32];
  const int scene = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *keys = idx + (size_t)scene * npos;
  int *row_ptr = ws + (size_t)scene * (nkeys + 1 + npos);
  int *list = row_ptr + nkeys + 1;
  for (int i = tid; i <= nkeys; i += 1024) row_ptr[i] = 0;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int p = tid; p < npos; p += 1024) atomicAdd(&row_ptr[keys[p]], 1);
  __syncthreads();
  for (int base = 0; base < nkeys; base += 1024) {
    const int key = base + tid;
    const int total = key < nkeys ? row_ptr[key] : 0;
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(PC_FULL_MASK, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = s_warp[lane], iv = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(PC_FULL_MASK, iv, o);
        if (lane >= o) iv += u;
      }
      s_warp[lane] = iv - v;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - total;
    if (key < nkeys) row_ptr[key] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + total;
    __syncthreads();
  }
  // row_ptr now holds starts; fill advances them, then they are shifted back.
  if (warp == 0) {
    const unsigned lt = lanemask_lt();
    for (int p0 = 0; p0 < npos; p0 += 32) {
      const int p = p0 + lane;
      const bool valid = p < npos;
      const unsigned act = __ballot_sync(PC_FULL_MASK, valid);
      if (valid) {
        const int key = keys[p];
        const unsigned peers = __match_any_sync(act, key);
        const int leader = __ffs(peers) - 1;
        int slot = 0;
        if (lane == leader) {
          slot = row_ptr[key];
          row_ptr[key] = slot + __popc(peers);
        }
        slot = __shfl_sync(peers, slot, leader);
        list[slot + __popc(peers & lt)] = p;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  // row_ptr[key] now = end of key = start of key+1: shift right by one (back to front, chunked).
  for (int top = nkeys; top > 0; top -= 1024) {
    const int i = top - tid;  // handles indices (top-1023 .. top)
    int v = 0;
    if (i >= 1) v = row_ptr[i - 1];
    __syncthreads();
    if (i >= 1) row_ptr[i] = v;
    __syncthreads();
  }
  if (tid == 0) row_ptr[0] = 0;
}

// out[scene, i, 4q..4q+3] = sum_{e in row i} src[scene, list[e]/div, 4q..] * (w ? w[scene, list[e]] : 1)
template <bool WEIGHTED>
__global__ void __launch_bounds__(256)
csr_reduce_vec4_kernel(int nkeys, int npos, int c4, int div, FastDiv fc, FastDiv fdiv, bool fits32,
                       const float4 *__restrict__ src, const float *__restrict__ w, const int *__restrict__ ws,
                       float4 *__restrict__ out) {
  const int scene = blockIdx.y;
  const int *row_ptr = ws + (size_t)scene * (nkeys + 1 + npos);
  const int *list = row_ptr + nkeys + 1;
  const float4 *s = src + (size_t)scene * (npos / div) * c4;
  const float *ww = WEIGHTED ? w + (size_t)scene * npos : nullptr;
  const size_t total = (size_t)nkeys * c4;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int i = fits32 ? (int)fc.div((uint32_t)t) : (int)(t / c4), q = (int)(t - (size_t)i * c4);
    const int lo = row_ptr[i], hi = row_ptr[i + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // U contributors' row loads in flight at a time, the NEXT batch's list entries fetched under them (a list -> row
    // chain per batch otherwise doubles the latency); the additions stay strictly in list order.  Rows are uneven --
    // ball query pads with its first hit, so low-index points collect hundreds of contributions -- and the longest
    // rows set the kernel's duration: U = 16 for the unweighted (group_point) form, 8 for the weighted (interpolation)
    // form whose rows are short (~3n/m) and which needs the registers for occupancy instead.
    constexpr int U = WEIGHTED ? 8 : 16;
    int pn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) pn[u] = (lo + u < hi) ? list[lo + u] : -1;
    for (int e = lo; e < hi; e += U) {
      float4 g[U];
      float wt[U];
      bool on[U];
      // loads are unconditional (a batch's missing entries re-read the row of position 0) so that nothing ties a load
      // to its use: all U are issued before the first addition.  A masked-out term is +0, and acc + (+0) == acc bit for
      // bit (acc starts at +0 and can never become -0)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        on[u] = pn[u] >= 0;
        const int pp = on[u] ? pn[u] : 0;
        g[u] = __ldg(s + (size_t)fdiv.div((uint32_t)pp) * c4 + q);
        wt[u] = WEIGHTED ? __ldg(ww + pp) : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) pn[u] = (e + U + u < hi) ? list[e + U + u] : -1;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 t = g[u];
        if (WEIGHTED) {
          t.x = __fmul_rn(t.x, wt[u]); t.y = __fmul_rn(t.y, wt[u]); t.z = __fmul_rn(t.z, wt[u]); t.w = __fmul_rn(t.w, wt[u]);
        }
        acc.x = __fadd_rn(acc.x, on[u] ? t.x : 0.0f); acc.y = __fadd_rn(acc.y, on[u] ? t.y : 0.0f);
        acc.z = __fadd_rn(acc.z, on[u] ? t.z : 0.0f); acc.w = __fadd_rn(acc.w, on[u] ? t.w : 0.0f);
      }
    }
    out[((size_t)scene * nkeys + i) * c4 + q] = acc;
  }
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(256)
csr_reduce_scalar_kernel(int nkeys, int npos, int c, int div, FastDiv fc, FastDiv fdiv, bool fits32,
                         const float *__restrict__ src, const float *__restrict__ w, const int *__restrict__ ws,
                         float *__restrict__ out) {
  const int scene = blockIdx.y;
  const int *row_ptr = ws + (size_t)scene * (nkeys + 1 + npos);
  const int *list = row_ptr + nkeys + 1;
  const float *s = src + (size_t)scene * (npos / div) * c;
  const float *ww = WEIGHTED ? w + (size_t)scene * npos : nullptr;
  const size_t total = (size_t)nkeys * c;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int i = fits32 ? (int)fc.div((uint32_t)t) : (int)(t / c), l = (int)(t - (size_t)i * c);
    const int lo = row_ptr[i], hi = row_ptr[i + 1];
    float acc = 0.f;
    for (int e = lo; e < hi; ++e) {
      const int p = list[e];
      float g = __ldg(s + (size_t)fdiv.div((uint32_t)p) * c + l);
      if (WEIGHTED) g = __fmul_rn(g, __ldg(ww + p));
      acc = __fadd_rn(acc, g);
    }
    out[((size_t)scene * nkeys + i) * c + l] = acc;
  }
}


// Entry-stream form of the segmented reduction (the default for c % 32 == 0).  The thread-per-(key, 4 channels) kernels
// above keep `row length` loads in flight per thread and most rows are short (a ball's real hits: 1-5 entries; only a
// first-hit key collects ~30), so their memory-level parallelism follows the row-length histogram.  Here a WARP owns KW
// consecutive keys = ONE contiguous run of list entries [row_ptr[k0], row_ptr[k0+KW]) and walks it as a stream: U
// contributor rows are in flight per lane at all times, across row boundaries (the next batch's list entries are
// fetched under the current batch's row loads), and the 32 lanes span 32*V consecutive channels of the row -- one
// coalesced 128*V-byte segment per contributor.  The additions stay strictly in list order, one accumulator per key,
// flushed at the row boundary (a warp-uniform compare against the row_ptr values the lanes hold): bit-identical to
// the serial reference loops.  Empty rows fall out as zero stores of the same flush.
template <int V, bool WEIGHTED>
struct VecT;
template <bool W> struct VecT<4, W> { typedef float4 T; };
template <bool W> struct VecT<2, W> { typedef float2 T; };
template <bool W> struct VecT<1, W> { typedef float T; };

template <int V>
__device__ __forceinline__ void vload(const float *p, float (&v)[V]) {
  if constexpr (V == 4) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else if constexpr (V == 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y; }
  else { v[0] = __ldg(p); }
}
template <int V>
__device__ __forceinline__ void vstore(float *p, const float (&v)[V]) {
  if constexpr (V == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else if constexpr (V == 2) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
  else *p = v[0];
}

constexpr int kStreamWarps = 8;
template <int V, bool WEIGHTED>
__global__ void __launch_bounds__(kStreamWarps * 32, 3)
csr_reduce_stream_kernel(int nkeys, int npos, int c, int KW, FastDiv fdiv, const float *__restrict__ src,
                         const float *__restrict__ w, const int *__restrict__ ws, float *__restrict__ out) {
  constexpr int U = 8;
  const int scene = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k0 = (blockIdx.x * kStreamWarps + warp) * KW;
  if (k0 >= nkeys) return;
  const int nk = min(KW, nkeys - k0);
  const int *row_ptr = ws + (size_t)scene * (nkeys + 1 + npos);
  const int *list = row_ptr + nkeys + 1;
  const int choff = blockIdx.z * 32 * V + lane * V;
  const float *s = src + (size_t)scene * (npos / fdiv.d) * c + choff;
  const float *ww = WEIGHTED ? w + (size_t)scene * npos : nullptr;
  float *o = out + ((size_t)scene * nkeys + k0) * c + choff;
  const int rp = row_ptr[k0 + min(lane, nk)];                  // lane l holds the start of local key l (l <= nk)
  const int E0 = __shfl_sync(PC_FULL_MASK, rp, 0), E1 = __shfl_sync(PC_FULL_MASK, rp, nk);
  int kk = 0, next_end = __shfl_sync(PC_FULL_MASK, rp, 1);
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  int pn[U];
#pragma unroll
  for (int u = 0; u < U; ++u) pn[u] = (E0 + u < E1) ? __ldg(list + E0 + u) : -1;
  for (int e = E0; e < E1; e += U) {
    float g[U][V], wt[U];
    int on[U];
    // unconditional loads (missing entries re-read position 0): all U issue before the first addition
#pragma unroll
    for (int u = 0; u < U; ++u) {
      on[u] = pn[u] >= 0;
      const int pp = on[u] ? pn[u] : 0;
      vload<V>(s + (size_t)fdiv.div((uint32_t)pp) * c, g[u]);
      wt[u] = WEIGHTED ? __ldg(ww + pp) : 1.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) pn[u] = (e + U + u < E1) ? __ldg(list + e + U + u) : -1;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (on[u]) {                                              // warp-uniform
        while (e + u == next_end) {                             // row boundary (possibly several empty rows): flush
          vstore<V>(o + (size_t)kk * c, acc);
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = 0.f;
          ++kk;
          next_end = __shfl_sync(PC_FULL_MASK, rp, kk + 1);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = __fadd_rn(acc[v], WEIGHTED ? __fmul_rn(g[u][v], wt[u]) : g[u][v]);
      }
    }
  }
  for (; kk < nk; ++kk) {                                       // the last row and any trailing empty rows
    vstore<V>(o + (size_t)kk * c, acc);
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.f;
  }
}

template <int V>
int launch_stream(int b, int nkeys, int npos, int c, int div, const float *src, const float *w, const int *workspace,
                  float *out, cudaStream_t st) {
  const int cb = c / (32 * V);
  // keys per warp: enough warps to fill the chip several times over, never more than 16 keys (a lane holds one start)
  long long warps_at_1 = (long long)b * nkeys * cb;
  int KW = (int)(warps_at_1 / ((long long)num_sms() * 48));
  KW = KW < 1 ? 1 : (KW > 16 ? 16 : KW);
  dim3 grid((unsigned)((nkeys + KW * kStreamWarps - 1) / (KW * kStreamWarps)), (unsigned)b, (unsigned)cb);
  if (w) csr_reduce_stream_kernel<V, true><<<grid, kStreamWarps * 32, 0, st>>>(nkeys, npos, c, KW, FastDiv((uint32_t)div), src, w, workspace, out);
  else   csr_reduce_stream_kernel<V, false><<<grid, kStreamWarps * 32, 0, st>>>(nkeys, npos, c, KW, FastDiv((uint32_t)div), src, w, workspace, out);
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace

size_t csr_workspace_bytes(int b, int nkeys, int npos) {
  if (b <= 0 || nkeys <= 0) return 0;
  return (size_t)b * ((size_t)nkeys + 1 + (size_t)(npos > 0 ? npos : 0)) * sizeof(int);
}

int csr_build(int b, int nkeys, int npos, const int *idx, int *workspace, cudaStream_t st) {
  if (b > 65535) return PC_ERR_UNSUPPORTED;
  const size_t per_part = (size_t)nkeys * sizeof(int);
  if (2 * per_part <= kCsrSmemBudget) {   // at least one part's counters + the totals fit shared memory
    // CTAs per scene: a cluster of 8 where a scene has enough positions to pay for the cluster's two barriers and the
    // DSMEM pass (measured under ncu: 24576 positions / 1024 keys 38 -> 29.5 us; 8192 positions 15 -> 22 us: stays 1 CTA)
    const int CL = (npos >= 16384) ? 8 : 1;
    int slice = ((npos + CL - 1) / CL + 31) / 32 * 32;
    if (slice < 32) slice = 32;
    int W = (int)((kCsrSmemBudget - per_part) / per_part);
    if (W > 32) W = 32;
    const int steps = (slice + 31) / 32;  // no more parts than 32-position steps
    if (W > steps) W = steps;
    int part = ((slice + W - 1) / W + 31) / 32 * 32;
    W = (slice + part - 1) / part;
    if (W < 1) W = 1;
    const size_t smem = (size_t)(W + 1) * per_part;
    int nbits = 0;
    while ((1 << nbits) < nkeys) ++nbits;
    if (CL == 8) return launch_csr_build<8>(b, nkeys, npos, slice, W, part, nbits, smem, idx, workspace, st);
    if (CL == 4) return launch_csr_build<4>(b, nkeys, npos, slice, W, part, nbits, smem, idx, workspace, st);
    if (CL == 2) return launch_csr_build<2>(b, nkeys, npos, slice, W, part, nbits, smem, idx, workspace, st);
    return launch_csr_build<1>(b, nkeys, npos, slice, W, part, nbits, smem, idx, workspace, st);
  }
  csr_build_big_kernel<<<b, 1024, 0, st>>>(nkeys, npos, idx, workspace);
  PC_RETURN_LAUNCH_STATUS();
}

int csr_reduce(int b, int nkeys, int npos, int c, int div, const float *src, const float *w, const int *workspace,
               float *out, cudaStream_t st) {
  const int sms = num_sms();
  if (b <= 65535 && c >= 32 && c / 32 <= 65535 * 4) {  // entry-stream kernel: 32 lanes x V floats per channel block
    if (c % 128 == 0 && aligned16(src) && aligned16(out)) return launch_stream<4>(b, nkeys, npos, c, div, src, w, workspace, out, st);
    if (c % 64 == 0 && aligned16(src) && aligned16(out)) return launch_stream<2>(b, nkeys, npos, c, div, src, w, workspace, out, st);
    if (c % 32 == 0) return launch_stream<1>(b, nkeys, npos, c, div, src, w, workspace, out, st);
  }
  if (c % 4 == 0 && aligned16(src) && aligned16(out)) {
    const size_t total = (size_t)nkeys * (c / 4);
    unsigned gx = (unsigned)((total + 255) / 256);
    if (gx > (unsigned)sms * 32) gx = (unsigned)sms * 32;
    dim3 grid(gx, b);
    if (w) csr_reduce_vec4_kernel<true><<<grid, 256, 0, st>>>(nkeys, npos, c / 4, div, FastDiv((uint32_t)(c / 4)), FastDiv((uint32_t)div), total < (1ull << 31), (const float4 *)src, w, workspace, (float4 *)out);
    else   csr_reduce_vec4_kernel<false><<<grid, 256, 0, st>>>(nkeys, npos, c / 4, div, FastDiv((uint32_t)(c / 4)), FastDiv((uint32_t)div), total < (1ull << 31), (const float4 *)src, w, workspace, (float4 *)out);
  } else {
    const size_t total = (size_t)nkeys * c;
    unsigned gx = (unsigned)((total + 255) / 256);
    if (gx > (unsigned)sms * 32) gx = (unsigned)sms * 32;
    dim3 grid(gx, b);
    if (w) csr_reduce_scalar_kernel<true><<<grid, 256, 0, st>>>(nkeys, npos, c, div, FastDiv((uint32_t)c), FastDiv((uint32_t)div), total < (1ull << 31), src, w, workspace, out);
    else   csr_reduce_scalar_kernel<false><<<grid, 256, 0, st>>>(nkeys, npos, c, div, FastDiv((uint32_t)c), FastDiv((uint32_t)div), total < (1ull << 31), src, w, workspace, out);
  }
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace pc

// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t pc_group_point_grad_workspace_bytes(int b, int n, int c, int m, int nsample) {
  (void)c;
  if (m <= 0 || nsample <= 0) return 0;
  return pc::csr_workspace_bytes(b, n, m * nsample);
}

extern "C" int pc_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx,
                                   float *grad_points, void *workspace, pc_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || m < 0 || nsample < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0 || c == 0) return PC_OK;
  if (!grad_points) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0 || nsample == 0) {  // nothing scatters: the op still zero-fills (tf_grouping.cpp:204)
    PC_CUDA_TRY(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * n * c, st));
    return PC_OK;
  }
  if (!grad_out || !idx) return PC_ERR_INVALID_ARGUMENT;
  if (!workspace) return PC_ERR_WORKSPACE;
  if ((long long)m * nsample > 0x7fffffffLL) return PC_ERR_UNSUPPORTED;
  int rc = pc::csr_build(b, n, m * nsample, idx, (int *)workspace, st);
  if (rc) return rc;
  return pc::csr_reduce(b, n, m * nsample, c, 1, grad_out, nullptr, (const int *)workspace, grad_points, st);
}

extern "C" size_t pc_gather_point_grad_workspace_bytes(int b, int n, int m) {
  if (m <= 0) return 0;
  return pc::csr_workspace_bytes(b, n, m);
}

extern "C" int pc_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g,
                                    void *workspace, pc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return PC_OK;
  if (!inp_g) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (m == 0) {
    PC_CUDA_TRY(cudaMemsetAsync(inp_g, 0, sizeof(float) * (size_t)b * n * 3, st));  // tf_sampling.cpp:174
    return PC_OK;
  }
  if (!out_g || !idx) return PC_ERR_INVALID_ARGUMENT;
  if (!workspace) return PC_ERR_WORKSPACE;
  int rc = pc::csr_build(b, n, m, idx, (int *)workspace, st);
  if (rc) return rc;
  return pc::csr_reduce(b, n, m, 3, 1, out_g, nullptr, (const int *)workspace, inp_g, st);
}

extern "C" size_t pc_three_interpolate_grad_workspace_bytes(int b, int n, int c, int m) {
  (void)c;
  if (n <= 0) return 0;
  return pc::csr_workspace_bytes(b, m, n * 3);
}

extern "C" int pc_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                         const float *weight, float *grad_points, void *workspace,
                                         pc_stream_t stream) {
  if (b < 0 || n < 0 || c < 0 || m < 0) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0 || c == 0) return PC_OK;
  if (!grad_points) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    PC_CUDA_TRY(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, st));  // tf_interpolate.cpp:258
    return PC_OK;
  }
  if (!grad_out || !idx || !weight) return PC_ERR_INVALID_ARGUMENT;
  if (!workspace) return PC_ERR_WORKSPACE;
  if ((long long)n * 3 > 0x7fffffffLL) return PC_ERR_UNSUPPORTED;
  int rc = pc::csr_build(b, m, n * 3, idx, (int *)workspace, st);
  if (rc) return rc;
  return pc::csr_reduce(b, m, n * 3, c, 3, grad_out, weight, (const int *)workspace, grad_points, st);
}
