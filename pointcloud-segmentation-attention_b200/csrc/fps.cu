// Farthest point sampling for sm_100a.
//
// Replaces farthestpointsamplingKernel / Launcher (reference tf_ops/sampling/tf_sampling_g.cu:105-170,203-205).
// Semantics kept bit-for-bit: out[0] = 0; running min-distance starts at 1e38; distance is the un-fused
// (dx*dx + dy*dy) + dz*dz; the next centre is the point of maximal min-distance with ties resolved to the
// smallest (k mod 512, k) -- what the reference's 512-thread strided partition and left-biased tree produce.
//
// Design (not the reference's): one scene's whole state lives ON CHIP.  Each thread owns P points in registers
// (xyz as packed fp32x2 pairs + running min-distance = 4P registers), a copy of xyz sits in shared memory only to
// broadcast the chosen centre.  A round is: P/2 packed distance updates per thread (FADD2/FMUL2/FFMA2, two points per
// instruction, each operation still rounded on its own) -> redux.sync max over the value bits (non-negative floats
// order like ints) -> the lane at the warp maximum resolves its tie-break key, redux.sync min -> each warp publishes
// (value, key) in shared memory -> ONE barrier -> every warp re-reduces the published pairs.  One barrier per round
// instead of the reference's ten (none when the scene fits one warp), no global traffic inside the m-1 dependent
// rounds, and the sampled coordinates can be written by the same kernel (pc_fps_gather).
//
// Measured anatomy of a round at n = 8192 on B200 (scripts/ubench/fps_rounds.cu): distance updates 620-660 cycles (the
// fp32 pipe needs 512: 8192 points x 8 un-fused flops / 128 lanes), winner resolution + barrier + broadcast ~300-450.
// A scene cannot go faster on one SM, and splitting it over a cluster buys less than the DSMEM exchange costs at this
// size -- so throughput comes from running scenes, and batches, side by side.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace pc {
namespace {

constexpr int kMaxRegPoints = 8192;  // T=256 threads x P=32 points (4 registers per point)

__device__ __forceinline__ int tie_key(int k) { return ((k & 511) << 22) | (k >> 9); }
__device__ __forceinline__ int tie_key_to_index(int t) { return ((t & 0x3fffff) << 9) | (t >> 22); }

// One CTA of T threads per scene (grid-stride over scenes); a thread owns P points (P even) held as P/2 packed fp32x2
// register pairs so that one FADD2/FMUL2/FFMA2 advances two points (common.cuh: sqdist3_x2 keeps every operation
// individually rounded).  FEW, FAT threads: the fp32 pipe needs n*8/128 cycles per round whatever the shape, but the
// per-round tail (two reductions, two barriers) is paid per warp, so 8192 points run as 8 warps x 32 points per thread.
//
// Point <-> thread mapping.  The reference's tie-break order is (k mod 512, k).  Thread `tid` owns the residues
// tid, tid+T, ... (mod 512) -- R = 512/T of them -- and within a residue every 512th point:
//     local slot i = r*A + a  <->  k = tid + T*r + 512*a      (A = P/R points per residue)
// so a thread's slots are in ascending tie-break order and "first slot at the maximum" is its best candidate.  With
// that, the winner scan after the block maximum is known costs 1 compare per group of 8 slots plus 8 compares
// inside the first matching group, instead of P key computations.  (Scenes of at most 512 points: key order is k
// order and the plain mapping k = tid + T*i is the same thing.)
template <int P, int T>
struct FpsMap {
  static constexpr int R = (T >= 512) ? 1 : ((512 / T < P) ? 512 / T : P);  // residues per thread
  static constexpr int A = P / R;
  __device__ static __forceinline__ int k_of(int tid, int i) { return tid + T * (i / A) + (R * T) * (i % A); }
};

// Register cap: the state itself is 4 registers per point; capping the rest leaves register file and warp slots on an
// FPS SM for CTAs of other kernels (a scene's FPS keeps the fp32 pipe < 50 % busy), which matters when several
// batches are in flight.
template <int P, int T, bool kOneBarrier, int kRegs>
__global__ void __launch_bounds__(T, 1) __maxnreg__(kRegs)
fps_onchip_kernel(int b, int n, int m, float one, const float *__restrict__ xyz, int *__restrict__ out,
                  float *__restrict__ out_xyz) {
  extern __shared__ float s_xyz[];  // n*3
  __shared__ int2 s_pair[2][32];    // per-warp (max value bits, tie-break key), double-buffered by round parity
  __shared__ int s_tb[2];           // two-barrier variant: winning tie-break key, double-buffered by round parity
  using Map = FpsMap<P, T>;
  static_assert(P % 2 == 0 && T % 32 == 0 && (T >= 512 || P % Map::R == 0), "bad FPS shape");
  constexpr int H = P / 2;
  constexpr int G = (P < 8) ? P : 8;  // slots per group
  constexpr int NG = P / G;
  constexpr int nwarps = T / 32;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const f32x2 one2 = pack2(one, one);

  for (int scene = blockIdx.x; scene < b; scene += gridDim.x) {
    const float *p = xyz + (size_t)scene * n * 3;
    int *o = out + (size_t)scene * m;
    __syncthreads();  // previous scene fully done with s_xyz / s_pair
    for (int i = tid; i < n * 3; i += T) s_xyz[i] = p[i];
    float *oxyz = out_xyz ? out_xyz + (size_t)scene * m * 3 : nullptr;
    if (tid == 0) {
      o[0] = 0;
      s_tb[0] = INT_MAX; s_tb[1] = INT_MAX;
      if (oxyz) { oxyz[0] = p[0]; oxyz[1] = p[1]; oxyz[2] = p[2]; }
    }
    __syncthreads();

    f32x2 px[H], py[H], pz[H];
    float td[P];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float x[2], y[2], z[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 2 * h + e, k = Map::k_of(tid, i);
        if (k < n) {
          x[e] = s_xyz[k * 3 + 0]; y[e] = s_xyz[k * 3 + 1]; z[e] = s_xyz[k * 3 + 2];
          td[i] = 1e38f;
        } else {  // padding never wins: real values are >= 0
          x[e] = y[e] = z[e] = 0.0f;
          td[i] = -1.0f;
        }
      }
      px[h] = pack2(x[0], x[1]); py[h] = pack2(y[0], y[1]); pz[h] = pack2(z[0], z[1]);
    }

    int old = 0;
    int2 *wslot = &s_pair[1][warp];         // slot this warp's lane 0 writes in round j (parity j&1), j starts at 1
    const int2 *rslot = &s_pair[1][lane < nwarps ? lane : 0];
    int par = 1;
    for (int j = 1; j < m; ++j) {
      const float cx = s_xyz[old * 3 + 0], cy = s_xyz[old * 3 + 1], cz = s_xyz[old * 3 + 2];
      const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
      float gm[NG];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        gm[g] = -1.0f;
#pragma unroll
        for (int h = g * G / 2; h < (g + 1) * G / 2; ++h) {
          float d0, d1;
          unpack2(sqdist3_x2(px[h], py[h], pz[h], cx2, cy2, cz2, one2), d0, d1);
          td[2 * h] = fminf(d0, td[2 * h]);
          td[2 * h + 1] = fminf(d1, td[2 * h + 1]);
          gm[g] = fmaxf(gm[g], fmaxf(td[2 * h], td[2 * h + 1]));
        }
      }
      float vmax = gm[0];
#pragma unroll
      for (int g = 1; g < NG; ++g) vmax = fmaxf(vmax, gm[g]);
      if constexpr (!kOneBarrier && nwarps > 1) {
        // Two-barrier variant: block maximum first, then only the thread(s) at the block maximum resolve a key.
        const int vb = __float_as_int(vmax);
        const int wmax = __reduce_max_sync(PC_FULL_MASK, vb);
        if (lane == 0) wslot->x = wmax;
        __syncthreads();
        const int gmax = __reduce_max_sync(PC_FULL_MASK, rslot->x);
        if (vb == gmax) {
          int tb = INT_MAX;
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (tb == INT_MAX && __float_as_int(gm[g]) == gmax) {
#pragma unroll
              for (int e = G - 1; e >= 0; --e)
                if (__float_as_int(td[g * G + e]) == gmax) tb = tie_key(Map::k_of(tid, g * G + e));
            }
          }
          atomicMin(&s_tb[par], tb);
        }
        if (tid == 0) s_tb[par ^ 1] = INT_MAX;
        __syncthreads();
        old = tie_key_to_index(s_tb[par]);
        par ^= 1;
        wslot += par ? 32 : -32;
        rslot += par ? 32 : -32;
      } else {
      // Warp winner: value by redux, then the lane(s) at the warp maximum resolve their first slot at that value (slots
      // are in tie-break order) and a second redux picks the smallest key.  ONE barrier per round: every warp publishes
      // (value, key) in a parity-double-buffered slot and re-reduces all warps' pairs after the barrier.
      const int vb = __float_as_int(vmax);
      const int wmax = __reduce_max_sync(PC_FULL_MASK, vb);
      int tb = INT_MAX;
      if (vb == wmax) {
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if (tb == INT_MAX && __float_as_int(gm[g]) == wmax) {
#pragma unroll
            for (int e = G - 1; e >= 0; --e)
              if (__float_as_int(td[g * G + e]) == wmax) tb = tie_key(Map::k_of(tid, g * G + e));
          }
        }
      }
      const int wkey = __reduce_min_sync(PC_FULL_MASK, tb);
      if (nwarps > 1) {
        if (lane == 0) *wslot = make_int2(wmax, wkey);
        __syncthreads();
        const int2 pr = *rslot;  // lanes >= nwarps re-read slot 0: harmless for max / min
        const int gmax = __reduce_max_sync(PC_FULL_MASK, pr.x);
        old = tie_key_to_index(__reduce_min_sync(PC_FULL_MASK, pr.x == gmax ? pr.y : INT_MAX));
        par ^= 1;
        wslot += par ? 32 : -32;
        rslot += par ? 32 : -32;
      } else {  // a single warp: no shared-memory hop, no barrier
        old = tie_key_to_index(wkey);
      }
      }
      if (tid == 0) {
        o[j] = old;
        if (oxyz) { oxyz[j * 3 + 0] = s_xyz[old * 3 + 0]; oxyz[j * 3 + 1] = s_xyz[old * 3 + 1]; oxyz[j * 3 + 2] = s_xyz[old * 3 + 2]; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pruned variant for 2048 < n <= 8192: same rounds, same arithmetic, same winners -- but most warps skip most rounds.
// At kernel start the scene is counting-sorted by the Morton code of a 16^3 cell grid (shared-memory histogram), so a
// warp's points form a compact patch with a bounding box.  A new centre c cannot lower any running minimum of a warp
// whose box is farther from c than the warp's largest running minimum:
//        dist_lb(c, box)^2 * (1 - 1e-5)  >  max_p td[p]     =>     d(p, c) >= td[p] for every p of the warp
// (the margin covers the fp32 rounding of both sides by two orders of magnitude), and then min(td, d) leaves every td,
// the warp's maximum and its tie-break key unchanged: the warp republishes its cached (value, key) pair and skips the
// P/2 packed updates.  After a few dozen samples a centre only touches the 2-3 patches around it, so the fp32 work of a
// round drops by 3-4x while every output index stays bit-identical (ties use the ORIGINAL index through a slot ->
// index table in shared memory; the sort order itself never shows in the result).
// Register cap of the default shape (16 warps x 16 points): the state is 64 registers per thread; ptxas needs 116 when
// left alone and compiles without spills down to 88.  A lower cap leaves register file on an FPS-occupied SM for
// co-resident CTAs of the streaming kernels (the FPS round issues in < 50 % of its cycles).  Measured on B200, 8 batches
// in flight: cap 120 -> 61.0 k scenes/s (FPS alone 543 us), 96 -> 63.0 k (553 us), 88 -> 63.9 k (567 us): the lone
// launch gets 4 % slower, the pipeline 5 % faster.
#ifndef PCOPS_FPS_PRUNED_REGS
#define PCOPS_FPS_PRUNED_REGS 88
#endif
template <int P, int T>
__global__ void __maxnreg__(T >= 1024 ? 64 : (T >= 512 ? PCOPS_FPS_PRUNED_REGS : 200))
fps_pruned_kernel(int b, int n, int m, float one, const float *__restrict__ xyz, int *__restrict__ out,
                  float *__restrict__ out_xyz) {
  extern __shared__ float s_dyn[];                       // xyz (n*3 floats) | hist (4096 ints) | perm (T*P ushorts)
  __shared__ int2 s_pair[2][32];
  __shared__ float s_red[32];
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  constexpr int H = P / 2, G = (P < 8) ? P : 8, NG = P / G, nwarps = T / 32, kBins = 4096;
  float *s_xyz = s_dyn;
  int *s_hist = reinterpret_cast<int *>(s_dyn + (size_t)n * 3);
  unsigned short *s_perm = reinterpret_cast<unsigned short *>(s_hist + kBins);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const f32x2 one2 = pack2(one, one);
  const float inf = __int_as_float(0x7f800000);

  for (int scene = blockIdx.x; scene < b; scene += gridDim.x) {
    const float *p = xyz + (size_t)scene * n * 3;
    int *o = out + (size_t)scene * m;
    float *oxyz = out_xyz ? out_xyz + (size_t)scene * m * 3 : nullptr;
    __syncthreads();
    for (int i = tid; i < n * 3; i += T) s_xyz[i] = p[i];
    for (int i = tid; i < kBins; i += T) s_hist[i] = 0;
    if (tid == 0) {
      o[0] = 0;
      s_carry = 0;
      if (oxyz) { oxyz[0] = p[0]; oxyz[1] = p[1]; oxyz[2] = p[2]; }
    }
    __syncthreads();
    // ---- bounding box of the scene
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    for (int k = tid; k < n; k += T) {
#pragma unroll
      for (int c = 0; c < 3; ++c) { const float v = s_xyz[k * 3 + c]; lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v = q ? hi[c] : lo[c];
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) {
          const float w = __shfl_xor_sync(PC_FULL_MASK, v, of);
          v = q ? fmaxf(v, w) : fminf(v, w);
        }
        __syncthreads();
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        float r = s_red[0];
        for (int i = 1; i < nwarps; ++i) r = q ? fmaxf(r, s_red[i]) : fminf(r, s_red[i]);
        if (q) hi[c] = r; else lo[c] = r;
      }
    }
    float scale[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { const float e = hi[c] - lo[c]; scale[c] = (e > 0.f && e < inf) ? 16.0f / e : 0.0f; }
    auto morton = [&](int k) {
      int key = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int q = (int)((s_xyz[k * 3 + c] - lo[c]) * scale[c]);
        q = min(max(q, 0), 15);
        key |= ((q & 1) | ((q & 2) << 2) | ((q & 4) << 4) | ((q & 8) << 6)) << c;  // bits 0,3,6,9 (+c)
      }
      return key;
    };
    // ---- counting sort of the point indices by Morton key (order inside a cell is irrelevant to the result)
    for (int k = tid; k < n; k += T) atomicAdd(&s_hist[morton(k)], 1);
    __syncthreads();
    for (int b0 = 0; b0 < kBins; b0 += T) {
      const int v = s_hist[b0 + tid];
      int incl = v;
#pragma unroll
      for (int of = 1; of < 32; of <<= 1) {
        const int u = __shfl_up_sync(PC_FULL_MASK, incl, of);
        if (lane >= of) incl += u;
      }
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      int wsum = 0;
      for (int i = 0; i < warp; ++i) wsum += s_warp[i];
      const int excl = s_carry + wsum + incl - v;
      s_hist[b0 + tid] = excl;
      __syncthreads();
      if (tid == T - 1) s_carry = excl + v;
      __syncthreads();
    }
    for (int k = tid; k < n; k += T) s_perm[atomicAdd(&s_hist[morton(k)], 1)] = (unsigned short)k;
    for (int k = n + tid; k < T * P; k += T) s_perm[k] = 0xffff;
    __syncthreads();

    // ---- slots: warp w owns sorted positions [w*32*P, (w+1)*32*P); slot i of lane l is position w*32*P + i*32 + l
    const int pos0 = warp * 32 * P + lane;
    f32x2 px[H], py[H], pz[H];
    float td[P];
    float blo[3] = {inf, inf, inf}, bhi[3] = {-inf, -inf, -inf};
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float x[2], y[2], z[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 2 * h + e;
        const int k = s_perm[pos0 + i * 32];
        if (k != 0xffff) {
          x[e] = s_xyz[k * 3 + 0]; y[e] = s_xyz[k * 3 + 1]; z[e] = s_xyz[k * 3 + 2];
          td[i] = 1e38f;
          blo[0] = fminf(blo[0], x[e]); bhi[0] = fmaxf(bhi[0], x[e]);
          blo[1] = fminf(blo[1], y[e]); bhi[1] = fmaxf(bhi[1], y[e]);
          blo[2] = fminf(blo[2], z[e]); bhi[2] = fmaxf(bhi[2], z[e]);
        } else {
          x[e] = y[e] = z[e] = 0.0f;
          td[i] = -1.0f;
        }
      }
      px[h] = pack2(x[0], x[1]); py[h] = pack2(y[0], y[1]); pz[h] = pack2(z[0], z[1]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) {
        blo[c] = fminf(blo[c], __shfl_xor_sync(PC_FULL_MASK, blo[c], of));
        bhi[c] = fmaxf(bhi[c], __shfl_xor_sync(PC_FULL_MASK, bhi[c], of));
      }
    }
    const bool empty = !(blo[0] <= bhi[0]);  // a warp of padding only

    int old = 0, par = 1;
    int cmax = empty ? __float_as_int(-1.0f) : __float_as_int(1e38f), ckey = INT_MAX;  // cached warp winner
    bool fresh = false;  // cmax / ckey describe the current td values
    for (int j = 1; j < m; ++j) {
      const float cx = s_xyz[old * 3 + 0], cy = s_xyz[old * 3 + 1], cz = s_xyz[old * 3 + 2];
      // squared distance from the centre to this warp's box (0 inside), un-fused like everything else
      const float ex = fmaxf(fmaxf(blo[0] - cx, cx - bhi[0]), 0.f), ey = fmaxf(fmaxf(blo[1] - cy, cy - bhi[1]), 0.f),
                  ez = fmaxf(fmaxf(blo[2] - cz, cz - bhi[2]), 0.f);
      const float lb = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
      const bool skip = empty || (fresh && __fmul_rn(lb, 0.99999f) > __int_as_float(cmax));
      if (!skip) {  // warp-uniform
        const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
        float gm[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          gm[g] = -1.0f;
#pragma unroll
          for (int h = g * G / 2; h < (g + 1) * G / 2; ++h) {
            float d0, d1;
            unpack2(sqdist3_x2(px[h], py[h], pz[h], cx2, cy2, cz2, one2), d0, d1);
            td[2 * h] = fminf(d0, td[2 * h]);
            td[2 * h + 1] = fminf(d1, td[2 * h + 1]);
            gm[g] = fmaxf(gm[g], fmaxf(td[2 * h], td[2 * h + 1]));
          }
        }
        float vmax = gm[0];
#pragma unroll
        for (int g = 1; g < NG; ++g) vmax = fmaxf(vmax, gm[g]);
        const int vb = __float_as_int(vmax);
        cmax = __reduce_max_sync(PC_FULL_MASK, vb);
        int tb = INT_MAX;
        if (vb == cmax) {  // every slot at the maximum competes with its ORIGINAL index (slots are in sorted order)
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (__float_as_int(gm[g]) == cmax) {
#pragma unroll
              for (int e = 0; e < G; ++e)
                if (__float_as_int(td[g * G + e]) == cmax) tb = min(tb, tie_key((int)s_perm[pos0 + (g * G + e) * 32]));
            }
          }
        }
        ckey = __reduce_min_sync(PC_FULL_MASK, tb);
        fresh = true;
      }
      if (lane == 0) s_pair[par][warp] = make_int2(cmax, ckey);
      __syncthreads();
      const int2 pr = s_pair[par][lane < nwarps ? lane : 0];
      const int gmax = __reduce_max_sync(PC_FULL_MASK, pr.x);
      old = tie_key_to_index(__reduce_min_sync(PC_FULL_MASK, pr.x == gmax ? pr.y : INT_MAX));
      par ^= 1;
      if (tid == 0) {
        o[j] = old;
        if (oxyz) { oxyz[j * 3 + 0] = s_xyz[old * 3 + 0]; oxyz[j * 3 + 1] = s_xyz[old * 3 + 1]; oxyz[j * 3 + 2] = s_xyz[old * 3 + 2]; }
      }
    }
  }
}

template <int P, int T>
int launch_pruned(int b, int n, int m, const float *xyz, int *out, float *out_xyz, cudaStream_t st) {
  const size_t smem = (size_t)n * 3 * sizeof(float) + 4096 * sizeof(int) + (size_t)T * P * sizeof(unsigned short);
  PC_CUDA_TRY(allow_smem(fps_pruned_kernel<P, T>, smem));
  fps_pruned_kernel<P, T><<<b, T, smem, st>>>(b, n, m, 1.0f, xyz, out, out_xyz);
  PC_RETURN_LAUNCH_STATUS();
}

// ---------------------------------------------------------------------------------------------------------------
// 8192 < n <= 16 * 8192: one THREAD-BLOCK CLUSTER per scene.  CTA r of the cluster keeps points [8192 r, 8192 (r+1))
// on chip exactly as the single-CTA kernel does (8 warps x 32 points, same slot order, same one-barrier CTA-level
// winner); the CTA winners then meet through distributed shared memory: every CTA sends its (value, key, x, y, z)
// record into slot [parity][r] of EVERY CTA of the cluster with st.async, which completes transaction bytes on the
// receiver's mbarrier -- no cluster-wide barrier in the loop -- and every CTA reduces the C records locally; the record
// carries the winner's coordinates, so the next round starts without another remote read.  No global-memory traffic
// inside the m-1 rounds; the alternative for these sizes (running minima streamed through L2, fps_stream_kernel) moves
// 16 n bytes per round.
//
// Slice size: 8192 points per CTA (32 per thread, coordinates and running minima in registers) up to 131072 points;
// up to 262144 points a CTA keeps 16384 (64 per thread): running minima for all of them in registers, coordinates in
// registers for the first 32 slots and re-read from the shared-memory copy of the slice for the other 32 (PS).
constexpr int kSlice = 8192;
constexpr int kMaxClusterPoints = 16 * 16384;
struct __align__(16) FpsRec { int value, key; float x, y; float z; int pad[3]; };

// CL > 0: one cluster of CL CTAs per scene, records exchanged through distributed shared memory (above).
// CL == 0 (n > 262144): `cps` CTAs per scene in a COOPERATIVE launch (all resident), records exchanged through global
// memory: CTA r publishes its (value, key, x, y, z) record of round j as three 8-byte words in grec[scene][parity][r],
// the last one carrying the round number with release semantics; warp 0 of every CTA polls the cps records of the
// round with acquire loads, reduces them and hands the next centre to the CTA through shared memory.  Slot reuse two
// rounds later is safe for the same reason as in the cluster protocol: nobody can publish round j+2 before it has
// read every peer's round j+1 record, which that peer wrote after reading round j.  One L2 round trip (~1.5 us) per
// round instead of streaming 16 n bytes of running minima through L2 (fps_stream_kernel): N = 1 M points, m = 1024,
// 64 scenes 730 ms -> see DESIGN.md.
template <int CL, int P, int PS>
__global__ void __maxnreg__((PS > 0) ? 255 : 200)
fps_cluster_kernel(int n, int m, float one, const float *__restrict__ xyz, int *__restrict__ out,
                   float *__restrict__ out_xyz, int cps, int scene0, unsigned long long *__restrict__ grec) {
  constexpr int T = 256, H = P / 2, HR = (P - PS) / 2, G = 8, NG = P / G, nwarps = T / 32, kSlice = P * T;
  constexpr bool kGlobal = (CL == 0);
  using Map = FpsMap<P, T>;
  extern __shared__ float s_xyz[];  // this CTA's slice, up to kSlice * 3
  __shared__ int2 s_pair[2][32];
  __shared__ FpsRec s_rec[2][16];
  __shared__ __align__(8) uint64_t s_mb[2];
  __shared__ float s_next[4];       // kGlobal: next centre (x, y, z) and its index, written by warp 0
  const int rank = kGlobal ? (int)(blockIdx.x % (unsigned)cps) : (int)cg::this_cluster().block_rank();
  const int scene = kGlobal ? scene0 + (int)(blockIdx.x / (unsigned)cps) : (int)(blockIdx.x / (CL > 0 ? CL : 1));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const f32x2 one2 = pack2(one, one);
  const float *p = xyz + (size_t)scene * n * 3;
  int *o = out + (size_t)scene * m;
  float *oxyz = out_xyz ? out_xyz + (size_t)scene * m * 3 : nullptr;
  const int base = rank * kSlice, sn = max(0, min(kSlice, n - base));  // this CTA's points [base, base + sn)
  for (int i = tid; i < sn * 3; i += T) s_xyz[i] = p[(size_t)base * 3 + i];
  if (PS > 0)
    for (int i = sn * 3 + tid; i < kSlice * 3; i += T) s_xyz[i] = 0.0f;
  float cx = p[0], cy = p[1], cz = p[2];  // first centre = point 0 (tf_sampling_g.cu:114-116)
  if (rank == 0 && tid == 0) {
    o[0] = 0;
    if (oxyz) { oxyz[0] = cx; oxyz[1] = cy; oxyz[2] = cz; }
  }
  __syncthreads();

  f32x2 px[HR], py[HR], pz[HR];
  float td[P];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float x[2], y[2], z[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * h + e, k = Map::k_of(tid, i);  // local index; global index = base + k, same (k mod 512)
      if (k < sn) {
        x[e] = s_xyz[k * 3 + 0]; y[e] = s_xyz[k * 3 + 1]; z[e] = s_xyz[k * 3 + 2];
        td[i] = 1e38f;
      } else {
        x[e] = y[e] = z[e] = 0.0f;
        td[i] = -1.0f;  // never a winner; its shared-memory slot (PS part) is zero-filled below
      }
    }
    if (h < HR) { px[h] = pack2(x[0], x[1]); py[h] = pack2(y[0], y[1]); pz[h] = pack2(z[0], z[1]); }
  }
  const uint32_t mb_local[2] = {(uint32_t)__cvta_generic_to_shared(&s_mb[0]), (uint32_t)__cvta_generic_to_shared(&s_mb[1])};
  if constexpr (!kGlobal) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb_local[0]), "r"(1u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb_local[1]), "r"(1u));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
    cg::this_cluster().sync();  // every CTA's s_rec / s_mb is initialised and addressable before the first remote store
  }
  unsigned long long *myrec = nullptr;
  if constexpr (kGlobal) myrec = grec + (size_t)(scene - scene0) * 2 * cps * 3;  // [parity][cps][3 words] per scene

  int par = 1;
  for (int j = 1; j < m; ++j) {
    const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
    float gm[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      gm[g] = -1.0f;
#pragma unroll
      for (int h = g * G / 2; h < (g + 1) * G / 2; ++h) {
        float d0, d1;
        if (h < HR) {
          unpack2(sqdist3_x2(px[h], py[h], pz[h], cx2, cy2, cz2, one2), d0, d1);
        } else {  // coordinates from the slice copy (stride-3 words across lanes: conflict-free); slots beyond sn read
                  // the zero fill -- their td stays -1 because fminf(d, -1) = -1 for every d >= 0
          const int k0 = Map::k_of(tid, 2 * h), k1 = Map::k_of(tid, 2 * h + 1);
          const f32x2 qx = pack2(s_xyz[k0 * 3 + 0], s_xyz[k1 * 3 + 0]), qy = pack2(s_xyz[k0 * 3 + 1], s_xyz[k1 * 3 + 1]),
                      qz = pack2(s_xyz[k0 * 3 + 2], s_xyz[k1 * 3 + 2]);
          unpack2(sqdist3_x2(qx, qy, qz, cx2, cy2, cz2, one2), d0, d1);
        }
        td[2 * h] = fminf(d0, td[2 * h]);
        td[2 * h + 1] = fminf(d1, td[2 * h + 1]);
        gm[g] = fmaxf(gm[g], fmaxf(td[2 * h], td[2 * h + 1]));
      }
    }
    float vmax = gm[0];
#pragma unroll
    for (int g = 1; g < NG; ++g) vmax = fmaxf(vmax, gm[g]);
    const int vb = __float_as_int(vmax);
    const int wmax = __reduce_max_sync(PC_FULL_MASK, vb);
    int tb = INT_MAX;
    if (vb == wmax) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (tb == INT_MAX && __float_as_int(gm[g]) == wmax) {
#pragma unroll
          for (int e = G - 1; e >= 0; --e)
            if (__float_as_int(td[g * G + e]) == wmax) tb = tie_key(base + Map::k_of(tid, g * G + e));
        }
      }
    }
    const int wkey = __reduce_min_sync(PC_FULL_MASK, tb);
    if (lane == 0) s_pair[par][warp] = make_int2(wmax, wkey);
    __syncthreads();
    const int2 pr = s_pair[par][lane < nwarps ? lane : 0];
    const int cmax = __reduce_max_sync(PC_FULL_MASK, pr.x);
    const int ckey = __reduce_min_sync(PC_FULL_MASK, pr.x == cmax ? pr.y : INT_MAX);
    if constexpr (kGlobal) {
      unsigned long long *slot = myrec + (size_t)par * cps * 3;
      if (tid == 0) {
        int rx = 0, ry = 0, rz = 0;
        if (cmax >= 0) {
          const int kl = tie_key_to_index(ckey) - base;
          rx = __float_as_int(s_xyz[kl * 3 + 0]); ry = __float_as_int(s_xyz[kl * 3 + 1]); rz = __float_as_int(s_xyz[kl * 3 + 2]);
        }
        unsigned long long *w = slot + (size_t)rank * 3;
        w[0] = ((unsigned long long)(unsigned)cmax << 32) | (unsigned)ckey;
        w[1] = ((unsigned long long)(unsigned)rx << 32) | (unsigned)ry;
        const unsigned long long w2 = ((unsigned long long)(unsigned)rz << 32) | (unsigned)j;
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(w + 2), "l"(w2) : "memory");
      }
      if (warp == 0) {
        int bv = INT_MIN, bk = INT_MAX;
        float bx = 0.f, by = 0.f, bz = 0.f;
        for (int r = lane; r < cps; r += 32) {
          const unsigned long long *w = slot + (size_t)r * 3;
          unsigned long long w2;
          unsigned spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w2) : "l"(w + 2) : "memory");
            if ((unsigned)w2 != (unsigned)j && ++spins > (1u << 24)) __trap();
          } while ((unsigned)w2 != (unsigned)j);
          const unsigned long long w0 = __ldcg(w), w1 = __ldcg(w + 1);  // L2, ordered after the acquire
          const int v = (int)(w0 >> 32), k = (int)(unsigned)w0;
          if (v > bv || (v == bv && k < bk)) {
            bv = v; bk = k;
            bx = __int_as_float((int)(w1 >> 32)); by = __int_as_float((int)(unsigned)w1); bz = __int_as_float((int)(w2 >> 32));
          }
        }
        const int gmax = __reduce_max_sync(PC_FULL_MASK, bv);
        const int gkey = __reduce_min_sync(PC_FULL_MASK, bv == gmax ? bk : INT_MAX);
        const int src = __ffs(__ballot_sync(PC_FULL_MASK, bv == gmax && bk == gkey)) - 1;
        bx = __shfl_sync(PC_FULL_MASK, bx, src); by = __shfl_sync(PC_FULL_MASK, by, src); bz = __shfl_sync(PC_FULL_MASK, bz, src);
        if (lane == 0) { s_next[0] = bx; s_next[1] = by; s_next[2] = bz; s_next[3] = __int_as_float(gkey); }
      }
      __syncthreads();
      cx = s_next[0]; cy = s_next[1]; cz = s_next[2];
      const int gkey = __float_as_int(s_next[3]);
      par ^= 1;
      if (rank == 0 && tid == 0) {
        o[j] = tie_key_to_index(gkey);
        if (oxyz) { oxyz[j * 3 + 0] = cx; oxyz[j * 3 + 1] = cy; oxyz[j * 3 + 2] = cz; }
      }
      continue;  // (the next round's first barrier orders the s_next reads before warp 0 overwrites it)
    }
    // Exchange without a cluster barrier: thread i sends this CTA's 32-byte record straight into slot [par][rank] of
    // CTA i with two st.async (16 bytes each) that complete transaction bytes on CTA i's mbarrier s_mb[par]; every CTA
    // armed that barrier for CL * 32 bytes and simply waits for its phase.
    if (tid == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_local[par]), "r"((uint32_t)(CL * 32)) : "memory");
    if (tid < CL) {
      int rx = 0, ry = 0, rz = 0;
      if (cmax >= 0) {  // a CTA without live points reports (-1.0f bits): never the cluster winner
        const int kl = tie_key_to_index(ckey) - base;
        rx = __float_as_int(s_xyz[kl * 3 + 0]); ry = __float_as_int(s_xyz[kl * 3 + 1]); rz = __float_as_int(s_xyz[kl * 3 + 2]);
      }
      const uint32_t slot = (uint32_t)__cvta_generic_to_shared(&s_rec[par][rank]);
      uint32_t rslot, rbar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rslot) : "r"(slot), "r"((uint32_t)tid));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(mb_local[par]), "r"((uint32_t)tid));
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rslot),
                   "r"(cmax), "r"(ckey), "r"(rx), "r"(ry), "r"(rbar)
                   : "memory");
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rslot + 16),
                   "r"(rz), "r"(0), "r"(0), "r"(0), "r"(rbar)
                   : "memory");
    }
    {
      const uint32_t parity = ((j - 1) >> 1) & 1;  // s_mb[par] is used every second round
      uint32_t ok = 0, spins = 0;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(mb_local[par]), "r"(parity)
                     : "memory");
        if (!ok && ++spins > (1u << 26)) __trap();
      } while (!ok);
    }
    const FpsRec *rr = &s_rec[par][lane < CL ? lane : 0];
    const int rv = rr->value, rk = rr->key;
    const float rx = rr->x, ry = rr->y, rz = rr->z;
    const int gmax = __reduce_max_sync(PC_FULL_MASK, rv);
    const int gkey = __reduce_min_sync(PC_FULL_MASK, rv == gmax ? rk : INT_MAX);
    const int src = __ffs(__ballot_sync(PC_FULL_MASK, rv == gmax && rk == gkey)) - 1;
    cx = __shfl_sync(PC_FULL_MASK, rx, src);
    cy = __shfl_sync(PC_FULL_MASK, ry, src);
    cz = __shfl_sync(PC_FULL_MASK, rz, src);
    par ^= 1;
    if (rank == 0 && tid == 0) {
      o[j] = tie_key_to_index(gkey);
      if (oxyz) { oxyz[j * 3 + 0] = cx; oxyz[j * 3 + 1] = cy; oxyz[j * 3 + 2] = cz; }
    }
  }
  if constexpr (!kGlobal) cg::this_cluster().sync();  // no CTA may exit while a peer can still write into its shared memory
}

template <int CL, int P = 32, int PS = 0>
int launch_cluster(int b, int n, int m, const float *xyz, int *out, float *out_xyz, cudaStream_t st) {
  constexpr int slice = P * 256;
  const size_t smem = (size_t)(PS > 0 ? slice : (n < slice ? n : slice)) * 3 * sizeof(float);
  auto kernel = fps_cluster_kernel<CL, P, PS>;
  PC_CUDA_TRY(allow_smem(kernel, smem));
  if (CL > 8) PC_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)b * CL);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, n, m, 1.0f, xyz, out, out_xyz, 0, 0, (unsigned long long *)nullptr));
  return PC_OK;
}

// n > 262144: cooperative launches of `spw` scenes x `cps` CTAs (16384-point slices, all resident), one after the other
// on the stream; the record area (workspace) is cleared first so that no stale round number can match.
constexpr int kCoopSlice = 16384;
inline int coop_cps(int n) { return (n + kCoopSlice - 1) / kCoopSlice; }
inline size_t coop_rec_bytes(int scenes, int cps) { return (size_t)scenes * 2 * cps * 3 * sizeof(unsigned long long); }

int launch_coop(int b, int n, int m, const float *xyz, int *out, float *out_xyz, void *workspace, cudaStream_t st) {
  auto kernel = fps_cluster_kernel<0, 64, 32>;
  const size_t smem = (size_t)kCoopSlice * 3 * sizeof(float);
  PC_CUDA_TRY(allow_smem(kernel, smem));
  const int cps = coop_cps(n);
  int spw = num_sms() / cps;  // scenes per launch: every CTA of a launch must be resident (1 CTA per SM: 192 KB smem)
  if (spw < 1) return PC_ERR_UNSUPPORTED;
  if (spw > b) spw = b;
  for (int s0 = 0; s0 < b; s0 += spw) {
    const int ns = (b - s0 < spw) ? b - s0 : spw;
    PC_CUDA_TRY(cudaMemsetAsync(workspace, 0, coop_rec_bytes(ns, cps), st));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ns * cps));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, n, m, 1.0f, xyz, out, out_xyz, cps, s0, (unsigned long long *)workspace));
  }
  return PC_OK;
}

// General sizes (n > 8192): running min-distances stream through a global workspace (b*n floats),
// xyz is read through L1/L2.  Same reduction protocol as above.
__global__ void __launch_bounds__(1024, 1)
fps_stream_kernel(int b, int n, int m, const float *__restrict__ xyz, float *__restrict__ temp,
                  int *__restrict__ out) {
  __shared__ int s_wmax[32];
  __shared__ int s_tb[2];
  const int tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  for (int scene = blockIdx.x; scene < b; scene += gridDim.x) {
    const float *p = xyz + (size_t)scene * n * 3;
    float *td = temp + (size_t)scene * n;
    int *o = out + (size_t)scene * m;
    __syncthreads();
    for (int k = tid; k < n; k += T) td[k] = 1e38f;
    if (tid == 0) { s_tb[0] = INT_MAX; s_tb[1] = INT_MAX; o[0] = 0; }
    __syncthreads();
    int old = 0;
    for (int j = 1; j < m; ++j) {
      const float cx = __ldg(p + old * 3 + 0), cy = __ldg(p + old * 3 + 1), cz = __ldg(p + old * 3 + 2);
      float vmax = -1.0f;
      for (int k = tid; k < n; k += T) {
        float d = sqdist3(__ldg(p + k * 3 + 0), __ldg(p + k * 3 + 1), __ldg(p + k * 3 + 2), cx, cy, cz);
        float t = td[k];
        float d2 = fminf(d, t);
        if (d2 != t) td[k] = d2;
        vmax = fmaxf(vmax, d2);
      }
      const int vb = __float_as_int(vmax);
      const int wmax = __reduce_max_sync(PC_FULL_MASK, vb);
      if (lane == 0) s_wmax[warp] = wmax;
      __syncthreads();
      const int g = (lane < nwarps) ? s_wmax[lane] : INT_MIN;
      const int gmax = __reduce_max_sync(PC_FULL_MASK, g);
      const int slot = j & 1;
      if (vb == gmax) {
        int tb = INT_MAX;
        for (int k = tid; k < n; k += T)
          if (__float_as_int(td[k]) == gmax) tb = min(tb, tie_key(k));
        atomicMin(&s_tb[slot], tb);
      }
      if (tid == 0) s_tb[slot ^ 1] = INT_MAX;
      __syncthreads();
      old = tie_key_to_index(s_tb[slot]);
      if (tid == 0) o[j] = old;
    }
  }
}

template <int P, int T, bool kOneBarrier = true, int kRegs = (P >= 32 ? 168 : 128)>
int launch_onchip(int b, int n, int m, const float *xyz, int *out, float *out_xyz, cudaStream_t st) {
  size_t smem = (size_t)n * 3 * sizeof(float);
  if (smem > 48 * 1024) PC_CUDA_TRY(allow_smem(fps_onchip_kernel<P, T, kOneBarrier, kRegs>, smem));
  int grid = b;  // one CTA per scene; more scenes than SMs simply queue (1 CTA/SM resident)
  fps_onchip_kernel<P, T, kOneBarrier, kRegs><<<grid, T, smem, st>>>(b, n, m, 1.0f, xyz, out, out_xyz);
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace
}  // namespace pc

extern "C" size_t pc_fps_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 0;
  if (n <= pc::kMaxClusterPoints) return 0;  // single CTA or one thread-block cluster per scene: all state on chip
  const int cps = pc::coop_cps(n);
  if (cps <= 148) return pc::coop_rec_bytes(b < 148 / cps ? b : 148 / cps, cps);  // cooperative grid: records only
  return (size_t)b * n * sizeof(float);
}

// FPS that also writes the sampled coordinates: out_xyz (b,m,3) = xyz[b, out_idx[b,j], :], i.e. the
// farthest_point_sample + gather_point pair every caller issues back to back (utils/pointnet_util.py:34), in one launch.
extern "C" int pc_fps_gather(int b, int n, int m, const float *xyz, void *workspace, int *out_idx, float *out_xyz,
                             pc_stream_t stream) {
  if (b < 0 || n < 0) return PC_ERR_INVALID_ARGUMENT;
  if (m <= 0 || b == 0) return PC_OK;  // tf_sampling_g.cu:106-107
  if (n == 0) return PC_ERR_INVALID_ARGUMENT;
  if (!xyz || !out_idx) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= pc::kMaxRegPoints) {
    // Few, fat threads (see the kernel comment).  n <= 256 runs in ONE warp (no barrier at all); up to 4096 points
    // take 4 warps (one per SM sub-partition); beyond that 8 warps x 32 points per thread.
    if (n <= 64) return pc::launch_onchip<2, 32>(b, n, m, xyz, out_idx, out_xyz, st);
    if (n <= 128) return pc::launch_onchip<4, 32>(b, n, m, xyz, out_idx, out_xyz, st);
    if (n <= 256) return pc::launch_onchip<8, 32>(b, n, m, xyz, out_idx, out_xyz, st);
    if (n <= 512) return pc::launch_onchip<4, 128>(b, n, m, xyz, out_idx, out_xyz, st);
    if (n <= 1024) return pc::launch_onchip<8, 128>(b, n, m, xyz, out_idx, out_xyz, st);
    if (n <= 2048) return pc::launch_onchip<16, 128>(b, n, m, xyz, out_idx, out_xyz, st);
    if (n <= 4096) return pc::launch_onchip<32, 128>(b, n, m, xyz, out_idx, out_xyz, st);
    // 4097..8192 points (measured on B200, 16 ScanNet-shaped scenes x 8192 -> 1024, one launch alone):
    //   pruned kernel, 16 warps x 16 points per thread (this one)                  539 us
    //   un-pruned, 8 warps x 32 points, one barrier, 168 registers                 645 us
    //   un-pruned, 16 warps x 16 points, two-barrier tail                          584 us
    //   pruned, 8 warps x 32 points                                                618 us
    // All give the same indices.
    return pc::launch_pruned<16, 512>(b, n, m, xyz, out_idx, out_xyz, st);
  }
  if (n <= pc::kMaxClusterPoints && (long long)b * 16 < 0x7fffffffLL) {  // one cluster of 2 / 4 / 8 / 16 CTAs per scene
    const int slices = (n + pc::kSlice - 1) / pc::kSlice;
    if (slices <= 2) return pc::launch_cluster<2>(b, n, m, xyz, out_idx, out_xyz, st);
    if (slices <= 4) return pc::launch_cluster<4>(b, n, m, xyz, out_idx, out_xyz, st);
    if (slices <= 8) return pc::launch_cluster<8>(b, n, m, xyz, out_idx, out_xyz, st);
    if (slices <= 16) return pc::launch_cluster<16>(b, n, m, xyz, out_idx, out_xyz, st);
    return pc::launch_cluster<16, 64, 32>(b, n, m, xyz, out_idx, out_xyz, st);  // 16384-point slices
  }
  if (!workspace) return PC_ERR_WORKSPACE;
  if (pc::coop_cps(n) <= 148 && pc::coop_cps(n) <= pc::num_sms())
    return pc::launch_coop(b, n, m, xyz, out_idx, out_xyz, workspace, st);
  pc::fps_stream_kernel<<<b, 1024, 0, st>>>(b, n, m, xyz, (float *)workspace, out_idx);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  if (out_xyz) return pc_gather_point(b, n, m, xyz, out_idx, out_xyz, stream);
  return PC_OK;
}

extern "C" int pc_fps(int b, int n, int m, const float *xyz, void *workspace, int *out_idx, pc_stream_t stream) {
  return pc_fps_gather(b, n, m, xyz, workspace, out_idx, nullptr, stream);
}
