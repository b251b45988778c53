// Per-neighbourhood attention contraction for sm_100a (fp32, HBM-streaming).
//
// Replaces the reshape / matmul / softmax / matmul tail of AttentionLayer.call (reference
// attention_points/attention_scannet/attention_layer.py:35-42), which TensorFlow runs as two batched cuBLAS GEMMs
// with M=1, K=key_dim=4 plus a softmax and three intermediate tensors.  Here one warp owns one (neighbourhood, head):
// thanks to the reference's RAW reshape (:35) a head's S pseudo-keys are one contiguous run of S*D floats in the K and
// V buffers, so lane s reads key s with a single 128-bit load (D=4), the softmax is two warp reductions and the
// weighted value sum is D warp reductions.  K and V are read exactly once; nothing is written but the (G,HD) output.
// Arithmetic intensity is 0.5 flop/byte -- this op is HBM-bound by construction, it is not tensor-core work.
#include <math.h>
#include "common.cuh"

namespace pc {
namespace {

constexpr int kAttWarps = 8;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(PC_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PC_FULL_MASK, v, o);
  return v;
}

template <int D, bool VEC>
__device__ __forceinline__ void load_row(const float *__restrict__ p, float (&r)[D]) {
  if (VEC && D % 4 == 0) {
#pragma unroll
    for (int d = 0; d < D; d += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4 *>(p + d));
      r[d] = t.x; r[d + 1] = t.y; r[d + 2] = t.z; r[d + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) r[d] = __ldg(p + d);
  }
}

template <int D, bool VEC>
__device__ __forceinline__ void store_row(float *__restrict__ p, const float (&r)[D]) {
  if (VEC && D % 4 == 0) {
#pragma unroll
    for (int d = 0; d < D; d += 4) *reinterpret_cast<float4 *>(p + d) = make_float4(r[d], r[d + 1], r[d + 2], r[d + 3]);
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) p[d] = r[d];
  }
}

// Softmax weights of one head for the lane's samples s = lane + 32*i (i < NS); returns them in a[].
template <int D, int NS, bool VEC>
__device__ __forceinline__ void head_softmax(int S, const float *__restrict__ kh, const float (&q)[D], float rsd,
                                             float (&a)[NS]) {
  const int lane = threadIdx.x & 31;
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int s = lane + 32 * i;
    a[i] = -INFINITY;
    if (s < S) {
      float k[D];
      load_row<D, VEC>(kh + (size_t)s * D, k);
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) acc = fmaf(q[d], k[d], acc);
      a[i] = acc * rsd;
    }
    mx = fmaxf(mx, a[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    a[i] = (lane + 32 * i < S) ? expf(a[i] - mx) : 0.f;
    sum += a[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < NS; ++i) a[i] *= inv;
}

// Sum of D per-lane values over the warp, result for component d valid in every lane.  For D == 4 the classic
// halving exchange is used: 6 shuffles instead of 4 x 5.
template <int D>
__device__ __forceinline__ void warp_sum_vec(float (&o)[D]) {
  if (D == 4) {
    const int lane = threadIdx.x & 31;
    // step 1 (xor 16): lanes 0-15 keep components 0,1; lanes 16-31 keep 2,3
    const bool hi = lane & 16;
    float s0 = hi ? o[0] : o[2], s1 = hi ? o[1] : o[3];          // what this lane gives away
    float k0 = hi ? o[2] : o[0], k1 = hi ? o[3] : o[1];          // what it keeps
    k0 += __shfl_xor_sync(PC_FULL_MASK, s0, 16);
    k1 += __shfl_xor_sync(PC_FULL_MASK, s1, 16);
    // step 2 (xor 8): keep one component
    const bool hi8 = lane & 8;
    float g = hi8 ? k0 : k1, k = hi8 ? k1 : k0;
    k += __shfl_xor_sync(PC_FULL_MASK, g, 8);
    k += __shfl_xor_sync(PC_FULL_MASK, k, 4);
    k += __shfl_xor_sync(PC_FULL_MASK, k, 2);
    k += __shfl_xor_sync(PC_FULL_MASK, k, 1);
    // lane l now holds component c(l) = 2*(l>>4 & 1) + (l>>3 & 1); lanes 0, 8, 16, 24 hold components 0, 1, 2, 3
    o[0] = __shfl_sync(PC_FULL_MASK, k, 0);
    o[1] = __shfl_sync(PC_FULL_MASK, k, 8);
    o[2] = __shfl_sync(PC_FULL_MASK, k, 16);
    o[3] = __shfl_sync(PC_FULL_MASK, k, 24);
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = warp_sum(o[d]);
  }
}

// One warp per (neighbourhood, head), warp-stride loop over heads.  When a head is at most one sample per lane
// (S <= 32, the reference's nsample) the K and V rows of the NEXT head are requested before the current head's
// softmax, so every lane keeps four 128-bit loads in flight: the kernel is a pure HBM stream of K and V.
template <int D, int NS, bool VEC>
__global__ void __launch_bounds__(kAttWarps * 32)
attention_fwd_kernel(size_t heads_total, int S, int H, const float *__restrict__ Q, const float *__restrict__ K,
                     const float *__restrict__ V, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * kAttWarps;
  size_t gh = (size_t)blockIdx.x * kAttWarps + (threadIdx.x >> 5);  // g*H + h
  const float rsd = 1.0f / sqrtf((float)D);
  if constexpr (NS == 1 && D <= 16) {   // (wider heads: the double-buffered rows would not fit the register file)
    // (g,h) chunk: K + g*S*H*D + h*S*D == K + gh*S*D ;  Q + g*H*D + h*D == Q + gh*D
    const bool has = lane < S;
    float kc[D], vc[D], qc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) kc[d] = vc[d] = qc[d] = 0.f;
    if (gh < heads_total) {
      if (has) {
        load_row<D, VEC>(K + gh * (size_t)S * D + (size_t)lane * D, kc);
        load_row<D, VEC>(V + gh * (size_t)S * D + (size_t)lane * D, vc);
      }
      load_row<D, VEC>(Q + gh * D, qc);
    }
    for (; gh < heads_total; gh += stride) {
      const size_t nx = gh + stride;
      float kn[D], vn[D], qn[D];
#pragma unroll
      for (int d = 0; d < D; ++d) kn[d] = vn[d] = qn[d] = 0.f;
      if (nx < heads_total) {  // prefetch the next head of this warp
        if (has) {
          load_row<D, VEC>(K + nx * (size_t)S * D + (size_t)lane * D, kn);
          load_row<D, VEC>(V + nx * (size_t)S * D + (size_t)lane * D, vn);
        }
        load_row<D, VEC>(Q + nx * D, qn);
      }
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) acc = fmaf(qc[d], kc[d], acc);
      const float logit = has ? acc * rsd : -INFINITY;
      const float mx = warp_max(logit);
      const float e = has ? expf(logit - mx) : 0.f;
      const float inv = 1.0f / warp_sum(e);
      const float a = e * inv;
      float o[D];
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] = fmaf(a, vc[d], 0.f);
      warp_sum_vec<D>(o);
      if (lane == 0) store_row<D, VEC>(out + gh * D, o);
#pragma unroll
      for (int d = 0; d < D; ++d) { kc[d] = kn[d]; vc[d] = vn[d]; qc[d] = qn[d]; }
    }
  } else
  for (; gh < heads_total; gh += stride) {
    const float *kh = K + gh * (size_t)S * D;
    const float *vh = V + gh * (size_t)S * D;
    float q[D];
    load_row<D, VEC>(Q + gh * D, q);
    float a[NS];
    head_softmax<D, NS, VEC>(S, kh, q, rsd, a);
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const int s = lane + 32 * i;
      if (s < S) {
        float v[D];
        load_row<D, VEC>(vh + (size_t)s * D, v);
#pragma unroll
        for (int d = 0; d < D; ++d) o[d] = fmaf(a[i], v[d], o[d]);
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = warp_sum(o[d]);
    if (lane == 0) store_row<D, VEC>(out + gh * D, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// S = 32, D = 4 (every attention level of the ScanNet models): LANE = HEAD, operands staged by bulk copies.
// The warp-per-head kernel above spends ~90 warp instructions per KB of K / V (two shuffle reductions for the softmax,
// a third for the weighted sum, expf, prefetch bookkeeping): ncu showed it at 74 % of the issue slots for 0.73 of the
// HBM peak -- an HBM stream that is nearly issue-bound, and that competes for issue slots with the FPS CTAs it shares
// SMs with in the pipeline.  Here a warp owns 32 CONSECUTIVE heads = one contiguous 16 KB run of K and one of V (the
// raw reshape, attention_layer.py:35, makes head gh the 128 floats at offset 128 gh).  Lane 0 issues one cp.async.bulk
// per run into the warp's shared-memory ring (mbarrier complete_tx), and lane l then walks the
// 32 pseudo-keys of head l out of shared memory with 128-bit loads: logits and maximum in a first pass over K, exp and
// the weighted sum in a second pass over V -- no shuffles, no reductions, ~16 warp instructions per KB.  Lane l visits
// its samples in the rotated order (s + l) mod 32, which keeps every quarter-warp's eight 16-byte accesses in eight
// different bank groups (the heads are 512 bytes apart).
// Footprint, measured in the pipelined step (8 batches in flight, FPS CTAs resident on 128 SMs with 128 KB of shared
// memory and 45 k registers each), scenes/s of the whole step:
//   warps/CTA x ring slots x blocks/warp   2x3 persistent 75.5 k | 1x3x16 77.8 k | 1x4x16 76.0 k | 1x6x16 72.0 k
//                                          1x2x16 80.1 k | 1x2x32 82.5 k | 1x2x64 83.6 k | 1x2x128 83.6 k | 1x2x256 78.8 k
// The kernel is an HBM stream that shares every SM with other kernels: what pays THERE is a small shared-memory
// reservation (32 KB CTAs find room beside an FPS CTA plus a tile-kernel CTA; 96 KB ones queued for the few free SMs)
// and FEW, long-lived warps (one 16 KB bulk copy in flight per warp is enough once a hundred warps stream; more warps
// only take issue slots and L2 bandwidth from the kernels in flight next to it).  A lone launch wants the opposite (SA1
// alone: 46 us with 592 warps x 3 slots, 58 us with 296 x 2, 86 us with 148 x 2).  So the shape follows the caller's
// concurrency hint (pc_set_concurrency_hint): 1 -> four single-warp CTAs per SM with three slots; 2..7 -> two per SM,
// two slots; >= 8 -> one per SM, two slots.  K / V copies carry an L2 evict-first policy (+1.5 k scenes/s: 500 MB read
// once no longer evict the gathers' working sets).
constexpr int kLhMaxSlots = 3;
constexpr int kLhSlotBytes = 32 * 128 * 4;  // one slot = the K (or V) run of 32 heads = 16 KB

__device__ __forceinline__ void lh_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();
  } while (!ok);
}

__global__ void __launch_bounds__(32)
attention_fwd_lanehead_kernel(size_t heads_total, int bpw, int nslots, const float *__restrict__ Q, const float *__restrict__ K,
                              const float *__restrict__ V, float *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char lh_smem[];
  __shared__ __align__(8) uint64_t s_bar[kLhMaxSlots];
  const int lane = threadIdx.x;
  unsigned char *ring = lh_smem;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
  if (lane == 0) {
    for (int i = 0; i < nslots; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s + 8 * i), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncwarp();
  const size_t nblocks = (heads_total + 31) / 32;
  // the warp owns bpw CONSECUTIVE blocks of 32 heads
  const size_t first = (size_t)blockIdx.x * (size_t)bpw;
  const int my_blocks = first < nblocks ? (int)(nblocks - first < (size_t)bpw ? nblocks - first : (size_t)bpw) : 0;
  const int nseq = 2 * my_blocks;   // the warp's copy sequence: K of block 0, V of block 0, K of block 1, ... round robin over the slots

  uint64_t policy;   // K and V are read exactly once: do not let 500 MB of them push the gathers' working sets out of L2
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  int issue_slot = 0;               // slot of the next copy to issue; copies are issued in sequence order
  auto issue = [&](int seq) {
    const size_t h0 = (first + (size_t)(seq >> 1)) * 32;
    const uint32_t nh = (uint32_t)((heads_total - h0 < 32) ? heads_total - h0 : 32);
    const uint32_t bytes = nh * 512u;
    const int slot = issue_slot;
    issue_slot = issue_slot + 1 == nslots ? 0 : issue_slot + 1;
    const float *src = ((seq & 1) ? V : K) + h0 * 128;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s + 8 * slot), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     ring_s + slot * kLhSlotBytes),
                 "l"(src), "r"(bytes), "r"(bar_s + 8 * slot), "l"(policy)
                 : "memory");
  };
  if (lane == 0)
    for (int seq = 0; seq < nslots && seq < nseq; ++seq) issue(seq);
  int slot = 0;                     // slot and phase of the next copy to consume
  uint32_t phase = 0;
  auto advance = [&]() {
    if (++slot == nslots) { slot = 0; phase ^= 1u; }
  };
  for (int b = 0; b < my_blocks; ++b) {
    const size_t gh = (first + (size_t)b) * 32 + lane;
    const bool live = gh < heads_total;
    const float4 q = live ? __ldg(reinterpret_cast<const float4 *>(Q) + gh) : make_float4(0.f, 0.f, 0.f, 0.f);
    const int seq_k = 2 * b, seq_v = 2 * b + 1;
    const float4 *ks = reinterpret_cast<const float4 *>(ring + (size_t)slot * kLhSlotBytes) + lane * 32;
    float p[32];
    float mx = -INFINITY;
    lh_wait(bar_s + 8 * slot, phase);
    advance();
    if (live) {
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float4 k = ks[(t + lane) & 31];
        // (q . k) / sqrt(4), accumulated like the warp-per-head kernel: fmaf chain from 0, then the scale
        const float lg = 0.5f * fmaf(q.w, k.w, fmaf(q.z, k.z, fmaf(q.y, k.y, fmaf(q.x, k.x, 0.f))));
        p[t] = lg;
        mx = fmaxf(mx, lg);
      }
    }
    __syncwarp();   // every lane is done with the K slot: refill it with the copy nslots steps ahead
    if (lane == 0 && seq_k + nslots < nseq) issue(seq_k + nslots);
    float sum = 0.f, o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    const float4 *vs = reinterpret_cast<const float4 *>(ring + (size_t)slot * kLhSlotBytes) + lane * 32;
    lh_wait(bar_s + 8 * slot, phase);
    advance();
    if (live) {
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float e = exp2f((p[t] - mx) * 1.4426950408889634f);
        const float4 v = vs[(t + lane) & 31];
        sum += e;
        o0 = fmaf(e, v.x, o0); o1 = fmaf(e, v.y, o1); o2 = fmaf(e, v.z, o2); o3 = fmaf(e, v.w, o3);
      }
      const float inv = 1.0f / sum;
      reinterpret_cast<float4 *>(out)[gh] = make_float4(o0 * inv, o1 * inv, o2 * inv, o3 * inv);
    }
    __syncwarp();
    if (lane == 0 && seq_v + nslots < nseq) issue(seq_v + nslots);
  }
}

int launch_fwd_lanehead(size_t heads, const float *Q, const float *K, const float *V, float *out, cudaStream_t st) {
  const int h = concurrency_hint();
  const int nslots = h == 1 ? 3 : 2;
  const size_t per_sm = h == 1 ? 4 : h < 8 ? 2 : 1;
  const size_t smem = (size_t)nslots * kLhSlotBytes;
  PC_CUDA_TRY(allow_smem(attention_fwd_lanehead_kernel, (size_t)kLhMaxSlots * kLhSlotBytes));
  const size_t nblocks = (heads + 31) / 32;
  const size_t ctas = (size_t)num_sms() * per_sm;        // equal contiguous shares of at most 64 blocks
  size_t bpw = (nblocks + ctas - 1) / ctas;
  bpw = bpw > 64 ? 64 : bpw < 1 ? 1 : bpw;
  const size_t grid = (nblocks + bpw - 1) / bpw;
  attention_fwd_lanehead_kernel<<<(unsigned)grid, 32, smem, st>>>(heads, (int)bpw, nslots, Q, K, V, out);
  PC_RETURN_LAUNCH_STATUS();
}

template <int D, int NS, bool VEC>
__global__ void __launch_bounds__(kAttWarps * 32)
attention_bwd_kernel(size_t heads_total, int S, int H, const float *__restrict__ Q, const float *__restrict__ K,
                     const float *__restrict__ V, const float *__restrict__ dout, float *__restrict__ dQ,
                     float *__restrict__ dK, float *__restrict__ dV) {
  const int lane = threadIdx.x & 31;
  const size_t gh = (size_t)blockIdx.x * kAttWarps + (threadIdx.x >> 5);
  if (gh >= heads_total) return;
  const float rsd = 1.0f / sqrtf((float)D);
  const float *kh = K + gh * (size_t)S * D;
  const float *vh = V + gh * (size_t)S * D;
  float q[D], go[D];
  load_row<D, VEC>(Q + gh * D, q);
  load_row<D, VEC>(dout + gh * D, go);
  float a[NS], da[NS];
  head_softmax<D, NS, VEC>(S, kh, q, rsd, a);
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int s = lane + 32 * i;
    da[i] = 0.f;
    if (s < S) {
      float v[D];
      load_row<D, VEC>(vh + (size_t)s * D, v);
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) acc = fmaf(go[d], v[d], acc);
      da[i] = acc;
      float gv[D];
#pragma unroll
      for (int d = 0; d < D; ++d) gv[d] = a[i] * go[d];
      store_row<D, VEC>(dV + gh * (size_t)S * D + (size_t)s * D, gv);
    }
    dot = fmaf(a[i], da[i], dot);
  }
  dot = warp_sum(dot);
  float gq[D];
#pragma unroll
  for (int d = 0; d < D; ++d) gq[d] = 0.f;
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const int s = lane + 32 * i;
    if (s < S) {
      const float dl = a[i] * (da[i] - dot) * rsd;
      float k[D], gk[D];
      load_row<D, VEC>(kh + (size_t)s * D, k);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        gq[d] = fmaf(dl, k[d], gq[d]);
        gk[d] = dl * q[d];
      }
      store_row<D, VEC>(dK + gh * (size_t)S * D + (size_t)s * D, gk);
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) gq[d] = warp_sum(gq[d]);
  if (lane == 0) store_row<D, VEC>(dQ + gh * D, gq);
}

// ---------------------------------------------------------------------------------------------------------------
// InnerAttentionLayer (attention_layer.py:48-78): attention ACROSS THE 5 HEADS of one point.  Per row: Q, K, V (5, kd)
// (a raw reshape of the (5 kd) Dense outputs), weights = softmax_j(Q_i . K_j / sqrt(kd)) (5 x 5), out_i = sum_j w_ij V_j.
// A thread per row; rows are 5 kd floats (kd multiple of 4: 128-bit loads).  Pure streaming of 3 reads + 1 write.
template <int KD>
__global__ void __launch_bounds__(128)
inner_attention_kernel(size_t rows, const float *__restrict__ Q, const float *__restrict__ K, const float *__restrict__ V,
                       float *__restrict__ out) {
  constexpr int H = 5, W = H * KD;
  const float rsd = 1.0f / sqrtf((float)KD);
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
    const float *q = Q + r * W, *k = K + r * W, *v = V + r * W;
    float w[H][H];
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int j = 0; j < H; ++j) w[i][j] = 0.f;
#pragma unroll
    for (int d = 0; d < KD; ++d) {
      float qd[H], kd_[H];
#pragma unroll
      for (int i = 0; i < H; ++i) { qd[i] = __ldg(q + i * KD + d); kd_[i] = __ldg(k + i * KD + d); }
#pragma unroll
      for (int i = 0; i < H; ++i)
#pragma unroll
        for (int j = 0; j < H; ++j) w[i][j] = fmaf(qd[i], kd_[j], w[i][j]);
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      float mx = -INFINITY, sum = 0.f;
#pragma unroll
      for (int j = 0; j < H; ++j) { w[i][j] *= rsd; mx = fmaxf(mx, w[i][j]); }
#pragma unroll
      for (int j = 0; j < H; ++j) { w[i][j] = expf(w[i][j] - mx); sum += w[i][j]; }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int j = 0; j < H; ++j) w[i][j] *= inv;
    }
    float *o = out + r * W;
#pragma unroll
    for (int d = 0; d < KD; ++d) {
      float vd[H];
#pragma unroll
      for (int j = 0; j < H; ++j) vd[j] = __ldg(v + j * KD + d);
#pragma unroll
      for (int i = 0; i < H; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < H; ++j) acc = fmaf(w[i][j], vd[j], acc);
        o[i * KD + d] = acc;
      }
    }
  }
}

template <int KD>
int launch_inner(size_t rows, const float *Q, const float *K, const float *V, float *out, cudaStream_t st) {
  size_t blocks = (rows + 127) / 128;
  const size_t cap = (size_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  inner_attention_kernel<KD><<<(unsigned)blocks, 128, 0, st>>>(rows, Q, K, V, out);
  PC_RETURN_LAUNCH_STATUS();
}

template <int D, int NS, bool VEC>
int launch_fwd(size_t heads, int S, int H, const float *Q, const float *K, const float *V, float *out,
               cudaStream_t st) {
  size_t blocks = (heads + kAttWarps - 1) / kAttWarps;
  const size_t cap = (size_t)num_sms() * 16;  // persistent-ish: 8 resident CTAs per SM, two waves
  if (blocks > cap) blocks = cap;
  attention_fwd_kernel<D, NS, VEC><<<(unsigned)blocks, kAttWarps * 32, 0, st>>>(heads, S, H, Q, K, V, out);
  PC_RETURN_LAUNCH_STATUS();
}
template <int D, int NS, bool VEC>
int launch_bwd(size_t heads, int S, int H, const float *Q, const float *K, const float *V, const float *dout,
               float *dQ, float *dK, float *dV, cudaStream_t st) {
  const size_t blocks = (heads + kAttWarps - 1) / kAttWarps;
  attention_bwd_kernel<D, NS, VEC><<<(unsigned)blocks, kAttWarps * 32, 0, st>>>(heads, S, H, Q, K, V, dout, dQ, dK, dV);
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace
}  // namespace pc

namespace {
int check_att(int G, int S, int H, int D) {
  if (G < 0 || S <= 0 || H <= 0 || D <= 0) return PC_ERR_INVALID_ARGUMENT;
  if (!(D == 1 || D == 2 || D == 4 || D == 8 || D == 16 || D == 32 || D == 64) || S > 128) return PC_ERR_UNSUPPORTED;
  if ((size_t)G * H > 0x7fffffffu * (size_t)pc::kAttWarps) return PC_ERR_UNSUPPORTED;
  return PC_OK;
}
}  // namespace

extern "C" int pc_attention_fwd(int G, int S, int H, int D, const float *Q, const float *K, const float *V,
                                float *out, pc_stream_t stream) {
  int rc = check_att(G, S, H, D);
  if (rc) return rc;
  if (G == 0) return PC_OK;
  if (!Q || !K || !V || !out) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t heads = (size_t)G * H;
  const bool vec = (D % 4 == 0) && pc::aligned16(Q) && pc::aligned16(K) && pc::aligned16(V) && pc::aligned16(out);
  if (D == 4 && S == 32 && vec) return pc::launch_fwd_lanehead(heads, Q, K, V, out, st);   // lane = head, bulk-copy staging
#define PC_FWD(DD, VV)                                                                        \
  do {                                                                                        \
    if (S <= 32) return pc::launch_fwd<DD, 1, VV>(heads, S, H, Q, K, V, out, st);             \
    if (S <= 64) return pc::launch_fwd<DD, 2, VV>(heads, S, H, Q, K, V, out, st);             \
    return pc::launch_fwd<DD, 4, VV>(heads, S, H, Q, K, V, out, st);                          \
  } while (0)
  switch (D) {
    case 1: PC_FWD(1, false);
    case 2: PC_FWD(2, false);
    case 4: if (vec) PC_FWD(4, true); else PC_FWD(4, false);
    case 8: if (vec) PC_FWD(8, true); else PC_FWD(8, false);
    case 16: if (vec) PC_FWD(16, true); else PC_FWD(16, false);
    case 32: if (vec) PC_FWD(32, true); else PC_FWD(32, false);   // key_dim of the experimental layers (attention_layer.py:128-210)
    default: if (vec) PC_FWD(64, true); else PC_FWD(64, false);
  }
#undef PC_FWD
}

extern "C" int pc_attention_bwd(int G, int S, int H, int D, const float *Q, const float *K, const float *V,
                                const float *dout, float *dQ, float *dK, float *dV, pc_stream_t stream) {
  int rc = check_att(G, S, H, D);
  if (rc) return rc;
  if (G == 0) return PC_OK;
  if (!Q || !K || !V || !dout || !dQ || !dK || !dV) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t heads = (size_t)G * H;
  const bool vec = (D % 4 == 0) && pc::aligned16(Q) && pc::aligned16(K) && pc::aligned16(V) && pc::aligned16(dout) &&
                   pc::aligned16(dQ) && pc::aligned16(dK) && pc::aligned16(dV);
#define PC_BWD(DD, VV)                                                                                  \
  do {                                                                                                  \
    if (S <= 32) return pc::launch_bwd<DD, 1, VV>(heads, S, H, Q, K, V, dout, dQ, dK, dV, st);          \
    if (S <= 64) return pc::launch_bwd<DD, 2, VV>(heads, S, H, Q, K, V, dout, dQ, dK, dV, st);          \
    return pc::launch_bwd<DD, 4, VV>(heads, S, H, Q, K, V, dout, dQ, dK, dV, st);                       \
  } while (0)
  switch (D) {
    case 1: PC_BWD(1, false);
    case 2: PC_BWD(2, false);
    case 4: if (vec) PC_BWD(4, true); else PC_BWD(4, false);
    case 8: if (vec) PC_BWD(8, true); else PC_BWD(8, false);
    case 16: if (vec) PC_BWD(16, true); else PC_BWD(16, false);
    case 32: if (vec) PC_BWD(32, true); else PC_BWD(32, false);
    default: if (vec) PC_BWD(64, true); else PC_BWD(64, false);
  }
#undef PC_BWD
}

extern "C" int pc_inner_attention_fwd(size_t rows, int key_dim, const float *Q, const float *K, const float *V, float *out,
                                      pc_stream_t stream) {
  if (key_dim <= 0) return PC_ERR_INVALID_ARGUMENT;
  if (rows == 0) return PC_OK;
  if (!Q || !K || !V || !out) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  switch (key_dim) {
    case 4: return pc::launch_inner<4>(rows, Q, K, V, out, st);
    case 8: return pc::launch_inner<8>(rows, Q, K, V, out, st);
    case 16: return pc::launch_inner<16>(rows, Q, K, V, out, st);
    case 32: return pc::launch_inner<32>(rows, Q, K, V, out, st);
    case 64: return pc::launch_inner<64>(rows, Q, K, V, out, st);
    default: return PC_ERR_UNSUPPORTED;
  }
}
