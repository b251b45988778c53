// Dense layers of the path on the tcgen05 tensor cores: Y = epilogue(X . W + b) in 3xTF32 split precision.
//
// What it replaces.  Everything between the geometry ops in the reference is a 1x1 convolution / Dense layer over
// (B*m*nsample) rows -- tf_util.conv2d with kernel [1,1] + batch norm + ReLU three times in pointnet_sa_module
// (utils/pointnet_util.py:119-131, tf_util.py:120-186,512-530) followed by a max over nsample (:134-135), and the three
// Dense projections of AttentionLayer (attention_layer.py:24-34) with their gradients under minimize
// (attention_points/train.py:337-339).  TensorFlow runs them as cuDNN / cuBLAS calls with every intermediate in HBM.
//
// One engine, three uses:
//   * pc_dense_fwd       Y (rows, N) = act(X (rows, K) . W (K, N) + bias)           shared-MLP layer (inference: batch norm
//                        folded into W / bias by the caller), Dense K | V projection of a training forward
//   * pc_dense_pool_fwd  Y (rows/32, N) = max over each group of 32 rows of the above   last MLP layer + the max-pool of
//                        pointnet_sa_module, pooled straight out of TMEM: the (rows, N) tensor is never written
//   * pc_dense_bwd_input dX (rows, K) = dY (rows, N) . W^T                              same kernel, weight image built
//                        from W with swapped strides
// and pc_dense_bwd_weight dW (K, N) = X^T . dY, db = column sums of dY             a second kernel whose producers
//                        transpose both operands on the way into shared memory (split over row ranges, partials
//                        summed in a fixed order: deterministic).
//
// Precision: fp32 in, fp32 out.  Plain TF32 (10-bit mantissa) misses the path's 1e-5 bound, so every operand is split
// x = hi + lo (both TF32) and each product is issued as hi*hi + hi*lo + lo*hi into an fp32 TMEM accumulator -- the
// same scheme, descriptors and swizzled layout as attention_layer_wide.cu (measured there: 2e-6 of the output scale).
//
// Schedule (one persistent CTA per SM, 288 threads): work item = (tile of 128 rows, chunk of <= 256 output columns).
//   warps 0-3  producers: per K block of 32 input columns, X rows global -> registers (two blocks ahead) -> hi / lo ->
//              128-byte-swizzled A stage; one thread starts the bulk copy (cp.async.bulk + mbarrier complete_tx) of the
//              pre-split, pre-swizzled B block of the weight image
//   warp  8    one thread issues 12 tcgen05.mma (M128 N<=256 K8 kind::tf32) per K block; tcgen05.commit frees the stage
//   warps 4-7  epilogue: tcgen05.ld 32 columns at a time, bias (+ ReLU), then either a transposed pass through a 4 KB
//              shared-memory patch so that the global stores are row-contiguous 128-bit transactions, or the warp-wide
//              maximum per column (an epilogue warp's 32 TMEM lanes ARE the 32 samples of one neighbourhood)
// Two TMEM accumulators (2 x 256 columns): the next item's MMAs run under this item's epilogue.
#include <math.h>
#include "common.cuh"

namespace pc {
namespace {

constexpr int kRows = 128;                 // M: rows per tile
constexpr int kMaxNc = 256;                // N per item
constexpr int kKB = 32;                    // tf32 elements per K block (one 128-byte swizzled row)
constexpr int kSBO = 1024;                 // bytes between 8-row groups
constexpr int kABlock = kRows * 128;       // 16 KB: A K-block (hi or lo)
constexpr int kBBlock = kMaxNc * 128;      // 32 KB: B K-block (hi or lo) at full width
constexpr int kStage = 2 * kABlock + 2 * kBBlock;   // 96 KB: A_hi | A_lo | B_hi | B_lo
constexpr int kThreads = 288;
#ifndef PCOPS_DENSE_WARP_ARRIVE
#define PCOPS_DENSE_WARP_ARRIVE 1   // 1: one mbarrier arrival per warp (after __syncwarp) instead of one per thread
#endif
#ifndef PCOPS_DENSE_PAIR
#define PCOPS_DENSE_PAIR 1
#endif
#ifndef PCOPS_DENSE_PWARPS
#define PCOPS_DENSE_PWARPS 8
#endif
#ifndef PCOPS_DENSE_AHEAD
#define PCOPS_DENSE_AHEAD 2
#endif
constexpr int kFwdPW = PCOPS_DENSE_PWARPS;          // producer warps of the forward kernel (4 or 8)
constexpr int kFwdAhead = PCOPS_DENSE_AHEAD;        // K blocks a producer thread holds in registers ahead of the ring
constexpr int kFwdThreads = (kFwdPW + 5) * 32;      // producers, four epilogue warps, the MMA issuer
constexpr int kF4 = 1024 / (kFwdPW * 32);           // float4s of a 128 x 32 block per producer thread
constexpr int kMaxN = 1024;                // bias staged in shared memory
constexpr int kPatch = 32 * 36;            // floats per epilogue warp: 32 rows x 32 columns, rows padded to 36
constexpr int kFixedBytes = 24 * 1024;     // bias, patches, barriers of the forward kernel in front of its operand ring

__host__ __device__ inline int block_offset(int row, int k) {  // byte offset of (row, k) inside one K block, k < 32
  return (row >> 3) * kSBO + (row & 7) * 128 + (((k >> 2) ^ (row & 7)) << 4) + (k & 3) * 4;
}

__device__ __forceinline__ float tf32_rna(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Weight image: for column chunk j and K block kb a 64 KB slot [B_hi | B_lo] (each nc x 128 bytes used) laid out as
// shared memory wants it.  Row n of a block is output column 256 j + n; element (n, kl) = W[(32 kb + kl) * sk + (256 j +
// n) * sn], zero beyond K / N.  (sk, sn) = (ldw, 1) for Y = X W with W stored [in][out]; (1, ldw) for dX = dY W^T.
__global__ void dense_prep_kernel(int K, int N, size_t sk, size_t sn, const float *__restrict__ w,
                                  unsigned char *__restrict__ image) {
  const int nkb = (K + kKB - 1) / kKB, nchunk = (N + kMaxNc - 1) / kMaxNc;
  const size_t total = (size_t)nchunk * nkb * kMaxNc * kKB;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int kl = (int)(t % kKB);
    const int n = (int)((t / kKB) % kMaxNc);
    const int kb = (int)((t / ((size_t)kKB * kMaxNc)) % nkb);
    const int j = (int)(t / ((size_t)kKB * kMaxNc * nkb));
    const int k = kb * kKB + kl, col = j * kMaxNc + n;
    const float v = (k < K && col < N) ? w[(size_t)k * sk + (size_t)col * sn] : 0.f;
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    unsigned char *blk = image + ((size_t)j * nkb + kb) * (2 * kBBlock);
    *reinterpret_cast<float *>(blk + block_offset(n, kl)) = hi;
    *reinterpret_cast<float *>(blk + kBBlock + block_offset(n, kl)) = lo;
  }
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO = 1024
  const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (1u << 16);
  const uint32_t hi = (uint32_t)(kSBO >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

#define PCG_TMEM_LD32(addr, v)                                                                                            \
  asm volatile(                                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                           \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),        \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),            \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),           \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                          \
      : "r"(addr))

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // bare try_wait loop: a suspend-time hint on try_wait and a nanosleep back-off were both measured, no difference
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // never hang the device on a lost arrival
  } while (!ok);
}

// Position of a warp role in the CTA's item sequence (item = blockIdx.x + w * gridDim.x = tile * nchunk + chunk) and in
// the K blocks of the item, advanced by additions only: the roles used to recompute (w, kb, tile, chunk, stage, phase)
// from a flat counter with six runtime integer divisions per block -- ~0.6 us of dependent latency per 128-row tile on
// the single-thread MMA issuer and every producer thread, more than the MMAs themselves (stub measurements).
struct ItemCursor {
  int tile, chunk, kb;
  __device__ __forceinline__ void next_block(int nkb, int dt, int dc, int nchunk) {
    if (++kb == nkb) {
      kb = 0;
      tile += dt;
      chunk += dc;
      if (chunk >= nchunk) { chunk -= nchunk; ++tile; }
    }
  }
  __device__ __forceinline__ void next_item(int dt, int dc, int nchunk) {
    tile += dt;
    chunk += dc;
    if (chunk >= nchunk) { chunk -= nchunk; ++tile; }
  }
};
struct RingCursor {   // slot index and mbarrier phase of a ring of n slots
  int s;
  uint32_t ph;
  __device__ __forceinline__ void next(int n) {
    if (++s == n) { s = 0; ph ^= 1u; }
  }
};

// kPool: 0 = store Y (rows, N) to out; 1 = store the maximum over each group of 32 rows, (rows / 32, N), to pooled;
// 2 = both (the attention-and-pooling module needs the activations for the attention layer and their maximum).
template <int kPool>
__global__ void __launch_bounds__(kFwdThreads, 1)
dense_tf32_kernel(size_t rows, int K, size_t ldx, int N, int sw, int nsplit, int nst, int nacc, int pair, size_t ldo, size_t ldp, int relu, int vec_x,
                  int vec_o, const float *__restrict__ x, const unsigned char *__restrict__ image, const float *__restrict__ bias,
                  float *__restrict__ out, float *__restrict__ pooled) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // fixed part first (kFixedBytes), then the operand ring: nst stages of [A_hi 16 KB | A_lo 16 KB | B_hi | B_lo], B parts
  // sw x 128 bytes each -- 2 stages at sw = 256, 3 at 128, 4 below: a narrow layer's block is handed over through the
  // same mbarrier / tcgen05.commit round trips as a wide one, so it needs more blocks in flight to hide them
  float *s_bias = reinterpret_cast<float *>(smem);                          // kMaxN
  float *s_patch = s_bias + kMaxN;                                          // 4 x kPatch
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_patch + 4 * kPatch);     // full[4], empty[4], t_full[8], t_empty[8]
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 24);
  unsigned char *stage_buf = smem + kFixedBytes;
  const int b_half = sw * 128, stage_bytes = 2 * kABlock + 2 * b_half;
  const int aw = pair ? 2 * sw : sw;                                        // TMEM columns of one accumulator

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar);
  const uint32_t full0 = bar0, empty0 = bar0 + 32;                          // full[s] = full0 + 8 s, empty[s] = empty0 + 8 s
  // nacc TMEM accumulators of sw columns each (2 at sw = 256 ... 8 at sw <= 64): an item's accumulator travels MMA ->
  // commit -> epilogue -> arrive -> MMA, and with two accumulators that round trip (about 4 us) capped a CTA at one item
  // per ~2 us whatever the item's size
  const uint32_t t_full0 = bar0 + 64, t_empty0 = bar0 + 128;                // t_full[a] = t_full0 + 8 a, ...

  if (warp == kFwdPW + 4) {  // the whole TMEM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(s_tmem)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s), "r"((uint32_t)(PCOPS_DENSE_WARP_ARRIVE ? kFwdPW : kFwdPW * 32)));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8 * s), "r"(1u));
    }
    for (int s = 0; s < 8; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_full0 + 8 * s), "r"(1u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_empty0 + 8 * s), "r"(PCOPS_DENSE_WARP_ARRIVE ? 4u : 128u));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = tid; i < N; i += kFwdThreads) s_bias[i] = bias ? __ldg(bias + i) : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;

  // item = (tile of 128 rows, unit of sw output columns); sw = 256 unless the launch would leave SMs without an item
  // (few rows): then 128 / 64 / 32, a contiguous slice of the 256-column image slot (8-row groups are 1 KB apart)
  const int nkb = (K + kKB - 1) / kKB, nchunk = (N + sw - 1) / sw;
  const int ntiles = (int)((rows + kRows - 1) / kRows);
  // item = tile * nchunk + u: a tile's units run on neighbouring CTAs
  const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage_buf);
  const int dt = (int)gridDim.x / nchunk, dc = (int)gridDim.x - dt * nchunk;     // item += gridDim.x in (tile, chunk) steps
  const int tile0 = (int)blockIdx.x / nchunk, chunk0 = (int)blockIdx.x - tile0 * nchunk;

  if (warp < kFwdPW) {
    // ---------------------------------------------------------------- producers
    // two cursors: the loads run kFwdAhead blocks ahead of the conversion into the ring
    ItemCursor fc{tile0, chunk0, 0}, pc_{tile0, chunk0, 0};
    RingCursor ring{0, 1u};                          // producers wait for "empty": parity starts flipped
    auto fetch = [&](float4 (&buf)[kF4]) {   // thread t takes float4 t + kFwdPW * 32 * i -> row (i4 >> 3), quad i4 & 7
      const int kb = fc.kb;
      const bool in_range = fc.tile < ntiles;
      const size_t row0 = (size_t)fc.tile * kRows;
      fc.next_block(nkb, dt, dc, nchunk);
      if (in_range && vec_x && row0 + kRows <= rows && kb * kKB + kKB <= K) {   // whole block in range: no per-element tests
        const float4 *src = reinterpret_cast<const float4 *>(x + (row0 + (tid >> 3)) * ldx + kb * kKB) + (tid & 7);
#pragma unroll
        for (int i = 0; i < kF4; ++i) buf[i] = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(src) + (size_t)(kFwdPW * 4 * i) * ldx));
        return;
      }
#pragma unroll
      for (int i = 0; i < kF4; ++i) {
        const int i4 = tid + kFwdPW * 32 * i, row = i4 >> 3, kq = i4 & 7;
        const int k0 = kb * kKB + kq * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_range && row0 + row < rows && k0 < K) {
          const float *src = x + (row0 + row) * ldx + k0;
          if (vec_x && k0 + 3 < K) {
            v = __ldg(reinterpret_cast<const float4 *>(src));
          } else {  // row stride or base not 16-byte aligned (K = 9, 67, 131, 259: the [xyz | features] inputs), or K tail
            v.x = __ldg(src);
            if (k0 + 1 < K) v.y = __ldg(src + 1);
            if (k0 + 2 < K) v.z = __ldg(src + 2);
            if (k0 + 3 < K) v.w = __ldg(src + 3);
          }
        }
        buf[i] = v;
      }
    };
    auto produce = [&](const float4 (&buf)[kF4]) {
      const int kb = pc_.kb;
      const int col0 = pc_.chunk * sw;
      const int nc = min(sw, N - col0);
      const int s = ring.s;
      mbar_wait(empty0 + 8 * s, ring.ph);   // the MMAs that read this stage nst blocks ago have completed
      pc_.next_block(nkb, dt, dc, nchunk);
      ring.next(nst);
      unsigned char *st = stage_buf + s * stage_bytes;
      if (tid == 0) {  // B block: hi and lo halves of the image slot, nc rows of 128 bytes each
        const unsigned char *src = image + ((size_t)(col0 / kMaxNc) * nkb + kb) * (2 * kBBlock) + (size_t)(col0 % kMaxNc) * 128;
        const uint32_t dst = stage_s + s * stage_bytes + 2 * kABlock;
        const uint32_t bytes = (uint32_t)((nc + 15) & ~15) * 128u;   // the MMA's N is a multiple of 16: zero rows beyond nc
        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s), "r"(2u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"(bytes), "r"(full0 + 8 * s)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + b_half),
                     "l"(src + kBBlock), "r"(bytes), "r"(full0 + 8 * s)
                     : "memory");
      }
#pragma unroll
      for (int i = 0; i < kF4; ++i) {
        const int i4 = tid + kFwdPW * 32 * i, row = i4 >> 3, kq = i4 & 7;
        const float4 v = buf[i];
        float4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
        const int off = block_offset(row, kq * 4);
        *reinterpret_cast<float4 *>(st + off) = h;
        *reinterpret_cast<float4 *>(st + kABlock + off) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's stores, visible to the tensor core's proxy
#if PCOPS_DENSE_WARP_ARRIVE
      __syncwarp();
      if (lane == 0)
#endif
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + 8 * s) : "memory");
    };
    // kFwdAhead K blocks ahead, in registers: kFwdPW * 32 threads x kFwdAhead x kF4 float4 = the bytes in flight per SM
    // (32 KB with 4 warps x 2 x 8; 64 KB with 8 warps x 4 x 4).  The loop is unrolled over the buffers so that every
    // buffer index is a compile-time constant (a dynamically indexed buffer array lives in local memory).
    float4 buf[kFwdAhead][kF4];
#pragma unroll
    for (int u = 0; u < kFwdAhead; ++u) fetch(buf[u]);
    while (pc_.tile < ntiles) {
#pragma unroll
      for (int u = 0; u < kFwdAhead; ++u) {
        if (pc_.tile < ntiles) {
          produce(buf[u]);
          fetch(buf[u]);
        }
      }
    }
  } else if (warp == kFwdPW + 4) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      ItemCursor c{tile0, chunk0, 0};
      RingCursor ring{0, 0u}, accr{0, 1u};             // waits for "full" stages; waits for "empty" accumulators
      for (; c.tile < ntiles; c.next_item(dt, dc, nchunk)) {
        const int a = accr.s;
        const int nc = min(sw, N - c.chunk * sw);
        // instruction descriptor: D = F32, A = B = TF32, both K-major, N = nc rounded up to 16, M = 128
        const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kRows >> 4) << 24);
        const uint32_t idesc = idesc0 | ((uint32_t)(((nc + 15) & ~15) >> 3) << 17);
        const uint32_t idesc2 = idesc0 | ((uint32_t)((2 * sw) >> 3) << 17);   // paired: N = [W_hi | W_lo], 2 sw columns
        mbar_wait(t_empty0 + 8 * a, accr.ph);      // the epilogue has drained this accumulator
        accr.next(nacc);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t acc = 0;
        const uint32_t d_tmem = tmem + a * aw;
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = ring.s;
          mbar_wait(full0 + 8 * s, ring.ph);
          ring.next(nst);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = stage_s + s * stage_bytes, a_lo = a_hi + kABlock, b_hi = a_hi + 2 * kABlock, b_lo = b_hi + b_half;
          // descriptors of the four operand parts; a K step of 8 (32 bytes) is +2 in the 16-byte address field
          const uint64_t d_ah = smem_desc(a_hi), d_al = smem_desc(a_lo), d_bh = smem_desc(b_hi), d_bl = smem_desc(b_lo);
          for (int split = 0; split < nsplit; ++split) {  // X_hi W_hi, X_hi W_lo, X_lo W_hi [, X_lo W_lo: fp32-grade]
            if (pair) {
              // An MMA of this shape costs ~85 cycles whatever its N (the 128 x 32-byte A operand comes out of shared
              // memory each time), so the W_hi and W_lo halves of the stage -- contiguous, sw rows each -- go in as ONE
              // 2 sw-column operand: X_hi [W_hi | W_lo] -> accumulator columns [0, sw) and [sw, 2 sw), then X_lo W_hi
              // (X_lo [W_hi | W_lo] with four products) on top; the epilogue adds the two column blocks.
              if (split >= 2) break;
              const uint64_t da = split ? d_al : d_ah;
              const uint32_t id = (split == 0 || nsplit == 4) ? idesc2 : idesc;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                mma_tf32(d_tmem, da + 2u * kk, d_bh + 2u * kk, id, acc | (uint32_t)split);
                acc = 1;
              }
              continue;
            }
            const uint64_t da = (split >= 2) ? d_al : d_ah, db = (split & 1) ? d_bl : d_bh;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              mma_tf32(d_tmem, da + 2u * kk, db + 2u * kk, idesc, acc);
              acc = 1;
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * s) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t_full0 + 8 * a) : "memory");
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue (four warps: TMEM lane quarters warp & 3)
    const int qtr = warp & 3;
    float *patch = s_patch + qtr * kPatch;
    ItemCursor c{tile0, chunk0, 0};
    RingCursor accr{0, 0u};
    for (; c.tile < ntiles; c.next_item(dt, dc, nchunk)) {
      const int a = accr.s, col0 = c.chunk * sw;
      const int nc = min(sw, N - col0);
      const size_t row0 = (size_t)c.tile * kRows + qtr * 32;    // this warp's 32 rows
      mbar_wait(t_full0 + 8 * a, accr.ph);
      accr.next(nacc);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(qtr * 32) << 16) + a * aw;
      for (int c0 = 0; c0 < nc; c0 += 32) {
        uint32_t v[32];
        PCG_TMEM_LD32(taddr + c0, v);
        if (pair) {   // second column block: the X_hi W_lo (+ X_lo W_lo) part
          uint32_t v2[32];
          PCG_TMEM_LD32(taddr + sw + c0, v2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = __float_as_uint(__uint_as_float(v[t]) + __uint_as_float(v2[t]));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        const float *cb = s_bias + col0 + c0;
        float y[32];
        if (relu) {   // one uniform branch instead of a predicate per element (NaN handling as before: fmaxf(NaN, 0) = 0)
#pragma unroll
          for (int t = 0; t < 32; ++t) y[t] = fmaxf(__uint_as_float(v[t]) + cb[t], 0.f);
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) y[t] = __uint_as_float(v[t]) + cb[t];
        }
        if (kPool) {
          // the warp's 32 lanes are the 32 samples of neighbourhood row0 / 32: maximum per column, lane t keeps column t
          const bool live = row0 + lane < rows;
          float mine = 0.f;
          if (relu) {  // values >= 0 order like their bit patterns: one redux per column
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const int m = __reduce_max_sync(PC_FULL_MASK, live ? __float_as_int(fmaxf(y[t], 0.f)) : 0);
              if (lane == t) mine = __int_as_float(m);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              float m = live ? y[t] : -INFINITY;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(PC_FULL_MASK, m, o));
              if (lane == t) mine = m;
            }
          }
          if (row0 < rows && c0 + lane < nc) pooled[(row0 >> 5) * ldp + (size_t)col0 + c0 + lane] = mine;
        }
        if (kPool != 1) {
          // transpose through the patch: thread = row writes its 32 columns, then 8 lanes cover one row's 128 bytes
          __syncwarp();
#pragma unroll
          for (int t = 0; t < 8; ++t)
            *reinterpret_cast<float4 *>(patch + lane * 36 + 4 * t) = make_float4(y[4 * t], y[4 * t + 1], y[4 * t + 2], y[4 * t + 3]);
          __syncwarp();
          const int rr = lane >> 3, cq = (lane & 7) * 4;
          if (vec_o && row0 + 32 <= rows && c0 + 32 <= nc) {   // whole 32 x 32 patch in range
            float *dst = out + (row0 + rr) * ldo + (size_t)col0 + c0 + cq;
#pragma unroll
            for (int t = 0; t < 8; ++t)
              *reinterpret_cast<float4 *>(dst + (size_t)(4 * t) * ldo) = *reinterpret_cast<const float4 *>(patch + (4 * t + rr) * 36 + cq);
            continue;
          }
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int r = 4 * t + rr;
            const float4 val = *reinterpret_cast<const float4 *>(patch + r * 36 + cq);
            if (row0 + r < rows && c0 + cq < nc) {
              float *dst = out + (row0 + r) * ldo + (size_t)col0 + c0 + cq;
              if (vec_o && c0 + cq + 3 < nc) {
                *reinterpret_cast<float4 *>(dst) = val;
              } else {
                dst[0] = val.x;
                if (c0 + cq + 1 < nc) dst[1] = val.y;
                if (c0 + cq + 2 < nc) dst[2] = val.z;
                if (c0 + cq + 3 < nc) dst[3] = val.w;
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // TMEM reads done before the accumulator is handed back
#if PCOPS_DENSE_WARP_ARRIVE
      __syncwarp();
      if (lane == 0)
#endif
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t_empty0 + 8 * a) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kFwdPW + 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient: dW (K, N) = X^T (K, rows) . dY (rows, N), db (N) = column sums of dY.
// The reduction dimension is the ROW index, i.e. both operands are "MN-major" in UMMA terms: a block of 32 rows x 128
// (or nc) columns of X / dY, row-major as it lies in HBM, IS a canonical MN-major operand (layout below).  So the
// producers load coalesced float4s exactly as
// the forward kernel does, split them into TF32 hi / lo and store 128 bits at a time; the instruction descriptor sets
// the a_major / b_major bits.  (A first version transposed both operands in the producers with lane = row: 32
// different cache lines per load instruction and 192 scalar stores per thread and block -- 887 us for 524288 x 64 x 64
// under ncu; this one: see DESIGN.md.)  Work item = (tile of 128 input channels, chunk of <= 256 output channels,
// split of the rows); each item leaves its partial product in the workspace, a second kernel adds the partials in
// ascending split order (deterministic; no float atomics).  db is accumulated by the same producers (items of channel
// tile 0 only).
// For 32-bit MN-major operands the tensor core accepts ONE shared-memory layout (CUTLASS sm100_common.inl:92: "for
// mn-major tf32 operands, SW128_32B is the only available smem layout"; with plain SWIZZLE_128B the MMAs ran and produced
// zeros): atoms of 4 k-rows x 128 bytes, the row's four 32-byte chunks XOR-ed by (row & 3) -- Swizzle<2,5,2>, descriptor
// layout type 1.  Atoms of one 4-row group are 512 bytes apart (LBO), groups follow each other (SBO); an MMA of K = 8
// spans two groups.
__device__ __forceinline__ int mn_offset(int krow, int quad, int atoms_per_group) {  // byte offset of float4 (k-row, column quad)
  return ((krow >> 2) * atoms_per_group + (quad >> 3)) * 512 + (krow & 3) * 128 + (((((quad & 7) >> 1) ^ (krow & 3)) << 5) | ((quad & 1) << 4));
}
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {  // MN-major, SWIZZLE_128B_BASE32B
  const uint32_t lo = ((saddr >> 4) & 0x3fffu) | ((lbo >> 4) << 16);
  const uint32_t hi = (sbo >> 4) | (1u << 14) | (1u << 29);
  return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(kThreads, 1)
dense_bwd_weight_kernel(size_t rows, int K, size_t ldx, int N, size_t ldy, int splits, size_t rows_per_split, int vec_x,
                        int vec_y, const float *__restrict__ x, const float *__restrict__ dy, float *__restrict__ part_w,
                        float *__restrict__ part_b) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *stage_buf = smem;
  float *s_patch = reinterpret_cast<float *>(smem + 2 * kStage);
  float *s_red = s_patch + 4 * kPatch;                                      // 64 float4: the producers' db exchange
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_red + 256);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar);
  const uint32_t full[2] = {bar0, bar0 + 8}, empty[2] = {bar0 + 16, bar0 + 24};
  const uint32_t t_full[2] = {bar0 + 32, bar0 + 40}, t_empty[2] = {bar0 + 48, bar0 + 56};
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(s_tmem)),
                 "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full[s]), "r"(128u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty[s]), "r"(1u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_full[s]), "r"(1u));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_empty[s]), "r"(128u));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;

  const int nmt = (K + kRows - 1) / kRows, nchunk = (N + kMaxNc - 1) / kMaxNc;
  const int nitems = nmt * nchunk * splits;            // item = (mt * nchunk + j) * splits + s
  const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage_buf);
  auto item_rows = [&](int item, size_t &r_lo, int &nblk) {
    const int sp = item % splits;
    r_lo = (size_t)sp * rows_per_split;
    const size_t r_hi = r_lo + rows_per_split < rows ? r_lo + rows_per_split : rows;
    nblk = r_lo < r_hi ? (int)((r_hi - r_lo + kKB - 1) / kKB) : 0;
  };

  if (warp < 4) {
    // ---------------------------------------------------------------- producers: coalesced loads, 128-bit stores
    // thread t takes float4 t + 128 i of the block: X (32 rows x 32 quads): row = i4 >> 5, quad = i4 & 31 (i < 8);
    // dY (32 rows x 64 quads): row = i4 >> 6, quad = i4 & 63 (i < 16).  The loads of block b + 1 are issued right after
    // block b has been stored, so they fly while the tensor core works on b.
    int it = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int mj = item / splits, mt = mj / nchunk, j = mj - mt * nchunk, sp = item - mj * splits;
      const int nc = min(kMaxNc, N - j * kMaxNc);
      size_t r_lo;
      int nblk;
      item_rows(item, r_lo, nblk);
      const size_t r_end = r_lo + rows_per_split < rows ? r_lo + rows_per_split : rows;
      float4 xa[8], yb[16];
      auto fetch = [&](int blk) {
        const size_t row0 = r_lo + (size_t)blk * kKB;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int i4 = tid + 128 * i, r = i4 >> 5, c0 = mt * kRows + 4 * (i4 & 31);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (blk < nblk && row0 + r < r_end && c0 < K) {
            const float *src = x + (row0 + r) * ldx + c0;
            if (vec_x && c0 + 3 < K) v = __ldg(reinterpret_cast<const float4 *>(src));
            else {
              v.x = __ldg(src);
              if (c0 + 1 < K) v.y = __ldg(src + 1);
              if (c0 + 2 < K) v.z = __ldg(src + 2);
              if (c0 + 3 < K) v.w = __ldg(src + 3);
            }
          }
          xa[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int i4 = tid + 128 * i, r = i4 >> 6, cl = 4 * (i4 & 63);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (blk < nblk && row0 + r < r_end && cl < nc) {
            const float *src = dy + (row0 + r) * ldy + j * kMaxNc + cl;
            if (vec_y && cl + 3 < nc) v = __ldg(reinterpret_cast<const float4 *>(src));
            else {
              v.x = __ldg(src);
              if (cl + 1 < nc) v.y = __ldg(src + 1);
              if (cl + 2 < nc) v.z = __ldg(src + 2);
              if (cl + 3 < nc) v.w = __ldg(src + 3);
            }
          }
          yb[i] = v;
        }
      };
      // column sums of dY: thread t always holds the same column quad (t & 63) of rows (t >> 6) + 2 i
      float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
      fetch(0);
      for (int blk = 0; blk < nblk; ++blk, ++it) {
        const int s = it & 1;
        mbar_wait(empty[s], ((it >> 1) & 1) ^ 1);
        unsigned char *st = stage_buf + s * kStage;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int i4 = tid + 128 * i;
          const float4 v = xa[i];
          float4 h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
          const int off = mn_offset(i4 >> 5, i4 & 31, 4);
          *reinterpret_cast<float4 *>(st + off) = h;
          *reinterpret_cast<float4 *>(st + kABlock + off) = l;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int i4 = tid + 128 * i;
          const float4 v = yb[i];
          bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
          float4 h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
          const int off = mn_offset(i4 >> 6, i4 & 63, 8);
          *reinterpret_cast<float4 *>(st + 2 * kABlock + off) = h;
          *reinterpret_cast<float4 *>(st + 2 * kABlock + kBBlock + off) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full[s]) : "memory");
        fetch(blk + 1);
      }
      if (mt == 0 && part_b) {
        // db partial of (chunk j, split sp): threads t and t + 64 hold the same column quad (rows of different parity):
        // combined through shared memory in a fixed order
        float4 *red = reinterpret_cast<float4 *>(s_red);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid >= 64) red[tid - 64] = bsum;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tid < 64) {
          const float4 o = red[tid];
          const int cl = 4 * tid;
          float *dst = part_b + (size_t)sp * N + j * kMaxNc + cl;
          if (cl < nc) dst[0] = bsum.x + o.x;
          if (cl + 1 < nc) dst[1] = bsum.y + o.y;
          if (cl + 2 < nc) dst[2] = bsum.z + o.z;
          if (cl + 3 < nc) dst[3] = bsum.w + o.w;
        }
      }
    }
  } else if (warp == 8) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int it = 0, w = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++w) {
        const int a = w & 1, j = (item / splits) % nchunk;
        const int nc = min(kMaxNc, N - j * kMaxNc);
        // D = F32, A = B = TF32, BOTH MN-major (bits 15, 16), N = nc rounded up to 16, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(((nc + 15) & ~15) >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
        size_t r_lo;
        int nblk;
        item_rows(item, r_lo, nblk);
        mbar_wait(t_empty[a], ((w >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t acc = 0;
        for (int blk = 0; blk < nblk; ++blk, ++it) {
          const int s = it & 1;
          mbar_wait(full[s], (it >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = stage_s + s * kStage, a_lo = a_hi + kABlock, b_hi = a_hi + 2 * kABlock, b_lo = b_hi + kBBlock;
#pragma unroll
          for (int split = 0; split < 3; ++split) {
            const uint32_t as = (split == 2) ? a_lo : a_hi, bs = (split == 1) ? b_lo : b_hi;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {   // rows 8 kk .. 8 kk + 7 = two 4-row groups; a group is 4 atoms (2 KB) of A, 8 (4 KB) of B
              mma_tf32(tmem + a * kMaxNc, smem_desc_mn(as + kk * 4096, 512, 2048), smem_desc_mn(bs + kk * 8192, 512, 4096),
                       idesc, acc);
              acc = 1;
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty[s]) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t_full[a]) : "memory");
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: partial (128 x nc) tile -> workspace
    const int qtr = warp & 3;
    float *patch = s_patch + qtr * kPatch;
    int w = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++w) {
      const int a = w & 1, mj = item / splits, mt = mj / nchunk, j = mj - mt * nchunk, sp = item - mj * splits;
      const int nc = min(kMaxNc, N - j * kMaxNc);
      size_t r_lo;
      int nblk;
      item_rows(item, r_lo, nblk);
      mbar_wait(t_full[a], (w >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(qtr * 32) << 16) + a * kMaxNc;
      float *dst_base = part_w + (size_t)sp * K * N;
      const int k0 = mt * kRows + qtr * 32;                // this warp's 32 input channels
      for (int c0 = 0; c0 < nc; c0 += 32) {
        uint32_t v[32];
        if (nblk > 0) {
          PCG_TMEM_LD32(taddr + c0, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {                                           // an empty split never touched the accumulator
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = 0u;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t)
          *reinterpret_cast<float4 *>(patch + lane * 36 + 4 * t) =
              make_float4(__uint_as_float(v[4 * t]), __uint_as_float(v[4 * t + 1]), __uint_as_float(v[4 * t + 2]),
                          __uint_as_float(v[4 * t + 3]));
        __syncwarp();
        const int rr = lane >> 3, cq = (lane & 7) * 4;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int r = 4 * t + rr;
          const float4 val = *reinterpret_cast<const float4 *>(patch + r * 36 + cq);
          if (k0 + r < K && c0 + cq < nc) {
            float *dst = dst_base + (size_t)(k0 + r) * N + (size_t)j * kMaxNc + c0 + cq;
            dst[0] = val.x;
            if (c0 + cq + 1 < nc) dst[1] = val.y;
            if (c0 + cq + 2 < nc) dst[2] = val.z;
            if (c0 + cq + 3 < nc) dst[3] = val.w;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t_empty[a]) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
  }
}

// dst[i] = sum over splits (ascending) of part[s][i]
__global__ void __launch_bounds__(256)
split_sum_kernel(size_t count, int splits, const float *__restrict__ part, float *__restrict__ dst) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += __ldg(part + (size_t)s * count + i);
    dst[i] = acc;
  }
}

int bwd_weight_splits(size_t rows, int K, int N) {
  const int nmt = (K + kRows - 1) / kRows, nchunk = (N + kMaxNc - 1) / kMaxNc;
  long long sp = (2LL * 148) / ((long long)nmt * nchunk);
  const long long max_sp = (long long)((rows + 4 * kKB - 1) / (4 * kKB));   // at least 4 row blocks per split
  if (sp > max_sp) sp = max_sp;
  if (sp > 512) sp = 512;
  return sp < 1 ? 1 : (int)sp;
}

size_t image_bytes(int K, int N) {
  return (size_t)((N + kMaxNc - 1) / kMaxNc) * ((K + kKB - 1) / kKB) * (2 * kBBlock);
}

int launch_prep(int K, int N, size_t sk, size_t sn, const float *w, unsigned char *image, cudaStream_t st) {
  const size_t elems = image_bytes(K, N) / 8;
  const unsigned blocks = (unsigned)((elems + 255) / 256 < 4096 ? (elems + 255) / 256 : 4096);
  dense_prep_kernel<<<blocks, 256, 0, st>>>(K, N, sk, sn, w, image);
  PC_RETURN_LAUNCH_STATUS();
}

int launch_dense(int pool, size_t rows, int K, size_t ldx, int N, size_t ldo, size_t ldp, int relu, const float *x,
                 const unsigned char *image, const float *bias, float *out, float *pooled, cudaStream_t st,
                 int nsplit = 3) {
  static_assert((kMaxN + 4 * kPatch) * sizeof(float) + 24 * sizeof(uint64_t) + 16 <= kFixedBytes, "fixed part");
  const int ntiles = (int)((rows + kRows - 1) / kRows);
  // output columns per item: narrower units only when the launch has very few (tile, unit) items -- every unit of a
  // tile stages the tile's A operand again, so splitting 64 items into 128 made SA4's layers slower (23 -> 33 us) while
  // splitting FP1's 8 items into 64 made them faster
  int sw = kMaxNc;
  while (sw > 32 && (long long)ntiles * ((N + sw - 1) / sw) <= num_sms() / 4) sw >>= 1;
  if (sw > N) sw = (N + 31) / 32 * 32 > 32 ? (N + 31) / 32 * 32 : 32;   // one unit narrower than 256: a smaller B block
  const int nitems = ntiles * ((N + sw - 1) / sw);
  const int grid = nitems < num_sms() ? nitems : num_sms();
  const size_t stage_bytes = 2 * (size_t)kABlock + 2 * (size_t)sw * 128;
  int nst = (int)((227 * 1024 - kFixedBytes) / stage_bytes);
  nst = nst > 4 ? 4 : nst;
  // [W_hi | W_lo] as one MMA operand where it measured faster (sw = 32: 40 -> 36 us at SA1; at sw = 64 the halved
  // number of accumulators and the doubled TMEM reads of the epilogue cost more than the four saved MMAs per block)
  const int pair = PCOPS_DENSE_PAIR && sw <= 32;
  int nacc = 512 / (pair ? 2 * sw : sw);
  nacc = nacc > 8 ? 8 : nacc;
  const size_t smem = kFixedBytes + (size_t)nst * stage_bytes;
  const int vec_x = (ldx % 4 == 0) && aligned16(x);
  const int vec_o = (ldo % 4 == 0) && aligned16(out);
  if (pool == 1) {
    PC_CUDA_TRY(allow_smem(dense_tf32_kernel<1>, smem));
    dense_tf32_kernel<1><<<grid, kFwdThreads, smem, st>>>(rows, K, ldx, N, sw, nsplit, nst, nacc, pair, ldo, ldp, relu, vec_x, vec_o, x, image, bias, out, pooled);
  } else if (pool == 2) {
    PC_CUDA_TRY(allow_smem(dense_tf32_kernel<2>, smem));
    dense_tf32_kernel<2><<<grid, kFwdThreads, smem, st>>>(rows, K, ldx, N, sw, nsplit, nst, nacc, pair, ldo, ldp, relu, vec_x, vec_o, x, image, bias, out, pooled);
  } else {
    PC_CUDA_TRY(allow_smem(dense_tf32_kernel<0>, smem));
    dense_tf32_kernel<0><<<grid, kFwdThreads, smem, st>>>(rows, K, ldx, N, sw, nsplit, nst, nacc, pair, ldo, ldp, relu, vec_x, vec_o, x, image, bias, out, pooled);
  }
  PC_RETURN_LAUNCH_STATUS();
}

bool dense_shape_ok(size_t rows, int K, int N) {
  return rows > 0 && rows < (1ull << 31) * 64 && K >= 1 && K <= 65536 && N >= 1 && N <= kMaxN;
}

}  // namespace

// used by the attention layers for their Q projection (attention_layer.cu)
size_t dense_image_bytes(int K, int N) { return image_bytes(K, N); }
int dense_prepare(int K, int N, size_t sk, size_t sn, const float *w, void *image, cudaStream_t st) {
  return launch_prep(K, N, sk, sn, w, (unsigned char *)image, st);
}
int dense_forward(size_t rows, int K, size_t ldx, int N, size_t ldo, int relu, const float *x, const void *image,
                  const float *bias, float *out, cudaStream_t st, bool four_products) {
  return launch_dense(0, rows, K, ldx, N, ldo, 0, relu, x, (const unsigned char *)image, bias, out, nullptr, st,
                      four_products ? 4 : 3);
}
}  // namespace pc

// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t pc_dense_image_bytes(int K, int N) {
  if (K <= 0 || N <= 0) return 0;
  return pc::image_bytes(K, N);
}

extern "C" int pc_dense_prepare(int K, int N, const float *w, int transpose, void *image, pc_stream_t stream) {
  if (K <= 0 || N <= 0 || !w || !image) return PC_ERR_INVALID_ARGUMENT;
  if (N > pc::kMaxN) return PC_ERR_UNSUPPORTED;
  // transpose = 0: w is (K, N) row-major, Y = X w.   transpose = 1: w is (N, K) row-major, Y = X w^T (input gradients)
  return pc::launch_prep(K, N, transpose ? 1 : (size_t)N, transpose ? (size_t)K : 1, w, (unsigned char *)image,
                         (cudaStream_t)stream);
}

extern "C" int pc_dense_fwd(size_t rows, int K, int N, const float *x, size_t ldx, const void *image, const float *bias,
                            int relu, float *y, size_t ldy, pc_stream_t stream) {
  if (rows == 0) return PC_OK;
  if (!x || !image || !y || ldx < (size_t)K || ldy < (size_t)N) return PC_ERR_INVALID_ARGUMENT;
  if (!pc::dense_shape_ok(rows, K, N)) return PC_ERR_UNSUPPORTED;
  return pc::launch_dense(0, rows, K, ldx, N, ldy, 0, relu, x, (const unsigned char *)image, bias, y, nullptr,
                          (cudaStream_t)stream);
}

extern "C" int pc_dense_pool_fwd(size_t groups, int group_size, int K, int N, const float *x, size_t ldx, const void *image,
                                 const float *bias, int relu, float *y_full, size_t ldy, float *y_pooled, size_t ldp,
                                 pc_stream_t stream) {
  if (groups == 0) return PC_OK;
  if (!x || !image || !y_pooled || ldx < (size_t)K || ldp < (size_t)N || (y_full && ldy < (size_t)N))
    return PC_ERR_INVALID_ARGUMENT;
  if (group_size != 32 || !pc::dense_shape_ok(groups * 32, K, N)) return PC_ERR_UNSUPPORTED;
  return pc::launch_dense(y_full ? 2 : 1, groups * 32, K, ldx, N, ldy, ldp, relu, x, (const unsigned char *)image, bias,
                          y_full, y_pooled, (cudaStream_t)stream);
}

extern "C" size_t pc_dense_bwd_weight_workspace_bytes(size_t rows, int K, int N) {
  if (rows == 0 || K <= 0 || N <= 0) return 0;
  const int sp = pc::bwd_weight_splits(rows, K, N);
  return ((size_t)sp * K * N + (size_t)sp * N) * sizeof(float);
}

extern "C" int pc_dense_bwd_weight(size_t rows, int K, int N, const float *x, size_t ldx, const float *dy, size_t ldy,
                                   float *dw, float *db, void *workspace, pc_stream_t stream) {
  if (K <= 0 || N <= 0 || !dw) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {   // no rows: the gradients are zero
    PC_CUDA_TRY(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)K * N, st));
    if (db) PC_CUDA_TRY(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, st));
    return PC_OK;
  }
  if (!x || !dy || ldx < (size_t)K || ldy < (size_t)N) return PC_ERR_INVALID_ARGUMENT;
  if (N > pc::kMaxN || K > 65536) return PC_ERR_UNSUPPORTED;
  if (!workspace) return PC_ERR_WORKSPACE;
  const int sp = pc::bwd_weight_splits(rows, K, N);
  size_t rps = (rows + sp - 1) / sp;
  rps = (rps + pc::kKB - 1) / pc::kKB * pc::kKB;
  float *part_w = (float *)workspace, *part_b = part_w + (size_t)sp * K * N;
  const size_t smem = 2 * (size_t)pc::kStage + (4 * pc::kPatch + 256) * sizeof(float) + 8 * sizeof(uint64_t) + 16;
  PC_CUDA_TRY(pc::allow_smem(pc::dense_bwd_weight_kernel, smem));
  const int nitems = ((K + pc::kRows - 1) / pc::kRows) * ((N + pc::kMaxNc - 1) / pc::kMaxNc) * sp;
  const int grid = nitems < pc::num_sms() ? nitems : pc::num_sms();
  pc::dense_bwd_weight_kernel<<<grid, pc::kThreads, smem, st>>>(rows, K, ldx, N, ldy, sp, rps, (ldx % 4 == 0) && pc::aligned16(x),
                                                              (ldy % 4 == 0) && pc::aligned16(dy), x, dy, part_w,
                                                              db ? part_b : nullptr);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  const size_t cnt = (size_t)K * N;
  pc::split_sum_kernel<<<(unsigned)((cnt + 255) / 256 < 2048 ? (cnt + 255) / 256 : 2048), 256, 0, st>>>(cnt, sp, part_w, dw);
  if (db) pc::split_sum_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>((size_t)N, sp, part_b, db);
  PC_RETURN_LAUNCH_STATUS();
}
