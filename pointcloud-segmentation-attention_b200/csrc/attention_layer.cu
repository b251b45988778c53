// The whole AttentionLayer.call of the SA1 attention level in ONE kernel, on the 5th-generation tensor cores.
//
// Reference (attention_points/attention_scannet/attention_layer.py:29-45, used by pointnet_sa_module_attention :255-261
// with output_dim = key_dim = 4, heads = C/4): three Dense(C) projections (Q from the query row, K and V from the 32
// grouped rows), a RAW reshape to heads, softmax(QK^T / 2) V.  TensorFlow runs that as three cuBLAS GEMMs, two batched
// GEMMs with M = 1 / K = 4, a softmax and five intermediate tensors; the K and V tensors alone are 268 MB at SA1.
//
// Here (C = 64; the wider levels, whose weights do not fit shared memory, are attention_layer_wide.cu):
//   * a persistent CTA per SM keeps W_k | W_v resident in shared memory as the B operand (N = 128, K = 64);
//   * per tile of 4 neighbourhoods (128 rows) the X rows are loaded once, split into TF32 hi / lo parts and written to
//     shared memory in the UMMA K-major 128-byte-swizzle layout;
//   * one elected thread issues 24 tcgen05.mma (kind::tf32, M = 128, N = 128, K = 8): X_hi W_hi + X_hi W_lo + X_lo W_hi,
//     i.e. 3xTF32 split accumulation in an fp32 TMEM accumulator (plain TF32 would miss the 1e-5 parity bound);
//   * the epilogue reads each accumulator row straight from TMEM (tcgen05.ld 32x32b): with the reference's raw
//     reshape a row of K (64 floats) is 16 pseudo-keys of one head, its neighbour row holds the other 16, so the
//     softmax over the 32 pseudo-keys is register-local plus one shuffle; K and V never exist in memory.
//   * Q (one row per neighbourhood) is a 64 x 64 GEMV done by a small CUDA-core kernel in front (0.5 % of the work);
//   * two worker groups with their own operand stage and TMEM accumulator run half a tile out of phase, so global
//     loads, tensor-core work and epilogues of neighbouring tiles overlap.
#include <math.h>
#include "common.cuh"

namespace pc {
namespace {

constexpr int kC = 64;            // layer width handled by this kernel
constexpr int kS = 32;            // samples per neighbourhood
constexpr int kRows = 128;        // rows per tile = 4 neighbourhoods
constexpr int kN = 2 * kC;        // K | V columns
// Operand layout in shared memory: UMMA K-major with 128-byte swizzle.  A row of K = 64 tf32 is two 128-byte segments
// ("K blocks" of 32 elements); a K block of the operand is [row / 8][row % 8][128 bytes] with the eight 16-byte chunks
// of a row XOR-permuted by row % 8 (Swizzle<3,4,3>); K blocks are 16 KB apart.  A warp's 128-bit stores of
// consecutive float4s of a row then spread over all banks (4 wavefronts per 512 bytes, the minimum).
constexpr int kSBO = 1024;                            // bytes between 8-row groups inside a K block
constexpr int kKBlockBytes = (kRows / 8) * kSBO;      // 16 KB: 128 rows x 128 bytes
constexpr int kOperandBytes = (kC / 32) * kKBlockBytes;  // 32 KB for a 128 x 64 tf32 operand
constexpr int kImageBytes = 2 * kOperandBytes + kC * kC * 4 + 3 * kC * 4;  // B_hi | B_lo | Wq | bq | bk | bv
constexpr int kThreads = 288;     // warps 0-3, 4-7: two worker groups (loader + epilogue); warp 8: MMA issuer

__host__ __device__ inline int operand_offset(int row, int k) {  // byte offset of element (row, k), k in tf32 elements
  return (k >> 5) * kKBlockBytes + (row >> 3) * kSBO + (row & 7) * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + (k & 3) * 4;
}

__device__ __forceinline__ float tf32_rna(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Builds the per-layer image the main kernel copies into shared memory: W_k | W_v transposed to [n][k], split into
// TF32 hi / lo, laid out as UMMA core matrices; then Wq and the three biases in fp32.
__global__ void attention_layer_prep_kernel(const float *__restrict__ wq, const float *__restrict__ bq,
                                            const float *__restrict__ wk, const float *__restrict__ bk,
                                            const float *__restrict__ wv, const float *__restrict__ bv,
                                            unsigned char *__restrict__ image) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < kN * kC) {
    const int n = t / kC, k = t - n * kC;
    const float w = (n < kC) ? wk[k * kC + n] : wv[k * kC + (n - kC)];  // Dense kernel is [in][out]
    const float hi = tf32_rna(w), lo = tf32_rna(w - hi);
    *reinterpret_cast<float *>(image + operand_offset(n, k)) = hi;
    *reinterpret_cast<float *>(image + kOperandBytes + operand_offset(n, k)) = lo;
  }
  float *tail = reinterpret_cast<float *>(image + 2 * kOperandBytes);
  if (t < kC * kC) tail[t] = wq[t];
  if (t < kC) {
    tail[kC * kC + t] = bq ? bq[t] : 0.f;
    tail[kC * kC + kC + t] = bk ? bk[t] : 0.f;
    tail[kC * kC + 2 * kC + t] = bv ? bv[t] : 0.f;
  }
}

// Q = xq Wq + bq for every neighbourhood (G x 64 outputs, 64 MACs each: 0.5 % of the layer's work) into a scratch the
// main kernel's epilogue reads back through L2; keeps the serial 64-step GEMV off the tile pipeline's critical path.
__global__ void __launch_bounds__(256)
attention_layer_q_kernel(int G, size_t ldq, const float *__restrict__ xq, const float *__restrict__ wq,
                         const float *__restrict__ bq, float *__restrict__ q) {
  __shared__ __align__(16) float s_w[kC * kC];
  __shared__ __align__(16) float s_x[64][kC + 4];  // 64 query rows per step; +4 keeps float4 rows aligned, spreads banks
  for (int i = threadIdx.x; i < kC * kC; i += 256) s_w[i] = __ldg(wq + i);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // thread -> rows 4*ty..+3, columns 4*tx..+3
  float bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bias[j] = bq ? __ldg(bq + 4 * tx + j) : 0.f;
  for (int g0 = blockIdx.x * 64; g0 < G; g0 += gridDim.x * 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * (kC / 4); i += 256) {
      const int r = i >> 4, c4 = i & 15;
      const float4 v = (g0 + r < G) ? __ldg(reinterpret_cast<const float4 *>(xq + (size_t)(g0 + r) * ldq) + c4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4 *>(&s_x[r][4 * c4]) = v;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = bias[j];
#pragma unroll 8
    for (int k = 0; k < kC; ++k) {  // every output accumulates k = 0 .. 63 in order
      const float4 w = *reinterpret_cast<const float4 *>(&s_w[k * kC + 4 * tx]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float xv = s_x[4 * ty + i][k];
        acc[i][0] = fmaf(xv, w.x, acc[i][0]); acc[i][1] = fmaf(xv, w.y, acc[i][1]);
        acc[i][2] = fmaf(xv, w.z, acc[i][2]); acc[i][3] = fmaf(xv, w.w, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = g0 + 4 * ty + i;
      if (g < G) *reinterpret_cast<float4 *>(q + (size_t)g * kC + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  // UMMA shared-memory descriptor, K-major, SWIZZLE_128B: start address (16-byte units), LBO = 1 (unused for swizzled
  // K-major), SBO = 1024 bytes, version 1, layout type 2
  const uint32_t lo = ((saddr >> 4) & 0x3fffu) | (1u << 16);
  const uint32_t hi = (uint32_t)(kSBO >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// start address of MMA K-step kk (8 tf32 = 32 bytes): K block kk / 4, then 32-byte steps inside the swizzled 128-byte row
__device__ __forceinline__ uint32_t kstep_addr(uint32_t base, int kk) { return base + (kk >> 2) * kKBlockBytes + (kk & 3) * 32; }

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

#define PC_TMEM_LD16(addr, v, o)                                                                                     \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),         \
                 "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),      \
                 "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])                                     \
               : "r"(addr))

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // never hang the device on a lost arrival
  } while (!ok);
}

// Pipeline (one CTA per SM, persistent over tiles of 4 neighbourhoods = 128 rows):
//   warps 0-3 and warps 4-7 are two identical worker groups; group g owns operand stage g and TMEM accumulator g and
//   takes every second tile of the CTA.  Per tile a group: splits the X rows it prefetched into registers into TF32
//   hi / lo, stores them into its operand stage, signals a_full[g], issues the global loads of its NEXT tile, waits
//   for t_full[g] and runs the epilogue -- while the other group is half a tile out of phase, so loads, tensor-core
//   work and epilogues of neighbouring tiles overlap;
//   warp 8: waits a_full[s], issues the 24 UMMAs of the tile into accumulator s, commits to t_full[s].
// Reuse hazards are ordered by each group's own program order (it observes t_full of its previous tile before it
// overwrites that tile's operands or lets the accumulator be overwritten).
__global__ void __launch_bounds__(kThreads, 1)
attention_layer_c64_kernel(int G, const float *__restrict__ qg, const float *__restrict__ x,
                           const unsigned char *__restrict__ image, float *__restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *a_buf = smem;                          // [2 stages][hi | lo][kOperandBytes]
  unsigned char *b_img = smem + 4 * kOperandBytes;      // B_hi | B_lo | Wq | bq | bk | bv  (kImageBytes)
  float *s_bk = reinterpret_cast<float *>(b_img + 2 * kOperandBytes) + kC * kC + kC, *s_bv = s_bk + kC;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(b_img + kImageBytes);  // a_full[2], t_full[2]
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar);
  const uint32_t a_full[2] = {bar0, bar0 + 8}, t_full[2] = {bar0 + 16, bar0 + 24};

  if (warp == 8) {  // TMEM: two 128-column fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(s_tmem)),
                 "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_full[0]), "r"((uint32_t)kRows));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a_full[1]), "r"((uint32_t)kRows));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_full[0]), "r"(1u));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t_full[1]), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  for (int i = tid; i < kImageBytes / 16; i += kThreads)
    reinterpret_cast<uint4 *>(b_img)[i] = __ldg(reinterpret_cast<const uint4 *>(image) + i);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s_tmem;

  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(a_buf);
  const uint32_t b_hi_s = (uint32_t)__cvta_generic_to_shared(b_img), b_lo_s = b_hi_s + kOperandBytes;
  const size_t total_rows = (size_t)G * kS;
  const int ntiles = (int)((total_rows + kRows - 1) / kRows);

  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int st = it & 1;
        mbar_wait(a_full[st], (it >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi_s = a_s + st * 2 * kOperandBytes, a_lo_s = a_hi_s + kOperandBytes;
        uint32_t acc = 0;
#pragma unroll
        for (int split = 0; split < 3; ++split) {  // X_hi W_hi, X_hi W_lo, X_lo W_hi
          const uint32_t as = (split == 2) ? a_lo_s : a_hi_s, bs = (split == 1) ? b_lo_s : b_hi_s;
#pragma unroll
          for (int kk = 0; kk < kC / 8; ++kk) {
            mma_tf32(tmem + st * kN, smem_desc(kstep_addr(as, kk)), smem_desc(kstep_addr(bs, kk)), idesc, acc);
            acc = 1;
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(t_full[st])
                     : "memory");
      }
    }
  } else {
    const int grp = tid >> 7, wt = tid & 127, wq4 = warp & 3;  // worker group, thread in group, TMEM lane quarter
    constexpr int kLd = (kRows * kC / 4) / kRows;               // 16 float4 per worker thread per tile
    float4 pre[kLd];
    auto fetch = [&](int tile) {
      const size_t row0 = (size_t)tile * kRows;
      const float4 *src = reinterpret_cast<const float4 *>(x + row0 * kC);
#pragma unroll
      for (int i = 0; i < kLd; ++i) {
        const int i4 = wt + kRows * i;
        pre[i] = (tile < ntiles && row0 + (i4 >> 4) < total_rows) ? __ldg(src + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    unsigned char *a_hi = a_buf + grp * 2 * kOperandBytes, *a_lo = a_hi + kOperandBytes;
    const int stride2 = 2 * gridDim.x;
    int tile = blockIdx.x + grp * gridDim.x;
    fetch(tile);
    for (int k = 0; tile < ntiles; tile += stride2, ++k) {
#pragma unroll
      for (int i = 0; i < kLd; ++i) {  // hi / lo split, swizzled stores
        const int i4 = wt + kRows * i, row = i4 >> 4, kq = i4 & 15;
        const float4 v = pre[i];
        float4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
        const int off = operand_offset(row, kq * 4);
        *reinterpret_cast<float4 *>(a_hi + off) = h;
        *reinterpret_cast<float4 *>(a_lo + off) = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_full[grp]) : "memory");
      fetch(tile + stride2);  // this group's next tile: in flight during the wait and the epilogue below

      // row r = wt of the tile: neighbourhood r / 32, sample row r % 32; with the raw reshape (attention_layer.py:35)
      // this row holds pseudo-keys 16*(r&1) .. +15 of head (r % 32) / 2, four consecutive columns each
      const int gl = wt >> 5, srow = wt & 31, head = srow >> 1;
      const size_t g = (size_t)tile * 4 + gl;
      const float4 q4 = (g < (size_t)G) ? __ldg(reinterpret_cast<const float4 *>(qg + g * kC + head * 4))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      mbar_wait(t_full[grp], k & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(wq4 * 32) << 16) + grp * kN;
      float a[16];
      float mx = -INFINITY;
      uint32_t kv[kC];
      PC_TMEM_LD16(taddr + 0, kv, 0);    // the row's 64 K columns = pseudo-keys 0..15 of this row, one wait
      PC_TMEM_LD16(taddr + 16, kv, 16);
      PC_TMEM_LD16(taddr + 32, kv, 32);
      PC_TMEM_LD16(taddr + 48, kv, 48);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float k0 = __uint_as_float(kv[4 * j + 0]) + s_bk[4 * j + 0], k1 = __uint_as_float(kv[4 * j + 1]) + s_bk[4 * j + 1],
                    k2 = __uint_as_float(kv[4 * j + 2]) + s_bk[4 * j + 2], k3 = __uint_as_float(kv[4 * j + 3]) + s_bk[4 * j + 3];
        a[j] = 0.5f * fmaf(q4.w, k3, fmaf(q4.z, k2, fmaf(q4.y, k1, q4.x * k0)));  // / sqrt(key_dim = 4)
        mx = fmaxf(mx, a[j]);
      }
      PC_TMEM_LD16(taddr + kC + 0, kv, 0);   // V columns: in flight during the softmax
      PC_TMEM_LD16(taddr + kC + 16, kv, 16);
      PC_TMEM_LD16(taddr + kC + 32, kv, 32);
      PC_TMEM_LD16(taddr + kC + 48, kv, 48);
      mx = fmaxf(mx, __shfl_xor_sync(PC_FULL_MASK, mx, 1));
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { a[j] = exp2f((a[j] - mx) * 1.4426950408889634f); sum += a[j]; }  // MUFU.EX2, rel. err ~1e-7
      sum += __shfl_xor_sync(PC_FULL_MASK, sum, 1);
      const float inv = 1.0f / sum;
      float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float w = a[j] * inv;
        o0 = fmaf(w, __uint_as_float(kv[4 * j + 0]) + s_bv[4 * j + 0], o0);
        o1 = fmaf(w, __uint_as_float(kv[4 * j + 1]) + s_bv[4 * j + 1], o1);
        o2 = fmaf(w, __uint_as_float(kv[4 * j + 2]) + s_bv[4 * j + 2], o2);
        o3 = fmaf(w, __uint_as_float(kv[4 * j + 3]) + s_bv[4 * j + 3], o3);
      }
      o0 += __shfl_xor_sync(PC_FULL_MASK, o0, 1); o1 += __shfl_xor_sync(PC_FULL_MASK, o1, 1);
      o2 += __shfl_xor_sync(PC_FULL_MASK, o2, 1); o3 += __shfl_xor_sync(PC_FULL_MASK, o3, 1);
      if ((srow & 1) == 0 && g < (size_t)G)
        *reinterpret_cast<float4 *>(out + g * kC + head * 4) = make_float4(o0, o1, o2, o3);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // orders these TMEM reads before later arrivals
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
  }
}

}  // namespace

// attention_layer_wide.cu: C = 128 / 256 / 512
size_t attention_layer_wide_image_bytes(int C);
size_t attention_layer_wide_workspace_bytes(int G, int C);
bool attention_layer_wide_supported(int S, int C);
int attention_layer_wide_fwd(int G, int C, const float *xq, const float *x, const float *wq, const float *bq,
                             const float *wk, const float *bk, const float *wv, const float *bv, float *out,
                             void *workspace, int mode, size_t ldq, cudaStream_t st);

// mode bit 0: build the operand image in the workspace; bit 1: run the layer on an image that is already there
int attention_layer_run(int G, int S, int C, const float *xq, const float *x, const float *wq, const float *bq,
                        const float *wk, const float *bk, const float *wv, const float *bv, float *out, void *workspace,
                        int mode, size_t ldq, cudaStream_t st) {
  const bool run = mode & 2;
  if (ldq == 0) ldq = (size_t)C;   // query rows packed; otherwise row stride in floats (e.g. S*C: sample 0 of every group)
  if (ldq < (size_t)C || (ldq & 3)) return PC_ERR_INVALID_ARGUMENT;
  if (G < 0 || S <= 0 || C <= 0) return PC_ERR_INVALID_ARGUMENT;
  const bool wide = attention_layer_wide_supported(S, C);
  if (!wide && (S != kS || C != kC)) return PC_ERR_UNSUPPORTED;  // other shapes: Dense GEMM + pc_attention_fwd
  if (run && G == 0 && !(mode & 1)) return PC_OK;
  if (!wq || !wk || !wv || (run && G > 0 && (!xq || !x || !out))) return PC_ERR_INVALID_ARGUMENT;
  if (!workspace) return PC_ERR_WORKSPACE;
  if (!aligned16(workspace) || (run && G > 0 && (!aligned16(x) || !aligned16(out)))) return PC_ERR_UNSUPPORTED;
  if (run && G == 0) mode &= ~2;
  if (wide) return attention_layer_wide_fwd(G, C, xq, x, wq, bq, wk, bk, wv, bv, out, workspace, mode, ldq, st);
  unsigned char *image = (unsigned char *)workspace;
  if (mode & 1) attention_layer_prep_kernel<<<(kN * kC + 255) / 256, 256, 0, st>>>(wq, bq, wk, bk, wv, bv, image);
  if (!(mode & 2)) PC_RETURN_LAUNCH_STATUS();
  float *qbuf = reinterpret_cast<float *>(image + ((kImageBytes + 255) / 256) * 256);
  const int qblocks = (G + 63) / 64 < num_sms() * 2 ? (G + 63) / 64 : num_sms() * 2;
  attention_layer_q_kernel<<<qblocks, 256, 0, st>>>(G, ldq, xq, wq, bq, qbuf);
  const size_t smem = 4 * kOperandBytes + kImageBytes + 64;
  PC_CUDA_TRY(allow_smem(attention_layer_c64_kernel, smem));
  const int ntiles = (int)(((size_t)G * kS + kRows - 1) / kRows);
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  attention_layer_c64_kernel<<<grid, kThreads, smem, st>>>(G, qbuf, x, image, out);
  PC_RETURN_LAUNCH_STATUS();
}
}  // namespace pc

extern "C" size_t pc_attention_layer_workspace_bytes(int G, int S, int C) {
  if (G > 0 && pc::attention_layer_wide_supported(S, C))
    return pc::attention_layer_wide_workspace_bytes(G, C);  // K | V operand image | Q scratch
  if (S != pc::kS || C != pc::kC || G <= 0) return 0;
  return (size_t)((pc::kImageBytes + 255) / 256) * 256 + (size_t)G * pc::kC * sizeof(float);  // operand image | Q scratch
}

extern "C" int pc_attention_layer_fwd(int G, int S, int C, const float *xq, const float *x, const float *wq,
                                      const float *bq, const float *wk, const float *bk, const float *wv,
                                      const float *bv, float *out, void *workspace, pc_stream_t stream) {
  return pc::attention_layer_run(G, S, C, xq, x, wq, bq, wk, bk, wv, bv, out, workspace, 3, 0, (cudaStream_t)stream);
}

extern "C" int pc_attention_layer_prepare(int S, int C, const float *wq, const float *bq, const float *wk,
                                          const float *bk, const float *wv, const float *bv, void *workspace,
                                          pc_stream_t stream) {
  return pc::attention_layer_run(0, S, C, nullptr, nullptr, wq, bq, wk, bk, wv, bv, nullptr, workspace, 1, 0,
                                 (cudaStream_t)stream);
}

extern "C" int pc_attention_layer_fwd_prepared(int G, int S, int C, const float *xq, size_t xq_stride, const float *x,
                                               const float *wq, const float *bq, const float *wk, const float *bk,
                                               const float *wv, const float *bv, float *out, void *workspace,
                                               pc_stream_t stream) {
  return pc::attention_layer_run(G, S, C, xq, x, wq, bq, wk, bk, wv, bv, out, workspace, 2, xq_stride,
                                 (cudaStream_t)stream);
}
