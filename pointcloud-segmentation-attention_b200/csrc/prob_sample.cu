// ProbSample for sm_100a: categorical sampling by inverse CDF.  Replaces probsampleLauncher(b,n,m,inp_p,inp_r,temp,out)
// (reference tf_ops/sampling/tf_sampling_g.cu:197-200 = cumsumKernel :7-88 + binarysearchKernel :90-104; op ProbSample
// tf_sampling.cpp:14-27,66-92).  No model calls it; it is here because it is part of the sampling library's op surface.
//
// Bit parity needs the reference's summation TREE, not just "a cumulative sum": per block of 8192 weights
//   group of four       p1 = v1, p2 = v2 + v1, p3 = v3 + p2, p4 = (v4 + v3) + p2; a trailing partial group is summed
//                       serially from zero
//   group totals        scanned in place by a reduce sweep  G[((2k+2)<<u)-1] += G[((2k+1)<<u)-1], u = 0, 1, ...
//                       and a fill sweep                    G[((2k+3)<<u)-1] += G[((2k+2)<<u)-1], u = ..., 1, 0
//   element             in-group prefix + total of the groups before it, + the running sum of earlier blocks, which is
//                       carried with a compensation term (t = total + c; s' = s + t; c = t - (s' - s))
// Every addition above is one fp32 rounding and fp32 addition commutes, so performing the same pairs in parallel gives
// the reference's bits.  One CTA per row; the sample for a uniform r is the smallest index whose cumulative value is
// >= r * total, by the reference's power-of-two descent (it differs from a textbook lower bound when weights are zero
// or negative, so the descent itself is reproduced).
#include "common.cuh"

namespace pc {
namespace {

constexpr int kScanBlock = 8192;   // weights per shared-memory block (the reference's BlockSize * 4)
constexpr int kScanThreads = 512;

__device__ __forceinline__ int gpad(int i) { return i + (i >> 5); }  // one pad word per 32 group totals: no bank conflicts

__global__ void __launch_bounds__(kScanThreads)
cumsum_kernel(int n, const float *__restrict__ inp, float *__restrict__ out) {
  __shared__ float p[kScanBlock];
  __shared__ float G[kScanBlock / 4 + kScanBlock / 128 + 1];
  const float *row = inp + (size_t)blockIdx.x * n;
  float *orow = out + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x;
  float running = 0.f, comp = 0.f;
  for (int j = 0; j < n; j += kScanBlock) {
    const int cnt = min(n - j, kScanBlock);
    const int n24 = (cnt + 3) & ~3, n2 = n24 >> 2;
    for (int g = tid; g < n2; g += kScanThreads) {
      const int k = g * 4;
      float total;
      if (k + 3 < cnt) {
        const float v1 = row[j + k], v2 = row[j + k + 1], v3 = row[j + k + 2], v4 = row[j + k + 3];
        const float p2 = __fadd_rn(v2, v1), p3 = __fadd_rn(v3, p2), p4 = __fadd_rn(__fadd_rn(v4, v3), p2);
        p[k] = v1; p[k + 1] = p2; p[k + 2] = p3; p[k + 3] = p4;
        total = p4;
      } else {
        float v = 0.f;
        for (int k2 = k; k2 < n24; ++k2) {
          if (k2 < cnt) v = __fadd_rn(v, row[j + k2]);
          p[k2] = v;
        }
        total = v;
      }
      G[gpad(g)] = total;
    }
    int u = 0;
    for (; (2 << u) <= n2; ++u) {  // reduce sweep
      __syncthreads();
      for (int k = tid; k < (n2 >> (u + 1)); k += kScanThreads) {
        const int a = gpad((((k << 1) + 2) << u) - 1), b = gpad((((k << 1) + 1) << u) - 1);
        G[a] = __fadd_rn(G[a], G[b]);
      }
    }
    for (--u; u >= 0; --u) {  // fill sweep
      __syncthreads();
      for (int k = tid; k < ((n2 - (1 << u)) >> (u + 1)); k += kScanThreads) {
        const int a = gpad((((k << 1) + 3) << u) - 1), b = gpad((((k << 1) + 2) << u) - 1);
        G[a] = __fadd_rn(G[a], G[b]);
      }
    }
    __syncthreads();
    for (int k = tid; k < cnt; k += kScanThreads) {
      const float v = k < 4 ? p[k] : __fadd_rn(p[k], G[gpad((k >> 2) - 1)]);
      orow[j + k] = __fadd_rn(v, running);
    }
    const float t = __fadd_rn(G[gpad(n2 - 1)], comp);
    const float r2 = __fadd_rn(running, t);
    comp = __fsub_rn(t, __fsub_rn(r2, running));
    running = r2;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
cdf_search_kernel(int n, int m, int base, const float *__restrict__ cdf, const float *__restrict__ rnd,
                  int *__restrict__ out) {
  const float *c = cdf + (size_t)blockIdx.y * n;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  const float q = __fmul_rn(rnd[(size_t)blockIdx.y * m + j], c[n - 1]);
  int r = n - 1;
  for (int k = base; k >= 1; k >>= 1)
    if (r >= k && c[r - k] >= q) r -= k;
  out[(size_t)blockIdx.y * m + j] = r;
}

}  // namespace
}  // namespace pc

extern "C" int pc_cumsum(int b, int n, const float *inp, float *out, pc_stream_t stream) {
  if (b < 0 || n < 0 || b > 65535) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || n == 0) return PC_OK;
  if (!inp || !out) return PC_ERR_INVALID_ARGUMENT;
  pc::cumsum_kernel<<<b, pc::kScanThreads, 0, (cudaStream_t)stream>>>(n, inp, out);
  PC_RETURN_LAUNCH_STATUS();
}

extern "C" int pc_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out,
                              pc_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || b > 65535) return PC_ERR_INVALID_ARGUMENT;
  if (b == 0 || m == 0) return PC_OK;
  if (n == 0) return PC_ERR_INVALID_ARGUMENT;
  if (!inp_p || !inp_r || !out) return PC_ERR_INVALID_ARGUMENT;
  if (!temp) return PC_ERR_WORKSPACE;
  int rc = pc_cumsum(b, n, inp_p, temp, stream);
  if (rc != PC_OK) return rc;
  int base = 1;
  while (base < n) base <<= 1;
  dim3 grid((m + 255) / 256, b);
  pc::cdf_search_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, base, temp, inp_r, out);
  PC_RETURN_LAUNCH_STATUS();
}
