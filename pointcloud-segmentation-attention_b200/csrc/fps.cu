// Farthest point sampling for sm_100a.
//
// Replaces farthestpointsamplingKernel / Launcher (reference tf_ops/sampling/tf_sampling_g.cu:105-170,203-205).
// Semantics kept bit-for-bit: out[0] = 0; running min-distance starts at 1e38; distance is the un-fused
// (dx*dx + dy*dy) + dz*dz; the next centre is the point of maximal min-distance with ties resolved to the
// smallest (k mod 512, k) -- what the reference's 512-thread strided partition and left-biased tree produce.
//
// Design (not the reference's): one scene's whole state lives ON CHIP.  Each thread owns P points in registers
// (xyz + running min-distance = 4P registers), a copy of xyz sits in shared memory only to broadcast the chosen
// centre.  A round is: P distance updates per thread -> redux.sync max over the value bits (non-negative floats
// order like ints) -> one shared-memory hop across warps -> only threads holding the maximal value compute their
// tie-break key and atomicMin it.  Two barriers per round instead of the reference's ten, no global traffic at all
// inside the m-1 dependent rounds.
#include "common.cuh"

namespace pc {
namespace {

constexpr int kMaxRegPoints = 8192;  // T=1024 threads x P=8 points

__device__ __forceinline__ int tie_key(int k) { return ((k & 511) << 22) | (k >> 9); }
__device__ __forceinline__ int tie_key_to_index(int t) { return ((t & 0x3fffff) << 9) | (t >> 22); }

// One CTA per scene (grid-stride over scenes), blockDim.x = T threads, thread owns points tid + i*T.
template <int P>
__global__ void __launch_bounds__(1024, 1)
fps_onchip_kernel(int b, int n, int m, const float *__restrict__ xyz, int *__restrict__ out) {
  extern __shared__ float s_xyz[];  // n*3
  __shared__ int s_wmax[32];
  __shared__ int s_tb[2];
  const int tid = threadIdx.x, T = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;

  for (int scene = blockIdx.x; scene < b; scene += gridDim.x) {
    const float *p = xyz + (size_t)scene * n * 3;
    int *o = out + (size_t)scene * m;
    __syncthreads();  // previous scene fully done with s_xyz / s_tb
    for (int i = tid; i < n * 3; i += T) s_xyz[i] = p[i];
    if (tid == 0) { s_tb[0] = INT_MAX; s_tb[1] = INT_MAX; o[0] = 0; }
    __syncthreads();

    float px[P], py[P], pz[P], td[P];
#pragma unroll
    for (int i = 0; i < P; ++i) {
      int k = tid + i * T;
      if (k < n) {
        px[i] = s_xyz[k * 3 + 0]; py[i] = s_xyz[k * 3 + 1]; pz[i] = s_xyz[k * 3 + 2];
        td[i] = 1e38f;
      } else {  // padding never wins: real values are >= 0
        px[i] = py[i] = pz[i] = 0.0f;
        td[i] = -1.0f;
      }
    }

    int old = 0;
    for (int j = 1; j < m; ++j) {
      const float cx = s_xyz[old * 3 + 0], cy = s_xyz[old * 3 + 1], cz = s_xyz[old * 3 + 2];
      float vmax = -1.0f;
#pragma unroll
      for (int i = 0; i < P; ++i) {
        float d = sqdist3(px[i], py[i], pz[i], cx, cy, cz);
        td[i] = fminf(d, td[i]);
        vmax = fmaxf(vmax, td[i]);
      }
      const int vb = __float_as_int(vmax);
      const int wmax = __reduce_max_sync(PC_FULL_MASK, vb);
      if (lane == 0) s_wmax[warp] = wmax;
      __syncthreads();
      const int g = (lane < nwarps) ? s_wmax[lane] : INT_MIN;
      const int gmax = __reduce_max_sync(PC_FULL_MASK, g);
      const int slot = j & 1;
      if (vb == gmax) {  // rare: this thread owns a point at the maximum
        int tb = INT_MAX;
#pragma unroll
        for (int i = 0; i < P; ++i)
          if (__float_as_int(td[i]) == gmax) tb = min(tb, tie_key(tid + i * T));
        atomicMin(&s_tb[slot], tb);
      }
      if (tid == 0) s_tb[slot ^ 1] = INT_MAX;
      __syncthreads();
      old = tie_key_to_index(s_tb[slot]);
      if (tid == 0) o[j] = old;
    }
  }
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

template <int P>
int launch_onchip(int b, int n, int m, int T, const float *xyz, int *out, cudaStream_t st) {
  size_t smem = (size_t)n * 3 * sizeof(float);
  if (smem > 48 * 1024) PC_CUDA_TRY(allow_smem(fps_onchip_kernel<P>, smem));
  int grid = b;  // one CTA per scene; more scenes than SMs simply queue (1 CTA/SM resident)
  fps_onchip_kernel<P><<<grid, T, smem, st>>>(b, n, m, xyz, out);
  PC_RETURN_LAUNCH_STATUS();
}

}  // namespace
}  // namespace pc

extern "C" size_t pc_fps_workspace_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 0 || m <= 0) return 0;
  if (n <= pc::kMaxRegPoints) return 0;
  return (size_t)b * n * sizeof(float);
}

extern "C" int pc_fps(int b, int n, int m, const float *xyz, void *workspace, int *out_idx, pc_stream_t stream) {
  if (b < 0 || n < 0) return PC_ERR_INVALID_ARGUMENT;
  if (m <= 0 || b == 0) return PC_OK;  // tf_sampling_g.cu:106-107
  if (n == 0) return PC_ERR_INVALID_ARGUMENT;
  if (!xyz || !out_idx) return PC_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= pc::kMaxRegPoints) {
    // ~4 points per thread until the CTA is full, then more points per thread.
    int T = ((n + 3) / 4 + 31) / 32 * 32;
    if (T > 1024) T = 1024;
    int P = (n + T - 1) / T;
    if (P <= 1) return pc::launch_onchip<1>(b, n, m, T, xyz, out_idx, st);
    if (P <= 2) return pc::launch_onchip<2>(b, n, m, T, xyz, out_idx, st);
    if (P <= 4) return pc::launch_onchip<4>(b, n, m, T, xyz, out_idx, st);
    return pc::launch_onchip<8>(b, n, m, T, xyz, out_idx, st);
  }
  if (!workspace) return PC_ERR_WORKSPACE;
  pc::fps_stream_kernel<<<b, 1024, 0, st>>>(b, n, m, xyz, (float *)workspace, out_idx);
  PC_RETURN_LAUNCH_STATUS();
}
