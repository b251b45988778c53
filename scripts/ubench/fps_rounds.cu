// Micro-benchmark: where do the cycles of one FPS round go?  Variants of the on-chip FPS round for n = 8192:
//   mode 0  body + full tail (as fps.cu, one-barrier tail)      mode 1  body only (centre = j, no reductions, no barrier)
//   mode 2  tail only (one distance update per thread)           mode 3  body + warp redux only (no barrier, no smem)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fps_rounds fps_rounds.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int P, int T, int MODE, bool SCALAR>
__global__ void __launch_bounds__(T, 1) k(int n, int m, float one, const float *xyz, int *out, long long *cyc) {
  extern __shared__ float s_xyz[];
  __shared__ int2 s_pair[2][32];
  constexpr int H = P / 2, nwarps = T / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *p = xyz + (size_t)blockIdx.x * n * 3;
  for (int i = tid; i < n * 3; i += T) s_xyz[i] = p[i];
  __syncthreads();
  f32x2 px[H], py[H], pz[H];
  float td[P];
  for (int h = 0; h < H; ++h) {
    int k0 = tid + 2 * h * T, k1 = tid + (2 * h + 1) * T;
    px[h] = pack2(s_xyz[k0 * 3], s_xyz[k1 * 3]); py[h] = pack2(s_xyz[k0 * 3 + 1], s_xyz[k1 * 3 + 1]); pz[h] = pack2(s_xyz[k0 * 3 + 2], s_xyz[k1 * 3 + 2]);
    td[2 * h] = td[2 * h + 1] = 1e38f;
  }
  const f32x2 one2 = pack2(one, one);
  int old = 0, par = 1;
  long long t0 = clock64();
  for (int j = 1; j < m; ++j) {
    const float cx = s_xyz[old * 3], cy = s_xyz[old * 3 + 1], cz = s_xyz[old * 3 + 2];
    const f32x2 cx2 = pack2(cx, cx), cy2 = pack2(cy, cy), cz2 = pack2(cz, cz);
    float vmax = -1.f;
    constexpr int HH = (MODE == 2) ? 1 : H;
#pragma unroll
    for (int h = 0; h < HH; ++h) {
      float d0, d1;
      if (SCALAR) {
        float x0, x1, y0, y1, z0, z1;
        unpack2(px[h], x0, x1); unpack2(py[h], y0, y1); unpack2(pz[h], z0, z1);
        float ax = __fsub_rn(x0, cx), ay = __fsub_rn(y0, cy), az = __fsub_rn(z0, cz);
        d0 = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
        float bx = __fsub_rn(x1, cx), by = __fsub_rn(y1, cy), bz = __fsub_rn(z1, cz);
        d1 = __fadd_rn(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by)), __fmul_rn(bz, bz));
      } else {
        f32x2 dx = sub2(px[h], cx2), dy = sub2(py[h], cy2), dz = sub2(pz[h], cz2);
        unpack2(fma2(fma2(mul2(dx, dx), one2, mul2(dy, dy)), one2, mul2(dz, dz)), d0, d1);
      }
      td[2 * h] = fminf(d0, td[2 * h]); td[2 * h + 1] = fminf(d1, td[2 * h + 1]);
      vmax = fmaxf(vmax, fmaxf(td[2 * h], td[2 * h + 1]));
    }
    if (MODE == 1) { old = (j * 37) & (n - 1); if (vmax == 123.f) old = 0; continue; }
    const int vb = __float_as_int(vmax);
    const int wmax = __reduce_max_sync(0xffffffffu, vb);
    int tb = 0x7fffffff;
    if (vb == wmax) {
#pragma unroll
      for (int i = P - 1; i >= 0; --i) if (__float_as_int(td[i]) == wmax) tb = tid + i * T;
    }
    const int wkey = __reduce_min_sync(0xffffffffu, tb);
    if (MODE == 3) { old = wkey & (n - 1); continue; }
    if (lane == 0) s_pair[par][warp] = make_int2(wmax, wkey);
    __syncthreads();
    const int2 pr = s_pair[par][lane < nwarps ? lane : 0];
    const int gmax = __reduce_max_sync(0xffffffffu, pr.x);
    old = __reduce_min_sync(0xffffffffu, pr.x == gmax ? pr.y : 0x7fffffff) & (n - 1);
    par ^= 1;
    if (tid == 0) out[blockIdx.x * m + j] = old;
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < P; ++i) s += td[i];
  if (s == 12345.f) out[0] = 1;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int P, int T, int MODE, bool SCALAR>
void run(const char *name, const float *xyz, int *out, long long *cyc) {
  const int n = 8192, m = 1024, b = 16;
  size_t smem = n * 3 * sizeof(float);
  cudaFuncSetAttribute(k<P, T, MODE, SCALAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<P, T, MODE, SCALAR><<<b, T, smem>>>(n, m, 1.0f, xyz, out, cyc);
  cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s %8.1f cycles/round  (%s)\n", name, (double)h[0] / (m - 1), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int n = 8192, b = 16;
  float *xyz; int *out; long long *cyc;
  cudaMalloc(&xyz, sizeof(float) * b * n * 3); cudaMalloc(&out, sizeof(int) * b * 1024); cudaMalloc(&cyc, 8 * 16);
  float *h = new float[b * n * 3];
  unsigned s = 1; for (int i = 0; i < b * n * 3; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) * (1.f / 16777216.f); }
  cudaMemcpy(xyz, h, sizeof(float) * b * n * 3, cudaMemcpyHostToDevice);
  run<32, 256, 0, false>("T256 P32 packed full", xyz, out, cyc);
  run<32, 256, 1, false>("T256 P32 packed body-only", xyz, out, cyc);
  run<32, 256, 3, false>("T256 P32 packed body+warp-redux", xyz, out, cyc);
  run<32, 256, 2, false>("T256 P32 tail-only", xyz, out, cyc);
  run<32, 256, 1, true>("T256 P32 scalar body-only", xyz, out, cyc);
  run<32, 256, 0, true>("T256 P32 scalar full", xyz, out, cyc);
  run<16, 512, 0, false>("T512 P16 packed full", xyz, out, cyc);
  run<16, 512, 1, false>("T512 P16 packed body-only", xyz, out, cyc);
  run<16, 512, 2, false>("T512 P16 tail-only", xyz, out, cyc);
  run<16, 512, 1, true>("T512 P16 scalar body-only", xyz, out, cyc);
  run<16, 512, 0, true>("T512 P16 scalar full", xyz, out, cyc);
  run<8, 1024, 0, false>("T1024 P8 packed full", xyz, out, cyc);
  run<8, 1024, 1, false>("T1024 P8 packed body-only", xyz, out, cyc);
  run<8, 1024, 2, false>("T1024 P8 tail-only", xyz, out, cyc);
  run<8, 1024, 1, true>("T1024 P8 scalar body-only", xyz, out, cyc);
  return 0;
}
