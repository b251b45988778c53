// Micro-benchmark: un-fused fp32 add/mul issue rate on sm_100a, scalar (FADD/FMUL) vs packed (FADD2/FMUL2/FFMA2).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_rate fp32_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int CH = 8;
// scalar: per iteration per chain: FADD(sub), FMUL, FADD  (3 un-fused ops)
__global__ void scalar_k(float *out, float c, int iters) {
  float v[CH];
  for (int i = 0; i < CH; ++i) v[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      float d = __fsub_rn(v[i], c);
      float m = __fmul_rn(d, d);
      v[i] = __fadd_rn(m, c);
    }
  }
  float s = 0;
  for (int i = 0; i < CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed: per iteration per chain: FADD2, FMUL2, FFMA2(x,1,c)  (3 packed ops = 6 un-fused fp32 ops)
__global__ void packed_k(float *out, float c, float one, int iters) {
  u64 v[CH];
  const u64 c2 = pk(c, c), one2 = pk(one, one);
  for (int i = 0; i < CH; ++i) v[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      u64 d = sub2(v[i], c2);
      u64 m = mul2(d, d);
      v[i] = fma2(m, one2, c2);
    }
  }
  float s = 0;
  for (int i = 0; i < CH; ++i) { float a, b; upk(v[i], a, b); s += a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// scalar FFMA (fused) for reference
__global__ void ffma_k(float *out, float c, int iters) {
  float v[CH];
  for (int i = 0; i < CH; ++i) v[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) { v[i] = fmaf(v[i], c, c); v[i] = fmaf(v[i], c, c); v[i] = fmaf(v[i], c, c); }
  }
  float s = 0;
  for (int i = 0; i < CH; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FMNMX rate (ALU pipe) mixed with FMUL (FMA pipe)
__global__ void mix_k(float *out, float c, int iters) {
  float v[CH], w[CH];
  for (int i = 0; i < CH; ++i) { v[i] = threadIdx.x * 0.001f + i; w[i] = 1e30f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) { v[i] = __fmul_rn(v[i], c); w[i] = fminf(w[i], v[i]); }
  }
  float s = 0;
  for (int i = 0; i < CH; ++i) s += v[i] + w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 256, blocks = sms * 8, iters = 4096;
  float *out;
  cudaMalloc(&out, sizeof(float) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char *name, auto launch, double ops_per_thread_iter) {
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    double ops = ops_per_thread_iter * iters * (double)threads * blocks;
    printf("%-10s %8.3f ms  %7.2f T fp32-ops/s  (%.1f ops/clk/SM at 1.965 GHz)\n", name, best, ops / best / 1e9,
           ops / (best * 1e-3) / sms / 1.965e9);
  };
  run("scalar", [&] { scalar_k<<<blocks, threads>>>(out, 0.5f, iters); }, 3.0 * CH);
  run("packed", [&] { packed_k<<<blocks, threads>>>(out, 0.5f, 1.0f, iters); }, 6.0 * CH);
  run("ffma", [&] { ffma_k<<<blocks, threads>>>(out, 0.5f, iters); }, 3.0 * CH);
  run("mul+min", [&] { mix_k<<<blocks, threads>>>(out, 0.999f, iters); }, 2.0 * CH);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
