// Microbenchmark: is packed FFMA2 (fma.rn.f32x2, sm_100+) faster than scalar FFMA per FLOP on B200?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/ffma2 profiles/microbench/ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int NACC>
__global__ void scalar_kernel(float* out, float a, float b, int iters) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>   // NACC packed accumulators = 2*NACC floats
__global__ void packed_kernel(float* out, float a, float b, int iters) {
  u64 acc[NACC];
  const u64 pa = pack(a, a * 1.0001f), pb = pack(b, b * 0.999f);
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = pack(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = ffma2(acc[i], pa, pb);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { float x, y; unpack(acc[i], x, y); s += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 20000, blocks = 148 * 8, threads = 256;
  const double threads_total = (double)blocks * threads;
  float ms1 = time_ms([&] { scalar_kernel<32><<<blocks, threads>>>(out, 0.999f, 0.001f, iters); });
  float ms2 = time_ms([&] { packed_kernel<16><<<blocks, threads>>>(out, 0.999f, 0.001f, iters); });
  float ms3 = time_ms([&] { packed_kernel<32><<<blocks, threads>>>(out, 0.999f, 0.001f, iters); });
  printf("scalar FFMA  : %.3f ms  %.1f TFLOP/s\n", ms1, 2.0 * 32 * iters * threads_total / ms1 * 1e-9);
  printf("FFMA2 (16x2) : %.3f ms  %.1f TFLOP/s\n", ms2, 2.0 * 32 * iters * threads_total / ms2 * 1e-9);
  printf("FFMA2 (32x2) : %.3f ms  %.1f TFLOP/s\n", ms3, 2.0 * 64 * iters * threads_total / ms3 * 1e-9);
  return 0;
}
