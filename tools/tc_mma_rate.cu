// Microbenchmark: cycles per tcgen05.mma (kind::tf32, M=128, K=8) as a function of N, A from TMEM (.ts) or shared memory (.ss).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I unopticalflow_b200/csrc tools/tc_mma_rate.cu -o tools/tc_mma_rate
#include <stdio.h>
#include "tcgen05.cuh"
using namespace uof;

template <int N, bool TS>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];      // B: [256 rows][16 floats] sw64 = 16 KB; A (ss): [128][16] = 8 KB
  __shared__ uint32_t tmem_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 6144; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (warp == 0) { tc::tmem_alloc(&tmem_s, 512); tc::tmem_relinquish(); }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  tc::fence_proxy_async();
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tmem = tmem_s;
  uint32_t z[16];
  for (int j = 0; j < 16; ++j) z[j] = __float_as_uint(1.0f);
  tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256, z);
  tc::wait_st();
  tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, N);
    const uint64_t bd = tc::smem_desc_kmajor(smem_u32(smem), 512, tc::kLayoutSw64);
    const uint64_t ad = tc::smem_desc_kmajor(smem_u32(smem) + 16384, 512, tc::kLayoutSw64);
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      if (TS) tc::mma_tf32_ts(tmem, tmem + 256 + 8 * (i & 1), bd, idesc, i > 0);
      else tc::mma_tf32_ss(tmem, ad, bd, idesc, i > 0);
    }
    tc::mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

template <int N, bool TS> void run(long long* d, int reps) {
  cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  long long h = 0;
  for (int k = 0; k < 2; ++k) { rate_kernel<N, TS><<<1, 128, 32768>>>(d, reps); cudaDeviceSynchronize(); }
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d %s: %d MMAs in %lld cycles = %.1f cycles/MMA  (%.0f MAC/clk)\n", N, TS ? "A in TMEM" : "A in smem", reps, h,
         (double)h / reps, 128.0 * N * 8 * reps / h);
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  const int reps = 512;
  run<16, true>(d, reps); run<32, true>(d, reps); run<64, true>(d, reps); run<96, true>(d, reps); run<128, true>(d, reps); run<256, true>(d, reps);
  run<32, false>(d, reps); run<64, false>(d, reps); run<128, false>(d, reps); run<256, false>(d, reps);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
