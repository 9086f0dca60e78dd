// Hardware probe for the tcgen05 building blocks of the tensor-core cost volume (run on a B200):
//   1. tcgen05.st / tcgen05.ld round trip through TMEM
//   2. tcgen05.mma kind::tf32, A from TMEM, B from shared memory in the K-major SWIZZLE_64B canonical layout written by hand;
//      how the tensor core converts fp32 inputs to tf32 (truncate vs round) and the error of 1xTF32 vs 3xTF32
//   3. the same B operand written by the TMA unit from an NCHW tensor through a tensor map with permuted dimensions
//      (x, c, y, b), CU_TENSOR_MAP_SWIZZLE_64B, negative box origin (zero fill)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I unopticalflow_b200/csrc tools/tc_probe.cu -o tools/tc_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tcgen05.cuh"

using namespace uof;

constexpr int M = 128, N = 32, K = 16;

struct Out {
  float roundtrip_err;
  float d1[M * N];      // 1 pass: A raw, B raw
  float d3[M * N];      // 3xTF32
  float dtma[M * N];    // 3xTF32 with B from TMA
  float btma[N * K];    // B tile as the TMA unit wrote it (de-swizzled by the kernel)
};

__device__ __forceinline__ int sw64_index(int row, int k) {      // float index inside a [rows][16] tile, SWIZZLE_64B
  const int chunk = (k >> 2) ^ ((row >> 1) & 3);
  return row * 16 + chunk * 4 + (k & 3);
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    const __grid_constant__ CUtensorMap fmap, int x0, int yrow, Out* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* b_hi = reinterpret_cast<float*>(smem_raw);            // [N][16] sw64, 2 KB
  float* b_lo = b_hi + N * K;                                   // 2 KB
  float* b_tma = b_lo + N * K;                                  // 2 KB (TMA destination)
  float* b_tma_lo = b_tma + N * K;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar_mma, bar_tma;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    tc::tmem_alloc(&tmem_base_s, 128);
    tc::tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_tma, 1);
    mbar_fence_init();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // columns: [0,32) D1, [32,64) D3, [64,96) Dtma, [96,112) A hi, [112,128) A lo
  const int m = tid;

  // ---- 1. round trip
  {
    uint32_t v[16], r[16];
    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint((float)(m * 100 + j));
    tc::tmem_st16(lane_base + 96, v);
    tc::wait_st();
    tc::tmem_ld16(lane_base + 96, r);
    tc::wait_ld();
    float e = 0;
    for (int j = 0; j < 16; ++j) e = fmaxf(e, fabsf(__uint_as_float(r[j]) - (float)(m * 100 + j)));
    if (e > 0 || m == 0) atomicMax(reinterpret_cast<int*>(&out->roundtrip_err), __float_as_int(e));
  }
  // ---- TMA load of the B tile (N channels x 16 x), permuted dims (x, c, y, b), swizzle 64B
  if (tid == 0) {
    mbar_expect_tx(&bar_tma, N * K * 4);
    tma_load_4d(b_tma, &fmap, &bar_tma, x0, 0, yrow, 0);
  }
  // ---- B by hand (raw fp32 in "hi", no pre-truncation, for the rounding-mode test) and split copies
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    b_hi[sw64_index(n, k)] = B[i];
  }
  __syncthreads();
  // A raw into TMEM cols [96,112)
  {
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(A[m * K + j]);
    tc::tmem_st16(lane_base + 96, v);
    tc::wait_st();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t idesc = tc::idesc_tf32(M, N);
  if (tid == 0) {
    for (int ks = 0; ks < 2; ++ks)
      tc::mma_tf32_ts(tmem + 0, tmem + 96 + 8 * ks, tc::smem_desc_kmajor(smem_u32(b_hi) + 32 * ks, 512, tc::kLayoutSw64), idesc,
                      ks > 0);
    tc::mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc::fence_after_sync();
  // ---- 3xTF32: split A (TMEM) and B (smem)
  {
    uint32_t hi[16], lo[16];
    for (int j = 0; j < 16; ++j) {
      const float x = A[m * K + j];
      hi[j] = tc::tf32_hi(x);
      lo[j] = tc::tf32_lo(x, hi[j]);
    }
    tc::tmem_st16(lane_base + 96, hi);
    tc::tmem_st16(lane_base + 112, lo);
    tc::wait_st();
  }
  for (int i = tid; i < N * K; i += 128) {
    const float x = b_hi[i];
    const uint32_t h = tc::tf32_hi(x);
    b_hi[i] = __uint_as_float(h);
    b_lo[i] = __uint_as_float(tc::tf32_lo(x, h));
  }
  // TMA tile: wait, de-swizzle copy out, split in place
  mbar_wait(&bar_tma, 0);
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    out->btma[i] = b_tma[sw64_index(n, k)];
  }
  __syncthreads();
  for (int i = tid; i < N * K; i += 128) {
    const float x = b_tma[i];
    const uint32_t h = tc::tf32_hi(x);
    b_tma[i] = __uint_as_float(h);
    b_tma_lo[i] = __uint_as_float(tc::tf32_lo(x, h));
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    for (int which = 0; which < 2; ++which) {
      const float* bh = which ? b_tma : b_hi;
      const float* bl = which ? b_tma_lo : b_lo;
      const uint32_t d = tmem + (which ? 64 : 32);
      bool acc = false;
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t dh = tc::smem_desc_kmajor(smem_u32(bh) + 32 * ks, 512, tc::kLayoutSw64);
        const uint64_t dl = tc::smem_desc_kmajor(smem_u32(bl) + 32 * ks, 512, tc::kLayoutSw64);
        tc::mma_tf32_ts(d, tmem + 96 + 8 * ks, dh, idesc, acc);       // hi * hi
        acc = true;
        tc::mma_tf32_ts(d, tmem + 112 + 8 * ks, dh, idesc, acc);      // lo * hi
        tc::mma_tf32_ts(d, tmem + 96 + 8 * ks, dl, idesc, acc);       // hi * lo
      }
    }
    tc::mma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 1);
  tc::fence_after_sync();
  for (int part = 0; part < 3; ++part) {
    float* dst = part == 0 ? out->d1 : (part == 1 ? out->d3 : out->dtma);
    for (int half = 0; half < 2; ++half) {
      uint32_t r[16];
      tc::tmem_ld16(lane_base + 32 * part + 16 * half, r);
      tc::wait_ld();
      for (int j = 0; j < 16; ++j) dst[m * N + 16 * half + j] = __uint_as_float(r[j]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

static float trunc_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float rna_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main(int argc, char** argv) {
  std::vector<float> A(M * K), B(N * K);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2 - 1;
  // feature tensor for the TMA test: (Bn=1, C=N, H=4, W=32); tile = channels x 16 columns starting at x0 (may be negative)
  const int C = N, H = 4, W = 32, x0 = argc > 1 ? atoi(argv[1]) : -4, yrow = 2;
  std::vector<float> F((size_t)C * H * W);
  for (auto& v : F) v = (float)rand() / RAND_MAX * 2 - 1;
  float *dA, *dB, *dF;
  Out* dO;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dF, F.size() * 4);
  cudaMalloc(&dO, sizeof(Out));
  cudaMemset(dO, 0, sizeof(Out));
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dF, F.data(), F.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap fmap;
  const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, 1};
  const cuuint64_t strides[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)C * H * W * 4};
  const cuuint32_t box[4] = {16, (cuuint32_t)N, 1, 1};
  if (!make_map_4d(&fmap, dF, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) {
    printf("tensor map with permuted dims REJECTED\n");
    return 1;
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  probe_kernel<<<1, 128, 16384>>>(dA, dB, fmap, x0, yrow, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  Out* o = new Out;
  cudaMemcpy(o, dO, sizeof(Out), cudaMemcpyDeviceToHost);
  printf("1. TMEM st/ld round trip max err: %g\n", o->roundtrip_err);
  // references
  double e_exact1 = 0, e_trunc = 0, e_rna = 0, e3 = 0, emax = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ex = 0, tr = 0, rn = 0;
      for (int k = 0; k < K; ++k) {
        ex += (double)A[m * K + k] * B[n * K + k];
        tr += (double)trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
        rn += (double)rna_tf32(A[m * K + k]) * rna_tf32(B[n * K + k]);
      }
      emax = fmax(emax, fabs(ex));
      e_exact1 = fmax(e_exact1, fabs(o->d1[m * N + n] - ex));
      e_trunc = fmax(e_trunc, fabs(o->d1[m * N + n] - tr));
      e_rna = fmax(e_rna, fabs(o->d1[m * N + n] - rn));
      e3 = fmax(e3, fabs(o->d3[m * N + n] - ex));
    }
  printf("2. MMA tf32 .ts (A in TMEM, B smem sw64 by hand), max |D| %.3f\n", emax);
  printf("   1xTF32 vs exact %.3e | vs truncated inputs %.3e | vs round-to-nearest inputs %.3e  => hardware %s\n", e_exact1,
         e_trunc, e_rna, e_trunc < e_rna ? "TRUNCATES" : "ROUNDS");
  printf("   3xTF32 vs exact %.3e (rel %.3e)\n", e3, e3 / emax);
  // TMA tile check
  double etile = 0, e3t = 0, emaxt = 0;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const int x = x0 + k;
      const float ref = (x >= 0 && x < W) ? F[((size_t)n * H + yrow) * W + x] : 0.0f;
      etile = fmax(etile, fabs(o->btma[n * K + k] - ref));
    }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ex = 0;
      for (int k = 0; k < K; ++k) {
        const int x = x0 + k;
        const float ref = (x >= 0 && x < W) ? F[((size_t)n * H + yrow) * W + x] : 0.0f;
        ex += (double)A[m * K + k] * ref;
      }
      emaxt = fmax(emaxt, fabs(ex));
      e3t = fmax(e3t, fabs(o->dtma[m * N + n] - ex));
    }
  printf("3. TMA permuted-dims sw64 tile vs expected (after de-swizzle): max err %.3e ; 3xTF32 MMA on it: %.3e (rel %.3e)\n", etile,
         e3t, e3t / emaxt);
  const bool ok = o->roundtrip_err == 0 && e3 / emax < 2e-6 && etile == 0 && e3t / emaxt < 2e-6;
  printf("%s\n", ok ? "PROBE OK" : "PROBE FAILED");
  return ok ? 0 : 2;
}
