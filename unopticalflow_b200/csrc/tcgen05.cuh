// tcgen05 (5th-generation tensor core) + TMEM helpers for sm_100a, inline PTX only.
// Encodings follow the CUTLASS/CuTe headers (cute/arch/mma_sm100_desc.hpp, mma_sm100_umma.hpp, copy_sm100.hpp,
// tmem_allocator_sm100.hpp) shipped with the CUDA 12.9 toolchain's libraries; nothing of CUTLASS is included.
#pragma once
#include <stdint.h>

#include "tma.cuh"

namespace uof {
namespace tc {

// ---- instruction descriptor (32 bit), kind::tf32, fp32 accumulate, A and B K-major -----------------------------
//   [4,6) c_format = 1 (F32)   [7,10) a_format = 2 (TF32)   [10,13) b_format = 2 (TF32)
//   [15] a_major = 0 (K)       [16] b_major = 0 (K)          [17,23) N >> 3     [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- shared-memory matrix descriptor (64 bit), K-major operand in a swizzled canonical layout ---------------------
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (1 for swizzled K-major)   [32,46) stride byte
//   offset >> 4 (distance between 8-row groups)   [46,48) version = 1 (Blackwell)   [61,64) layout: 6 = SWIZZLE_32B,
//   4 = SWIZZLE_64B, 2 = SWIZZLE_128B.  Rows are 32/64/128 bytes apart inside an 8-row group; the hardware XORs address
//   bits [4,4+n) with bits [7,7+n) -- the same pattern the TMA unit writes with CU_TENSOR_MAP_SWIZZLE_{32,64,128}B --
//   so the tile base must be aligned to 8 rows (256/512/1024 bytes).  Stepping along K inside a row = adding the byte
//   offset to the start address.
constexpr uint64_t kLayoutSw32 = 6, kLayoutSw64 = 4, kLayoutSw128 = 2;
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t smem_addr, uint32_t sbo_bytes, uint64_t layout) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
         ((uint64_t)1 << 46) | (layout << 61);
}

// ---- TMEM allocation (one warp, .sync.aligned) ---------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA: D[tmem] (+)= A * B^T, issued by ONE thread ------------------------------------------------------------------
// A from TMEM (M = 128 lanes x 8 consecutive 32-bit columns per instruction), B from shared memory (N rows x 8 tf32).
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// A and B from shared memory
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers: shape 32x32b, a warp touches the 32 lanes of its quadrant (warp id % 4), N columns per lane ----
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// tf32 split for 3xTF32: hi keeps the top 19 bits (sign, 8 exponent, 10 mantissa -- exactly representable whatever
// rounding the tensor core applies to its inputs), lo = x - hi is exact in fp32 and itself enters the MMA as tf32
// (relative error 2^-11 of a term that is already 2^-11 of x).  a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b.
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return __float_as_uint(x - __uint_as_float(hi)); }

}  // namespace tc

// fp32 tensor map with explicit dims / strides (innermost first) and a swizzle mode; out-of-bounds elements read as zero.
inline bool make_map_4d(CUtensorMap* map, const float* base, const cuuint64_t dims[4], const cuuint64_t strides_bytes[3],
                        const cuuint32_t box[4], CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || (reinterpret_cast<uintptr_t>(base) & 15u)) return false;
  for (int i = 0; i < 3; ++i)
    if (strides_bytes[i] % 16) return false;
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides_bytes, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace uof
