// Shared device/host helpers for libuof_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/uof_b200.h"

namespace uof {

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

#define UOF_REQUIRE(cond, ...)                     \
  do {                                             \
    if (!(cond)) {                                 \
      ::uof::set_error(__VA_ARGS__);               \
      return UOF_ERR_INVALID_ARGUMENT;             \
    }                                              \
  } while (0)

#define UOF_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      ::uof::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return UOF_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- warp-level reductions --------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// Per-sample reduction epilogue of the loss kernels: every warp of the block holds K partial sums (already warp-reduced,
// valid in lane 0) for the K consecutive floats at `dst` (nullptr: the warp had no work).  Warps that share `dst` are
// combined through shared memory and ONE atomic per quantity is issued for the group: all warps of a launch hit the same
// handful of addresses (one L2 slice), and ncu showed the serialised same-address REDs, not the streaming, bounding
// consis_fwd (30 us with one RED pair per warp vs 15 us for the backward kernel moving 1.4x the bytes).
// Must be reached by every thread of the block.
template <int K, int WARPS>
__device__ __forceinline__ void block_accumulate(const float (&v)[K], float* dst) {
  __shared__ float part[WARPS][K];
  __shared__ float* dsts[WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    dsts[warp] = dst;
#pragma unroll
    for (int k = 0; k < K; ++k) part[warp][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < WARPS * K) {
    const int w = threadIdx.x / K, k = threadIdx.x - w * K;
    float* d = dsts[w];
    bool leader = d != nullptr;
    for (int j = 0; j < w; ++j) leader = leader && dsts[j] != d;
    if (leader) {
      float s = part[w][k];
      for (int j = w + 1; j < WARPS; ++j)
        if (dsts[j] == d) s += part[j][k];
      atomicAdd(d + k, s);
    }
  }
}

// "Last block" epilogue: call after block_accumulate (every thread of the block).  Returns true in exactly one block
// of the launch -- the one whose counter increment arrives last -- after which every block's REDs are visible, so
// that block can form the final per-sample losses from the sums (read them with __ldcg).  Saves the separate
// finalize launch (~5 us of launch + dependent-load latency for a one-block kernel).
__device__ __forceinline__ bool last_block_done(unsigned* counter) {
  __shared__ bool is_last;
  __threadfence();                 // this thread's REDs are ordered before the counter increment below
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y * gridDim.z - 1;
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// The one-block "finalize" kernels that turn the (levels, B, K) partial sums of a fused loss into (B) losses are launched
// as programmatic dependents of the kernel that produces the sums: the producer's blocks call pdl_trigger() on entry, so
// the dependent grid is scheduled while the producer's last wave is still running, and it blocks in pdl_wait() until the
// producer grid has completed and its memory operations (the REDs into `sums`) are visible.  This hides the launch
// latency of the second kernel (~4 us of a 17 us consis_loss_fwd) without the per-block fence + counter of a "last
// block" epilogue, which was measured slower for these kernels.  Both instructions are no-ops when
// the kernel is launched without the attribute.  Works under stream capture (programmatic graph edges, CUDA >= 12.3).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();      // runtime.cu: false when UOF_NO_PDL is set

template <class... KArgs, class... Args>
inline cudaError_t launch_dependent(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- streaming loads/stores -------------------------------------------------------------------
__device__ __forceinline__ float ldg_f(const float* p) { return __ldg(p); }

// ---- cp.async (LDGSTS) with zero fill ---------------------------------------------------------
__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gmem_src, bool valid) {
  unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  int bytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gmem_src, bool valid) {
  unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
// "memory" clobbers: without them the compiler may move ordinary shared-memory loads across the wait (reading a slot
// before its copy has landed) or across the next copy into the same slot.
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- bilinear sampling coordinates of warp_flow (net_utils.py:39-46 + ATen grid_sampler) ---------
//   v = x + fx;  g = 2*v/max(W-1,1) - 1;  ix = ((g+1)*W - 1)/2        (align_corners = False)
//                                          ix = (g+1)/2*(W-1)          (align_corners = True)
// The fp32 rounding of this chain is part of the reference's result: at x ~ 800 one ulp of g is 2.5e-5 px of ix, the
// bilinear weights inherit it, and the gradient w.r.t. the flow (a difference of corner values times those weights) moves
// by ~1e-4 relative -- the reference run on a CPU and the same reference run on a GPU differ by 1.15e-4 in the parameter
// gradient at 256x832 for exactly this reason.  Both roundings are therefore implemented, selected by
// bit 1 of the `align_corners` flags word of the C ABI:
//   UOF_COORD_CUDA (default, bit clear): what ATen's CUDA kernels compute, i.e. what train.py runs --
//     `/ (W-1)` of a tensor by a CPU scalar multiplies by the fp32 reciprocal (BinaryDivTrueKernel.cu), and
//     grid_sampler_unnormalize's `(g+1)*W - 1` is contracted into one FMA by nvcc;
//   UOF_COORD_HOST (bit set): what ATen's CPU kernels compute -- a true division and no contraction; this is the
//     arithmetic of the CPU oracle and of the golden fixtures, used where floor() must land on the same integer.
// __f*_rn intrinsics keep ptxas from (re)contracting the chain.
constexpr int kAlignCorners = 1, kCoordHost = 2;
__device__ __forceinline__ float sample_coord(float pos, float f, int size, int flags) {
  const float v = __fadd_rn(pos, f);
  const float denom = (float)max(size - 1, 1);
  float g, ix;
  if (flags & kCoordHost) {
    g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, v), denom), 1.0f);
    // (x / 2) == (x * 0.5f) exactly in binary floating point
    if (flags & kAlignCorners) return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    ix = __fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)size), 1.0f);
  } else {
    g = __fsub_rn(__fmul_rn(__fmul_rn(2.0f, v), __frcp_rn(denom)), 1.0f);
    if (flags & kAlignCorners) return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    ix = __fmaf_rn(__fadd_rn(g, 1.0f), (float)size, -1.0f);
  }
  return __fmul_rn(ix, 0.5f);
}
// d(ix)/d(flow) for the mapping above.
inline float coord_scale(int size, int flags) {
  const bool align_corners = (flags & 1) != 0;
  float dg = 2.0f / (float)(size - 1 > 1 ? size - 1 : 1);
  return align_corners ? dg * 0.5f * (float)(size - 1) : dg * 0.5f * (float)size;
}

struct Bilinear {
  int x0, y0;            // top-left corner (may be out of range)
  float w00, w01, w10, w11;   // weights for (y0,x0) (y0,x1) (y1,x0) (y1,x1), zeroed when the corner is out of bounds
  float tx, ty;
  bool in00, in01, in10, in11;
};

__device__ __forceinline__ Bilinear make_bilinear(float ix, float iy, int H, int W) {
  Bilinear b;
  float fx0 = floorf(ix), fy0 = floorf(iy);
  b.tx = ix - fx0;
  b.ty = iy - fy0;
  // ATen forms the "one minus" weights as (x0 + 1) - ix, not 1 - tx
  float ux = (fx0 + 1.0f) - ix, uy = (fy0 + 1.0f) - iy;
  // clamp before the int conversion so that huge/NaN coordinates cannot overflow
  fx0 = fminf(fmaxf(fx0, -2.0f), (float)W + 1.0f);
  fy0 = fminf(fmaxf(fy0, -2.0f), (float)H + 1.0f);
  b.x0 = (int)fx0;
  b.y0 = (int)fy0;
  bool xin0 = (b.x0 >= 0) && (b.x0 < W), xin1 = (b.x0 + 1 >= 0) && (b.x0 + 1 < W);
  bool yin0 = (b.y0 >= 0) && (b.y0 < H), yin1 = (b.y0 + 1 >= 0) && (b.y0 + 1 < H);
  b.in00 = xin0 && yin0;
  b.in01 = xin1 && yin0;
  b.in10 = xin0 && yin1;
  b.in11 = xin1 && yin1;
  b.w00 = b.in00 ? ux * uy : 0.0f;
  b.w01 = b.in01 ? b.tx * uy : 0.0f;
  b.w10 = b.in10 ? ux * b.ty : 0.0f;
  b.w11 = b.in11 ? b.tx * b.ty : 0.0f;
  return b;
}

}  // namespace uof
