// Shared constants / internal interface of the cost-volume kernels (cost_volume.cu, cost_volume_tma.cu).
#pragma once
#include "common.cuh"

namespace uof {
namespace cv {

constexpr int RAD = 4;                 // max displacement (pwc_tf.py:16 md=4)
constexpr int ND = 2 * RAD + 1;        // 9
constexpr int CK = 8;                  // channels per pipeline stage
constexpr int PX = 4;                  // pixels per thread along x
constexpr int DYG = 3;                 // displacement rows per thread
constexpr int NGROUP = ND / DYG;       // 3 thread groups over dy
constexpr int NT = 192;                // threads per CTA: 64 pixel quads x 3 dy-groups (every tile shape has 256 px)

// split the channel loop so that small pyramid levels still put >= ~2 CTAs on every SM
inline int pick_split(long long ctas, int nchunks) {
  int s = 1;
  while (ctas * s < 2 * kNumSMs && s < nchunks) ++s;
  return s;
}

// Forward split-K goes through fp32 atomics into a zeroed output, which is expensive (measured: the 32x104 level takes
// 25 us unsplit vs 47 us split in two, profiles/r1_cv_ksplit_sweep.txt): split only when there are fewer tiles than
// ~SMs, and only up to ~120 CTAs.
inline int pick_ksplit_atomic(long long ctas, int nchunks) {
  int s = (int)((120 + ctas - 1) / ctas);
  if (s < 1) s = 1;
  return s < nchunks ? s : nchunks;
}

// TMA + mbarrier implementations.  Return false when the TMA path does not apply (W % 4 != 0, unaligned
// pointers, no driver entry point, UOF_DISABLE_TMA=1); otherwise launch and store the status in *rc.
bool fwd_tma(const float* f1, const float* f2, float* out, int B, int C, int H, int W, long long out_bs,
             cudaStream_t stream, int* rc);
bool bwd_tma(const float* gout, long long gout_bs, const float* f1, const float* f2, float* gf1, float* gf2, int B, int C,
             int H, int W, cudaStream_t stream, int* rc);

}  // namespace cv
}  // namespace uof
