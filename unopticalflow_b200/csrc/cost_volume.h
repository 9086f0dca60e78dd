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

// Backward: split the channel loop of a tile over `s` CTAs (each writes its own output channels, no atomics) so that small
// pyramid levels still fill the 148 SMs.  The split is chosen by a wave model, not just "enough CTAs": 2 CTAs are
// resident per SM, a launch costs waves(ctas * s) x (slabs per CTA + prologue), the prologue (108 coefficient gathers per
// thread, barrier init, first TMA round trip) being worth ~1.5 slabs.  The old rule (smallest s with ctas * s >= 296)
// put the 16x52 / 8x26 / 4x13 levels at 384 / 320 / 320 CTAs = two waves of which the second is nearly empty.
inline int pick_split(long long ctas, int nchunks) {
  const long long slots = 2ll * kNumSMs;
  int best = 1;
  double best_cost = 1e300;
  for (int s = 1; s <= nchunks; ++s) {
    const long long waves = (ctas * s + slots - 1) / slots;
    const int per = (nchunks + s - 1) / s;
    const double cost = (double)waves * ((double)per + 1.5);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

// Forward split-K goes through fp32 atomics into a zeroed output, which is expensive (measured: the 32x104 level takes
// 25 us unsplit vs 47 us split in two, profiles/r1_cv_ksplit_sweep.txt): split only when there are fewer tiles than
// ~SMs, and only up to ~120 CTAs.
inline int pick_ksplit_atomic(long long ctas, int nchunks) {
  int s = (int)((120 + ctas - 1) / ctas);
  if (s < 1) s = 1;
  return s < nchunks ? s : nchunks;
}

// Smallest pyramid levels (<= 64 pixel quads per image): one CTA per (image, displacement row), no split-K atomics.
// Returns false when the level is not small (or UOF_CV_NO_SMALL is set); otherwise launches and stores the status in *rc.
// `f1_bs`: batch stride of f1 in floats (C*H*W when dense; larger when f1 is the `c1` slice of the decoder's concat buffer).
bool fwd_small(const float* f1, long long f1_bs, const float* f2, float* out, int B, int C, int H, int W, long long out_bs,
               cudaStream_t stream, int* rc);

// TMA + mbarrier implementations.  Return false when the TMA path does not apply (W % 4 != 0, unaligned
// pointers, no driver entry point, UOF_DISABLE_TMA=1); otherwise launch and store the status in *rc.
// `gadd` (nullable, batch stride `gadd_bs`): a (B,C,H,W) view added to gf1 in the epilogue -- the gradient of the concat
// buffer's `c1` slice, so that d/d c1 of cat((corr(c1, .), c1, .)) needs no separate add (pwc_tf.py:122-123).
bool fwd_tma(const float* f1, long long f1_bs, const float* f2, float* out, int B, int C, int H, int W, long long out_bs,
             cudaStream_t stream, int* rc);
bool bwd_tma(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
             long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, cudaStream_t stream, int* rc);
// Smallest levels (<= 64 pixel quads per image), cost_volume_small.cu.  Same contract.
bool bwd_small(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
               long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, cudaStream_t stream, int* rc);

// Tensor-core (tcgen05 + TMEM) backward, cost_volume_tc.cu.  Same contract as bwd_tma.
bool bwd_tc(const float* gout, long long gout_bs, const float* f1, const float* f2, float* gf1, float* gf2, int B, int C,
            int H, int W, cudaStream_t stream, int* rc);

}  // namespace cv
}  // namespace uof
