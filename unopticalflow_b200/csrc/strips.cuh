// Work decomposition shared by the "marching warp" kernels (photo loss, SSIM map, smoothness):
// every warp owns a strip of 32 image columns (32 - 2*HALO of them outputs) x `rows` output rows of
// one sample of one pyramid level; a launch covers all levels via a prefix table.
#pragma once
#include "common.cuh"

namespace uof {

struct StripTable {
  int H[UOF_MAX_LEVELS], W[UOF_MAX_LEVELS];
  int warp_begin[UOF_MAX_LEVELS + 1];
  int strips_x[UOF_MAX_LEVELS], strips_y[UOF_MAX_LEVELS];
  int rows;
  int nlevels, B;
};

struct Strip {
  int level, b, col, y0, y1;
};

template <int HALO>
__device__ __forceinline__ bool locate_strip(const StripTable& T, int gw, int lane, Strip& s) {
  if (gw >= T.warp_begin[T.nlevels]) return false;
  int l = 0;
  while (l + 1 < T.nlevels && gw >= T.warp_begin[l + 1]) ++l;
  const int local = gw - T.warp_begin[l];
  const int sx = local % T.strips_x[l];
  const int sy = (local / T.strips_x[l]) % T.strips_y[l];
  s.level = l;
  s.b = local / (T.strips_x[l] * T.strips_y[l]);
  s.col = sx * (32 - 2 * HALO) - HALO + lane;
  s.y0 = sy * T.rows;
  s.y1 = min(s.y0 + T.rows, T.H[l]);
  return true;
}

// Host side: choose the strip height so that the launch has >= ~12 warps per SM, then lay out the
// per-level prefix table.  Returns the total number of warps (strips), or -1 if it overflows.
inline long long build_strip_table(StripTable& T, const int* H, const int* W, int nlevels, int B, int halo) {
  T.nlevels = nlevels;
  T.B = B;
  long long px = 0;
  for (int l = 0; l < nlevels; ++l) px += (long long)B * H[l] * W[l];
  const int outw = 32 - 2 * halo;
  int rows = 32;
  while (rows > 8 && px / ((long long)outw * rows) < 12ll * kNumSMs) rows >>= 1;
  T.rows = rows;
  long long total = 0;
  for (int l = 0; l < nlevels; ++l) {
    T.H[l] = H[l];
    T.W[l] = W[l];
    T.strips_x[l] = ceil_div(W[l], outw);
    T.strips_y[l] = ceil_div(H[l], rows);
    T.warp_begin[l] = (int)total;
    total += (long long)T.strips_x[l] * T.strips_y[l] * B;
    if (total >= (1ll << 30)) return -1;
  }
  T.warp_begin[nlevels] = (int)total;
  return total;
}

}  // namespace uof
