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

// PPL = pixels per lane: a lane owns PPL adjacent columns, s.col is its first one.
template <int HALO, int PPL = 1>
__device__ __forceinline__ bool locate_strip(const StripTable& T, int gw, int lane, Strip& s) {
  if (gw >= T.warp_begin[T.nlevels]) return false;
  int l = 0;
  while (l + 1 < T.nlevels && gw >= T.warp_begin[l + 1]) ++l;
  const int local = gw - T.warp_begin[l];
  const int sx = local % T.strips_x[l];
  const int sy = (local / T.strips_x[l]) % T.strips_y[l];
  s.level = l;
  s.b = local / (T.strips_x[l] * T.strips_y[l]);
  s.col = sx * (32 * PPL - 2 * HALO) - HALO + lane * PPL;
  s.y0 = sy * T.rows;
  s.y1 = min(s.y0 + T.rows, T.H[l]);
  return true;
}

// Host side: choose the strip height and lay out the per-level prefix table.
// The launch runs `grid_mult` x ceil(warps / warps_per_block) blocks on 148 * blocks_per_sm resident slots; the strip
// height is the one that minimises (number of waves) x (rows + 2*halo), i.e. whole waves with the least halo
// re-computation (a 1.3-wave launch costs as much as a 2-wave one).  Returns the number of warps, or -1 on overflow.
// `halo` is the row halo (and the column halo unless `col_halo` >= 0 is given: the quad kernels use whole lanes).
inline long long build_strip_table(StripTable& T, const int* H, const int* W, int nlevels, int B, int halo,
                                   int grid_mult = 1, int blocks_per_sm = 4, int warps_per_block = 4, int ppl = 1,
                                   int col_halo = -1) {
  T.nlevels = nlevels;
  T.B = B;
  const int outw = 32 * ppl - 2 * (col_halo >= 0 ? col_halo : halo);
  const long long slots = (long long)kNumSMs * (blocks_per_sm > 0 ? blocks_per_sm : 1);
  int best_rows = 8;
  double best_cost = 1e300;
  for (int rows = 8; rows <= 64; rows += 4) {
    long long warps = 0;
    for (int l = 0; l < nlevels; ++l) warps += (long long)ceil_div(W[l], outw) * ceil_div(H[l], rows) * B;
    const long long blocks = ceil_div_ll(warps, warps_per_block) * grid_mult;
    const double cost = (double)ceil_div_ll(blocks, slots) * (double)(rows + 2 * halo);
    if (cost <= best_cost) {
      best_cost = cost;
      best_rows = rows;
    }
  }
  T.rows = best_rows;
  long long total = 0;
  for (int l = 0; l < nlevels; ++l) {
    T.H[l] = H[l];
    T.W[l] = W[l];
    T.strips_x[l] = ceil_div(W[l], outw);
    T.strips_y[l] = ceil_div(H[l], best_rows);
    T.warp_begin[l] = (int)total;
    total += (long long)T.strips_x[l] * T.strips_y[l] * B;
    if (total >= (1ll << 30)) return -1;
  }
  T.warp_begin[nlevels] = (int)total;
  return total;
}

// Resident blocks per SM of a kernel (cached by the caller in a static).
template <class K>
inline int resident_blocks(K kernel, int threads) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) {
    (void)cudaGetLastError();
    n = 4;
  }
  return n;
}

}  // namespace uof
