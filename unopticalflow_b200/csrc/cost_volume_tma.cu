// a1 (Blackwell path): cost volume forward/backward with TMA staging and an mbarrier pipeline.
// Replaces PWC_tf.corr_naive, /root/reference/core/networks/structures/pwc_tf.py:97-106.
//
// Compared with the cp.async kernels in cost_volume.cu (kept for W % 4 != 0):
//   * one elected thread issues two cp.async.bulk.tensor.4d loads per 8-channel slab straight from the NCHW
//     tensors; the box origin is signed and out-of-bounds elements (halo, channel tail) are zero-filled by the
//     TMA unit, so padding costs no instructions (the cp.async staging loop was ~30 % of issued instructions);
//   * slabs move through a 3-stage ring guarded by full/empty mbarriers; consumer warps never meet at a CTA
//     barrier in the forward kernel and meet once per slab in the backward kernel (double-buffered reduction);
//   * the 256-pixel CTA tile is either 8x32 or 16x16, chosen per shape so that the grid fills whole waves of
//     the 148 SMs (e.g. W = 208: 16x16 tiles divide the row exactly, 2.8 waves instead of 3.03 -> 4).
// Thread/register layout is the one described in cost_volume.cu: 192 threads = 64 pixel quads x 3 dy-groups,
// 108 accumulators (forward) or 108 coefficients (backward) per thread, 12-float sliding register window.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "cost_volume.h"
#include "tma.cuh"

namespace uof {
namespace cv {
namespace {

constexpr int NSTAGE = 3;      // backward: 3 slabs + two reduction buffers = 110 KB per CTA, 2 CTAs per SM
constexpr int NSTAGE_F = 3;    // forward; a 4th stage (106-112 KB per CTA, still 2 CTAs per SM) measured no gain (41.4 vs 41.3 us)
constexpr int NWARP = NT / 32;

template <int TH_, int TW_>
struct Tile {
  static constexpr int TH = TH_, TW = TW_;
  static constexpr int HTH = TH + 2 * RAD, HTW = TW + 2 * RAD;
  static constexpr int QX = TW / PX;                  // pixel quads per tile row
  static constexpr int QUADS = QX * TH;               // 64
  static constexpr int S1 = CK * TH * TW;             // floats of the f1 slab
  static constexpr int S2 = CK * HTH * HTW;           // floats of the f2 halo slab
  static constexpr int kRed = NGROUP * CK * TH * TW;  // floats of one backward reduction buffer
  static_assert(QUADS * NGROUP == NT, "tile must hold 256 pixels");
  // Row owned by a thread.  For 16-wide tiles the halo rows are 24 floats apart, so the two rows served by one
  // quarter-warp must be 2 apart (48 floats = 16 banks) to keep the LDS.128 window loads conflict-free:
  // swap the two low bits of the row slot.
  __device__ static __forceinline__ int row_of(int slot) {
    return TW == 16 ? ((slot & ~3) | ((slot & 1) << 1) | ((slot >> 1) & 1)) : slot;
  }
};

// Round the dynamic shared-memory base up to 128 B (TMA destination alignment) with POINTER arithmetic on the
// __shared__ array: going through uintptr_t hid the address space from the compiler, and every window load of the
// inner loop became a generic LD.E.128 with 64-bit address math and long-scoreboard stalls (ncu source page,
// profiles/r2_cv_generic_ld.txt) instead of LDS.128.
__device__ __forceinline__ float* align128(unsigned char* p) {
  return reinterpret_cast<float*>(p + ((128u - (smem_u32(p) & 127u)) & 127u));
}

// ------------------------------------------------------------------------------------------- forward
// CLUSTER: the `ksplit` CTAs that share a tile (split over channels) form a thread-block cluster (1,1,ksplit); their
// partial accumulators are combined through distributed shared memory by the cluster's rank-0 CTA, which alone writes the
// tile -- no fp32 atomics into a zero-filled output (the non-cluster split-K path, kept for ksplit > 8).
constexpr int kStashFloats = DYG * ND * PX * NT;          // one CTA's accumulators: 108 x 192 floats = 81 KB

template <class T, bool CLUSTER>
__global__ void __launch_bounds__(NT, 2)
cost_volume_fwd_tma_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                           float* __restrict__ out, int C, int H, int W, long long out_bs, int ksplit, float inv_c) {
  constexpr int TH = T::TH, TW = T::TW, HTH = T::HTH, HTW = T::HTW, S1 = T::S1, S2 = T::S2;
  constexpr int kStage = S1 + S2;
  constexpr int kBarOff = (CLUSTER && kStashFloats > NSTAGE_F * kStage) ? kStashFloats : NSTAGE_F * kStage;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* smem = align128(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* empty = full + NSTAGE_F;

  const int tid = threadIdx.x, lane = tid & 31;
  const int gx = tid % T::QX, ty = T::row_of((tid / T::QX) % TH), dg = tid / T::QUADS;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int b = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  const int nchunks = (C + CK - 1) / CK;
  const int per = (nchunks + ksplit - 1) / ksplit;
  const int k_begin = ks * per, k_end = min(nchunks, k_begin + per);
  if (!CLUSTER && k_begin >= k_end) return;
  const int n = max(k_end - k_begin, 0);              // cluster: a CTA without channels still joins the reduction (zeros)

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE_F; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NWARP);
    }
    mbar_fence_init();
    tma_prefetch_desc(&map1);
    tma_prefetch_desc(&map2);
  }
  __syncthreads();

  auto issue = [&](int j) {
    const int s = j % NSTAGE_F;
    float* dst = smem + s * kStage;
    mbar_expect_tx(full + s, kStage * (unsigned)sizeof(float));
    tma_load_4d(dst, &map1, full + s, x0, y0, (k_begin + j) * CK, b);
    tma_load_4d(dst + S1, &map2, full + s, x0 - RAD, y0 - RAD, (k_begin + j) * CK, b);
  };
  if (tid == 0)
    for (int j = 0; j < min(NSTAGE_F - 1, n); ++j) issue(j);

  float acc[DYG][ND][PX];
#pragma unroll
  for (int r = 0; r < DYG; ++r)
#pragma unroll
    for (int j = 0; j < ND; ++j)
#pragma unroll
      for (int p = 0; p < PX; ++p) acc[r][j][p] = 0.0f;

  for (int j = 0; j < n; ++j) {
    if (tid == 0) {   // refill the stage consumed one iteration ago
      const int jj = j + NSTAGE_F - 1;
      if (jj < n) {
        if (j >= 1) mbar_wait(empty + (j - 1) % NSTAGE_F, ((j - 1) / NSTAGE_F) & 1);
        issue(jj);
      }
    }
    const int s = j % NSTAGE_F;
    mbar_wait(full + s, (j / NSTAGE_F) & 1);
    const float* a_base = smem + s * kStage + ty * TW + PX * gx;
    const float* w_base = smem + s * kStage + S1 + (ty + dg * DYG) * HTW + PX * gx;
#pragma unroll 2
    for (int cc = 0; cc < CK; ++cc) {
      const float4 a4 = *reinterpret_cast<const float4*>(a_base + cc * TH * TW);
      const float a[PX] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int r = 0; r < DYG; ++r) {
        const float4* wp = reinterpret_cast<const float4*>(w_base + cc * HTH * HTW + r * HTW);
        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int jd = 0; jd < ND; ++jd)
#pragma unroll
          for (int p = 0; p < PX; ++p) acc[r][jd][p] = fmaf(a[p], win[p + jd], acc[r][jd][p]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }

  const int y = y0 + ty, x = x0 + PX * gx;
  if (CLUSTER) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();          // == ks: the cluster spans the split
    float* stash = smem;                                 // [accumulator][thread]; the stage buffers are free now
    __syncthreads();
    if (rank != 0) {
#pragma unroll
      for (int r = 0; r < DYG; ++r)
#pragma unroll
        for (int jd = 0; jd < ND; ++jd)
#pragma unroll
          for (int p = 0; p < PX; ++p) stash[((r * ND + jd) * PX + p) * NT + tid] = acc[r][jd][p];
    }
    cluster.sync();
    if (rank == 0) {
      for (unsigned rr = 1; rr < (unsigned)ksplit; ++rr) {
        const float* remote = cluster.map_shared_rank(stash, rr);
#pragma unroll
        for (int r = 0; r < DYG; ++r)
#pragma unroll
          for (int jd = 0; jd < ND; ++jd)
#pragma unroll
            for (int p = 0; p < PX; ++p) acc[r][jd][p] += remote[((r * ND + jd) * PX + p) * NT + tid];
      }
      if (y < H && x < W) {
        float* ob = out + (size_t)b * out_bs + (size_t)y * W + x;
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int r = 0; r < DYG; ++r)
#pragma unroll
          for (int jd = 0; jd < ND; ++jd)
            *reinterpret_cast<float4*>(ob + (size_t)((dg * DYG + r) * ND + jd) * plane) =
                make_float4(acc[r][jd][0] * inv_c, acc[r][jd][1] * inv_c, acc[r][jd][2] * inv_c, acc[r][jd][3] * inv_c);
      }
    }
    cluster.sync();                                      // the other CTAs' shared memory stays alive until rank 0 has read it
    return;
  }
  if (y >= H || x >= W) return;
  float* ob = out + (size_t)b * out_bs + (size_t)y * W + x;
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int r = 0; r < DYG; ++r) {
#pragma unroll
    for (int jd = 0; jd < ND; ++jd) {
      float* o = ob + (size_t)((dg * DYG + r) * ND + jd) * plane;
      if (ksplit == 1) {
        *reinterpret_cast<float4*>(o) =
            make_float4(acc[r][jd][0] * inv_c, acc[r][jd][1] * inv_c, acc[r][jd][2] * inv_c, acc[r][jd][3] * inv_c);
      } else {
#pragma unroll
        for (int p = 0; p < PX; ++p) atomicAdd(o + p, acc[r][jd][p] * inv_c);
      }
    }
  }
}

// ----------------------------------------------------------------------------------- forward, persistent
// Same math and tiling, but a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (grid = 2 CTAs per SM) and the
// slab pipeline runs across tile boundaries: the producer keeps issuing the next tile's TMA loads while the consumers
// finish the current tile and write it out, so the per-tile pipeline fill (barrier init, descriptor prefetch, first TMA
// round trip -- ~10-20 % of a 4-slab tile at C = 32, during which the SM's other CTA computes alone with 1.5 warps per
// scheduler) is paid once per CTA instead of once per tile.  Used when the level has more tiles than resident CTAs and
// no split-K.
template <class T>
__global__ void __launch_bounds__(NT, 2)
cost_volume_fwd_tma_persistent_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                                      float* __restrict__ out, int C, int H, int W, long long out_bs, float inv_c,
                                      int tiles_x, int tiles_y, int total_tiles) {
  constexpr int TH = T::TH, TW = T::TW, HTH = T::HTH, HTW = T::HTW, S1 = T::S1, S2 = T::S2;
  constexpr int kStage = S1 + S2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* smem = align128(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE_F * kStage);
  uint64_t* empty = full + NSTAGE_F;

  const int tid = threadIdx.x, lane = tid & 31;
  const int gx = tid % T::QX, ty = T::row_of((tid / T::QX) % TH), dg = tid / T::QUADS;
  const int n = (C + CK - 1) / CK;                                   // slabs per tile
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * n;                                    // slabs this CTA consumes
  if (total <= 0) return;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE_F; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NWARP);
    }
    mbar_fence_init();
    tma_prefetch_desc(&map1);
    tma_prefetch_desc(&map2);
  }
  __syncthreads();

  // tile index -> origin; tiles of one image are consecutive (x fastest) so neighbouring CTAs share halos in L2
  auto origin = [&](int t, int& x0, int& y0, int& b) {
    const int bx = t % tiles_x;
    const int r = t / tiles_x;
    x0 = bx * TW;
    y0 = (r % tiles_y) * TH;
    b = r / tiles_y;
  };
  auto issue = [&](int q) {                                            // q-th slab of this CTA's sequence
    int x0, y0, b;
    origin((int)blockIdx.x + (q / n) * (int)gridDim.x, x0, y0, b);
    const int j = q % n, s = q % NSTAGE_F;
    float* dst = smem + s * kStage;
    mbar_expect_tx(full + s, kStage * (unsigned)sizeof(float));
    tma_load_4d(dst, &map1, full + s, x0, y0, j * CK, b);
    tma_load_4d(dst + S1, &map2, full + s, x0 - RAD, y0 - RAD, j * CK, b);
  };
  if (tid == 0)
    for (int q = 0; q < min(NSTAGE_F - 1, total); ++q) issue(q);

  float acc[DYG][ND][PX];
  int q = 0;
  for (int it = 0; it < my_tiles; ++it) {
#pragma unroll
    for (int r = 0; r < DYG; ++r)
#pragma unroll
      for (int j = 0; j < ND; ++j)
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[r][j][p] = 0.0f;

    for (int j = 0; j < n; ++j, ++q) {
      if (tid == 0) {   // refill the stage consumed one iteration ago (possibly with a slab of the NEXT tile)
        const int qq = q + NSTAGE_F - 1;
        if (qq < total) {
          if (q >= 1) mbar_wait(empty + (q - 1) % NSTAGE_F, ((q - 1) / NSTAGE_F) & 1);
          issue(qq);
        }
      }
      const int s = q % NSTAGE_F;
      mbar_wait(full + s, (q / NSTAGE_F) & 1);
      const float* a_base = smem + s * kStage + ty * TW + PX * gx;
      const float* w_base = smem + s * kStage + S1 + (ty + dg * DYG) * HTW + PX * gx;
#pragma unroll 2
      for (int cc = 0; cc < CK; ++cc) {
        const float4 a4 = *reinterpret_cast<const float4*>(a_base + cc * TH * TW);
        const float a[PX] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int r = 0; r < DYG; ++r) {
          const float4* wp = reinterpret_cast<const float4*>(w_base + cc * HTH * HTW + r * HTW);
          const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
          const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
          for (int jd = 0; jd < ND; ++jd)
#pragma unroll
            for (int p = 0; p < PX; ++p) acc[r][jd][p] = fmaf(a[p], win[p + jd], acc[r][jd][p]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }

    int x0, y0, b;
    origin((int)blockIdx.x + it * (int)gridDim.x, x0, y0, b);
    const int y = y0 + ty, x = x0 + PX * gx;
    if (y < H && x < W) {
      float* ob = out + (size_t)b * out_bs + (size_t)y * W + x;
      const size_t plane = (size_t)H * W;
#pragma unroll
      for (int r = 0; r < DYG; ++r)
#pragma unroll
        for (int jd = 0; jd < ND; ++jd)
          *reinterpret_cast<float4*>(ob + (size_t)((dg * DYG + r) * ND + jd) * plane) =
              make_float4(acc[r][jd][0] * inv_c, acc[r][jd][1] * inv_c, acc[r][jd][2] * inv_c, acc[r][jd][3] * inv_c);
    }
  }
}

// ------------------------------------------------------------------------------------------ backward
// blockIdx.y selects the role:
//   0: gf1[c,p] = 1/C * sum_d gout[d,p]          * f2[c,p+d]
//   1: gf2[c,q] = 1/C * sum_d gout[flip(d),q+d]  * f1[c,q+d]     (d -> -d re-indexed as flip)
template <class T, bool ADD>
__global__ void __launch_bounds__(NT, 2)
cost_volume_bwd_tma_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                           const float* __restrict__ gout, long long gout_bs, const float* __restrict__ gadd,
                           long long gadd_bs, float* __restrict__ gf1, float* __restrict__ gf2, int C, int H, int W,
                           int tiles_x, int csplit, float inv_c) {
  constexpr int TH = T::TH, TW = T::TW, HTH = T::HTH, HTW = T::HTW, S2 = T::S2, kRed = T::kRed;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* smem = align128(smem_raw);
  float* red = smem + NSTAGE * S2;                       // 2 x [NGROUP][CK][TH][TW]
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * kRed);
  uint64_t* empty = full + NSTAGE;

  const int tid = threadIdx.x, lane = tid & 31;
  const int gx = tid % T::QX, ty = T::row_of((tid / T::QX) % TH), dg = tid / T::QUADS;
  const int x0 = (blockIdx.x % tiles_x) * TW, y0 = (blockIdx.x / tiles_x) * TH;
  const bool mirror = blockIdx.y != 0;
  const int b = blockIdx.z / csplit, cs = blockIdx.z % csplit;
  const int nchunks = (C + CK - 1) / CK;
  const int per = (nchunks + csplit - 1) / csplit;
  const int k_begin = cs * per, k_end = min(nchunks, k_begin + per);
  if (k_begin >= k_end) return;
  const int n = k_end - k_begin;
  const CUtensorMap* src_map = mirror ? &map1 : &map2;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NWARP);
    }
    mbar_fence_init();
    tma_prefetch_desc(src_map);
  }
  __syncthreads();

  auto issue = [&](int j) {
    const int s = j % NSTAGE;
    mbar_expect_tx(full + s, S2 * (unsigned)sizeof(float));
    tma_load_4d(smem + s * S2, src_map, full + s, x0 - RAD, y0 - RAD, (k_begin + j) * CK, b);
  };
  if (tid == 0)
    for (int j = 0; j < min(NSTAGE - 1, n); ++j) issue(j);

  float* dst_b = (mirror ? gf2 : gf1) + (size_t)b * C * H * W;
  const float* gb = gout + (size_t)b * gout_bs;
  // ADD: the concat-slice gradient `gadd` is folded into gf1 (role 0 only)
  const float* add_b = (ADD && !mirror) ? gadd + (size_t)b * gadd_bs : nullptr;
  const size_t plane = (size_t)H * W;
  const int y = y0 + ty, x = x0 + PX * gx;

  float kc[DYG][ND][PX];
#pragma unroll
  for (int r = 0; r < DYG; ++r) {
    const int i = dg * DYG + r;
#pragma unroll
    for (int jd = 0; jd < ND; ++jd) {
      if (!mirror) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y < H && x < W) v = __ldg(reinterpret_cast<const float4*>(gb + (size_t)(i * ND + jd) * plane + (size_t)y * W + x));
        kc[r][jd][0] = v.x; kc[r][jd][1] = v.y; kc[r][jd][2] = v.z; kc[r][jd][3] = v.w;
      } else {
#pragma unroll
        for (int p = 0; p < PX; ++p) {
          const int ys = y + i - RAD, xs = x + p + jd - RAD;
          float v = 0.0f;
          if (y < H && x + p < W && ys >= 0 && ys < H && xs >= 0 && xs < W)
            v = __ldg(gb + (size_t)((ND - 1 - i) * ND + (ND - 1 - jd)) * plane + (size_t)ys * W + xs);
          kc[r][jd][p] = v;
        }
      }
    }
  }

  for (int j = 0; j < n; ++j) {
    if (tid == 0) {
      const int jj = j + NSTAGE - 1;
      if (jj < n) {
        if (j >= 1) mbar_wait(empty + (j - 1) % NSTAGE, ((j - 1) / NSTAGE) & 1);
        issue(jj);
      }
    }
    const int s = j % NSTAGE;
    mbar_wait(full + s, (j / NSTAGE) & 1);
    const float* w_base = smem + s * S2 + (ty + dg * DYG) * HTW + PX * gx;
    float* rbuf = red + (j & 1) * kRed;
    const int k = k_begin + j;
    // the slice gradient of this thread's output elements is fetched at the top of the slab, so its latency hides behind
    // the slab's FMAs (loaded inside the reduction loop it cost 19 us at 16x32x64x208: every slab ended on an exposed L2
    // round trip)
    constexpr int kPerThread = (CK * TH * (TW / 4) + NT - 1) / NT;
    float4 addv[ADD ? kPerThread : 1];
    if (ADD) {
#pragma unroll
      for (int u = 0; u < kPerThread; ++u) {
        const int e = tid + u * NT;
        const int q = e % (TW / 4), r = (e / (TW / 4)) % TH, cc = e / ((TW / 4) * TH);
        const int c = k * CK + cc, yy = y0 + r, xx = x0 + 4 * q;
        addv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (add_b && e < CK * TH * (TW / 4) && c < C && yy < H && xx < W)
          addv[u] = __ldg(reinterpret_cast<const float4*>(add_b + ((size_t)c * H + yy) * W + xx));
      }
    }

#pragma unroll
    for (int cc = 0; cc < CK; ++cc) {
      float part[PX] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int r = 0; r < DYG; ++r) {
        const float4* wp = reinterpret_cast<const float4*>(w_base + cc * HTH * HTW + r * HTW);
        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int jd = 0; jd < ND; ++jd)
#pragma unroll
          for (int p = 0; p < PX; ++p) part[p] = fmaf(kc[r][jd][p], win[p + jd], part[p]);
      }
      *reinterpret_cast<float4*>(rbuf + ((dg * CK + cc) * TH + ty) * TW + PX * gx) =
          make_float4(part[0], part[1], part[2], part[3]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
    __syncthreads();   // partial sums of the three dy-groups are in rbuf (double-buffered: one barrier per slab)

#pragma unroll
    for (int u = 0; u < kPerThread; ++u) {
      const int e = tid + u * NT;
      if (e >= CK * TH * (TW / 4)) break;
      const int q = e % (TW / 4), r = (e / (TW / 4)) % TH, cc = e / ((TW / 4) * TH);
      const int c = k * CK + cc, yy = y0 + r, xx = x0 + 4 * q;
      if (c >= C || yy >= H || xx >= W) continue;
      const float* rp = rbuf + (cc * TH + r) * TW + 4 * q;
      float4 sum = *reinterpret_cast<const float4*>(rp);
#pragma unroll
      for (int g = 1; g < NGROUP; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(rp + g * CK * TH * TW);
        sum.x += t.x; sum.y += t.y; sum.z += t.z; sum.w += t.w;
      }
      float4 res = make_float4(sum.x * inv_c, sum.y * inv_c, sum.z * inv_c, sum.w * inv_c);
      if (ADD) {
        res.x += addv[u].x; res.y += addv[u].y; res.z += addv[u].z; res.w += addv[u].w;
      }
      *reinterpret_cast<float4*>(dst_b + ((size_t)c * H + yy) * W + xx) = res;
    }
  }
}

template <class T>
constexpr size_t fwd_smem() {
  return NSTAGE_F * (T::S1 + T::S2) * sizeof(float) + 2 * NSTAGE_F * sizeof(uint64_t) + 128;
}
template <class T>
constexpr size_t fwd_cluster_smem() {
  constexpr size_t stage = NSTAGE_F * (T::S1 + T::S2), stash = kStashFloats;
  return (stage > stash ? stage : stash) * sizeof(float) + 2 * NSTAGE_F * sizeof(uint64_t) + 128;
}
template <class T>
constexpr size_t bwd_smem() {
  return (NSTAGE * T::S2 + 2 * T::kRed) * sizeof(float) + 2 * NSTAGE * sizeof(uint64_t) + 128;
}

bool tma_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("UOF_DISABLE_TMA");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Cost of a tiling = waves of the grid over the 2-CTA-per-SM slots x pixels staged per CTA (tile + halo).
// Returns true when 16x16 tiles beat 8x32 tiles for this shape.
bool prefer_square(int H, int W, long long ctas_per_tile) {
  const char* force = getenv("UOF_CV_TILE");
  if (force && force[0] == 's') return true;
  if (force && force[0] == 'w') return false;
  const long long slots = 2ll * kNumSMs;
  auto cost = [&](int th, int tw) {
    const long long ctas = (long long)ceil_div(W, tw) * ceil_div(H, th) * ctas_per_tile;
    const long long waves = ceil_div_ll(ctas, slots);
    return (double)waves * (double)((th + 2 * RAD) * (tw + 2 * RAD) + th * tw);
  };
  return cost(16, 16) < 0.9 * cost(8, 32);   // square tiles pay ~10 % in shared-memory bank conflicts
}

template <class T>
int launch_fwd(const float* f1, long long f1_bs, const float* f2, float* out, int B, int C, int H, int W, long long out_bs,
               cudaStream_t stream) {
  CUtensorMap m1, m2;
  if (!make_nchw_map(&m1, f1, B, C, H, W, T::TW, T::TH, CK, f1_bs) || !make_nchw_map(&m2, f2, B, C, H, W, T::HTW, T::HTH, CK))
    return -1;
  const int tx = ceil_div(W, T::TW), ty = ceil_div(H, T::TH);
  int ksplit = pick_ksplit_atomic((long long)tx * ty * B, ceil_div(C, CK));
  {
    static const char* force = getenv("UOF_CV_KSPLIT");
    if (force) ksplit = max(1, min(atoi(force), ceil_div(C, CK)));
  }
  UOF_REQUIRE((long long)B * ksplit <= 65535 && ty <= 65535, "cost_volume_fwd: grid too large");
  static const bool no_cluster = getenv("UOF_CV_NO_CLUSTER") != nullptr;
  if (ksplit >= 2 && ksplit <= 8 && !no_cluster) {     // split-K combined through distributed shared memory: no atomics, no zero-fill
    auto ck = cost_volume_fwd_tma_kernel<T, true>;
    UOF_CUDA(cudaFuncSetAttribute(ck, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_cluster_smem<T>()));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tx, ty, B * ksplit);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = fwd_cluster_smem<T>();
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = (unsigned)ksplit;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const float inv_c = 1.0f / (float)C;
    UOF_CUDA(cudaLaunchKernelEx(&cfg, ck, m1, m2, out, C, H, W, out_bs, ksplit, inv_c));
    count_launch();
    return check_launch("cost_volume_fwd (tma, cluster split-K)");
  }
  if (ksplit > 1)
    UOF_CUDA(cudaMemset2DAsync(out, out_bs * sizeof(float), 0, (size_t)UOF_NUM_DISPLACEMENTS * H * W * sizeof(float), B, stream));
  const long long tiles = (long long)tx * ty * B;
  static const bool no_persist = getenv("UOF_CV_NO_PERSIST") != nullptr;
  if (ksplit == 1 && tiles > 2 * kNumSMs && tiles < (1ll << 30) && !no_persist) {
    auto pk = cost_volume_fwd_tma_persistent_kernel<T>;
    UOF_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<T>()));
    pk<<<2 * kNumSMs, NT, fwd_smem<T>(), stream>>>(m1, m2, out, C, H, W, out_bs, 1.0f / (float)C, tx, ty, (int)tiles);
    count_launch();
    return check_launch("cost_volume_fwd (tma, persistent)");
  }
  auto kern = cost_volume_fwd_tma_kernel<T, false>;
  UOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<T>()));
  kern<<<dim3(tx, ty, B * ksplit), NT, fwd_smem<T>(), stream>>>(m1, m2, out, C, H, W, out_bs, ksplit, 1.0f / (float)C);
  count_launch();
  return check_launch("cost_volume_fwd (tma)");
}

template <class T>
int launch_bwd(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
               long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, cudaStream_t stream) {
  CUtensorMap m1, m2;
  if (!make_nchw_map(&m1, f1, B, C, H, W, T::HTW, T::HTH, CK, f1_bs) || !make_nchw_map(&m2, f2, B, C, H, W, T::HTW, T::HTH, CK))
    return -1;
  const int tx = ceil_div(W, T::TW), ty = ceil_div(H, T::TH);
  const int csplit = pick_split((long long)tx * ty * B * 2, ceil_div(C, CK));
  UOF_REQUIRE((long long)B * csplit <= 65535, "cost_volume_bwd: grid too large");
  auto kern = gadd ? cost_volume_bwd_tma_kernel<T, true> : cost_volume_bwd_tma_kernel<T, false>;
  UOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem<T>()));
  kern<<<dim3(tx * ty, 2, B * csplit), NT, bwd_smem<T>(), stream>>>(m1, m2, gout, gout_bs, gadd, gadd_bs, gf1, gf2, C, H, W,
                                                                    tx, csplit, 1.0f / (float)C);
  count_launch();
  return check_launch("cost_volume_bwd (tma)");
}

}  // namespace

bool fwd_tma(const float* f1, long long f1_bs, const float* f2, float* out, int B, int C, int H, int W, long long out_bs,
             cudaStream_t stream, int* rc) {
  if (!tma_enabled() || W % 4 != 0 || out_bs % 4 != 0 || f1_bs % 4 != 0 || !aligned16(f1) || !aligned16(f2) || !aligned16(out))
    return false;
  const int r = prefer_square(H, W, B) ? launch_fwd<Tile<16, 16>>(f1, f1_bs, f2, out, B, C, H, W, out_bs, stream)
                                       : launch_fwd<Tile<8, 32>>(f1, f1_bs, f2, out, B, C, H, W, out_bs, stream);
  if (r < 0) return false;    // tensor-map encoding unavailable: let the caller fall back
  *rc = r;
  return true;
}

bool bwd_tma(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
             long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, cudaStream_t stream, int* rc) {
  if (!tma_enabled() || W % 4 != 0 || gout_bs % 4 != 0 || f1_bs % 4 != 0 || !aligned16(f1) || !aligned16(f2) ||
      !aligned16(gf1) || !aligned16(gf2) || !aligned16(gout) || (gadd && (!aligned16(gadd) || gadd_bs % 4 != 0)))
    return false;
  // measured (profiles/): the backward kernel is faster with 8x32 tiles at every level of the 256x832 pyramid
  // (its reduction buffer and coefficient gathers favour wide rows); the square variant is kept behind UOF_CV_TILE=s
  const char* force = getenv("UOF_CV_TILE");
  const int r = (force && force[0] == 's')
                    ? launch_bwd<Tile<16, 16>>(gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, B, C, H, W, stream)
                    : launch_bwd<Tile<8, 32>>(gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, B, C, H, W, stream);
  if (r < 0) return false;
  *rc = r;
  return true;
}

}  // namespace cv
}  // namespace uof
