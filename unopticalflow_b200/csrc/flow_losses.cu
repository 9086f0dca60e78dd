// a7 + a8: edge-aware second-order smoothness and flow-direction consistency losses.
// Replaces, in /root/reference/core/networks/model_flow_paper.py, gradients / cal_grad2_error /
// compute_loss_flow_smooth (:152-177, 154 slice/sub launches per direction) and
// get_flow_normalization / compute_loss_flow_consis (:44-51,180-195, 50 launches).
//
// Smoothness uses the marching-warp scheme (strips.cuh): a lane loads the 2 flow + 3 image values of
// its pixel once per row, x-neighbours come from warp shuffles, y-neighbours from a 3-row register
// ring; all pyramid levels and both flow directions (flow batch B, image batch Bimg, image index
// b % Bimg) go in one launch.  Consistency is a pure streaming kernel.
#include <stdlib.h>

#include "strips.cuh"

namespace uof {
namespace {

constexpr int kWarpsPerBlock = 4;
constexpr float kEps = 1e-12f;

struct SmoothParams {
  uof_smooth_level lv[UOF_MAX_LEVELS];
  StripTable T;
  int Bimg;
  int accumulate;      // backward: gflow += instead of gflow = (the *_bwd_acc entry points)
};

__device__ void smooth_finalize(const SmoothParams& P, const float* sums, float* loss);

__device__ __forceinline__ float sgn(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }

// second difference (hi - mid) - (mid - lo) with the reference's roundings (model_flow_paper.py:153-155 applied twice)
__device__ __forceinline__ float d2(float lo, float mid, float hi) { return __fsub_rn(__fsub_rn(hi, mid), __fsub_rn(mid, lo)); }

// exp(-10 * mean_c |a_c - b_c|)   (model_flow_paper.py:159-160)
__device__ __forceinline__ float edge_weight(const float* a, const float* b) {
  return __expf((-10.0f / 3.0f) * (fabsf(a[0] - b[0]) + fabsf(a[1] - b[1]) + fabsf(a[2] - b[2])));
}

// Row pipeline of the smoothness kernels: every warp owns a ring of kDepth rows x 5 planes x 32 lanes in shared
// memory and keeps kDepth-1 rows in flight with cp.async (LDGSTS, 4 bytes per lane, zero-filled outside the image).
// Register prefetching was tried first: ptxas hoists the next row's LDGs above the last use of the registers they
// replace and then copies the just-loaded values, so each row waited for its own prefetch (long-scoreboard stalls on
// MOVs: 39 % of the samples of smooth_fwd in the ncu source page of round 2, whatever the ring shape).  The
// asynchronous copies take the prefetch out of the register allocator's hands and cost no registers.
constexpr int kDepth = 4;
constexpr int kRowVals = 5;     // flow x, flow y, image r, g, b

// start the copy of row r (lane's pixel) into ring slot `slot`, one commit group per row
__device__ __forceinline__ void fetch_row(float* ring, int slot, const float* __restrict__ fb, const float* __restrict__ ib,
                                          size_t plane, int W, int r, int H, int colc, bool col_in) {
  const bool inb = col_in && r >= 0 && r < H;
  const size_t off = (size_t)min(max(r, 0), H - 1) * W + colc;
  float* d = ring + slot * (kRowVals * 32);
  cp_async_4(d, fb + off, inb);
  cp_async_4(d + 32, fb + off + plane, inb);
#pragma unroll
  for (int c = 0; c < 3; ++c) cp_async_4(d + (2 + c) * 32, ib + off + c * plane, inb);
  cp_async_commit();
}

// wait for the oldest row in flight and read it
__device__ __forceinline__ void take_row(const float* ring, int slot, float* v) {
  cp_async_wait<kDepth - 1>();
  const float* d = ring + slot * (kRowVals * 32);
#pragma unroll
  for (int k = 0; k < kRowVals; ++k) v[k] = d[k * 32];
}

// --------------------------------------------------------------------------------- smooth fwd
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
smooth_fwd_kernel(const __grid_constant__ SmoothParams P, float* __restrict__ sums, float* __restrict__ loss) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Strip sc;
  const bool live = locate_strip<1>(P.T, gw, lane, sc);   // idle warps still join the block reduction below
  if (!live) sc.level = sc.b = sc.col = sc.y0 = sc.y1 = 0;
  const uof_smooth_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const size_t plane = (size_t)H * W;
  const float* fb = L.flow + (size_t)sc.b * 2 * plane;
  const float* ib = L.img + (size_t)(sc.b % P.Bimg) * 3 * plane;
  const bool col_in = sc.col >= 0 && sc.col < W;
  const bool col_out = col_in && lane >= 1 && lane <= 30;

  float f[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};   // flow/20 at rows r-2, r-1, r
  float im[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};    // image at rows r-1, r
  float sum_x = 0.0f, sum_y = 0.0f;
  if (live) {

  __shared__ float ring_s[kWarpsPerBlock][kDepth * kRowVals * 32];
  float* ring = ring_s[threadIdx.x >> 5] + lane;
  const int colc = min(max(sc.col, 0), W - 1);
#pragma unroll
  for (int i = 0; i < kDepth - 1; ++i) fetch_row(ring, i, fb, ib, plane, W, sc.y0 - 1 + i, H, colc, col_in);
  int slot = 0;
  for (int r = sc.y0 - 1; r <= sc.y1; ++r) {
    fetch_row(ring, slot == 0 ? kDepth - 1 : slot - 1, fb, ib, plane, W, r + kDepth - 1, H, colc, col_in);
    float nxt[kRowVals];
    take_row(ring, slot, nxt);
    slot = slot + 1 == kDepth ? 0 : slot + 1;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      f[0][k] = f[1][k];
      f[1][k] = f[2][k];
      f[2][k] = __fmul_rn(nxt[k], 0.05f);   // :174 flow/20.0 (no FMA contraction, see take_quad)
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      im[0][c] = im[1][c];
      im[1][c] = nxt[2 + c];
    }
    // x term centred on (r, col): needs col-1 and col+1 from the neighbouring lanes
    float ir[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ir[c] = __shfl_down_sync(kFullMask, im[1][c], 1);
    float d2x = 0.0f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float fl = __shfl_up_sync(kFullMask, f[2][k], 1), fr = __shfl_down_sync(kFullMask, f[2][k], 1);
      d2x += fabsf(d2(fl, f[2][k], fr));   // :153-155,163-164
    }
    if (col_out && r >= sc.y0 && r < sc.y1 && sc.col >= 1 && sc.col <= W - 2) sum_x = fmaf(edge_weight(ir, im[1]), d2x, sum_x);
    // y term centred on row m = r-1: rows m-1, m, m+1 are f[0], f[1], f[2]
    const int m = r - 1;
    if (col_out && m >= sc.y0 && m >= 1 && m <= H - 2) {
      const float d2y = fabsf(d2(f[0][0], f[1][0], f[2][0])) + fabsf(d2(f[0][1], f[1][1], f[2][1]));
      sum_y = fmaf(edge_weight(im[1], im[0]), d2y, sum_y);
    }
  }
  cp_async_wait<0>();
  }
  const float acc[2] = {warp_sum(sum_x), warp_sum(sum_y)};
  block_accumulate<2, kWarpsPerBlock>(acc, live ? sums + ((size_t)sc.level * P.T.B + sc.b) * 2 : nullptr);
  if (last_block_done(reinterpret_cast<unsigned*>(sums + (size_t)P.T.nlevels * P.T.B * 2))) smooth_finalize(P, sums, loss);
}

// Runs in the last block of the forward launch (last_block_done).
__device__ void smooth_finalize(const SmoothParams& P, const float* sums, float* loss) {
  for (int b = threadIdx.x; b < P.T.B; b += blockDim.x) {
    float acc = 0.0f;
    for (int l = 0; l < P.T.nlevels; ++l) {
      const float H = (float)P.lv[l].H, W = (float)P.lv[l].W;
      const float* s = sums + ((size_t)l * P.T.B + b) * 2;
      acc += (__ldcg(s) / (2.0f * H * (W - 2.0f)) + __ldcg(s + 1) / (2.0f * (H - 2.0f) * W)) / 2.0f;   // :165-166
    }
    loss[b] = acc;
  }
}

// --------------------------------------------------------------------------------- smooth bwd
// gflow[p] = ( sx[p-1] - 2 sx[p] + sx[p+1]  +  sy[p-1] - 2 sy[p] + sy[p+1] ) / 20 with
// sx[m] = cx * wx[m] * sign(d2x[m]) for 1 <= m <= W-2 (0 elsewhere), sy likewise along rows.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
smooth_bwd_kernel(const __grid_constant__ SmoothParams P, const float* __restrict__ g_loss) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Strip sc;
  if (!locate_strip<2>(P.T, gw, lane, sc)) return;
  const uof_smooth_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const size_t plane = (size_t)H * W;
  const float* fb = L.flow + (size_t)sc.b * 2 * plane;
  const float* ib = L.img + (size_t)(sc.b % P.Bimg) * 3 * plane;
  float* gb = L.gflow + (size_t)sc.b * 2 * plane;
  const bool col_in = sc.col >= 0 && sc.col < W;
  const bool col_out = col_in && lane >= 2 && lane <= 29;
  const float g = __ldg(g_loss + sc.b);
  const float cx = g * 0.5f / (2.0f * (float)H * ((float)W - 2.0f));
  const float cy = g * 0.5f / (2.0f * ((float)H - 2.0f) * (float)W);

  float f[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};     // flow/20 rows r-2, r-1, r
  float im[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};      // image rows r-1, r
  float sy[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};    // sy rows m-2, m-1, m   (m = r-1)
  float gxr[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};   // x-part of the gradient, rows r-2, r-1, r

  __shared__ float ring_s[kWarpsPerBlock][kDepth * kRowVals * 32];
  float* ring = ring_s[threadIdx.x >> 5] + lane;
  const int colc = min(max(sc.col, 0), W - 1);
#pragma unroll
  for (int i = 0; i < kDepth - 1; ++i) fetch_row(ring, i, fb, ib, plane, W, sc.y0 - 2 + i, H, colc, col_in);
  int slot = 0;
  for (int r = sc.y0 - 2; r <= sc.y1 + 1; ++r) {
    const bool inb = col_in && r >= 0 && r < H;
    fetch_row(ring, slot == 0 ? kDepth - 1 : slot - 1, fb, ib, plane, W, r + kDepth - 1, H, colc, col_in);
    float nxt[kRowVals];
    take_row(ring, slot, nxt);
    slot = slot + 1 == kDepth ? 0 : slot + 1;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      f[0][k] = f[1][k];
      f[1][k] = f[2][k];
      f[2][k] = __fmul_rn(nxt[k], 0.05f);
      gxr[0][k] = gxr[1][k];
      gxr[1][k] = gxr[2][k];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      im[0][c] = im[1][c];
      im[1][c] = nxt[2 + c];
    }
    // x part for row r
    float ir[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ir[c] = __shfl_down_sync(kFullMask, im[1][c], 1);
    const bool mx_ok = inb && sc.col >= 1 && sc.col <= W - 2;
    const float wx = mx_ok ? cx * edge_weight(ir, im[1]) : 0.0f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float fl = __shfl_up_sync(kFullMask, f[2][k], 1), fr = __shfl_down_sync(kFullMask, f[2][k], 1);
      const float s = wx * sgn(d2(fl, f[2][k], fr));
      gxr[2][k] = __shfl_up_sync(kFullMask, s, 1) - 2.0f * s + __shfl_down_sync(kFullMask, s, 1);
    }
    // y part: sy at row m = r-1
    const int m = r - 1;
    const bool my_ok = col_in && m >= 1 && m <= H - 2;
    const float wy = my_ok ? cy * edge_weight(im[1], im[0]) : 0.0f;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      sy[0][k] = sy[1][k];
      sy[1][k] = sy[2][k];
      sy[2][k] = wy * sgn(d2(f[0][k], f[1][k], f[2][k]));
    }
    // output row p = r-2: sy rows p-1, p, p+1 are sy[0..2]; x part is gxr[0]
    const int p = r - 2;
    if (p >= sc.y0 && p < sc.y1 && col_out) {
      const size_t o = (size_t)p * W + sc.col;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float v = (gxr[0][k] + (sy[0][k] - 2.0f * sy[1][k] + sy[2][k])) * 0.05f;
        gb[o + k * plane] = P.accumulate ? gb[o + k * plane] + v : v;
      }
    }
  }
  cp_async_wait<0>();
}


// ------------------------------------------------------------------------------ smooth, quad layout
// Same math, four adjacent columns per lane (W % 4 == 0, 16-byte aligned planes): a strip is 128 columns wide, lanes
// 1..30 own outputs (120 columns) and lanes 0 / 31 are the column halo.  The one-pixel-per-lane kernels above are
// instruction-issue bound (ncu, round 2: ~140 instructions per warp-row of 30 pixels, 43 % of them integer address
// and predicate work, issue slots 75 % busy, DRAM at 25 %); a quad amortises addressing, predicates and the
// neighbour shuffles over four pixels (7 SHFL per 4 px instead of 7 per px) and moves rows with 16-byte cp.async
// (five LDGSTS.128 + five LDS.128 per quad-row, zero-filled outside the image) through a per-warp three-slot ring.
// The row loop is unrolled by 3 so that ring slots and the 3-row register rings are statically indexed.
constexpr int kQDepth = 3;
constexpr float kEdgeLog2 = (-10.0f / 3.0f) * 1.4426950408889634f;   // exp(-10/3 s) = 2^(kEdgeLog2 s)

__device__ __forceinline__ float ex2_ftz(float x) {    // arguments are <= 0: results in (0, 1], tiny ones flush to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct QuadRow {
  float f[4][2];     // flow / 20
  float im[4][3];
};

__device__ __forceinline__ void fetch_quad(float4* ring, int slot, const float* __restrict__ fb, const float* __restrict__ ib,
                                           size_t plane, int W, int r, int H, int colc, bool qin) {
  const bool inb = qin && r >= 0 && r < H;
  const size_t off = (size_t)min(max(r, 0), H - 1) * W + colc;
  float4* d = ring + slot * (kRowVals * 32);
  cp_async_16(reinterpret_cast<float*>(d), fb + off, inb);
  cp_async_16(reinterpret_cast<float*>(d + 32), fb + off + plane, inb);
#pragma unroll
  for (int c = 0; c < 3; ++c) cp_async_16(reinterpret_cast<float*>(d + (2 + c) * 32), ib + off + c * plane, inb);
  cp_async_commit();
}

__device__ __forceinline__ void take_quad(const float4* ring, int slot, QuadRow& q) {
  cp_async_wait<kQDepth - 1>();
  const float4* d = ring + slot * (kRowVals * 32);
  const float4 a = d[0], b = d[32];
  // :174 flow/20.0.  __fmul_rn / d2(): no FMA contraction -- the decoder's flows are bilinearly up-sampled, i.e. piecewise
  // linear, so their second differences are pure rounding noise over large areas and the backward pass takes the SIGN of
  // that noise: it has to round exactly like the reference's separate mul / sub / sub.
  q.f[0][0] = __fmul_rn(a.x, 0.05f); q.f[1][0] = __fmul_rn(a.y, 0.05f); q.f[2][0] = __fmul_rn(a.z, 0.05f); q.f[3][0] = __fmul_rn(a.w, 0.05f);
  q.f[0][1] = __fmul_rn(b.x, 0.05f); q.f[1][1] = __fmul_rn(b.y, 0.05f); q.f[2][1] = __fmul_rn(b.z, 0.05f); q.f[3][1] = __fmul_rn(b.w, 0.05f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 t = d[(2 + c) * 32];
    q.im[0][c] = t.x; q.im[1][c] = t.y; q.im[2][c] = t.z; q.im[3][c] = t.w;
  }
}

__device__ __forceinline__ float edge_l1(const float* a, const float* b) {
  return fabsf(a[0] - b[0]) + fabsf(a[1] - b[1]) + fabsf(a[2] - b[2]);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 5)
smooth_fwd_quad_kernel(const __grid_constant__ SmoothParams P, float* __restrict__ sums, float* __restrict__ loss) {
  __shared__ float4 ring_s[kWarpsPerBlock][kQDepth * kRowVals * 32];
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Strip sc;
  const bool live = locate_strip<4, 4>(P.T, gw, lane, sc);   // idle warps still join the block reduction below
  if (!live) sc.level = sc.b = sc.col = sc.y0 = sc.y1 = 0;
  const uof_smooth_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const size_t plane = (size_t)H * W;
  const float* fb = L.flow + (size_t)sc.b * 2 * plane;
  const float* ib = L.img + (size_t)(sc.b % P.Bimg) * 3 * plane;
  const bool qin = sc.col >= 0 && sc.col < W;             // W % 4 == 0: the quad is entirely inside or outside
  const bool lane_out = qin && lane >= 1 && lane <= 30;
  const int colc = min(max(sc.col, 0), W - 4);
  float mx[4], my;                                         // column validity of the x / y terms as multipliers
#pragma unroll
  for (int j = 0; j < 4; ++j) mx[j] = (lane_out && sc.col + j >= 1 && sc.col + j <= W - 2) ? 1.0f : 0.0f;
  my = lane_out ? 1.0f : 0.0f;
  float sum_x = 0.0f, sum_y = 0.0f;
  if (live) {
    float4* ring = ring_s[threadIdx.x >> 5] + lane;
    QuadRow q[3];                                          // ring slots: row r -> u, r-1 -> (u+2)%3, r-2 -> (u+1)%3
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q[a].f[j][0] = q[a].f[j][1] = 0.0f;
        q[a].im[j][0] = q[a].im[j][1] = q[a].im[j][2] = 0.0f;
      }
    const int r_begin = sc.y0 - 1, r_end = sc.y1;
    fetch_quad(ring, 0, fb, ib, plane, W, r_begin, H, colc, qin);
    fetch_quad(ring, 1, fb, ib, plane, W, r_begin + 1, H, colc, qin);
    for (int rb = r_begin; rb <= r_end; rb += 3) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int r = rb + u;
        if (r > r_end) break;
        fetch_quad(ring, (u + 2) % 3, fb, ib, plane, W, r + 2, H, colc, qin);
        QuadRow& c = q[u];
        const QuadRow& p1 = q[(u + 2) % 3];
        const QuadRow& p2 = q[(u + 1) % 3];
        take_quad(ring, u, c);
        // neighbours across the quad boundary: column -1 from lane-1's last pixel, column 4 from lane+1's first pixel
        float fl[2], fr[2], ir[3];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          fl[k] = __shfl_up_sync(kFullMask, c.f[3][k], 1);
          fr[k] = __shfl_down_sync(kFullMask, c.f[0][k], 1);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) ir[ch] = __shfl_down_sync(kFullMask, c.im[0][ch], 1);
        if (r >= sc.y0 && r < sc.y1) {          // x term centred on (r, col)                       :153-155,163-164
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* lf = j == 0 ? fl : c.f[j - 1];
            const float* rf = j == 3 ? fr : c.f[j + 1];
            const float* ri = j == 3 ? ir : c.im[j + 1];
            const float w = ex2_ftz(kEdgeLog2 * edge_l1(ri, c.im[j])) * mx[j];
            const float dd = fabsf(d2(lf[0], c.f[j][0], rf[0])) + fabsf(d2(lf[1], c.f[j][1], rf[1]));
            sum_x = fmaf(w, dd, sum_x);
          }
        }
        const int m = r - 1;                    // y term centred on row m: rows m-1, m, m+1 are p2, p1, c
        if (m >= sc.y0 && m >= 1 && m <= H - 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float w = ex2_ftz(kEdgeLog2 * edge_l1(c.im[j], p1.im[j])) * my;
            const float dd = fabsf(d2(p2.f[j][0], p1.f[j][0], c.f[j][0])) + fabsf(d2(p2.f[j][1], p1.f[j][1], c.f[j][1]));
            sum_y = fmaf(w, dd, sum_y);
          }
        }
      }
    }
    cp_async_wait<0>();
  }
  const float acc[2] = {warp_sum(sum_x), warp_sum(sum_y)};
  block_accumulate<2, kWarpsPerBlock>(acc, live ? sums + ((size_t)sc.level * P.T.B + sc.b) * 2 : nullptr);
  if (last_block_done(reinterpret_cast<unsigned*>(sums + (size_t)P.T.nlevels * P.T.B * 2))) smooth_finalize(P, sums, loss);
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
smooth_bwd_quad_kernel(const __grid_constant__ SmoothParams P, const float* __restrict__ g_loss) {
  __shared__ float4 ring_s[kWarpsPerBlock][kQDepth * kRowVals * 32];
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Strip sc;
  if (!locate_strip<4, 4>(P.T, gw, lane, sc)) return;
  const uof_smooth_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const size_t plane = (size_t)H * W;
  const float* fb = L.flow + (size_t)sc.b * 2 * plane;
  const float* ib = L.img + (size_t)(sc.b % P.Bimg) * 3 * plane;
  float* gb = L.gflow + (size_t)sc.b * 2 * plane;
  const bool qin = sc.col >= 0 && sc.col < W;
  const bool lane_out = qin && lane >= 1 && lane <= 30;
  const int colc = min(max(sc.col, 0), W - 4);
  const float g = __ldg(g_loss + sc.b);
  const float cx = g * 0.5f / (2.0f * (float)H * ((float)W - 2.0f));
  const float cy = qin ? g * 0.5f / (2.0f * ((float)H - 2.0f) * (float)W) : 0.0f;
  float cxm[4];                                  // cx where the x term exists (1 <= col <= W-2), else 0
#pragma unroll
  for (int j = 0; j < 4; ++j) cxm[j] = (qin && sc.col + j >= 1 && sc.col + j <= W - 2) ? cx : 0.0f;

  float4* ring = ring_s[threadIdx.x >> 5] + lane;
  QuadRow q[3];            // rows r, r-1, r-2 in slots u, (u+2)%3, (u+1)%3
  float sy[3][4][2];       // sy at rows m = r-1 (slot u), m-1, m-2
  float gx[3][4][2];       // x part of the gradient at rows r (slot u), r-1, r-2
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q[a].f[j][0] = q[a].f[j][1] = 0.0f;
      q[a].im[j][0] = q[a].im[j][1] = q[a].im[j][2] = 0.0f;
      sy[a][j][0] = sy[a][j][1] = gx[a][j][0] = gx[a][j][1] = 0.0f;
    }
  const int r_begin = sc.y0 - 2, r_end = sc.y1 + 1;
  fetch_quad(ring, 0, fb, ib, plane, W, r_begin, H, colc, qin);
  fetch_quad(ring, 1, fb, ib, plane, W, r_begin + 1, H, colc, qin);
  for (int rb = r_begin; rb <= r_end; rb += 3) {
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int r = rb + u;
      if (r > r_end) break;
      fetch_quad(ring, (u + 2) % 3, fb, ib, plane, W, r + 2, H, colc, qin);
      QuadRow& c = q[u];
      const QuadRow& p1 = q[(u + 2) % 3];
      const QuadRow& p2 = q[(u + 1) % 3];
      take_quad(ring, u, c);
      float fl[2], fr[2], ir[3];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        fl[k] = __shfl_up_sync(kFullMask, c.f[3][k], 1);
        fr[k] = __shfl_down_sync(kFullMask, c.f[0][k], 1);
      }
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) ir[ch] = __shfl_down_sync(kFullMask, c.im[0][ch], 1);
      // x part of row r: s = cx wx sign(d2x) (zero rows outside the image give sign(0) = 0), then its second difference
      float s[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* lf = j == 0 ? fl : c.f[j - 1];
        const float* rf = j == 3 ? fr : c.f[j + 1];
        const float* ri = j == 3 ? ir : c.im[j + 1];
        const float w = ex2_ftz(kEdgeLog2 * edge_l1(ri, c.im[j])) * cxm[j];
#pragma unroll
        for (int k = 0; k < 2; ++k) s[j][k] = w * sgn(d2(lf[k], c.f[j][k], rf[k]));
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float sl = __shfl_up_sync(kFullMask, s[3][k], 1), sr = __shfl_down_sync(kFullMask, s[0][k], 1);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          gx[u][j][k] = (j == 0 ? sl : s[j - 1][k]) - 2.0f * s[j][k] + (j == 3 ? sr : s[j + 1][k]);
      }
      // y part: sy at row m = r-1 from rows m-1, m, m+1 = p2, p1, c
      const int m = r - 1;
      const float cym = (m >= 1 && m <= H - 2) ? cy : 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = ex2_ftz(kEdgeLog2 * edge_l1(c.im[j], p1.im[j])) * cym;
#pragma unroll
        for (int k = 0; k < 2; ++k) sy[u][j][k] = w * sgn(d2(p2.f[j][k], p1.f[j][k], c.f[j][k]));
      }
      // output row p = r-2: sy rows p-1, p, p+1 are slots (u+1)%3, (u+2)%3, u; its x part is gx slot (u+1)%3
      const int p = r - 2;
      if (p >= sc.y0 && p < sc.y1 && lane_out) {
        const size_t o = (size_t)p * W + sc.col;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          float4 v;
          float* vp = reinterpret_cast<float*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            vp[j] = (gx[(u + 1) % 3][j][k] + (sy[(u + 1) % 3][j][k] - 2.0f * sy[(u + 2) % 3][j][k] + sy[u][j][k])) * 0.05f;
          if (P.accumulate) {
            const float4 t = *reinterpret_cast<const float4*>(gb + o + k * plane);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
          }
          *reinterpret_cast<float4*>(gb + o + k * plane) = v;
        }
      }
    }
  }
  cp_async_wait<0>();
}

// true when every level can use the quad kernels: W % 4 == 0 and 16-byte aligned planes
bool smooth_quad_ok(const uof_smooth_level* levels, int nlevels, bool bwd) {
  static const bool off = getenv("UOF_SMOOTH_NO_QUAD") != nullptr;
  if (off || !levels) return false;
  for (int l = 0; l < nlevels && l < UOF_MAX_LEVELS; ++l) {
    uintptr_t bits = reinterpret_cast<uintptr_t>(levels[l].flow) | reinterpret_cast<uintptr_t>(levels[l].img);
    if (bwd) bits |= reinterpret_cast<uintptr_t>(levels[l].gflow);
    if ((bits & 15u) || levels[l].W % 4 != 0 || levels[l].W < 4) return false;
  }
  return true;
}

int fill_smooth(SmoothParams& P, const uof_smooth_level* levels, int nlevels, int B, int Bimg, int halo, bool bwd,
                int blocks_per_sm, bool quad = false) {
  UOF_REQUIRE(levels && nlevels >= 1 && nlevels <= UOF_MAX_LEVELS, "smooth_loss: nlevels must be 1..%d", UOF_MAX_LEVELS);
  UOF_REQUIRE(B > 0 && Bimg > 0 && B % Bimg == 0, "smooth_loss: flow batch %d must be a multiple of image batch %d", B, Bimg);
  int H[UOF_MAX_LEVELS], W[UOF_MAX_LEVELS];
  for (int l = 0; l < nlevels; ++l) {
    UOF_REQUIRE(levels[l].flow && levels[l].img && levels[l].H >= 3 && levels[l].W >= 3, "smooth_loss: level %d incomplete (needs H,W >= 3)", l);
    if (bwd) UOF_REQUIRE(levels[l].gflow, "smooth_loss_bwd: level %d has no gradient buffer", l);
    P.lv[l] = levels[l];
    H[l] = levels[l].H;
    W[l] = levels[l].W;
  }
  P.Bimg = Bimg;
  UOF_REQUIRE(build_strip_table(P.T, H, W, nlevels, B, halo, 1, blocks_per_sm, kWarpsPerBlock, quad ? 4 : 1, quad ? 4 : -1) > 0,
              "smooth_loss: problem too large");
  return UOF_OK;
}

// ------------------------------------------------------------------------------------- consis
struct ConsisParams {
  uof_consis_level lv[UOF_MAX_LEVELS];
  int warp_begin[UOF_MAX_LEVELS + 1];
  int warps_per_sample[UOF_MAX_LEVELS];
  int nlevels, B;
  int vec4;          // every plane is a multiple of 4 pixels and every pointer 16-byte aligned
  int px_per_warp;   // pixels owned by a warp: kConsisPxPerWarp (backward) or kConsisFwdChunks times that (forward)
  int accumulate;    // backward: gflow_fwd += instead of =
};
constexpr int kConsisPxPerWarp = 32 * 8;
constexpr int kConsisFwdChunks = 1;      // measured: 4 chunks per warp (4x fewer blocks, barriers, REDs) is not faster (19.0 vs 18.5 us)

__device__ __forceinline__ bool locate_chunk(const ConsisParams& P, int gw, int& level, int& b, int& px0) {
  if (gw >= P.warp_begin[P.nlevels]) return false;
  int l = 0;
  while (l + 1 < P.nlevels && gw >= P.warp_begin[l + 1]) ++l;
  const int local = gw - P.warp_begin[l];
  level = l;
  b = local / P.warps_per_sample[l];
  px0 = (local % P.warps_per_sample[l]) * P.px_per_warp;
  return true;
}

// One warp handles kConsisPxPerWarp = 256 pixels of one sample.  VEC = 4: two float4 loads per plane and lane
// (all ten loads of a lane are issued before any arithmetic); VEC = 1: scalar fallback for planes that are not a
// multiple of 4 pixels.
template <int VEC>
__device__ __forceinline__ void consis_load(const float* __restrict__ base, int p, int plane, float* v) {
  if (VEC == 4) {
    const float4 t = (p < plane) ? __ldg(reinterpret_cast<const float4*>(base + p)) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    v[0] = (p < plane) ? __ldg(base + p) : 0.0f;
  }
}

constexpr int kConsisFwdWarps = 4;   // forward: the warps of a block share one RED pair (block_accumulate)

template <int VEC>
__global__ void __launch_bounds__(kConsisFwdWarps * 32)
consis_fwd_kernel(const __grid_constant__ ConsisParams P, float* __restrict__ sums) {
  pdl_trigger();      // let the finalize grid (a programmatic dependent) be scheduled behind this grid's last wave
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kConsisFwdWarps + (threadIdx.x >> 5);
  int level = 0, b = 0, px0 = 0;
  const bool live_warp = locate_chunk(P, gw, level, b, px0);
  const uof_consis_level& L = P.lv[level];
  const int plane = L.H * L.W;
  const float* ff = L.flow_fwd + (size_t)b * 2 * plane;
  const float* fb = L.flow_bwd + (size_t)b * 2 * plane;
  const float* wf = L.weight_fwd + (size_t)b * plane;
  constexpr int NIT = kConsisPxPerWarp / (32 * VEC);
  const int plane_live = live_warp ? plane : 0;      // idle warps load nothing and contribute zeros
  float num = 0.0f, den = 0.0f;
  // a warp walks kConsisFwdChunks chunks of 256 pixels
  for (int ch = 0; ch < kConsisFwdChunks; ++ch, px0 += kConsisPxPerWarp) {
    if (px0 >= plane_live) break;
    float ax[NIT][VEC], ay[NIT][VEC], bx[NIT][VEC], by[NIT][VEC], w[NIT][VEC];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int p = px0 + (it * 32 + lane) * VEC;
      consis_load<VEC>(ff, p, plane_live, ax[it]);
      consis_load<VEC>(ff + plane, p, plane_live, ay[it]);
      consis_load<VEC>(fb, p, plane_live, bx[it]);
      consis_load<VEC>(fb + plane, p, plane_live, by[it]);
      consis_load<VEC>(wf, p, plane_live, w[it]);
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const bool live = px0 + (it * 32 + lane) * VEC < plane_live;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float occ = live ? 1.0f - w[it][v] : 0.0f;                                               // :187
        const float ia = __fdividef(1.0f, sqrtf(fmaf(ax[it][v], ax[it][v], ay[it][v] * ay[it][v])) + kEps);   // :49
        const float ib = __fdividef(1.0f, sqrtf(fmaf(bx[it][v], bx[it][v], by[it][v] * by[it][v])) + kEps);
        num = fmaf(fabsf(fmaf(ax[it][v], ia, bx[it][v] * ib)) + fabsf(fmaf(ay[it][v], ia, by[it][v] * ib)), occ, num);   // :191
        den += occ;
      }
    }
  }
  const float acc[2] = {warp_sum(num), warp_sum(den)};
  block_accumulate<2, kConsisFwdWarps>(acc, live_warp ? sums + ((size_t)level * P.B + b) * 2 : nullptr);
}

// Separate one-block launch (folding it into the forward kernel like smooth_fwd does brought nothing here: 18.5 -> 19.1 us,
// the block counter is one more same-address atomic for each of the ~1100 short-lived blocks), but as a programmatic
// dependent of the forward kernel: it is already resident and waiting when the last block of that grid retires.
__global__ void consis_finalize_kernel(const __grid_constant__ ConsisParams P, const float* __restrict__ sums,
                                       float* __restrict__ loss) {
  pdl_wait();
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < P.B; b += gridDim.x * blockDim.x) {
    float acc = 0.0f;
    for (int l = 0; l < P.nlevels; ++l) {
      const float n = (float)P.lv[l].H * (float)P.lv[l].W;
      const float* s = sums + ((size_t)l * P.B + b) * 2;
      acc += (__ldcg(s) / (2.0f * n)) / (__ldcg(s + 1) / n + kEps);   // :189-192
    }
    loss[b] = acc;
  }
}

template <int VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
consis_bwd_kernel(const __grid_constant__ ConsisParams P, const float* __restrict__ sums, const float* __restrict__ g_loss) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  int level, b, px0;
  if (!locate_chunk(P, gw, level, b, px0)) return;
  const uof_consis_level& L = P.lv[level];
  const int plane = L.H * L.W;
  const float* ff = L.flow_fwd + (size_t)b * 2 * plane;
  const float* fb = L.flow_bwd + (size_t)b * 2 * plane;
  const float* wf = L.weight_fwd + (size_t)b * plane;
  float* gf = L.gflow_fwd + (size_t)b * 2 * plane;
  const float n = (float)plane;
  const float coef = __ldg(g_loss + b) / (2.0f * n) / (sums[((size_t)level * P.B + b) * 2 + 1] / n + kEps);
  constexpr int NIT = kConsisPxPerWarp / (32 * VEC);
  float ax[NIT][VEC], ay[NIT][VEC], bx[NIT][VEC], by[NIT][VEC], w[NIT][VEC];
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int p = px0 + (it * 32 + lane) * VEC;
    consis_load<VEC>(ff, p, plane, ax[it]);
    consis_load<VEC>(ff + plane, p, plane, ay[it]);
    consis_load<VEC>(fb, p, plane, bx[it]);
    consis_load<VEC>(fb + plane, p, plane, by[it]);
    consis_load<VEC>(wf, p, plane, w[it]);
  }
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int p = px0 + (it * 32 + lane) * VEC;
    if (p >= plane) continue;
    float gx[VEC], gy[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float occ = 1.0f - w[it][v];
      const float ra = sqrtf(fmaf(ax[it][v], ax[it][v], ay[it][v] * ay[it][v]));
      const float ia = __fdividef(1.0f, ra + kEps);
      const float ib = __fdividef(1.0f, sqrtf(fmaf(bx[it][v], bx[it][v], by[it][v] * by[it][v])) + kEps);
      const float ux = sgn(fmaf(ax[it][v], ia, bx[it][v] * ib)) * coef * occ;
      const float uy = sgn(fmaf(ay[it][v], ia, by[it][v] * ib)) * coef * occ;
      // d(a_i/na)/d a_j = delta_ij/na - a_i a_j/(ra*na^2); torch.norm's subgradient at 0 is 0
      const float k = ra > 0.0f ? (ux * ax[it][v] + uy * ay[it][v]) * ia * ia * __fdividef(1.0f, ra) : 0.0f;
      gx[v] = fmaf(-k, ax[it][v], ux * ia);
      gy[v] = fmaf(-k, ay[it][v], uy * ia);
    }
    if (VEC == 4) {
      float4 vx = make_float4(gx[0], gx[1], gx[2], gx[3]), vy = make_float4(gy[0], gy[1], gy[2], gy[3]);
      if (P.accumulate) {
        const float4 tx = *reinterpret_cast<const float4*>(gf + p), ty = *reinterpret_cast<const float4*>(gf + plane + p);
        vx.x += tx.x; vx.y += tx.y; vx.z += tx.z; vx.w += tx.w;
        vy.x += ty.x; vy.y += ty.y; vy.z += ty.z; vy.w += ty.w;
      }
      *reinterpret_cast<float4*>(gf + p) = vx;
      *reinterpret_cast<float4*>(gf + plane + p) = vy;
    } else {
      gf[p] = P.accumulate ? gf[p] + gx[0] : gx[0];
      gf[plane + p] = P.accumulate ? gf[plane + p] + gy[0] : gy[0];
    }
  }
}

int fill_consis(ConsisParams& P, const uof_consis_level* levels, int nlevels, int B, bool bwd) {
  P.px_per_warp = bwd ? kConsisPxPerWarp : kConsisPxPerWarp * kConsisFwdChunks;
  UOF_REQUIRE(levels && nlevels >= 1 && nlevels <= UOF_MAX_LEVELS, "consis_loss: nlevels must be 1..%d", UOF_MAX_LEVELS);
  UOF_REQUIRE(B > 0, "consis_loss: bad batch %d", B);
  long long total = 0;
  for (int l = 0; l < nlevels; ++l) {
    const uof_consis_level& L = levels[l];
    UOF_REQUIRE(L.flow_fwd && L.flow_bwd && L.weight_fwd && L.H > 0 && L.W > 0, "consis_loss: level %d incomplete", l);
    if (bwd) UOF_REQUIRE(L.gflow_fwd, "consis_loss_bwd: level %d has no gradient buffer", l);
    UOF_REQUIRE((long long)L.H * L.W < (1ll << 30), "consis_loss: level %d too large", l);
    P.lv[l] = L;
    P.warps_per_sample[l] = ceil_div(L.H * L.W, P.px_per_warp);
    P.warp_begin[l] = (int)total;
    total += (long long)P.warps_per_sample[l] * B;
    UOF_REQUIRE(total < (1ll << 30), "consis_loss: problem too large");
  }
  P.warp_begin[nlevels] = (int)total;
  P.nlevels = nlevels;
  P.B = B;
  P.vec4 = 1;
  for (int l = 0; l < nlevels; ++l) {
    const uof_consis_level& L = levels[l];
    const uintptr_t bits = reinterpret_cast<uintptr_t>(L.flow_fwd) | reinterpret_cast<uintptr_t>(L.flow_bwd) |
                           reinterpret_cast<uintptr_t>(L.weight_fwd) | reinterpret_cast<uintptr_t>(L.gflow_fwd);
    if ((L.H * L.W) % 4 != 0 || (bits & 15u)) P.vec4 = 0;
  }
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_smooth_loss_fwd(const uof_smooth_level* levels, int nlevels, int B, int Bimg, float* sums, float* loss,
                                   uof_stream_t stream_) {
  UOF_REQUIRE(sums && loss, "smooth_loss_fwd: null output");
  SmoothParams P;
  P.accumulate = 0;
  static const int occ = resident_blocks(smooth_fwd_kernel, kWarpsPerBlock * 32);
  static const int occ_quad = resident_blocks(smooth_fwd_quad_kernel, kWarpsPerBlock * 32);
  const bool quad = smooth_quad_ok(levels, nlevels, false);
  if (int rc = fill_smooth(P, levels, nlevels, B, Bimg, 1, false, quad ? occ_quad : occ, quad)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UOF_CUDA(cudaMemsetAsync(sums, 0, ((size_t)nlevels * B * 2 + UOF_SUMS_EXTRA) * sizeof(float), stream));
  if (quad)
    smooth_fwd_quad_kernel<<<ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(P, sums, loss);
  else
    smooth_fwd_kernel<<<ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(P, sums, loss);
  count_launch();
  return check_launch("smooth_loss_fwd");
}

extern "C" int uof_smooth_loss_bwd(const uof_smooth_level* levels, int nlevels, int B, int Bimg, const float* g_loss,
                                   uof_stream_t stream_) {
  return uof_smooth_loss_bwd_acc(levels, nlevels, B, Bimg, g_loss, 0, stream_);
}

extern "C" int uof_smooth_loss_bwd_acc(const uof_smooth_level* levels, int nlevels, int B, int Bimg, const float* g_loss,
                                       int accumulate, uof_stream_t stream_) {
  UOF_REQUIRE(g_loss, "smooth_loss_bwd: null input");
  SmoothParams P;
  P.accumulate = accumulate ? 1 : 0;
  static const int occ = resident_blocks(smooth_bwd_kernel, kWarpsPerBlock * 32);
  static const int occ_quad = resident_blocks(smooth_bwd_quad_kernel, kWarpsPerBlock * 32);
  const bool quad = smooth_quad_ok(levels, nlevels, true);
  if (int rc = fill_smooth(P, levels, nlevels, B, Bimg, 2, true, quad ? occ_quad : occ, quad)) return rc;
  if (quad)
    smooth_bwd_quad_kernel<<<ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock), kWarpsPerBlock * 32, 0,
                             static_cast<cudaStream_t>(stream_)>>>(P, g_loss);
  else
    smooth_bwd_kernel<<<ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock), kWarpsPerBlock * 32, 0,
                        static_cast<cudaStream_t>(stream_)>>>(P, g_loss);
  count_launch();
  return check_launch("smooth_loss_bwd");
}

extern "C" int uof_consis_loss_fwd(const uof_consis_level* levels, int nlevels, int B, float* sums, float* loss,
                                   uof_stream_t stream_) {
  UOF_REQUIRE(sums && loss, "consis_loss_fwd: null output");
  ConsisParams P;
  P.accumulate = 0;
  if (int rc = fill_consis(P, levels, nlevels, B, false)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UOF_CUDA(cudaMemsetAsync(sums, 0, ((size_t)nlevels * B * 2 + UOF_SUMS_EXTRA) * sizeof(float), stream));
  if (P.vec4)
    consis_fwd_kernel<4><<<ceil_div(P.warp_begin[nlevels], kConsisFwdWarps), kConsisFwdWarps * 32, 0, stream>>>(P, sums);
  else
    consis_fwd_kernel<1><<<ceil_div(P.warp_begin[nlevels], kConsisFwdWarps), kConsisFwdWarps * 32, 0, stream>>>(P, sums);
  UOF_CUDA(launch_dependent(consis_finalize_kernel, dim3(ceil_div(B, 64)), dim3(64), stream, P, (const float*)sums, loss));
  count_launch(2);
  return check_launch("consis_loss_fwd");
}

extern "C" int uof_consis_loss_bwd(const uof_consis_level* levels, int nlevels, int B, const float* sums,
                                   const float* g_loss, uof_stream_t stream_) {
  return uof_consis_loss_bwd_acc(levels, nlevels, B, sums, g_loss, 0, stream_);
}

extern "C" int uof_consis_loss_bwd_acc(const uof_consis_level* levels, int nlevels, int B, const float* sums,
                                       const float* g_loss, int accumulate, uof_stream_t stream_) {
  UOF_REQUIRE(sums && g_loss, "consis_loss_bwd: null input");
  ConsisParams P;
  P.accumulate = accumulate ? 1 : 0;
  if (int rc = fill_consis(P, levels, nlevels, B, true)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (P.vec4)
    consis_bwd_kernel<4><<<ceil_div(P.warp_begin[nlevels], kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(P, sums, g_loss);
  else
    consis_bwd_kernel<1><<<ceil_div(P.warp_begin[nlevels], kWarpsPerBlock), kWarpsPerBlock * 32, 0, stream>>>(P, sums, g_loss);
  count_launch();
  return check_launch("consis_loss_bwd");
}
