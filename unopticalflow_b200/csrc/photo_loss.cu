// a4+a5+a6 fused: photometric weight maps, masked L1 and masked SSIM losses, both directions,
// every pyramid level in ONE launch (forward) / ONE launch (backward).
// Replaces, in /root/reference/core/networks/model_flow_paper.py, compute_diff_weight (:101-134),
// compute_loss_with_mask x2 (:90-99), compute_loss_ssim x2 (:137-148) and
// pytorch_ssim/ssim.py:4-19 -- ~350 ATen launches and ~40 full-size temporaries forward.
//
// Design ("marching warp"): a warp owns a strip of 32 image columns (30 or 28 of them outputs, the
// rest halo) and marches down R rows.  Each lane loads its pixel's 9 values (I, W_l, W_r) straight
// from global memory (128 B coalesced per plane per row, next row prefetched into registers while the
// current one is processed), computes the weights in registers, gets its left/right neighbours with warp
// shuffles (horizontal 3-tap sums) and keeps a 3-row register ring (statically indexed: the row loop
// is unrolled by 3) for the vertical 3-tap sums.  No shared memory, no temporaries in HBM; per-sample
// reductions are warp-shuffle trees + one fp32 atomic per warp and quantity.  The backward pass
// chains two such box filters (5x5 footprint) with a second ring.
//
// The kernels are instruction-issue bound, not HBM bound (ncu: profiles/): SSIM on 2 directions x 3
// channels costs ~50 instructions per map and pixel, so the arithmetic is kept lean: SSIM is evaluated
// on raw 3x3 sums (the 1/9 factors cancel), one reciprocal per map, and the softmax/Gaussian weight
// pair needs two exp2-based exponentials per pixel (a_l - 0.5 == -(a_r - 0.5)).
#include <stdlib.h>

#include "strips.cuh"

namespace uof {
namespace {

constexpr float C1x81 = 81.0f * 0.01f * 0.01f;      // 81 * C1, ssim.py:5
constexpr float C2x81 = 81.0f * 0.03f * 0.03f;      // 81 * C2, ssim.py:6
constexpr float kEps = 1e-12f;                      // model_flow_paper.py:97,145
constexpr float kThird = 1.0f / 3.0f;
constexpr float kInvSigma2 = 1.0f / 0.03f;          // model_flow_paper.py:126
constexpr int kWarpsPerBlock = 4;

struct PhotoParams {
  uof_photo_level lv[UOF_MAX_LEVELS];
  StripTable T;
};

struct PixelWeights {
  float dl, dr, wl, wr;
};

// model_flow_paper.py:111-129 for one pixel.  px = {I0,I1,I2, L0,L1,L2, R0,R1,R2}.
__device__ __forceinline__ PixelWeights pixel_weights(const float* px) {
  PixelWeights o;
  o.dl = (fabsf(px[0] - px[3]) + fabsf(px[1] - px[4]) + fabsf(px[2] - px[5])) * kThird;
  o.dr = (fabsf(px[0] - px[6]) + fabsf(px[1] - px[7]) + fabsf(px[2] - px[8])) * kThird;
  const float vl = (px[3] == 0.0f && px[4] == 0.0f && px[5] == 0.0f) ? 0.0f : 1.0f;   // :112
  const float vr = (px[6] == 0.0f && px[7] == 0.0f && px[8] == 0.0f) ? 0.0f : 1.0f;   // :111
  // 1 - softmax over (dl, dr): the larger difference gets t/(1+t), the smaller 1/(1+t), t = exp(-|dl-dr|);
  // both are equally far from 0.5, so the Gaussian weight 2*exp(-(a-0.5)^2/0.03) is shared (:120-126)
  const float t = __expf(-fabsf(o.dl - o.dr));
  const float h = __fdividef(1.0f, 1.0f + t) - 0.5f;
  const float g = 2.0f * __expf(-(h * h) * kInvSigma2);
  o.wl = g * vl;
  o.wr = g * vr;
  return o;
}

// Unconditional loads from a clamped (always valid) offset: the caller zeroes the weights of out-of-image pixels, which
// zeroes everything derived from px.  (Predicated zero-filling loads made ptxas copy the just-loaded registers, so the
// prefetch of the next row stalled the current one.)
__device__ __forceinline__ void load_pixel(const uof_photo_level& L, unsigned off, unsigned plane, float* px) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    px[c] = __ldg(L.img + off + c * plane);
    px[3 + c] = __ldg(L.warped_l + off + c * plane);
    px[6 + c] = __ldg(L.warped_r + off + c * plane);
  }
}

// horizontal 3-tap sums of (x, y, x^2, y^2, xy) across adjacent lanes
__device__ __forceinline__ void hsum_moments(float xv, float yv, float* m) {
  const float xl = __shfl_up_sync(kFullMask, xv, 1), xr = __shfl_down_sync(kFullMask, xv, 1);
  const float yl = __shfl_up_sync(kFullMask, yv, 1), yr = __shfl_down_sync(kFullMask, yv, 1);
  m[0] = xl + xv + xr;
  m[1] = yl + yv + yr;
  m[2] = fmaf(xr, xr, fmaf(xv, xv, xl * xl));
  m[3] = fmaf(yr, yr, fmaf(yv, yv, yl * yl));
  m[4] = fmaf(xr, yr, fmaf(xv, yv, xl * yl));
}

// SSIM on raw 3x3 sums: with Sx = 9 mu_x etc. every factor of ssim.py:15-16 is scaled by 81, which cancels.
struct SsimTerms {
  float Sx, Sy, A1, A2, B1, B2, invD, S;
};

// S = {Sx, Sy, Sxx, Syy, Sxy}: raw sums over the 3x3 window
__device__ __forceinline__ SsimTerms ssim_from_total(float Sx, float Sy, float Sxx, float Syy, float Sxy) {
  SsimTerms t;
  t.Sx = Sx;
  t.Sy = Sy;
  const float pxy = t.Sx * t.Sy, pxx = t.Sx * t.Sx, pyy = t.Sy * t.Sy;
  t.A1 = fmaf(2.0f, pxy, C1x81);
  t.A2 = fmaf(2.0f, fmaf(9.0f, Sxy, -pxy), C2x81);
  t.B1 = pxx + pyy + C1x81;
  t.B2 = fmaf(9.0f, Sxx, -pxx) + fmaf(9.0f, Syy, -pyy) + C2x81;
  t.invD = __fdividef(1.0f, t.B1 * t.B2);
  t.S = (t.A1 * t.A2) * t.invD;
  return t;
}

__device__ __forceinline__ SsimTerms ssim_from_sums(const float* s0, const float* s1, const float* s2) {
  return ssim_from_total(s0[0] + s1[0] + s2[0], s0[1] + s1[1] + s2[1], s0[2] + s1[2] + s2[2], s0[3] + s1[3] + s2[3],
                         s0[4] + s1[4] + s2[4]);
}

// -------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
photo_loss_fwd_kernel(const __grid_constant__ PhotoParams P, float* __restrict__ sums) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Strip sc;
  const bool live = locate_strip<1>(P.T, gw, lane, sc);   // idle warps still join the block reduction at the end
  if (!live) sc.level = sc.b = sc.col = sc.y0 = sc.y1 = 0;
  const uof_photo_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const unsigned plane = (unsigned)(H * W);
  const unsigned img_base = (unsigned)sc.b * 3u * plane, map_base = (unsigned)sc.b * plane;
  const bool col_in = sc.col >= 0 && sc.col < W;
  const bool col_out = col_in && lane >= 1 && lane <= 30;
  const int colc = min(max(sc.col, 0), W - 1);

  // Vertical 3-tap sums without a 3-row ring: per map keep last[] = h-sums of row r-1 and pair[] = h(r-2) + h(r-1);
  // row r then gives the window total pair + h(r) -- the same (s0 + s1) + s2 association as a ring, so results are
  // bit-identical -- with 60 live registers instead of 90 and no slot rotation (the ring version spilled at the
  // 128-register cap of 4 blocks/SM).
  float last[2][3][5], pair[2][3][5];   // [direction][channel][moment]
#pragma unroll
  for (int d = 0; d < 2; ++d)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 5; ++k) last[d][c][k] = pair[d][c][k] = 0.0f;
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

  const int r_begin = sc.y0 - 1, r_end = live ? sc.y1 : r_begin - 1;
  float nxt[9];         // next row, prefetched one iteration ahead
  load_pixel(L, img_base + (unsigned)max(r_begin, 0) * W + colc, plane, nxt);

  for (int r = r_begin; r <= r_end; ++r) {
    {
      const bool inb = col_in && r >= 0 && r < H;
      const unsigned off = (unsigned)max(r, 0) * W + colc;
      float px[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) px[k] = nxt[k];
      load_pixel(L, img_base + (unsigned)min(r + 1, H - 1) * W + colc, plane, nxt);

      PixelWeights pw = {0.f, 0.f, 0.f, 0.f};
      if (inb) pw = pixel_weights(px);
      if (inb && col_out && r >= sc.y0 && r < sc.y1) {
        acc[0] = fmaf(pw.dl, pw.wl, acc[0]);
        acc[1] += pw.wl;
        acc[2] = fmaf(pw.dr, pw.wr, acc[2]);
        acc[3] += pw.wr;
        if (L.weight_l) L.weight_l[map_base + off] = pw.wl;
        if (L.weight_r) L.weight_r[map_base + off] = pw.wr;
        if (L.diff_l) L.diff_l[map_base + off] = pw.dl;
        if (L.diff_r) L.diff_r[map_base + off] = pw.dr;
      }
      const bool emit = r - 1 >= sc.y0 && col_out;      // row q = r-1 now has its full 3x3 window
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const float wd = d ? pw.wr : pw.wl;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float h[5];
          hsum_moments(px[c] * wd, px[3 + 3 * d + c] * wd, h);
          float* pr = pair[d][c];
          float* la = last[d][c];
          if (emit) {
            const SsimTerms t = ssim_from_total(pr[0] + h[0], pr[1] + h[1], pr[2] + h[2], pr[3] + h[3], pr[4] + h[4]);
            acc[4 + d] += __saturatef(fmaf(-0.5f, t.S, 0.5f));   // clamp((1-S)/2, 0, 1), model_flow_paper.py:144
          }
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            pr[k] = la[k] + h[k];
            la[k] = h[k];
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
  block_accumulate<6, kWarpsPerBlock>(acc, live ? sums + ((size_t)sc.level * P.T.B + sc.b) * 6 : nullptr);
}

// loss_pixel[b] = sum_l sum_d mean(d*w)/(mean(w)+eps);  loss_ssim[b] likewise (model_flow_paper.py:94-98,141-147).
// A separate one-block launch: folding it into the forward kernel ("last block" pattern, as smooth_fwd does) was measured
// SLOWER here (59.7 -> 69.6 us): the per-block __threadfence has to drain the block's weight-map stores and the counter
// is one more same-address atomic per block.
__global__ void photo_loss_finalize_kernel(const __grid_constant__ PhotoParams P, const float* __restrict__ sums,
                                           float* __restrict__ loss_pixel, float* __restrict__ loss_ssim) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < P.T.B; b += gridDim.x * blockDim.x) {
    float lp = 0.0f, ls = 0.0f;
    for (int l = 0; l < P.T.nlevels; ++l) {
      const float n = (float)P.lv[l].H * (float)P.lv[l].W;
      const float* s = sums + ((size_t)l * P.T.B + b) * 6;
      const float s0 = __ldcg(s), s1 = __ldcg(s + 1), s2 = __ldcg(s + 2), s3 = __ldcg(s + 3), s4 = __ldcg(s + 4), s5 = __ldcg(s + 5);
      // reference order: forward/right term first, then backward/left (:241-245)
      lp += (s2 / n) / (s3 / n + kEps) + (s0 / n) / (s1 / n + kEps);
      ls += (s5 / (3.0f * n)) / (s3 / n + kEps) + (s4 / (3.0f * n)) / (s1 / n + kEps);
    }
    loss_pixel[b] = lp;
    loss_ssim[b] = ls;
  }
}

// ------------------------------------------------------------------------------------- backward
// blockIdx.y = direction (0: left/"bwd", 1: right/"fwd").  Output: d loss / d warped_{l,r}.
//   y = W*w, raw sums Sy, Syy, Sxy over the 3x3 window of q;  S = S(Sx, Sy, Sxx, Syy, Sxy)
//   dL/dy[p] = sum_{q in N3(p)} gS[q] * ( dS/dSy[q] + 2 y[p] dS/dSyy[q] + x[p] dS/dSxy[q] )
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
photo_loss_bwd_kernel(const __grid_constant__ PhotoParams P, const float* __restrict__ sums,
                      const float* __restrict__ g_pixel, const float* __restrict__ g_ssim) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int dir = blockIdx.y;
  Strip sc;
  if (!locate_strip<2>(P.T, gw, lane, sc)) return;
  const uof_photo_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const unsigned plane = (unsigned)(H * W);
  const unsigned img_base = (unsigned)sc.b * 3u * plane;
  const bool col_in = sc.col >= 0 && sc.col < W;
  const bool col_out = col_in && lane >= 2 && lane <= 29;
  const int colc = min(max(sc.col, 0), W - 1);
  float* gout = dir ? L.gwarped_r : L.gwarped_l;

  const float n = (float)H * (float)W;
  const float* s = sums + ((size_t)sc.level * P.T.B + sc.b) * 6;
  const float inv_div = 1.0f / (s[dir ? 3 : 1] / n + kEps);
  const float coef_l1 = __ldg(g_pixel + sc.b) * inv_div / n / 3.0f;           // d loss / d |I_c - W_c| per unit weight
  const float coef_ss = -0.5f * __ldg(g_ssim + sc.b) * inv_div / (3.0f * n);  // d loss / d S where the clamp passes

  float mom[3][3][5];    // [row slot][channel][moment]         rows r-2, r-1, r
  float abc[3][3][3];    // [row slot][channel][a,b,c] h-sums   rows q-2, q-1, q   (q = r-1)
  float xy[3][3][2];     // [row slot][channel][x,y]
  float wl1[3][4];       // [row slot][w, l1grad_c0..2]
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 5; ++k) mom[a][c][k] = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; ++k) abc[a][c][k] = 0.0f;
      xy[a][c][0] = xy[a][c][1] = 0.0f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) wl1[a][k] = 0.0f;
  }

  const int r_begin = sc.y0 - 2, r_end = sc.y1 + 1;
  float nxt[9];
  load_pixel(L, img_base + (unsigned)max(r_begin, 0) * W + colc, plane, nxt);

  for (int rb = r_begin; rb <= r_end; rb += 3) {
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int r = rb + u;
      if (r > r_end) break;
      // ring slots: row r -> u, q = r-1 -> (u+2)%3, p = r-2 -> (u+1)%3
      const bool inb = col_in && r >= 0 && r < H;
      float px[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) px[k] = nxt[k];
      if (r < r_end) load_pixel(L, img_base + (unsigned)min(max(r + 1, 0), H - 1) * W + colc, plane, nxt);
      PixelWeights pw = {0.f, 0.f, 0.f, 0.f};
      if (inb) pw = pixel_weights(px);
      const float wd = dir ? pw.wr : pw.wl;

      wl1[u][0] = wd;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float wc = dir ? px[6 + c] : px[3 + c];
        const float df = px[c] - wc;
        const float sg = df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f);
        wl1[u][1 + c] = -sg * coef_l1 * wd;          // d(masked L1)/d W_c
        xy[u][c][0] = px[c] * wd;
        xy[u][c][1] = wc * wd;
        hsum_moments(xy[u][c][0], xy[u][c][1], mom[u][c]);
      }

      // row q = r-1: coefficients of the SSIM derivative w.r.t. the raw sums (Sy, Syy, Sxy) at (q, col)
      const int q = r - 1;
      if (q >= sc.y0 - 1) {
        const bool q_in = col_in && q >= 0 && q < H;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float a = 0.0f, bb = 0.0f, cc = 0.0f;
          if (q_in) {
            const SsimTerms t = ssim_from_sums(mom[0][c], mom[1][c], mom[2][c]);
            const float term = fmaf(-0.5f, t.S, 0.5f);
            if (term >= 0.0f && term <= 1.0f) {        // clamp passes gradient on the closed interval
              const float k = coef_ss * t.invD;
              const float dN = 2.0f * t.Sx * (t.A2 - t.A1);
              const float dD = 2.0f * t.Sy * (t.B2 - t.B1);
              a = k * fmaf(-t.S, dD, dN);
              bb = -9.0f * k * t.S * t.B1;             // dS/dSyy = -9 S / B2 = -9 S B1 / D
              cc = 18.0f * k * t.A1;                   // dS/dSxy = 18 A1 / D
            }
          }
          float* dst = abc[(u + 2) % 3][c];
          dst[0] = __shfl_up_sync(kFullMask, a, 1) + a + __shfl_down_sync(kFullMask, a, 1);
          dst[1] = __shfl_up_sync(kFullMask, bb, 1) + bb + __shfl_down_sync(kFullMask, bb, 1);
          dst[2] = __shfl_up_sync(kFullMask, cc, 1) + cc + __shfl_down_sync(kFullMask, cc, 1);
        }
      }

      // row p = r-2: all nine coefficient neighbours are in the ring
      const int p = r - 2;
      if (p >= sc.y0 && p < sc.y1 && col_out) {
        const unsigned o = img_base + (unsigned)p * W + sc.col;
        const int sp = (u + 1) % 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float A = abc[0][c][0] + abc[1][c][0] + abc[2][c][0];
          const float Bq = abc[0][c][1] + abc[1][c][1] + abc[2][c][1];
          const float Cq = abc[0][c][2] + abc[1][c][2] + abc[2][c][2];
          const float gy = fmaf(xy[sp][c][0], Cq, fmaf(2.0f * xy[sp][c][1], Bq, A));
          gout[o + c * plane] = fmaf(gy, wl1[sp][0], wl1[sp][1 + c]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------- backward, split variant
// Used when the weight maps written by the forward pass are available (the Python layer saves them): one block of
// six warps per strip, warp = (direction, channel).  A warp loads only I_c, W_dc and w_d (3 instead of 9 values per
// pixel), skips the weight math entirely and keeps one channel's rings (36 registers instead of 102), so the kernel
// runs at ~70 registers / 28+ resident warps per SM instead of 151 / 12 -- the fused-direction kernel above was
// latency/issue bound at 60 % issue utilisation (ncu, profiles/).
constexpr int kSplitWarps = 6;
constexpr int kPairDepth = 4;      // rows in the cp.async ring of the pixel-pair kernel

__global__ void __launch_bounds__(kSplitWarps * 32)
photo_loss_bwd_split_kernel(const __grid_constant__ PhotoParams P, const float* __restrict__ sums,
                            const float* __restrict__ g_pixel, const float* __restrict__ g_ssim) {
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;            // 0..5
  const int dir = role / 3, c = role - 3 * dir;
  Strip sc;
  if (!locate_strip<2>(P.T, blockIdx.x, lane, sc)) return;
  const uof_photo_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const unsigned plane = (unsigned)(H * W);
  const unsigned ch_base = ((unsigned)sc.b * 3u + (unsigned)c) * plane, map_base = (unsigned)sc.b * plane;
  const bool col_in = sc.col >= 0 && sc.col < W;
  const bool col_out = col_in && lane >= 2 && lane <= 29;
  const int colc = min(max(sc.col, 0), W - 1);
  const float* __restrict__ img = L.img + ch_base;
  const float* __restrict__ wrp = (dir ? L.warped_r : L.warped_l) + ch_base;
  const float* __restrict__ wgt = (dir ? L.weight_r : L.weight_l) + map_base;
  float* __restrict__ gout = (dir ? L.gwarped_r : L.gwarped_l) + ch_base;

  const float n = (float)H * (float)W;
  const float* s = sums + ((size_t)sc.level * P.T.B + sc.b) * 6;
  const float inv_div = 1.0f / (s[dir ? 3 : 1] / n + kEps);
  const float coef_l1 = __ldg(g_pixel + sc.b) * inv_div / n / 3.0f;
  const float coef_ss = -0.5f * __ldg(g_ssim + sc.b) * inv_div / (3.0f * n);

  float mom[3][5], abc[3][3], xy[3][2], wl1[3][2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int k = 0; k < 5; ++k) mom[a][k] = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) abc[a][k] = 0.0f;
    xy[a][0] = xy[a][1] = wl1[a][0] = wl1[a][1] = 0.0f;
  }

  const int r_begin = sc.y0 - 2, r_end = sc.y1 + 1;
  // three rows of (I, W, w) are kept in flight: a row is ~100 instructions of work, far less than a DRAM round trip
  float pre[3][3];   // two rows in flight in a three-slot ring (see pair kernel)
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const unsigned o = (unsigned)min(max(r_begin + u, 0), H - 1) * W + colc;    // unconditional, clamped (see pair kernel)
    pre[u][0] = __ldg(img + o);
    pre[u][1] = __ldg(wrp + o);
    pre[u][2] = __ldg(wgt + o);
  }
  for (int rb = r_begin; rb <= r_end; rb += 3) {
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int r = rb + u;
      if (r > r_end) break;
      // ring slots: row r -> u, q = r-1 -> (u+2)%3, p = r-2 -> (u+1)%3
      const bool row_in = col_in && r >= 0 && r < H;
      const float vI = row_in ? pre[u][0] : 0.0f, vW = row_in ? pre[u][1] : 0.0f, vw = row_in ? pre[u][2] : 0.0f;
      {
        const unsigned o = (unsigned)min(r + 2, H - 1) * W + colc;
        float* nx = pre[(u + 2) % 3];
        nx[0] = __ldg(img + o);
        nx[1] = __ldg(wrp + o);
        nx[2] = __ldg(wgt + o);
      }
      const float df = vI - vW;
      const float sg = df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f);
      wl1[u][0] = vw;
      wl1[u][1] = -sg * coef_l1 * vw;
      xy[u][0] = vI * vw;
      xy[u][1] = vW * vw;
      hsum_moments(xy[u][0], xy[u][1], mom[u]);

      const int q = r - 1;
      if (q >= sc.y0 - 1) {
        float a = 0.0f, bb = 0.0f, cc = 0.0f;
        if (col_in && q >= 0 && q < H) {
          const SsimTerms t = ssim_from_sums(mom[0], mom[1], mom[2]);
          const float term = fmaf(-0.5f, t.S, 0.5f);
          if (term >= 0.0f && term <= 1.0f) {
            const float k = coef_ss * t.invD;
            a = k * fmaf(-t.S, 2.0f * t.Sy * (t.B2 - t.B1), 2.0f * t.Sx * (t.A2 - t.A1));
            bb = -9.0f * k * t.S * t.B1;
            cc = 18.0f * k * t.A1;
          }
        }
        float* dst = abc[(u + 2) % 3];
        dst[0] = __shfl_up_sync(kFullMask, a, 1) + a + __shfl_down_sync(kFullMask, a, 1);
        dst[1] = __shfl_up_sync(kFullMask, bb, 1) + bb + __shfl_down_sync(kFullMask, bb, 1);
        dst[2] = __shfl_up_sync(kFullMask, cc, 1) + cc + __shfl_down_sync(kFullMask, cc, 1);
      }

      const int p = r - 2;
      if (p >= sc.y0 && p < sc.y1 && col_out) {
        const int sp = (u + 1) % 3;
        const float A = abc[0][0] + abc[1][0] + abc[2][0];
        const float Bq = abc[0][1] + abc[1][1] + abc[2][1];
        const float Cq = abc[0][2] + abc[1][2] + abc[2][2];
        const float gy = fmaf(xy[sp][0], Cq, fmaf(2.0f * xy[sp][1], Bq, A));
        gout[(unsigned)p * W + sc.col] = fmaf(gy, wl1[sp][0], wl1[sp][1]);
      }
    }
  }
}

// ------------------------------------------------------------------- backward, split + pixel-pair variant
// Same decomposition as the split kernel, but every lane owns TWO adjacent columns (float2 loads/stores, 64
// columns per warp of which 60 are outputs).  ncu showed the split kernel issue-bound at 161 instructions per
// warp-row, ~75 of them addressing/predication/control: the pair layout amortises that overhead over two pixels,
// halves the shuffles per pixel (2 instead of 4 per quantity and pair) and shrinks the column halo from 4/32 to 4/64.
// Requires even W at every level (pairs are then entirely inside or outside the image).
__device__ __forceinline__ void pair_moments(float x0, float y0, float x1, float y1, float* m0, float* m1) {
  const float xl = __shfl_up_sync(kFullMask, x1, 1), yl = __shfl_up_sync(kFullMask, y1, 1);      // left neighbour of px0
  const float xr = __shfl_down_sync(kFullMask, x0, 1), yr = __shfl_down_sync(kFullMask, y0, 1);  // right neighbour of px1
  const float sx = x0 + x1, sy = y0 + y1;
  const float sxx = fmaf(x1, x1, x0 * x0), syy = fmaf(y1, y1, y0 * y0), sxy = fmaf(x1, y1, x0 * y0);
  m0[0] = sx + xl;            m1[0] = sx + xr;
  m0[1] = sy + yl;            m1[1] = sy + yr;
  m0[2] = fmaf(xl, xl, sxx);  m1[2] = fmaf(xr, xr, sxx);
  m0[3] = fmaf(yl, yl, syy);  m1[3] = fmaf(yr, yr, syy);
  m0[4] = fmaf(xl, yl, sxy);  m1[4] = fmaf(xr, yr, sxy);
}

__device__ __forceinline__ void ssim_coeffs(const float* s0, const float* s1, const float* s2, bool live, float coef_ss,
                                            float* abc) {
  abc[0] = abc[1] = abc[2] = 0.0f;
  if (!live) return;
  const SsimTerms t = ssim_from_sums(s0, s1, s2);
  const float term = fmaf(-0.5f, t.S, 0.5f);
  if (term >= 0.0f && term <= 1.0f) {        // clamp passes gradient on the closed interval
    const float k = coef_ss * t.invD;
    abc[0] = k * fmaf(-t.S, 2.0f * t.Sy * (t.B2 - t.B1), 2.0f * t.Sx * (t.A2 - t.A1));
    abc[1] = -9.0f * k * t.S * t.B1;
    abc[2] = 18.0f * k * t.A1;
  }
}

__global__ void __launch_bounds__(kSplitWarps * 32)
photo_loss_bwd_pair_kernel(const __grid_constant__ PhotoParams P, const float* __restrict__ sums,
                           const float* __restrict__ g_pixel, const float* __restrict__ g_ssim) {
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;
  const int dir = role / 3, c = role - 3 * dir;
  Strip sc;
  if (!locate_strip<2, 2>(P.T, blockIdx.x, lane, sc)) return;
  const uof_photo_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const unsigned plane = (unsigned)(H * W);
  const bool pin = sc.col >= 0 && sc.col < W;            // W even: the pair is entirely inside or outside
  const bool pout = pin && lane >= 1 && lane <= 30;
  const unsigned colc = (unsigned)min(max(sc.col, 0), W - 2);   // clamped: loads are unconditional
  const unsigned ch_base = ((unsigned)sc.b * 3u + (unsigned)c) * plane + colc, map_base = (unsigned)sc.b * plane + colc;
  const float* __restrict__ img = L.img + ch_base;
  const float* __restrict__ wrp = (dir ? L.warped_r : L.warped_l) + ch_base;
  const float* __restrict__ wgt = (dir ? L.weight_r : L.weight_l) + map_base;
  float* __restrict__ gout = (dir ? L.gwarped_r : L.gwarped_l) + ch_base;

  const float n = (float)H * (float)W;
  const float* s = sums + ((size_t)sc.level * P.T.B + sc.b) * 6;
  const float inv_div = 1.0f / (s[dir ? 3 : 1] / n + kEps);
  const float coef_l1 = __ldg(g_pixel + sc.b) * inv_div / n / 3.0f;
  const float coef_ss = -0.5f * __ldg(g_ssim + sc.b) * inv_div / (3.0f * n);

  float mom[3][2][5], abc[3][2][3], xy[3][2][2], wl1[3][2][2];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
      for (int j = 0; j < 5; ++j) mom[a][k][j] = 0.0f;
#pragma unroll
      for (int j = 0; j < 3; ++j) abc[a][k][j] = 0.0f;
      xy[a][k][0] = xy[a][k][1] = wl1[a][k][0] = wl1[a][k][1] = 0.0f;
    }

  const int r_begin = sc.y0 - 2, r_end = sc.y1 + 1;
  // Row pipeline: a per-warp shared-memory ring of kPairDepth rows x (I, W, w) float2 pairs, kPairDepth-1 rows kept in
  // flight with 8-byte cp.async (zero-filled outside the image).  Register prefetching (three rows of float2 in
  // `pre[3][3]`) left 32 % of the samples on long-scoreboard stalls: ptxas hoists the next LDGs above the last use of the
  // registers they replace and then copies the just-loaded values (ncu source page, round 2).
  __shared__ float2 ring_s[kSplitWarps][kPairDepth * 3 * 32];
  float2* ring = ring_s[role] + lane;
  auto fetch = [&](int r, int slot) {
    const bool inb = pin && r >= 0 && r < H;
    const unsigned o = (unsigned)min(max(r, 0), H - 1) * W;
    float2* d = ring + slot * (3 * 32);
    cp_async_8(d, img + o, inb);
    cp_async_8(d + 32, wrp + o, inb);
    cp_async_8(d + 64, wgt + o, inb);
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < kPairDepth - 1; ++i) fetch(r_begin + i, i);
  int slot = 0;
  for (int rb = r_begin; rb <= r_end; rb += 3) {
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int r = rb + u;
      if (r > r_end) break;
      // ring slots of the moments: row r -> u, q = r-1 -> (u+2)%3, p = r-2 -> (u+1)%3
      fetch(r + kPairDepth - 1, slot == 0 ? kPairDepth - 1 : slot - 1);
      cp_async_wait<kPairDepth - 1>();
      const float2 vI = ring[slot * 96], vW = ring[slot * 96 + 32], vw = ring[slot * 96 + 64];
      slot = slot + 1 == kPairDepth ? 0 : slot + 1;
      const float d0 = vI.x - vW.x, d1 = vI.y - vW.y;
      wl1[u][0][0] = vw.x;
      wl1[u][1][0] = vw.y;
      wl1[u][0][1] = (d0 > 0.0f ? -coef_l1 : (d0 < 0.0f ? coef_l1 : 0.0f)) * vw.x;   // d(masked L1)/dW = -sign(I-W) w coef
      wl1[u][1][1] = (d1 > 0.0f ? -coef_l1 : (d1 < 0.0f ? coef_l1 : 0.0f)) * vw.y;
      xy[u][0][0] = vI.x * vw.x; xy[u][0][1] = vW.x * vw.x;
      xy[u][1][0] = vI.y * vw.y; xy[u][1][1] = vW.y * vw.y;
      pair_moments(xy[u][0][0], xy[u][0][1], xy[u][1][0], xy[u][1][1], mom[u][0], mom[u][1]);

      const int q = r - 1;
      if (q >= sc.y0 - 1) {
        const bool q_in = pin && q >= 0 && q < H;
        float c0[3], c1[3];
        ssim_coeffs(mom[0][0], mom[1][0], mom[2][0], q_in, coef_ss, c0);
        ssim_coeffs(mom[0][1], mom[1][1], mom[2][1], q_in, coef_ss, c1);
        float* dst0 = abc[(u + 2) % 3][0];
        float* dst1 = abc[(u + 2) % 3][1];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float lft = __shfl_up_sync(kFullMask, c1[j], 1), rgt = __shfl_down_sync(kFullMask, c0[j], 1);
          const float mid = c0[j] + c1[j];
          dst0[j] = mid + lft;
          dst1[j] = mid + rgt;
        }
      }

      const int p = r - 2;
      if (p >= sc.y0 && p < sc.y1 && pout) {
        const int sp = (u + 1) % 3;
        float g[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float A = abc[0][k][0] + abc[1][k][0] + abc[2][k][0];
          const float Bq = abc[0][k][1] + abc[1][k][1] + abc[2][k][1];
          const float Cq = abc[0][k][2] + abc[1][k][2] + abc[2][k][2];
          const float gy = fmaf(xy[sp][k][0], Cq, fmaf(2.0f * xy[sp][k][1], Bq, A));
          g[k] = fmaf(gy, wl1[sp][k][0], wl1[sp][k][1]);
        }
        *reinterpret_cast<float2*>(gout + (unsigned)p * W) = make_float2(g[0], g[1]);
      }
    }
  }
  cp_async_wait<0>();
}

int fill_params(PhotoParams& P, const uof_photo_level* levels, int nlevels, int B, int halo, bool bwd, int blocks_per_sm,
                bool split = false, int ppl = 1) {
  UOF_REQUIRE(levels && nlevels >= 1 && nlevels <= UOF_MAX_LEVELS, "photo_loss: nlevels must be 1..%d", UOF_MAX_LEVELS);
  UOF_REQUIRE(B > 0, "photo_loss: bad batch %d", B);
  int H[UOF_MAX_LEVELS], W[UOF_MAX_LEVELS];
  for (int l = 0; l < nlevels; ++l) {
    const uof_photo_level& L = levels[l];
    UOF_REQUIRE(L.img && L.warped_l && L.warped_r && L.H > 0 && L.W > 0, "photo_loss: level %d incomplete", l);
    UOF_REQUIRE((long long)B * 3 * L.H * L.W < (1ll << 32), "photo_loss: level %d too large for 32-bit offsets", l);
    if (bwd) UOF_REQUIRE(L.gwarped_l && L.gwarped_r, "photo_loss_bwd: level %d has no gradient buffers", l);
    P.lv[l] = L;
    H[l] = L.H;
    W[l] = L.W;
  }
  // split backward: one block per strip; fused backward: two grid rows (directions) of 4-warp blocks
  UOF_REQUIRE(build_strip_table(P.T, H, W, nlevels, B, halo, (bwd && !split) ? 2 : 1, blocks_per_sm,
                                split ? 1 : kWarpsPerBlock, ppl) > 0,
              "photo_loss: problem too large");
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_photo_loss_fwd(const uof_photo_level* levels, int nlevels, int B, float* sums, float* loss_pixel,
                                  float* loss_ssim, uof_stream_t stream_) {
  UOF_REQUIRE(sums && loss_pixel && loss_ssim, "photo_loss_fwd: null output");
  PhotoParams P;
  static const int occ = resident_blocks(photo_loss_fwd_kernel, kWarpsPerBlock * 32);
  if (int rc = fill_params(P, levels, nlevels, B, 1, false, occ)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UOF_CUDA(cudaMemsetAsync(sums, 0, ((size_t)nlevels * B * 6 + UOF_SUMS_EXTRA) * sizeof(float), stream));
  const int blocks = ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock);
  photo_loss_fwd_kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(P, sums);
  photo_loss_finalize_kernel<<<ceil_div(B, 64), 64, 0, stream>>>(P, sums, loss_pixel, loss_ssim);
  count_launch(2);
  return check_launch("photo_loss_fwd");
}

extern "C" int uof_photo_loss_bwd(const uof_photo_level* levels, int nlevels, int B, const float* sums,
                                  const float* g_loss_pixel, const float* g_loss_ssim, uof_stream_t stream_) {
  UOF_REQUIRE(sums && g_loss_pixel && g_loss_ssim, "photo_loss_bwd: null input");
  PhotoParams P;
  bool have_weights = levels != nullptr && nlevels >= 1 && nlevels <= UOF_MAX_LEVELS;
  for (int l = 0; have_weights && l < nlevels; ++l) have_weights = levels[l].weight_l && levels[l].weight_r;
  bool even_w = have_weights;
  for (int l = 0; even_w && l < nlevels; ++l) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(levels[l].img) | reinterpret_cast<uintptr_t>(levels[l].warped_l) |
                           reinterpret_cast<uintptr_t>(levels[l].warped_r) | reinterpret_cast<uintptr_t>(levels[l].weight_l) |
                           reinterpret_cast<uintptr_t>(levels[l].weight_r) | reinterpret_cast<uintptr_t>(levels[l].gwarped_l) |
                           reinterpret_cast<uintptr_t>(levels[l].gwarped_r);
    even_w = (levels[l].W % 2 == 0) && (bits & 7u) == 0;
  }
  static const bool no_pair = getenv("UOF_PHOTO_NO_PAIR") != nullptr;
  if (even_w && !no_pair) {     // two pixels per lane, float2 accesses
    static const int occ_pair = resident_blocks(photo_loss_bwd_pair_kernel, kSplitWarps * 32);
    if (int rc = fill_params(P, levels, nlevels, B, 2, true, occ_pair, /*split=*/true, /*ppl=*/2)) return rc;
    photo_loss_bwd_pair_kernel<<<P.T.warp_begin[nlevels], kSplitWarps * 32, 0, static_cast<cudaStream_t>(stream_)>>>(
        P, sums, g_loss_pixel, g_loss_ssim);
    count_launch();
    return check_launch("photo_loss_bwd (pair)");
  }
  if (have_weights) {     // weight maps of the forward pass are available: (strip, direction, channel) warps
    static const int occ_split = resident_blocks(photo_loss_bwd_split_kernel, kSplitWarps * 32);
    if (int rc = fill_params(P, levels, nlevels, B, 2, true, occ_split, /*split=*/true)) return rc;
    photo_loss_bwd_split_kernel<<<P.T.warp_begin[nlevels], kSplitWarps * 32, 0, static_cast<cudaStream_t>(stream_)>>>(
        P, sums, g_loss_pixel, g_loss_ssim);
    count_launch();
    return check_launch("photo_loss_bwd (split)");
  }
  static const int occ = resident_blocks(photo_loss_bwd_kernel, kWarpsPerBlock * 32);
  if (int rc = fill_params(P, levels, nlevels, B, 2, true, occ)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dim3 grid(ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock), 2);
  photo_loss_bwd_kernel<<<grid, kWarpsPerBlock * 32, 0, stream>>>(P, sums, g_loss_pixel, g_loss_ssim);
  count_launch();
  return check_launch("photo_loss_bwd");
}
