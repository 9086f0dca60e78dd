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

#include "photo_math.cuh"
#include "strips.cuh"

namespace uof {
namespace {

constexpr int kWarpsPerBlock = 4;

struct PhotoParams {
  uof_photo_level lv[UOF_MAX_LEVELS];
  StripTable T;
};


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


// -------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
photo_loss_fwd_kernel(const __grid_constant__ PhotoParams P, float* __restrict__ sums) {
  pdl_trigger();      // the finalize grid is a programmatic dependent (common.cuh)
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



// ------------------------------------------------------- forward, direction-split pixel-pair variant (default)
// The fused-direction kernel above is instruction-issue bound (ncu, round 2: ~490 instructions per warp-row of 30 pixels,
// issue slots 74 % busy, FMA pipe 45 %, DRAM 21 %).  This variant halves the issued instructions per pixel:
//   * a warp owns ONE direction and TWO adjacent columns per lane (64 columns per strip, 60 of them outputs), so the
//     addressing / predicate / shuffle overhead is amortised over two pixels and the 3x5 moment state of one direction
//     (last + pair: 60 registers for the pixel pair) fits without spills;
//   * everything that is the same arithmetic on both pixels -- the vertical window sums and the whole SSIM formula -- runs
//     as packed fp32 (fma.rn.f32x2 / add.f32x2 / mul.f32x2, Blackwell's FFMA2 / FADD2 / FMUL2): one issue slot for two
//     pixels.  FFMA2 has the same FLOP throughput as FFMA (profiles/microbench), which is why it did not help the
//     FMA-bound cost volume, but it halves the issue pressure of an issue-bound kernel;
//   * rows arrive through a per-warp cp.async ring (9 float2 planes, zero-filled outside the image -- a zero pixel has
//     valid = 0, hence weight 0, so no further masking is needed).
// The two direction warps of a strip sit in the same block (shared L1 lines); both evaluate the weight pair, which needs
// all nine values of a pixel.  Requires even W and 8-byte aligned planes at every level.
constexpr int kFwdPairDepth = 3;

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
photo_loss_fwd_pair_kernel(const __grid_constant__ PhotoParams P, float* __restrict__ sums) {
  pdl_trigger();      // the finalize grid is a programmatic dependent (common.cuh)
  __shared__ float2 ring_s[kWarpsPerBlock][kFwdPairDepth * 9 * 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int dir = wid & 1;                                   // 0: left / "bwd", 1: right / "fwd"
  Strip sc;
  const bool live = locate_strip<2, 2>(P.T, blockIdx.x * (kWarpsPerBlock / 2) + (wid >> 1), lane, sc);
  if (!live) sc.level = sc.b = sc.col = sc.y0 = sc.y1 = 0;   // idle warps still join the block reduction at the end
  const uof_photo_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const unsigned plane = (unsigned)(H * W);
  const bool pin = sc.col >= 0 && sc.col < W;                // W even: the pair is entirely inside or outside
  const bool pout = pin && lane >= 1 && lane <= 30;
  const unsigned colc = (unsigned)min(max(sc.col, 0), W - 2);
  const unsigned img_base = (unsigned)sc.b * 3u * plane + colc, map_base = (unsigned)sc.b * plane + colc;
  const float* __restrict__ img = L.img + img_base;
  const float* __restrict__ wpl = L.warped_l + img_base;
  const float* __restrict__ wpr = L.warped_r + img_base;
  float* __restrict__ wmap = dir ? L.weight_r : L.weight_l;
  float* __restrict__ dmap = dir ? L.diff_r : L.diff_l;

  f2 last[3][5], pair[3][5];                                 // [channel][moment], the two pixels packed
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 5; ++k) last[c][k] = pair[c][k] = splat2(0.0f);
  float acc_l1 = 0.0f, acc_w = 0.0f;
  f2 acc_ss = splat2(0.0f);

  if (live) {
    float2* ring = ring_s[wid] + lane;
    auto fetch = [&](int r, int slot) {
      const bool inb = pin && r >= 0 && r < H;
      const unsigned o = (unsigned)min(max(r, 0), H - 1) * W;
      float2* d = ring + slot * (9 * 32);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cp_async_8(d + c * 32, img + c * plane + o, inb);
        cp_async_8(d + (3 + c) * 32, wpl + c * plane + o, inb);
        cp_async_8(d + (6 + c) * 32, wpr + c * plane + o, inb);
      }
      cp_async_commit();
    };
    const int r_begin = sc.y0 - 1, r_end = sc.y1;
#pragma unroll
    for (int i = 0; i < kFwdPairDepth - 1; ++i) fetch(r_begin + i, i);
    int slot = 0;
    for (int r = r_begin; r <= r_end; ++r) {
      fetch(r + kFwdPairDepth - 1, slot == 0 ? kFwdPairDepth - 1 : slot - 1);
      cp_async_wait<kFwdPairDepth - 1>();
      f2 v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = ring[slot * (9 * 32) + k * 32];
      slot = slot + 1 == kFwdPairDepth ? 0 : slot + 1;

      // weights of both pixels (model_flow_paper.py:111-129); an all-zero (out-of-image) pixel gets weight 0
      const float p0[9] = {v[0].x, v[1].x, v[2].x, v[3].x, v[4].x, v[5].x, v[6].x, v[7].x, v[8].x};
      const float p1[9] = {v[0].y, v[1].y, v[2].y, v[3].y, v[4].y, v[5].y, v[6].y, v[7].y, v[8].y};
      const PixelWeights w0 = pixel_weights(p0), w1 = pixel_weights(p1);
      const f2 w = dir ? make_float2(w0.wr, w1.wr) : make_float2(w0.wl, w1.wl);
      if (pout && r >= sc.y0 && r < sc.y1) {
        const f2 dd = dir ? make_float2(w0.dr, w1.dr) : make_float2(w0.dl, w1.dl);
        acc_l1 = fmaf(dd.y, w.y, fmaf(dd.x, w.x, acc_l1));
        acc_w += w.x + w.y;
        const unsigned off = map_base + (unsigned)r * W;
        if (wmap) *reinterpret_cast<float2*>(wmap + off) = w;
        if (dmap) *reinterpret_cast<float2*>(dmap + off) = dd;
      }
      const bool emit = r - 1 >= sc.y0 && pout;              // row q = r-1 now has its full 3x3 window
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const f2 x = mul2(v[c], w), y = mul2(v[3 + 3 * dir + c], w);
        float m0[5], m1[5];
        pair_moments(x.x, y.x, x.y, y.y, m0, m1);
        f2 h[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) h[k] = make_float2(m0[k], m1[k]);
        if (emit) {
          const f2 S = ssim2(add2(pair[c][0], h[0]), add2(pair[c][1], h[1]), add2(pair[c][2], h[2]), add2(pair[c][3], h[3]),
                             add2(pair[c][4], h[4]));
          const f2 t = fma2(splat2(-0.5f), S, splat2(0.5f));   // clamp((1-S)/2, 0, 1), model_flow_paper.py:144
          acc_ss = add2(acc_ss, make_float2(__saturatef(t.x), __saturatef(t.y)));
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          pair[c][k] = add2(last[c][k], h[k]);
          last[c][k] = h[k];
        }
      }
    }
    cp_async_wait<0>();
  }
  // sums layout: [0]=sum d_l*w_l [1]=sum w_l [2]=sum d_r*w_r [3]=sum w_r [4]=ssim_l [5]=ssim_r
  const float l1 = warp_sum(acc_l1), ws = warp_sum(acc_w), ss = warp_sum(acc_ss.x + acc_ss.y);
  const float acc[6] = {dir ? 0.0f : l1, dir ? 0.0f : ws, dir ? l1 : 0.0f, dir ? ws : 0.0f, dir ? 0.0f : ss, dir ? ss : 0.0f};
  block_accumulate<6, kWarpsPerBlock>(acc, live ? sums + ((size_t)sc.level * P.T.B + sc.b) * 6 : nullptr);
}

// loss_pixel[b] = sum_l sum_d mean(d*w)/(mean(w)+eps);  loss_ssim[b] likewise (model_flow_paper.py:94-98,141-147).
// A separate one-block launch: folding it into the forward kernel ("last block" pattern, as smooth_fwd does) was measured
// SLOWER here (59.7 -> 69.6 us): the per-block __threadfence has to drain the block's weight-map stores and the counter
// is one more same-address atomic per block.
__global__ void photo_loss_finalize_kernel(const __grid_constant__ PhotoParams P, const float* __restrict__ sums,
                                           float* __restrict__ loss_pixel, float* __restrict__ loss_ssim) {
  pdl_wait();         // launched while the forward grid is still running; its sums are complete and visible after this
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

  // state of the pixel pair, packed (.x = first column, .y = second): the vertical sums, the SSIM derivative and the
  // final blend are identical arithmetic on both pixels and run as packed fp32, one issue slot for the pair
  f2 mom[3][5], abc[3][3], xy[3][2], wl1[3][2];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int j = 0; j < 5; ++j) mom[a][j] = splat2(0.0f);
#pragma unroll
    for (int j = 0; j < 3; ++j) abc[a][j] = splat2(0.0f);
    xy[a][0] = xy[a][1] = wl1[a][0] = wl1[a][1] = splat2(0.0f);
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
      wl1[u][0] = vw;
      // d(masked L1)/dW = -sign(I-W) w coef
      wl1[u][1] = mul2(make_float2(d0 > 0.0f ? -coef_l1 : (d0 < 0.0f ? coef_l1 : 0.0f), d1 > 0.0f ? -coef_l1 : (d1 < 0.0f ? coef_l1 : 0.0f)), vw);
      xy[u][0] = mul2(vI, vw);
      xy[u][1] = mul2(vW, vw);
      {
        float m0[5], m1[5];
        pair_moments(xy[u][0].x, xy[u][1].x, xy[u][0].y, xy[u][1].y, m0, m1);
#pragma unroll
        for (int j = 0; j < 5; ++j) mom[u][j] = make_float2(m0[j], m1[j]);
      }

      const int q = r - 1;
      if (q >= sc.y0 - 1) {
        const bool q_in = pin && q >= 0 && q < H;
        f2 cf[3];
        ssim_coeffs2(mom[0], mom[1], mom[2], q_in, coef_ss, cf);
        f2* dst = abc[(u + 2) % 3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float lft = __shfl_up_sync(kFullMask, cf[j].y, 1), rgt = __shfl_down_sync(kFullMask, cf[j].x, 1);
          const float mid = cf[j].x + cf[j].y;
          dst[j] = make_float2(mid + lft, mid + rgt);
        }
      }

      const int p = r - 2;
      if (p >= sc.y0 && p < sc.y1 && pout) {
        const int sp = (u + 1) % 3;
        const f2 A = add2(add2(abc[0][0], abc[1][0]), abc[2][0]);
        const f2 Bq = add2(add2(abc[0][1], abc[1][1]), abc[2][1]);
        const f2 Cq = add2(add2(abc[0][2], abc[1][2]), abc[2][2]);
        const f2 gy = fma2(xy[sp][0], Cq, fma2(add2(xy[sp][1], xy[sp][1]), Bq, A));
        *reinterpret_cast<float2*>(gout + (unsigned)p * W) = fma2(gy, wl1[sp][0], wl1[sp][1]);
      }
    }
  }
  cp_async_wait<0>();
}

int fill_params(PhotoParams& P, const uof_photo_level* levels, int nlevels, int B, int halo, bool bwd, int blocks_per_sm,
                bool split = false, int ppl = 1, int strips_per_block = 0, int col_halo = -1) {
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
                                strips_per_block > 0 ? strips_per_block : (split ? 1 : kWarpsPerBlock), ppl, col_halo) > 0,
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static const bool no_fwd_pair = getenv("UOF_PHOTO_FWD_NO_PAIR") != nullptr;
  bool pair_ok = levels != nullptr && nlevels >= 1 && nlevels <= UOF_MAX_LEVELS && !no_fwd_pair;
  for (int l = 0; pair_ok && l < nlevels; ++l) {
    const uintptr_t bits = reinterpret_cast<uintptr_t>(levels[l].img) | reinterpret_cast<uintptr_t>(levels[l].warped_l) |
                           reinterpret_cast<uintptr_t>(levels[l].warped_r) | reinterpret_cast<uintptr_t>(levels[l].weight_l) |
                           reinterpret_cast<uintptr_t>(levels[l].weight_r) | reinterpret_cast<uintptr_t>(levels[l].diff_l) |
                           reinterpret_cast<uintptr_t>(levels[l].diff_r);
    pair_ok = levels[l].W % 2 == 0 && levels[l].W >= 2 && (bits & 7u) == 0;
  }
  if (pair_ok) {      // direction-split pixel-pair kernel: two warps (directions) per strip, two strips per block
    static const int occ_pair = resident_blocks(photo_loss_fwd_pair_kernel, kWarpsPerBlock * 32);
    if (int rc = fill_params(P, levels, nlevels, B, 1, false, occ_pair, false, 2, kWarpsPerBlock / 2, 2)) return rc;
    UOF_CUDA(cudaMemsetAsync(sums, 0, ((size_t)nlevels * B * 6 + UOF_SUMS_EXTRA) * sizeof(float), stream));
    photo_loss_fwd_pair_kernel<<<ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock / 2), kWarpsPerBlock * 32, 0, stream>>>(P, sums);
    UOF_CUDA(launch_dependent(photo_loss_finalize_kernel, dim3(ceil_div(B, 64)), dim3(64), stream, P, (const float*)sums, loss_pixel, loss_ssim));
    count_launch(2);
    return check_launch("photo_loss_fwd (pair)");
  }
  static const int occ = resident_blocks(photo_loss_fwd_kernel, kWarpsPerBlock * 32);
  if (int rc = fill_params(P, levels, nlevels, B, 1, false, occ)) return rc;      // argument validation before any CUDA call
  UOF_CUDA(cudaMemsetAsync(sums, 0, ((size_t)nlevels * B * 6 + UOF_SUMS_EXTRA) * sizeof(float), stream));
  const int blocks = ceil_div(P.T.warp_begin[nlevels], kWarpsPerBlock);
  photo_loss_fwd_kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(P, sums);
  UOF_CUDA(launch_dependent(photo_loss_finalize_kernel, dim3(ceil_div(B, 64)), dim3(64), stream, P, (const float*)sums, loss_pixel, loss_ssim));
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
