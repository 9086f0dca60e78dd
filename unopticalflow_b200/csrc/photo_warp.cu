// a3+a4+a5+a6 fused: the masked image warps of Model_flow.forward evaluated inside the photometric kernels.
// Replaces, in /root/reference/core/networks/model_flow_paper.py, warp_flow_pyramid x2 (:62-66, :236-237 ->
// structures/net_utils.py:16-54), compute_diff_weight (:101-134), compute_loss_with_mask x2 (:90-99) and
// compute_loss_ssim x2 (:137-148): ONE launch forward and ONE backward for all pyramid levels and both directions
// (before: 3 image-warp launches + photo_loss each way, with the warped pyramids making a round trip through HBM in
// between and d loss / d warped making another one on the way back).
//
// Forward = photo_loss_fwd_pair_kernel (photo_loss.cu: a warp owns ONE direction and two adjacent columns per lane and
// marches down a strip of rows; packed-fp32 SSIM) whose warped operand is not loaded but gathered: each lane forms the
// bilinear footprints of its two pixels from the flow (same coordinate chain, same validity mask, same blend order as
// warp.cu, so the values are bit-identical to uof_warp_fwd's) and gathers 4 corners x 3 channels per pixel.  The weight
// pair of compute_diff_weight needs the mean absolute difference of BOTH directions; the two direction warps of a strip
// sit in the same block, march in lock step and exchange that one float2 per row through shared memory (double buffered,
// one 64-thread named barrier per row).
//
// Backward = photo_loss_bwd_pair_kernel (block = strip, warp = direction x channel) with the warp backward as its
// epilogue: d loss / d flow = mask * sum_c gW_c * d(bilinear)/d(ix, iy).  The three channel warps of a direction leave
// their gW pairs of three consecutive rows in shared memory, meet at one 96-thread named barrier per three rows, and
// warp c then handles row c of the triple for all three channels (footprint, 24 gathers, reduction over channels, one
// float2 store per flow component), so the per-pixel coordinate work is done once per pixel and direction, not once
// per channel warp.  The kernels are instruction-issue bound, which is why the forward pass
// still STORES the warped values (3 x STG.64 per lane and row) for the backward pass instead of re-deriving them in
// each of its channel warps (a footprint is ~80 instructions per pixel).
#include <stdlib.h>

#include "photo_math.cuh"
#include "strips.cuh"

namespace uof {
namespace {

constexpr float kMaskThreshold = 0.9999f;   // net_utils.py:50
constexpr int kFwdWarps = 4;                // 2 strips x 2 directions
constexpr int kFwdDepth = 3;                // rows in the forward cp.async ring (the pipeline below assumes 3)
constexpr int kFwdPlanes = 5;               // img x3, flow x2
constexpr int kBwdWarps = 6;                // direction x channel
#ifndef UOF_PW_BWD_DEPTH
#define UOF_PW_BWD_DEPTH 4
#endif
constexpr int kBwdDepth = UOF_PW_BWD_DEPTH;
constexpr int kBwdSmem = kBwdWarps * (6 * 32 * 16 + kBwdDepth * 3 * 32 * 8 + 10 * 32 * 4) + 2 * 2 * 3 * 3 * 32 * 8;

struct PWParams {
  uof_photo_warp_level lv[UOF_MAX_LEVELS];
  StripTable T;
};

// barrier ids are immediates: with a register id ptxas reserves all 16 hardware barriers for the CTA
template <int COUNT>
__device__ __forceinline__ void named_barrier(int which) {
  if (which == 0)
    asm volatile("bar.sync 1, %0;" ::"n"(COUNT) : "memory");
  else
    asm volatile("bar.sync 2, %0;" ::"n"(COUNT) : "memory");
}

// Bilinear footprint of one pixel: clamped plane offsets of the four corners (always valid addresses) and their weights,
// zero for out-of-bounds corners and -- masked -- for pixels whose in-bounds weight sum is below 0.9999 (net_utils.py:47-52).
struct Foot {
  int o00, o01, o10, o11;
  float w00, w01, w10, w11;
};

template <int FLAGS>
__device__ __forceinline__ Foot make_foot(float px, float py, float fx, float fy, int H, int W, bool inside) {
  const Bilinear bl = make_bilinear(sample_coord(px, fx, W, FLAGS), sample_coord(py, fy, H, FLAGS), H, W);
  // grid_sample of a ones image accumulates nw, ne, sw, se in this order (warp.cu:cover_of)
  const float cover = ((bl.w00 + bl.w01) + bl.w10) + bl.w11;
  const float m = (!inside || cover < kMaskThreshold) ? 0.0f : 1.0f;
  const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
  const int ya = min(max(bl.y0, 0), H - 1) * W, yb = min(max(bl.y0 + 1, 0), H - 1) * W;
  Foot f;
  f.o00 = ya + xa; f.o01 = ya + xb; f.o10 = yb + xa; f.o11 = yb + xb;
  // (v*w)*m == v*(w*m) exactly for m in {0,1}
  f.w00 = bl.w00 * m; f.w01 = bl.w01 * m; f.w10 = bl.w10 * m; f.w11 = bl.w11 * m;
  return f;
}

// -------------------------------------------------------------------------------------- forward
template <int FLAGS, int MINB>
__global__ void __launch_bounds__(kFwdWarps * 32, MINB)
photo_warp_fwd_kernel(const __grid_constant__ PWParams P, float* __restrict__ sums) {
  pdl_trigger();      // the finalize grid is a programmatic dependent (common.cuh)
  __shared__ float2 ring_s[kFwdWarps][kFwdDepth * kFwdPlanes * 32];
  __shared__ float4 gat_s[kFwdWarps][2][6 * 32];             // gathered corners, see below
  __shared__ float2 xch_s[kFwdWarps][2][32];                 // [warp][row parity][lane]: this warp's mean |I - W| pair
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int dir = wid & 1;                                   // 0: left / "bwd", 1: right / "fwd"
  Strip sc;
  const bool live = locate_strip<2, 2>(P.T, blockIdx.x * (kFwdWarps / 2) + (wid >> 1), lane, sc);
  if (!live) sc.level = sc.b = sc.col = sc.y0 = sc.y1 = 0;   // idle warps still join the block reduction at the end
  const uof_photo_warp_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const unsigned plane = (unsigned)(H * W);
  const bool pin = sc.col >= 0 && sc.col < W;                // W even: the pair is entirely inside or outside
  const bool pout = pin && lane >= 1 && lane <= 30;
  const unsigned colc = (unsigned)min(max(sc.col, 0), W - 2);
  const unsigned img_base = (unsigned)sc.b * 3u * plane, map_base = (unsigned)sc.b * plane + colc;
  const float* __restrict__ img = L.img + img_base + colc;
  const float* __restrict__ flo = (dir ? L.flow_r : L.flow_l) + (size_t)sc.b * 2u * plane + colc;
  const float* __restrict__ src = (dir ? L.src_r : L.src_l) + img_base;
  float* __restrict__ wout = dir ? L.warped_r : L.warped_l;
  float* __restrict__ wmap = dir ? L.weight_r : L.weight_l;
  float* __restrict__ dmap = dir ? L.diff_r : L.diff_l;
  const float px0 = (float)sc.col, px1 = (float)(sc.col + 1);

  f2 last[3][5], pair[3][5];                                 // [channel][moment], the two pixels packed
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 5; ++k) last[c][k] = pair[c][k] = splat2(0.0f);
  float acc_l1 = 0.0f, acc_w = 0.0f;
  f2 acc_ss = splat2(0.0f);

  if (live) {
    float2* ring = ring_s[wid] + lane;
    float2* xmine = &xch_s[wid][0][lane];
    const float2* xother = &xch_s[wid ^ 1][0][lane];
    const int bar_id = wid >> 1;
    auto fetch = [&](int r, int slot) {
      const bool inb = pin && r >= 0 && r < H;
      const unsigned o = (unsigned)min(max(r, 0), H - 1) * W;
      float2* d = ring + slot * (kFwdPlanes * 32);
#pragma unroll
      for (int c = 0; c < 3; ++c) cp_async_8(d + c * 32, img + c * plane + o, inb);
      cp_async_8(d + 3 * 32, flo + o, inb);
      cp_async_8(d + 4 * 32, flo + plane + o, inb);
      cp_async_commit();
    };
    // The warp (net_utils.py:39-54), software-pipelined by one row: while row r is processed, the footprints of row r+1 are
    // formed from its flow (already in the ring) and its 4 corners x 3 channels x 2 pixels are gathered with 4-byte
    // cp.async into a per-warp shared-memory slot; only the 8 blend weights stay in registers.  (Gathering into registers
    // right before the blend left 27 % of the samples on long-scoreboard stalls at the first use, ncu round 2.)
    float4* gat = gat_s[wid][0] + lane;                      // [slot][channel * 2 + pixel][lane] = 4 corners
    float wq[8];
    auto gather = [&](int r, int rslot, int gslot) {
      const f2 fx = ring[rslot * (kFwdPlanes * 32) + 3 * 32], fy = ring[rslot * (kFwdPlanes * 32) + 4 * 32];
      const bool inb = pin && r >= 0 && r < H;
      const Foot f0 = make_foot<FLAGS>(px0, (float)r, fx.x, fy.x, H, W, inb);
      const Foot f1 = make_foot<FLAGS>(px1, (float)r, fx.y, fy.y, H, W, inb);
      float* d = reinterpret_cast<float*>(gat + gslot * (6 * 32));
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* sp = src + c * plane;
        float* d0 = d + (c * 2) * (32 * 4);
        float* d1 = d0 + 32 * 4;
        cp_async_4(d0, sp + f0.o00, true); cp_async_4(d0 + 1, sp + f0.o01, true);
        cp_async_4(d0 + 2, sp + f0.o10, true); cp_async_4(d0 + 3, sp + f0.o11, true);
        cp_async_4(d1, sp + f1.o00, true); cp_async_4(d1 + 1, sp + f1.o01, true);
        cp_async_4(d1 + 2, sp + f1.o10, true); cp_async_4(d1 + 3, sp + f1.o11, true);
      }
      cp_async_commit();
      wq[0] = f0.w00; wq[1] = f0.w01; wq[2] = f0.w10; wq[3] = f0.w11;
      wq[4] = f1.w00; wq[5] = f1.w01; wq[6] = f1.w10; wq[7] = f1.w11;
    };
    const int r_begin = sc.y0 - 1, r_end = sc.y1;
    fetch(r_begin, 0);
    fetch(r_begin + 1, 1);
    cp_async_wait<1>();
    gather(r_begin, 0, 0);
    int slot = 0;
    for (int r = r_begin; r <= r_end; ++r) {
      // groups in flight: [rows r+1] [gathers of row r] -> + [row r+2]; all but the newest must have landed
      const int slot1 = slot + 1 == kFwdDepth ? 0 : slot + 1;
      fetch(r + 2, slot == 0 ? kFwdDepth - 1 : slot - 1);
      cp_async_wait<1>();
      f2 v[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) v[k] = ring[slot * (kFwdPlanes * 32) + k * 32];
      f2 wv[3];
      {
        const float4* gs = gat + ((r - r_begin) & 1) * (6 * 32);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float4 a = gs[(c * 2) * 32], b = gs[(c * 2 + 1) * 32];
          wv[c].x = fmaf(a.w, wq[3], fmaf(a.z, wq[2], fmaf(a.y, wq[1], a.x * wq[0])));
          wv[c].y = fmaf(b.w, wq[7], fmaf(b.z, wq[6], fmaf(b.y, wq[5], b.x * wq[4])));
        }
      }
      if (r < r_end) gather(r + 1, slot1, (r + 1 - r_begin) & 1);
      slot = slot1;

      // weights of both pixels (model_flow_paper.py:111-129): this direction's mean |I - W| and validity are local, the other
      // direction's mean difference comes from the partner warp
      f2 dd;
      dd.x = (fabsf(v[0].x - wv[0].x) + fabsf(v[1].x - wv[1].x) + fabsf(v[2].x - wv[2].x)) * kThird;
      dd.y = (fabsf(v[0].y - wv[0].y) + fabsf(v[1].y - wv[1].y) + fabsf(v[2].y - wv[2].y)) * kThird;
      const int par = (r - r_begin) & 1;
      xmine[par * 32] = dd;
#ifndef UOF_PW_NO_SYNCWARP
      __syncwarp();
#endif
      named_barrier<64>(bar_id);
      const f2 od = xother[par * 32];
      const float val0 = (wv[0].x == 0.0f && wv[1].x == 0.0f && wv[2].x == 0.0f) ? 0.0f : 1.0f;   // :111-112
      const float val1 = (wv[0].y == 0.0f && wv[1].y == 0.0f && wv[2].y == 0.0f) ? 0.0f : 1.0f;
      f2 w;
      {
        const float t0 = __expf(-fabsf(dd.x - od.x)), t1 = __expf(-fabsf(dd.y - od.y));
        const float h0 = __fdividef(1.0f, 1.0f + t0) - 0.5f, h1 = __fdividef(1.0f, 1.0f + t1) - 0.5f;
        w.x = 2.0f * __expf(-(h0 * h0) * kInvSigma2) * val0;
        w.y = 2.0f * __expf(-(h1 * h1) * kInvSigma2) * val1;
      }
      if (pout && r >= sc.y0 && r < sc.y1) {
        acc_l1 = fmaf(dd.y, w.y, fmaf(dd.x, w.x, acc_l1));
        acc_w += w.x + w.y;
        const unsigned off = map_base + (unsigned)r * W;
        if (wmap) *reinterpret_cast<float2*>(wmap + off) = w;
        if (dmap) *reinterpret_cast<float2*>(dmap + off) = dd;
        if (wout) {
          float* wo = wout + img_base + colc + (unsigned)r * W;
#pragma unroll
          for (int c = 0; c < 3; ++c) *reinterpret_cast<float2*>(wo + c * plane) = wv[c];
        }
      }
      const bool emit = r - 1 >= sc.y0 && pout;              // row q = r-1 now has its full 3x3 window
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const f2 x = mul2(v[c], w), y = mul2(wv[c], w);
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
  block_accumulate<6, kFwdWarps>(acc, live ? sums + ((size_t)sc.level * P.T.B + sc.b) * 6 : nullptr);
}

// loss_pixel[b] = sum_l sum_d mean(d*w)/(mean(w)+eps);  loss_ssim[b] likewise (model_flow_paper.py:94-98,141-147);
// programmatic dependent of the forward grid (see photo_loss.cu:photo_loss_finalize_kernel)
__global__ void photo_warp_finalize_kernel(const __grid_constant__ PWParams P, const float* __restrict__ sums,
                                           float* __restrict__ loss_pixel, float* __restrict__ loss_ssim) {
  pdl_wait();
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
// See photo_loss.cu:photo_loss_bwd_pair_kernel for the photometric part (identical arithmetic); what is new is the
// epilogue: instead of storing gW = d loss / d warped, the channel warps hand it over in shared memory and the flow
// gradient is formed here (warp.cu:warp_bwd_nchw_kernel without the scatter to the image).
template <int FLAGS, int MINB>
__global__ void __launch_bounds__(kBwdWarps * 32) __maxnreg__(MINB == 3 ? 112 : (MINB == 2 ? 168 : 255))
photo_warp_bwd_kernel(const __grid_constant__ PWParams P, const float* __restrict__ sums,
                      const float* __restrict__ g_pixel, const float* __restrict__ g_ssim) {
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;
  const int dir = role / 3, c = role - 3 * dir;
  Strip sc;
  if (!locate_strip<2, 2>(P.T, blockIdx.x, lane, sc)) return;
  const uof_photo_warp_level& L = P.lv[sc.level];
  const int H = L.H, W = L.W;
  const unsigned plane = (unsigned)(H * W);
  const bool pin = sc.col >= 0 && sc.col < W;            // W even: the pair is entirely inside or outside
  const bool pout = pin && lane >= 1 && lane <= 30;
  const unsigned colc = (unsigned)min(max(sc.col, 0), W - 2);   // clamped: loads are unconditional
  const unsigned img_base = (unsigned)sc.b * 3u * plane;
  const unsigned ch_base = img_base + (unsigned)c * plane + colc, map_base = (unsigned)sc.b * plane + colc;
  const float* __restrict__ img = L.img + ch_base;
  const float* __restrict__ wrp = (dir ? L.warped_r : L.warped_l) + ch_base;
  const float* __restrict__ wgt = (dir ? L.weight_r : L.weight_l) + map_base;
  const float* __restrict__ src = (dir ? L.src_r : L.src_l) + img_base;
  const float* __restrict__ flo = (dir ? L.flow_r : L.flow_l) + (size_t)sc.b * 2u * plane + colc;
  float* __restrict__ gfl = (dir ? L.gflow_r : L.gflow_l) + (size_t)sc.b * 2u * plane + colc;

  const float n = (float)H * (float)W;
  const float* s = sums + ((size_t)sc.level * P.T.B + sc.b) * 6;
  const float inv_div = 1.0f / (s[dir ? 3 : 1] / n + kEps);
  const float coef_l1 = __ldg(g_pixel + sc.b) * inv_div / n / 3.0f;
  const float coef_ss = -0.5f * __ldg(g_ssim + sc.b) * inv_div / (3.0f * n);
  // d(ix)/d(flow x), d(iy)/d(flow y) of the coordinate chain (common.cuh:coord_scale)
  const float dgx = 2.0f / (float)(W - 1 > 1 ? W - 1 : 1), dgy = 2.0f / (float)(H - 1 > 1 ? H - 1 : 1);
  const float sx = (FLAGS & kAlignCorners) ? dgx * 0.5f * (float)(W - 1) : dgx * 0.5f * (float)W;
  const float sy = (FLAGS & kAlignCorners) ? dgy * 0.5f * (float)(H - 1) : dgy * 0.5f * (float)H;

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
  // dynamic shared memory (52.5 KB > the 48 KB static limit), carved with typed pointer arithmetic on the array itself
  extern __shared__ float4 bwd_smem[];
  float4* const gat_all = bwd_smem;                                                        // [warp][6 * 32] float4
  float2* const ring_all = reinterpret_cast<float2*>(gat_all + kBwdWarps * 6 * 32);        // [warp][depth * 3 * 32] float2
  // gW hand-over: [triple parity][direction][row of the triple][channel][lane]
  float2 (*gw_s)[2][3][3][32] = reinterpret_cast<float2 (*)[2][3][3][32]>(ring_all + kBwdWarps * kBwdDepth * 3 * 32);
  float* const fsc_all = reinterpret_cast<float*>(&gw_s[2][0][0][0][0]);                   // [warp][10 * 32] float
  float2* ring = ring_all + role * (kBwdDepth * 3 * 32) + lane;
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
  for (int i = 0; i < kBwdDepth - 1; ++i) fetch(r_begin + i, i);
  // Epilogue pipeline: the flow of this warp's row of the NEXT triple is loaded one triple ahead (registers); at the top
  // of a triple the footprints of its row are formed, the 4 corners x 3 channels x 2 pixels are gathered with cp.async
  // (zero-filled for out-of-bounds corners) and the five scalars per pixel the flow gradient needs are parked in shared
  // memory; the three rows of photometric work then hide the gather latency.
  float* gat = reinterpret_cast<float*>(gat_all + role * (6 * 32) + lane);   // [channel * 2 + pixel][lane] = corners 00, 01, 10, 11
  float* fsc = fsc_all + role * (10 * 32) + lane;                            // [pixel * 5 + {tx, ty, ux, uy, mask}][lane]
  auto flow_of = [&](int p, float2& fx, float2& fy) {
    fx = fy = make_float2(0.0f, 0.0f);
    if (pout && p >= sc.y0 && p < sc.y1) {
      fx = __ldg(reinterpret_cast<const float2*>(flo + (unsigned)p * W));
      fy = __ldg(reinterpret_cast<const float2*>(flo + plane + (unsigned)p * W));
    }
  };
  float2 fxn, fyn;
  flow_of(r_begin + c - 2, fxn, fyn);
  int slot = 0, par = 0;
  for (int rb = r_begin; rb <= r_end; rb += 3, par ^= 1) {
    // this warp's row of the triple: p = rb + c - 2
    const int pmine = rb + c - 2;
    const bool mine = pout && pmine >= sc.y0 && pmine < sc.y1;
    if (mine) {
#ifdef UOF_PW_TOP_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
      for (int k = 0; k < 2; ++k) {       // not unrolled: 12 instead of 24 64-bit gather addresses live on top of the row state
        const float ix = sample_coord((float)(sc.col + k), k ? fxn.y : fxn.x, W, FLAGS);
        const float iy = sample_coord((float)pmine, k ? fyn.y : fyn.x, H, FLAGS);
        const Bilinear bl = make_bilinear(ix, iy, H, W);
        const float cover = ((bl.w00 + bl.w01) + bl.w10) + bl.w11;
        const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
        const int ya = min(max(bl.y0, 0), H - 1) * W, yb = min(max(bl.y0 + 1, 0), H - 1) * W;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const float* sp = src + cc * plane;
          float* d = gat + (cc * 2 + k) * (32 * 4);
          cp_async_4(d, sp + ya + xa, bl.in00);
          cp_async_4(d + 1, sp + ya + xb, bl.in01);
          cp_async_4(d + 2, sp + yb + xa, bl.in10);
          cp_async_4(d + 3, sp + yb + xb, bl.in11);
        }
        fsc[(k * 5 + 0) * 32] = bl.tx;
        fsc[(k * 5 + 1) * 32] = bl.ty;
        fsc[(k * 5 + 2) * 32] = (floorf(ix) + 1.0f) - ix;
        fsc[(k * 5 + 3) * 32] = (floorf(iy) + 1.0f) - iy;
        fsc[(k * 5 + 4) * 32] = cover < kMaskThreshold ? 0.0f : 1.0f;
      }
    }
    cp_async_commit();                 // every warp commits the same number of groups per triple (see the wait below)
    flow_of(pmine + 3, fxn, fyn);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int r = rb + u;
      f2 gw = splat2(0.0f);
      fetch(r + kBwdDepth - 1, slot == 0 ? kBwdDepth - 1 : slot - 1);     // also past r_end: keeps the group count uniform
      if (r <= r_end) {
        // ring slots of the moments: row r -> u, q = r-1 -> (u+2)%3, p = r-2 -> (u+1)%3
        cp_async_wait<kBwdDepth - 1>();
        const float2 vI = ring[slot * 96], vW = ring[slot * 96 + 32], vw = ring[slot * 96 + 64];
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
          gw = fma2(gy, wl1[sp][0], wl1[sp][1]);           // d loss / d warped_c at (p, pair)
        }
      }
      slot = slot + 1 == kBwdDepth ? 0 : slot + 1;
      gw_s[par][dir][u][c][lane] = gw;
    }
#ifndef UOF_PW_NO_SYNCWARP
    __syncwarp();
#endif
    named_barrier<96>(dir);      // the three channel warps of this direction have left their gW of rows rb-2 .. rb
    cp_async_wait<3>();              // the three row groups of this triple may still be in flight, the gathers have landed
    if (mine) {
      // warp_flow backward for row pmine, all three channels (warp.cu:warp_bwd_nchw_kernel, NEED_GX = false)
      float gx2[2], gy2[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float tx = fsc[(k * 5 + 0) * 32], ty = fsc[(k * 5 + 1) * 32], ux = fsc[(k * 5 + 2) * 32], uy = fsc[(k * 5 + 3) * 32];
        const float msk = fsc[(k * 5 + 4) * 32];
        float gix = 0.0f, giy = 0.0f;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          const float4 v = *reinterpret_cast<const float4*>(gat + (cc * 2 + k) * (32 * 4));     // 00, 01, 10, 11
          const float2 g2 = gw_s[par][dir][c][cc][lane];
          const float gm = (k ? g2.y : g2.x) * msk;
          gix = fmaf(gm, (v.y - v.x) * uy + (v.w - v.z) * ty, gix);
          giy = fmaf(gm, (v.z - v.x) * ux + (v.w - v.y) * tx, giy);
        }
        gx2[k] = gix * sx;
        gy2[k] = giy * sy;
      }
      float* go = gfl + (unsigned)pmine * W;
      *reinterpret_cast<float2*>(go) = make_float2(gx2[0], gx2[1]);
      *reinterpret_cast<float2*>(go + plane) = make_float2(gy2[0], gy2[1]);
    }
  }
  cp_async_wait<0>();
}

int fill_params(PWParams& P, const uof_photo_warp_level* levels, int nlevels, int B, bool bwd, int blocks_per_sm) {
  UOF_REQUIRE(levels && nlevels >= 1 && nlevels <= UOF_MAX_LEVELS, "photo_warp_loss: nlevels must be 1..%d", UOF_MAX_LEVELS);
  UOF_REQUIRE(B > 0, "photo_warp_loss: bad batch %d", B);
  int H[UOF_MAX_LEVELS], W[UOF_MAX_LEVELS];
  for (int l = 0; l < nlevels; ++l) {
    const uof_photo_warp_level& L = levels[l];
    UOF_REQUIRE(L.img && L.src_l && L.src_r && L.flow_l && L.flow_r && L.H > 0 && L.W > 0, "photo_warp_loss: level %d incomplete", l);
    UOF_REQUIRE((long long)B * 3 * L.H * L.W < (1ll << 31), "photo_warp_loss: level %d too large for 32-bit offsets", l);
    if (bwd)
      UOF_REQUIRE(L.warped_l && L.warped_r && L.weight_l && L.weight_r && L.gflow_l && L.gflow_r,
                  "photo_warp_loss_bwd: level %d needs the warped images and weight maps of the forward pass and gflow buffers", l);
    const uintptr_t bits = reinterpret_cast<uintptr_t>(L.img) | reinterpret_cast<uintptr_t>(L.flow_l) | reinterpret_cast<uintptr_t>(L.flow_r) |
                           reinterpret_cast<uintptr_t>(L.warped_l) | reinterpret_cast<uintptr_t>(L.warped_r) |
                           reinterpret_cast<uintptr_t>(L.weight_l) | reinterpret_cast<uintptr_t>(L.weight_r) |
                           reinterpret_cast<uintptr_t>(L.diff_l) | reinterpret_cast<uintptr_t>(L.diff_r) |
                           reinterpret_cast<uintptr_t>(L.gflow_l) | reinterpret_cast<uintptr_t>(L.gflow_r);
    if (L.W % 2 != 0 || L.W < 2 || (bits & 7u) != 0) {
      set_error("photo_warp_loss: level %d needs even W and 8-byte aligned planes (W=%d); use uof_warp_* + uof_photo_loss_*", l, L.W);
      return UOF_ERR_UNSUPPORTED;
    }
    P.lv[l] = L;
    H[l] = L.H;
    W[l] = L.W;
  }
  // forward: two strips (x two direction warps) per block, 1-row halo; backward: one block of six warps per strip, 2-row halo
  UOF_REQUIRE(build_strip_table(P.T, H, W, nlevels, B, bwd ? 2 : 1, 1, blocks_per_sm, bwd ? 1 : kFwdWarps / 2, 2, 2) > 0,
              "photo_warp_loss: problem too large");
  return UOF_OK;
}

// Register budget of the two kernels (resident blocks per SM), chosen by measurement (profiles/r2_photo_warp.md); the
// environment variables exist for that sweep.
int env_minb(const char* name, int lo, int hi, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return v < lo ? lo : (v > hi ? hi : v);
}

template <int MINB>
int launch_fwd(PWParams& P, const uof_photo_warp_level* levels, int nlevels, int B, int flags, float* sums, cudaStream_t stream) {
  static const int occ[4] = {resident_blocks(photo_warp_fwd_kernel<0, MINB>, kFwdWarps * 32), resident_blocks(photo_warp_fwd_kernel<1, MINB>, kFwdWarps * 32),
                             resident_blocks(photo_warp_fwd_kernel<2, MINB>, kFwdWarps * 32), resident_blocks(photo_warp_fwd_kernel<3, MINB>, kFwdWarps * 32)};
  if (int rc = fill_params(P, levels, nlevels, B, false, occ[flags])) return rc;
  UOF_CUDA(cudaMemsetAsync(sums, 0, ((size_t)nlevels * B * 6 + UOF_SUMS_EXTRA) * sizeof(float), stream));
  const int blocks = ceil_div(P.T.warp_begin[nlevels], kFwdWarps / 2);
  switch (flags) {
    case 0: photo_warp_fwd_kernel<0, MINB><<<blocks, kFwdWarps * 32, 0, stream>>>(P, sums); break;
    case 1: photo_warp_fwd_kernel<1, MINB><<<blocks, kFwdWarps * 32, 0, stream>>>(P, sums); break;
    case 2: photo_warp_fwd_kernel<2, MINB><<<blocks, kFwdWarps * 32, 0, stream>>>(P, sums); break;
    default: photo_warp_fwd_kernel<3, MINB><<<blocks, kFwdWarps * 32, 0, stream>>>(P, sums); break;
  }
  return UOF_OK;
}

// opt in to > 48 KB of dynamic shared memory: a per-device, per-function attribute, set before every launch like the
// cost-volume kernels do (a host-side call of ~1 us; a cache keyed by the function TYPE would conflate the instantiations)
template <class K>
void bwd_opt_in(K kernel) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem) != cudaSuccess) (void)cudaGetLastError();
}
// resident blocks per SM (for the strip table's wave model)
template <class K>
int bwd_occupancy(K kernel) {
  int n = 0;
  bwd_opt_in(kernel);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kBwdWarps * 32, kBwdSmem) != cudaSuccess || n < 1) {
    (void)cudaGetLastError();
    n = 2;
  }
  return n;
}

template <int MINB>
int launch_bwd(PWParams& P, const uof_photo_warp_level* levels, int nlevels, int B, int flags, const float* sums, const float* gp,
               const float* gs, cudaStream_t stream) {
  static const int occ[4] = {bwd_occupancy(photo_warp_bwd_kernel<0, MINB>), bwd_occupancy(photo_warp_bwd_kernel<1, MINB>),
                             bwd_occupancy(photo_warp_bwd_kernel<2, MINB>), bwd_occupancy(photo_warp_bwd_kernel<3, MINB>)};
  if (int rc = fill_params(P, levels, nlevels, B, true, occ[flags])) return rc;
  const int blocks = P.T.warp_begin[nlevels];
  switch (flags) {
    case 0: bwd_opt_in(photo_warp_bwd_kernel<0, MINB>); photo_warp_bwd_kernel<0, MINB><<<blocks, kBwdWarps * 32, kBwdSmem, stream>>>(P, sums, gp, gs); break;
    case 1: bwd_opt_in(photo_warp_bwd_kernel<1, MINB>); photo_warp_bwd_kernel<1, MINB><<<blocks, kBwdWarps * 32, kBwdSmem, stream>>>(P, sums, gp, gs); break;
    case 2: bwd_opt_in(photo_warp_bwd_kernel<2, MINB>); photo_warp_bwd_kernel<2, MINB><<<blocks, kBwdWarps * 32, kBwdSmem, stream>>>(P, sums, gp, gs); break;
    default: bwd_opt_in(photo_warp_bwd_kernel<3, MINB>); photo_warp_bwd_kernel<3, MINB><<<blocks, kBwdWarps * 32, kBwdSmem, stream>>>(P, sums, gp, gs); break;
  }
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_photo_warp_loss_fwd(const uof_photo_warp_level* levels, int nlevels, int B, int coord_flags, float* sums,
                                       float* loss_pixel, float* loss_ssim, uof_stream_t stream_) {
  UOF_REQUIRE(sums && loss_pixel && loss_ssim, "photo_warp_loss_fwd: null output");
  UOF_REQUIRE(coord_flags >= 0 && coord_flags <= 3, "photo_warp_loss_fwd: bad coord_flags %d", coord_flags);
  PWParams P;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static const int minb = env_minb("UOF_PW_FWD_MINB", 3, 4, 3);
  if (int rc = minb == 4 ? launch_fwd<4>(P, levels, nlevels, B, coord_flags, sums, stream)
                         : launch_fwd<3>(P, levels, nlevels, B, coord_flags, sums, stream))
    return rc;
  UOF_CUDA(launch_dependent(photo_warp_finalize_kernel, dim3(ceil_div(B, 64)), dim3(64), stream, P, (const float*)sums, loss_pixel, loss_ssim));
  count_launch(2);
  return check_launch("photo_warp_loss_fwd");
}

extern "C" int uof_photo_warp_loss_bwd(const uof_photo_warp_level* levels, int nlevels, int B, int coord_flags, const float* sums,
                                       const float* g_loss_pixel, const float* g_loss_ssim, uof_stream_t stream_) {
  UOF_REQUIRE(sums && g_loss_pixel && g_loss_ssim, "photo_warp_loss_bwd: null input");
  UOF_REQUIRE(coord_flags >= 0 && coord_flags <= 3, "photo_warp_loss_bwd: bad coord_flags %d", coord_flags);
  PWParams P;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static const int minb = env_minb("UOF_PW_BWD_MINB", 1, 3, 2);
  if (int rc = minb == 3   ? launch_bwd<3>(P, levels, nlevels, B, coord_flags, sums, g_loss_pixel, g_loss_ssim, stream)
               : minb == 2 ? launch_bwd<2>(P, levels, nlevels, B, coord_flags, sums, g_loss_pixel, g_loss_ssim, stream)
                           : launch_bwd<1>(P, levels, nlevels, B, coord_flags, sums, g_loss_pixel, g_loss_ssim, stream))
    return rc;
  count_launch();
  return check_launch("photo_warp_loss_bwd");
}
