// Per-pixel arithmetic shared by the photometric kernels (photo_loss.cu, photo_warp.cu): the weight pair of
// compute_diff_weight (model_flow_paper.py:111-129), SSIM on raw 3x3 sums (pytorch_ssim/ssim.py:4-19) and its derivative,
// scalar and packed-fp32 (two pixels per lane) forms.
#pragma once
#include "common.cuh"

namespace uof {
namespace {

constexpr float C1x81 = 81.0f * 0.01f * 0.01f;      // 81 * C1, ssim.py:5
constexpr float C2x81 = 81.0f * 0.03f * 0.03f;      // 81 * C2, ssim.py:6
constexpr float kEps = 1e-12f;                      // model_flow_paper.py:97,145
constexpr float kThird = 1.0f / 3.0f;
constexpr float kInvSigma2 = 1.0f / 0.03f;          // model_flow_paper.py:126

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

// horizontal 3-tap moments of a pixel pair: left neighbour of px0 / right neighbour of px1 come from the adjacent lanes
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

typedef float2 f2;
__device__ __forceinline__ f2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }

// SSIM of two pixels at once from their raw window sums (see ssim_from_total)
__device__ __forceinline__ f2 ssim2(f2 Sx, f2 Sy, f2 Sxx, f2 Syy, f2 Sxy) {
  const f2 m1 = splat2(-1.0f), nine = splat2(9.0f);
  const f2 pxy = mul2(Sx, Sy), pxx = mul2(Sx, Sx), pyy = mul2(Sy, Sy);
  const f2 A1 = fma2(splat2(2.0f), pxy, splat2(C1x81));
  const f2 A2 = fma2(splat2(2.0f), fma2(nine, Sxy, mul2(pxy, m1)), splat2(C2x81));
  const f2 B1 = add2(add2(pxx, pyy), splat2(C1x81));
  const f2 B2 = add2(add2(fma2(nine, Sxx, mul2(pxx, m1)), fma2(nine, Syy, mul2(pyy, m1))), splat2(C2x81));
  const f2 D = mul2(B1, B2);
  const f2 invD = make_float2(__fdividef(1.0f, D.x), __fdividef(1.0f, D.y));
  return mul2(mul2(A1, A2), invD);
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

// ssim_coeffs for the two pixels of a pair at once, in packed fp32 (FFMA2 / FADD2 / FMUL2: one issue slot per pixel pair)
__device__ __forceinline__ void ssim_coeffs2(const f2* s0, const f2* s1, const f2* s2, bool live, float coef_ss, f2* abc) {
  const f2 m1 = splat2(-1.0f), nine = splat2(9.0f);
  const f2 Sx = add2(add2(s0[0], s1[0]), s2[0]), Sy = add2(add2(s0[1], s1[1]), s2[1]);
  const f2 Sxx = add2(add2(s0[2], s1[2]), s2[2]), Syy = add2(add2(s0[3], s1[3]), s2[3]), Sxy = add2(add2(s0[4], s1[4]), s2[4]);
  const f2 pxy = mul2(Sx, Sy), pxx = mul2(Sx, Sx), pyy = mul2(Sy, Sy);
  const f2 A1 = fma2(splat2(2.0f), pxy, splat2(C1x81));
  const f2 A2 = fma2(splat2(2.0f), fma2(nine, Sxy, mul2(pxy, m1)), splat2(C2x81));
  const f2 B1 = add2(add2(pxx, pyy), splat2(C1x81));
  const f2 B2 = add2(add2(fma2(nine, Sxx, mul2(pxx, m1)), fma2(nine, Syy, mul2(pyy, m1))), splat2(C2x81));
  const f2 D = mul2(B1, B2);
  const f2 invD = make_float2(__fdividef(1.0f, D.x), __fdividef(1.0f, D.y));
  const f2 S = mul2(mul2(A1, A2), invD);
  const f2 term = fma2(splat2(-0.5f), S, splat2(0.5f));
  const f2 k = mul2(splat2(coef_ss), invD);
  const f2 dN = mul2(add2(Sx, Sx), fma2(A1, m1, A2));            // 2 Sx (A2 - A1)
  const f2 dD = mul2(add2(Sy, Sy), fma2(B1, m1, B2));            // 2 Sy (B2 - B1)
  const f2 a = mul2(k, fma2(mul2(S, m1), dD, dN));               // dS/dSy  = (dN - S dD) / D
  const f2 b = mul2(mul2(splat2(-9.0f), k), mul2(S, B1));        // dS/dSyy = -9 S B1 / D
  const f2 c = mul2(mul2(splat2(18.0f), k), A1);                 // dS/dSxy = 18 A1 / D
  const bool on0 = live && term.x >= 0.0f && term.x <= 1.0f;     // clamp passes gradient on the closed interval
  const bool on1 = live && term.y >= 0.0f && term.y <= 1.0f;
  abc[0] = make_float2(on0 ? a.x : 0.0f, on1 ? a.y : 0.0f);
  abc[1] = make_float2(on0 ? b.x : 0.0f, on1 ? b.y : 0.0f);
  abc[2] = make_float2(on0 ? c.x : 0.0f, on1 ? c.y : 0.0f);
}

}  // namespace
}  // namespace uof
