// a6 seam: SSIM(x, y) map with 3x3 box moments, forward and backward.
// Replaces /root/reference/core/networks/pytorch_ssim/ssim.py:4-19 (5 avg_pool2d + ~25 elementwise
// launches, ~10 full-size temporaries) with one launch each way.  Same marching-warp scheme as the
// fused photometric kernel (photo_loss.cu); planes = B*C are independent.
#include "strips.cuh"

namespace uof {
namespace {

constexpr float C1 = 0.01f * 0.01f;
constexpr float C2 = 0.03f * 0.03f;
constexpr float kInv9 = 1.0f / 9.0f;
constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ float sum3_lanes(float v) {
  return __shfl_up_sync(kFullMask, v, 1) + v + __shfl_down_sync(kFullMask, v, 1);
}

struct Moments {
  float mux, muy, A1, A2, B1, B2, S;
};

__device__ __forceinline__ Moments moments_of(const float* s0, const float* s1, const float* s2) {
  Moments t;
  t.mux = (s0[0] + s1[0] + s2[0]) * kInv9;
  t.muy = (s0[1] + s1[1] + s2[1]) * kInv9;
  const float sxx = (s0[2] + s1[2] + s2[2]) * kInv9;
  const float syy = (s0[3] + s1[3] + s2[3]) * kInv9;
  const float sxy = (s0[4] + s1[4] + s2[4]) * kInv9;
  t.A1 = 2.0f * t.mux * t.muy + C1;
  t.A2 = 2.0f * (sxy - t.mux * t.muy) + C2;
  t.B1 = t.mux * t.mux + t.muy * t.muy + C1;
  t.B2 = (sxx - t.mux * t.mux) + (syy - t.muy * t.muy) + C2;
  t.S = (t.A1 * t.A2) / (t.B1 * t.B2);
  return t;
}

__device__ __forceinline__ void push_row(float (*ring)[5], float xv, float yv) {
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    ring[0][k] = ring[1][k];
    ring[1][k] = ring[2][k];
  }
  const float xl = __shfl_up_sync(kFullMask, xv, 1), xr = __shfl_down_sync(kFullMask, xv, 1);
  const float yl = __shfl_up_sync(kFullMask, yv, 1), yr = __shfl_down_sync(kFullMask, yv, 1);
  ring[2][0] = xl + xv + xr;
  ring[2][1] = yl + yv + yr;
  ring[2][2] = fmaf(xr, xr, fmaf(xv, xv, xl * xl));
  ring[2][3] = fmaf(yr, yr, fmaf(yv, yv, yl * yl));
  ring[2][4] = fmaf(xr, yr, fmaf(xv, yv, xl * yl));
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ssim_fwd_kernel(const __grid_constant__ StripTable T, const float* __restrict__ x, const float* __restrict__ y,
                float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Strip sc;
  if (!locate_strip<1>(T, gw, lane, sc)) return;
  const int H = T.H[0], W = T.W[0];
  const size_t base = (size_t)sc.b * H * W;
  const bool col_in = sc.col >= 0 && sc.col < W;
  const bool col_out = col_in && lane >= 1 && lane <= 30;
  float ring[3][5];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int k = 0; k < 5; ++k) ring[a][k] = 0.0f;
  for (int r = sc.y0 - 1; r <= sc.y1; ++r) {
    const bool inb = col_in && r >= 0 && r < H;
    const size_t off = base + (size_t)max(r, 0) * W + max(sc.col, 0);
    push_row(ring, inb ? __ldg(x + off) : 0.0f, inb ? __ldg(y + off) : 0.0f);
    const int q = r - 1;
    if (q >= sc.y0 && col_out) out[base + (size_t)q * W + sc.col] = moments_of(ring[0], ring[1], ring[2]).S;
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ssim_bwd_kernel(const __grid_constant__ StripTable T, const float* __restrict__ gout, const float* __restrict__ x,
                const float* __restrict__ y, float* __restrict__ gx, float* __restrict__ gy) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Strip sc;
  if (!locate_strip<2>(T, gw, lane, sc)) return;
  const int H = T.H[0], W = T.W[0];
  const size_t base = (size_t)sc.b * H * W;
  const bool col_in = sc.col >= 0 && sc.col < W;
  const bool col_out = col_in && lane >= 2 && lane <= 29;
  float mom[3][5], co[3][4], xy[3][2];   // co: h-sums of dS/d(mu_x, mu_y, E[x^2]=E[y^2] coefficient, E[xy]) * gout
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int k = 0; k < 5; ++k) mom[a][k] = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) co[a][k] = 0.0f;
    xy[a][0] = xy[a][1] = 0.0f;
  }
  for (int r = sc.y0 - 2; r <= sc.y1 + 1; ++r) {
    const bool inb = col_in && r >= 0 && r < H;
    const size_t off = base + (size_t)max(r, 0) * W + max(sc.col, 0);
    const float xv = inb ? __ldg(x + off) : 0.0f, yv = inb ? __ldg(y + off) : 0.0f;
    xy[0][0] = xy[1][0]; xy[0][1] = xy[1][1];
    xy[1][0] = xy[2][0]; xy[1][1] = xy[2][1];
    xy[2][0] = xv; xy[2][1] = yv;
    push_row(mom, xv, yv);
    const int q = r - 1;
    if (q >= sc.y0 - 1) {
      float ax = 0.0f, ay = 0.0f, bb = 0.0f, cc = 0.0f;
      if (col_in && q >= 0 && q < H) {
        const Moments t = moments_of(mom[0], mom[1], mom[2]);
        const float g = __ldg(gout + base + (size_t)q * W + sc.col);
        const float D = t.B1 * t.B2;
        const float dA = t.A2 - t.A1, dB = t.B2 - t.B1;
        ax = g * (2.0f * t.muy * dA - t.S * 2.0f * t.mux * dB) / D;
        ay = g * (2.0f * t.mux * dA - t.S * 2.0f * t.muy * dB) / D;
        bb = g * (-t.S / t.B2);
        cc = g * (2.0f * t.A1 / D);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        co[0][k] = co[1][k];
        co[1][k] = co[2][k];
      }
      co[2][0] = sum3_lanes(ax);
      co[2][1] = sum3_lanes(ay);
      co[2][2] = sum3_lanes(bb);
      co[2][3] = sum3_lanes(cc);
    }
    const int p = r - 2;
    if (p >= sc.y0 && p < sc.y1 && col_out) {
      const float Ax = co[0][0] + co[1][0] + co[2][0], Ay = co[0][1] + co[1][1] + co[2][1];
      const float Bq = co[0][2] + co[1][2] + co[2][2], Cq = co[0][3] + co[1][3] + co[2][3];
      const size_t o = base + (size_t)p * W + sc.col;
      gx[o] = (Ax + 2.0f * xy[0][0] * Bq + xy[0][1] * Cq) * kInv9;
      gy[o] = (Ay + 2.0f * xy[0][1] * Bq + xy[0][0] * Cq) * kInv9;
    }
  }
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_ssim_fwd(const float* x, const float* y, float* out, int N, int H, int W, uof_stream_t stream_) {
  UOF_REQUIRE(x && y && out, "ssim_fwd: null pointer");
  UOF_REQUIRE(N > 0 && H > 0 && W > 0, "ssim_fwd: bad shape N=%d H=%d W=%d", N, H, W);
  StripTable T;
  static const int occ = resident_blocks(ssim_fwd_kernel, kWarpsPerBlock * 32);
  UOF_REQUIRE(build_strip_table(T, &H, &W, 1, N, 1, 1, occ, kWarpsPerBlock) > 0, "ssim_fwd: problem too large");
  ssim_fwd_kernel<<<ceil_div(T.warp_begin[1], kWarpsPerBlock), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream_)>>>(
      T, x, y, out);
  count_launch();
  return check_launch("ssim_fwd");
}

extern "C" int uof_ssim_bwd(const float* gout, const float* x, const float* y, float* gx, float* gy, int N, int H, int W,
                            uof_stream_t stream_) {
  UOF_REQUIRE(gout && x && y && gx && gy, "ssim_bwd: null pointer");
  UOF_REQUIRE(N > 0 && H > 0 && W > 0, "ssim_bwd: bad shape N=%d H=%d W=%d", N, H, W);
  StripTable T;
  static const int occ = resident_blocks(ssim_bwd_kernel, kWarpsPerBlock * 32);
  UOF_REQUIRE(build_strip_table(T, &H, &W, 1, N, 2, 1, occ, kWarpsPerBlock) > 0, "ssim_bwd: problem too large");
  ssim_bwd_kernel<<<ceil_div(T.warp_begin[1], kWarpsPerBlock), kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream_)>>>(
      T, gout, x, y, gx, gy);
  count_launch();
  return check_launch("ssim_bwd");
}
