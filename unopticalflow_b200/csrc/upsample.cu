// Decoder glue: bilinear up-sampling of the flow fields fused with their rescaling.
// Replaces, in /root/reference/core/networks/structures/pwc_tf.py, `F.interpolate(flow, scale_factor=2.0, mode='bilinear') * 2.0`
// (:119,132,144,157) and `F.interpolate(flow * 4.0, [h, w], mode='bilinear')` (:174-177): per call one ATen up-sampling
// kernel + one multiply forward, and a scatter-with-atomics kernel + one multiply backward (8 + 7 + 23 launches, ~0.43 ms
// of the step in the ncu launch list of round 2).  Here: one streaming kernel each way, the backward pass a gather.
//
// Semantics = ATen upsample_bilinear2d, align_corners = False: src = max((dst + 0.5) * (in / out) - 0.5, 0),
// i0 = floor(src), i1 = min(i0 + 1, in - 1), lambda = src - i0; out = scale * sum of the four weighted neighbours
// (a power-of-two scale commutes exactly with the interpolation, so pre- and post-scaling give the same bits).
#include "common.cuh"

namespace uof {
namespace {

struct Tap {
  int i0, i1;
  float l0, l1;
};

__device__ __forceinline__ Tap tap_of(int dst, float ratio, int in) {
  Tap t;
  const float src = fmaxf(((float)dst + 0.5f) * ratio - 0.5f, 0.0f);
  t.i0 = min((int)src, in - 1);
  t.i1 = min(t.i0 + 1, in - 1);
  t.l1 = src - (float)t.i0;
  t.l0 = 1.0f - t.l1;
  return t;
}

// thread = one output pixel of one plane; planes = B * C
__global__ void __launch_bounds__(256)
upsample_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int H, int W, float ry, float rx,
                    float scale, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int x = (int)(t % W), y = (int)((t / W) % H);
  const long long plane = t / ((long long)W * H);
  const Tap ty = tap_of(y, ry, h), tx = tap_of(x, rx, w);
  const float* p = in + plane * ((long long)h * w);
  const float v00 = __ldg(p + ty.i0 * w + tx.i0), v01 = __ldg(p + ty.i0 * w + tx.i1);
  const float v10 = __ldg(p + ty.i1 * w + tx.i0), v11 = __ldg(p + ty.i1 * w + tx.i1);
  // same association as ATen: l0y * (l0x v00 + l1x v01) + l1y * (l0x v10 + l1x v11)
  out[t] = scale * (ty.l0 * (tx.l0 * v00 + tx.l1 * v01) + ty.l1 * (tx.l0 * v10 + tx.l1 * v11));
}

// Backward as a gather: thread = one INPUT pixel; it visits the outputs whose taps can include it
// (src in (i - 1, i + 1), i.e. dst in ((i - 0.5) / ratio - 0.5, (i + 1.5) / ratio - 0.5)) and re-derives their weights.
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const float* __restrict__ gout, float* __restrict__ gin, int h, int w, int H, int W, float ry, float rx,
                    float scale, long long total) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int j = (int)(t % w), i = (int)((t / w) % h);
  const long long plane = t / ((long long)w * h);
  const float inv_ry = 1.0f / ry, inv_rx = 1.0f / rx;
  const int y_lo = max((int)floorf(((float)i - 0.5f) * inv_ry - 0.5f) - 1, 0), y_hi = min((int)ceilf(((float)i + 1.5f) * inv_ry - 0.5f) + 1, H - 1);
  const int x_lo = max((int)floorf(((float)j - 0.5f) * inv_rx - 0.5f) - 1, 0), x_hi = min((int)ceilf(((float)j + 1.5f) * inv_rx - 0.5f) + 1, W - 1);
  const float* g = gout + plane * ((long long)H * W);
  float acc = 0.0f;
  constexpr int kMaxTaps = 16;                 // candidate columns whose weights are kept in registers (x4: 14)
  if (x_hi - x_lo < kMaxTaps) {
    float wx[kMaxTaps];
#pragma unroll
    for (int k = 0; k < kMaxTaps; ++k) {
      const Tap tx = tap_of(min(x_lo + k, W - 1), rx, w);
      wx[k] = (x_lo + k <= x_hi) ? (tx.i0 == j ? tx.l0 : 0.0f) + (tx.i1 == j ? tx.l1 : 0.0f) : 0.0f;
    }
    for (int y = y_lo; y <= y_hi; ++y) {
      const Tap ty = tap_of(y, ry, h);
      const float wy = (ty.i0 == i ? ty.l0 : 0.0f) + (ty.i1 == i ? ty.l1 : 0.0f);
      if (wy == 0.0f) continue;
      const float* gr = g + (long long)y * W + x_lo;
      float row = 0.0f;
#pragma unroll
      for (int k = 0; k < kMaxTaps; ++k)
        if (wx[k] != 0.0f) row = fmaf(wx[k], __ldg(gr + k), row);
      acc = fmaf(wy, row, acc);
    }
  } else {
    for (int y = y_lo; y <= y_hi; ++y) {
      const Tap ty = tap_of(y, ry, h);
      const float wy = (ty.i0 == i ? ty.l0 : 0.0f) + (ty.i1 == i ? ty.l1 : 0.0f);
      if (wy == 0.0f) continue;
      float row = 0.0f;
      for (int x = x_lo; x <= x_hi; ++x) {
        const Tap tx = tap_of(x, rx, w);
        const float wx = (tx.i0 == j ? tx.l0 : 0.0f) + (tx.i1 == j ? tx.l1 : 0.0f);
        row = fmaf(wx, __ldg(g + (long long)y * W + x), row);
      }
      acc = fmaf(wy, row, acc);
    }
  }
  gin[t] = scale * acc;
}

int check(const char* who, const void* a, const void* b, int planes, int h, int w, int H, int W) {
  UOF_REQUIRE(a && b, "%s: null pointer", who);
  UOF_REQUIRE(planes > 0 && h > 0 && w > 0 && H >= h && W >= w, "%s: bad shape planes=%d %dx%d -> %dx%d (up-sampling only)", who,
              planes, h, w, H, W);
  UOF_REQUIRE((long long)planes * H * W < (1ll << 40), "%s: tensor too large", who);
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_upsample_bilinear_fwd(const float* in, float* out, int planes, int h, int w, int H, int W, float scale,
                                         uof_stream_t stream_) {
  if (int rc = check("upsample_bilinear_fwd", in, out, planes, h, w, H, W)) return rc;
  const long long total = (long long)planes * H * W;
  upsample_fwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      in, out, h, w, H, W, (float)h / (float)H, (float)w / (float)W, scale, total);
  count_launch();
  return check_launch("upsample_bilinear_fwd");
}

extern "C" int uof_upsample_bilinear_bwd(const float* gout, float* gin, int planes, int h, int w, int H, int W, float scale,
                                         uof_stream_t stream_) {
  if (int rc = check("upsample_bilinear_bwd", gout, gin, planes, h, w, H, W)) return rc;
  const long long total = (long long)planes * h * w;
  upsample_bwd_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      gout, gin, h, w, H, W, (float)h / (float)H, (float)w / (float)W, scale, total);
  count_launch();
  return check_launch("upsample_bilinear_bwd");
}
