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

// grid = (x chunks, H, planes): a thread produces 4 adjacent outputs of one row (one float4 store when W % 4 == 0);
// no integer division anywhere (the first version decoded a 64-bit linear index: 40 us for 29 MB).
// `out2` (nullable): a second destination for the same values -- the `up_flow` slice of the decoder's concat buffer
// (pwc_tf.py:122-123), addressed as sample = plane / C2, channel = plane % C2 with batch stride `out2_bs`.
__global__ void __launch_bounds__(128)
upsample_fwd_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out2, int C2, long long out2_bs,
                    int h, int w, int H, int W, float ry, float rx, float scale) {
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
  if (x0 >= W) return;
  const size_t plane = blockIdx.z;
  const Tap ty = tap_of(y, ry, h);
  const float* p0 = in + plane * ((size_t)h * w) + (size_t)ty.i0 * w;
  const float* p1 = in + plane * ((size_t)h * w) + (size_t)ty.i1 * w;
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const Tap tx = tap_of(min(x0 + k, W - 1), rx, w);
    // same association as ATen: l0y * (l0x v00 + l1x v01) + l1y * (l0x v10 + l1x v11)
    r[k] = scale * (ty.l0 * (tx.l0 * __ldg(p0 + tx.i0) + tx.l1 * __ldg(p0 + tx.i1)) +
                    ty.l1 * (tx.l0 * __ldg(p1 + tx.i0) + tx.l1 * __ldg(p1 + tx.i1)));
  }
  float* o = out + (plane * H + y) * (size_t)W + x0;
  if ((W & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x0 + k < W) o[k] = r[k];
  }
  if (out2) {
    const size_t b = plane / (unsigned)C2, c = plane - b * (unsigned)C2;
    float* o2 = out2 + b * (size_t)out2_bs + (c * H + y) * (size_t)W + x0;
    if ((W & 3) == 0 && (out2_bs & 3) == 0 && (reinterpret_cast<uintptr_t>(out2) & 15u) == 0) {
      *reinterpret_cast<float4*>(o2) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (x0 + k < W) o2[k] = r[k];
    }
  }
}

// Backward as a gather: grid = (x chunks, h, planes), thread = one INPUT pixel; it visits the outputs whose taps can
// include it (src in (i - 1, i + 1), i.e. dst in ((i - 0.5) / ratio - 0.5, (i + 1.5) / ratio - 0.5)) and re-derives
// their weights (separably: the column weights once, then one pass over the candidate rows).
// The incoming gradient is gout (+ g2 + g3, nullable): g2 is a channel slice of a wider buffer (sample = plane / C2, batch
// stride g2_bs -- the concat buffer's `up_flow` slice gradient), g3 dense like gout (the `res + up_flow` branch), summed on
// the fly so the three consumers of an up-sampled flow (warp, concat, residual; pwc_tf.py:119-125) need no add kernels.
template <bool MULTI>
__global__ void __launch_bounds__(128)
upsample_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ g2, int C2, long long g2_bs,
                    const float* __restrict__ g3, float* __restrict__ gin, int h, int w, int H, int W, float ry, float rx,
                    float scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (j >= w) return;
  const size_t plane = blockIdx.z;
  const float inv_ry = 1.0f / ry, inv_rx = 1.0f / rx;
  const int y_lo = max((int)floorf(((float)i - 0.5f) * inv_ry - 0.5f) - 1, 0), y_hi = min((int)ceilf(((float)i + 1.5f) * inv_ry - 0.5f) + 1, H - 1);
  const int x_lo = max((int)floorf(((float)j - 0.5f) * inv_rx - 0.5f) - 1, 0), x_hi = min((int)ceilf(((float)j + 1.5f) * inv_rx - 0.5f) + 1, W - 1);
  const float* g = gout + plane * ((size_t)H * W);
  const float* ga = nullptr;
  const float* gb3 = nullptr;
  if (MULTI) {
    const size_t b = plane / (unsigned)C2, c = plane - b * (unsigned)C2;
    if (g2) ga = g2 + b * (size_t)g2_bs + c * ((size_t)H * W);
    if (g3) gb3 = g3 + plane * ((size_t)H * W);
  }
  auto at = [&](size_t o) {
    float v = __ldg(g + o);
    if (MULTI) {
      if (ga) v += __ldg(ga + o);
      if (gb3) v += __ldg(gb3 + o);
    }
    return v;
  };
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
      const size_t gr = (size_t)y * W + x_lo;
      float row = 0.0f;
#pragma unroll
      for (int k = 0; k < kMaxTaps; ++k)
        if (wx[k] != 0.0f) row = fmaf(wx[k], at(gr + k), row);
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
        row = fmaf(wx, at((size_t)y * W + x), row);
      }
      acc = fmaf(wy, row, acc);
    }
  }
  gin[(plane * h + i) * (size_t)w + j] = scale * acc;
}

int check(const char* who, const void* a, const void* b, int planes, int h, int w, int H, int W) {
  UOF_REQUIRE(a && b, "%s: null pointer", who);
  UOF_REQUIRE(planes > 0 && h > 0 && w > 0 && H >= h && W >= w, "%s: bad shape planes=%d %dx%d -> %dx%d (up-sampling only)", who,
              planes, h, w, H, W);
  UOF_REQUIRE(planes <= 65535 && H <= 65535, "%s: too many planes / rows for one launch (planes=%d H=%d)", who, planes, H);
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_upsample_bilinear_fwd2(const float* in, float* out, float* out2, int C2, long long out2_batch_stride,
                                          int planes, int h, int w, int H, int W, float scale, uof_stream_t stream_) {
  if (int rc = check("upsample_bilinear_fwd", in, out, planes, h, w, H, W)) return rc;
  if (out2)
    UOF_REQUIRE(C2 > 0 && planes % C2 == 0 && out2_batch_stride >= (long long)C2 * H * W,
                "upsample_bilinear_fwd2: bad second destination (C2=%d planes=%d)", C2, planes);
  upsample_fwd_kernel<<<dim3(ceil_div(ceil_div(W, 4), 128), H, planes), 128, 0, static_cast<cudaStream_t>(stream_)>>>(
      in, out, out2, out2 ? C2 : 1, out2_batch_stride, h, w, H, W, (float)h / (float)H, (float)w / (float)W, scale);
  count_launch();
  return check_launch("upsample_bilinear_fwd");
}

extern "C" int uof_upsample_bilinear_fwd(const float* in, float* out, int planes, int h, int w, int H, int W, float scale,
                                         uof_stream_t stream) {
  return uof_upsample_bilinear_fwd2(in, out, nullptr, 1, 0, planes, h, w, H, W, scale, stream);
}

extern "C" int uof_upsample_bilinear_bwd3(const float* gout, const float* g2, int C2, long long g2_batch_stride, const float* g3,
                                          float* gin, int planes, int h, int w, int H, int W, float scale,
                                          uof_stream_t stream_) {
  if (int rc = check("upsample_bilinear_bwd", gout, gin, planes, h, w, H, W)) return rc;
  if (g2)
    UOF_REQUIRE(C2 > 0 && planes % C2 == 0 && g2_batch_stride >= (long long)C2 * H * W,
                "upsample_bilinear_bwd3: bad strided gradient (C2=%d planes=%d)", C2, planes);
  const dim3 grid(ceil_div(w, 128), h, planes);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (g2 || g3)
    upsample_bwd_kernel<true><<<grid, 128, 0, stream>>>(gout, g2, g2 ? C2 : 1, g2_batch_stride, g3, gin, h, w, H, W,
                                                        (float)h / (float)H, (float)w / (float)W, scale);
  else
    upsample_bwd_kernel<false><<<grid, 128, 0, stream>>>(gout, nullptr, 1, 0, nullptr, gin, h, w, H, W, (float)h / (float)H,
                                                         (float)w / (float)W, scale);
  count_launch();
  return check_launch("upsample_bilinear_bwd");
}

extern "C" int uof_upsample_bilinear_bwd(const float* gout, float* gin, int planes, int h, int w, int H, int W, float scale,
                                         uof_stream_t stream) {
  return uof_upsample_bilinear_bwd3(gout, nullptr, 1, 0, nullptr, gin, planes, h, w, H, W, scale, stream);
}
