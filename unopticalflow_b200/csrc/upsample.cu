// Decoder glue: bilinear up-sampling of the flow fields fused with their rescaling.
// Replaces, in /root/reference/core/networks/structures/pwc_tf.py, `F.interpolate(flow, scale_factor=2.0, mode='bilinear') * 2.0`
// (:119,132,144,157) and `F.interpolate(flow * 4.0, [h, w], mode='bilinear')` (:174-177): per call one ATen up-sampling
// kernel + one multiply forward, and a scatter-with-atomics kernel + one multiply backward (8 + 7 + 23 launches, ~0.43 ms
// of the step in the ncu launch list of round 2).  Here: one streaming kernel each way, the backward pass a gather.
//
// Semantics = ATen upsample_bilinear2d, align_corners = False: src = max((dst + 0.5) * (in / out) - 0.5, 0),
// i0 = floor(src), i1 = min(i0 + 1, in - 1), lambda = src - i0; out = scale * sum of the four weighted neighbours
// (a power-of-two scale commutes exactly with the interpolation, so pre- and post-scaling give the same bits).
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------------------------------
// Integer ratios (H = R h, W = R w, R = 2 or 4): every up-sampling of the step (pwc_tf.py:119,132,144,157 are x2, :174 is
// x4).  The generic kernels above spend 16 loads per 4 outputs (forward) and visit 2R x 2R candidates per input pixel with
// one load each (backward): 18 / 26 us for the 29 MB of the x4 case, 0.24 / 0.17 of the HBM roofline.  With an integer
// ratio the taps have a fixed shape -- output (R i + m, R j + k) reads input rows {i-1, i} (m < R/2) or {i, i+1} and
// columns {j-1, j} (k < R/2) or {j, j+1}, indices clamped at the borders, where the clamped coordinate makes the far
// weight exactly 0 -- so a thread can own one input column, keep a 3-row x 3-column window in registers and emit R x R
// outputs per input pixel with vector stores.  The weights still come from tap_of(): (dst + 0.5) / R - 0.5 is exact in
// fp32, so indices and lambdas are the ones ATen computes, and the blend keeps ATen's association.
constexpr int kIntRows = 4;          // input rows per block

template <int R>
__global__ void __launch_bounds__(128)
upsample_fwd_int_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ out2, int C2, long long out2_bs,
                        int h, int w, float scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= w) return;
  const int H = R * h, W = R * w;
  const size_t plane = blockIdx.z;
  const int i_begin = blockIdx.y * kIntRows, i_end = min(i_begin + kIntRows, h);
  const float ratio = 1.0f / (float)R;
  float l0x[R], l1x[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const Tap t = tap_of(R * j + k, ratio, w);
    l0x[k] = t.l0;
    l1x[k] = t.l1;
  }
  const int jl = max(j - 1, 0), jr = min(j + 1, w - 1);
  const float* ip = in + plane * ((size_t)h * w);
  // horizontal blends of one input row: hx[k] = l0x v[i0(k)] + l1x v[i1(k)], shared by the output rows that read the row
  auto hrow = [&](int i, float* hx) {
    const float* rp = ip + (size_t)min(max(i, 0), h - 1) * w;
    const float a = __ldg(rp + jl), b = __ldg(rp + j), c = __ldg(rp + jr);
#pragma unroll
    for (int k = 0; k < R; ++k) hx[k] = k < R / 2 ? l0x[k] * a + l1x[k] * b : l0x[k] * b + l1x[k] * c;
  };
  float* o1 = out + (plane * H) * (size_t)W + (size_t)R * j;
  float* o2 = nullptr;
  if (out2) {
    const size_t b = plane / (unsigned)C2, c = plane - b * (unsigned)C2;
    o2 = out2 + b * (size_t)out2_bs + (c * H) * (size_t)W + (size_t)R * j;
  }
  // all kIntRows + 2 input rows of the strip are loaded up front (18 independent loads in flight per thread: with a
  // row-by-row window every iteration waited for its own three loads, 8.3 us instead of 6 for the x4 case)
  float hx[kIntRows + 2][R];
#pragma unroll
  for (int s = 0; s < kIntRows + 2; ++s) hrow(i_begin - 1 + s, hx[s]);
#pragma unroll
  for (int s = 0; s < kIntRows; ++s) {
    const int i = i_begin + s;
    if (i >= i_end) break;
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const int y = R * i + m;
      const Tap ty = tap_of(y, ratio, h);
      const float* ra = m < R / 2 ? hx[s] : hx[s + 1];
      const float* rb = m < R / 2 ? hx[s + 1] : hx[s + 2];
      float r[R];
#pragma unroll
      for (int k = 0; k < R; ++k) r[k] = scale * (ty.l0 * ra[k] + ty.l1 * rb[k]);
      if (R == 4) {
        *reinterpret_cast<float4*>(o1 + (size_t)y * W) = make_float4(r[0], r[1], r[2], r[3]);
        if (o2) *reinterpret_cast<float4*>(o2 + (size_t)y * W) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
        *reinterpret_cast<float2*>(o1 + (size_t)y * W) = make_float2(r[0], r[1]);
        if (o2) *reinterpret_cast<float2*>(o2 + (size_t)y * W) = make_float2(r[0], r[1]);
      }
    }
  }
}

// Backward, integer ratio: thread = input column j, block = kIntRows input rows.  The thread walks down the candidate
// output rows; per row it loads its own cell of R gradients with one vector load and the R/2 nearest values of the two
// neighbouring cells with one more load each (L1 hits: the neighbouring lanes load them as their own cells), forms the weighted row sum over its 2R candidate
// columns (weights from tap_of(), evaluated once) and adds it to the accumulators of the (at most two) input rows the
// output row reads.  Same summation order as the generic kernel: columns ascending, then rows ascending.
template <int R, bool MULTI>
__global__ void __launch_bounds__(128)
upsample_bwd_int_kernel(const float* __restrict__ gout, const float* __restrict__ g2, int C2, long long g2_bs,
                        const float* __restrict__ g3, float* __restrict__ gin, int h, int w, float scale) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = R * h, W = R * w;
  const size_t plane = blockIdx.z;
  const int i_begin = blockIdx.y * kIntRows;
  const float ratio = 1.0f / (float)R;
  const bool live = j < w;
  const int jc = min(j, w - 1);
  // candidate columns x = R j - R/2 + k, k = 0 .. 2R-1
  float wx[2 * R];
#pragma unroll
  for (int k = 0; k < 2 * R; ++k) {
    const int x = R * jc - R / 2 + k;
    const Tap t = tap_of(min(max(x, 0), W - 1), ratio, w);
    wx[k] = (x >= 0 && x < W) ? (t.i0 == jc ? t.l0 : 0.0f) + (t.i1 == jc ? t.l1 : 0.0f) : 0.0f;
  }
  const float* g = gout + plane * ((size_t)H * W);
  const float* ga = nullptr;
  const float* gb3 = nullptr;
  if (MULTI) {
    const size_t b = plane / (unsigned)C2, c = plane - b * (unsigned)C2;
    if (g2) ga = g2 + b * (size_t)g2_bs + c * ((size_t)H * W);
    if (g3) gb3 = g3 + plane * ((size_t)H * W);
  }
  const int xl = max(R * jc - R / 2, 0), xr = min(R * jc + R, W - R / 2);      // clamped (always valid) halo addresses
  float acc[kIntRows];
#pragma unroll
  for (int s = 0; s < kIntRows; ++s) acc[s] = 0.0f;
  const int y_lo = max(R * i_begin - R / 2, 0), y_hi = min(R * (i_begin + kIntRows) + R / 2 - 1, H - 1);
  constexpr int G = 4;               // rows per batch: their loads are all issued before the first use
  for (int y0 = y_lo; y0 <= y_hi; y0 += G) {
    float v[G][2 * R];
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const size_t ro = (size_t)min(y0 + q, H - 1) * W;      // rows past y_hi are loaded (clamped) and ignored below
      if (R == 4) {
        float4 c4 = __ldg(reinterpret_cast<const float4*>(g + ro + R * jc));
        float2 l2 = __ldg(reinterpret_cast<const float2*>(g + ro + xl)), r2 = __ldg(reinterpret_cast<const float2*>(g + ro + xr));
        if (MULTI) {
          if (ga) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(ga + ro + R * jc));
            const float2 tl = __ldg(reinterpret_cast<const float2*>(ga + ro + xl)), tr = __ldg(reinterpret_cast<const float2*>(ga + ro + xr));
            c4.x += t.x; c4.y += t.y; c4.z += t.z; c4.w += t.w; l2.x += tl.x; l2.y += tl.y; r2.x += tr.x; r2.y += tr.y;
          }
          if (gb3) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(gb3 + ro + R * jc));
            const float2 tl = __ldg(reinterpret_cast<const float2*>(gb3 + ro + xl)), tr = __ldg(reinterpret_cast<const float2*>(gb3 + ro + xr));
            c4.x += t.x; c4.y += t.y; c4.z += t.z; c4.w += t.w; l2.x += tl.x; l2.y += tl.y; r2.x += tr.x; r2.y += tr.y;
          }
        }
        v[q][0] = l2.x; v[q][1] = l2.y; v[q][2] = c4.x; v[q][3] = c4.y; v[q][4] = c4.z; v[q][5] = c4.w; v[q][6] = r2.x; v[q][7] = r2.y;
      } else {
        float2 c2 = __ldg(reinterpret_cast<const float2*>(g + ro + R * jc));
        float l1 = __ldg(g + ro + xl), r1 = __ldg(g + ro + xr);
        if (MULTI) {
          if (ga) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(ga + ro + R * jc));
            c2.x += t.x; c2.y += t.y; l1 += __ldg(ga + ro + xl); r1 += __ldg(ga + ro + xr);
          }
          if (gb3) {
            const float2 t = __ldg(reinterpret_cast<const float2*>(gb3 + ro + R * jc));
            c2.x += t.x; c2.y += t.y; l1 += __ldg(gb3 + ro + xl); r1 += __ldg(gb3 + ro + xr);
          }
        }
        v[q][0] = l1; v[q][1] = c2.x; v[q][2] = c2.y; v[q][3] = r1;
      }
    }
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const int y = y0 + q;
      float row = 0.0f;
#pragma unroll
      for (int k = 0; k < 2 * R; ++k) row = fmaf(wx[k], v[q][k], row);
      const Tap ty = tap_of(min(y, H - 1), ratio, h);
      const bool use = y <= y_hi;
#pragma unroll
      for (int s = 0; s < kIntRows; ++s) {
        const int i = i_begin + s;
        const float wy = use ? (ty.i0 == i ? ty.l0 : 0.0f) + (ty.i1 == i ? ty.l1 : 0.0f) : 0.0f;
        acc[s] = fmaf(wy, row, acc[s]);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int s = 0; s < kIntRows; ++s)
      if (i_begin + s < h) gin[(plane * h + i_begin + s) * (size_t)w + j] = scale * acc[s];
  }
}

// integer-ratio fast path applies: exact ratio 2 or 4 and vector-aligned rows
inline int int_ratio(int h, int w, int H, int W) {
  if (H == 2 * h && W == 2 * w) return 2;
  if (H == 4 * h && W == 4 * w) return 4;
  return 0;
}
inline bool aligned_to(const void* p, unsigned bytes) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & (bytes - 1)) == 0; }

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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static const bool no_int = getenv("UOF_UPSAMPLE_GENERIC") != nullptr;
  const int R = no_int ? 0 : int_ratio(h, w, H, W);
  const unsigned vec = R == 4 ? 16u : 8u;
  if (R && aligned_to(out, vec) && aligned_to(out2, vec) && (!out2 || out2_batch_stride % (vec / 4) == 0) && ceil_div(h, kIntRows) <= 65535) {
    const dim3 grid(ceil_div(w, 128), ceil_div(h, kIntRows), planes);
    if (R == 4)
      upsample_fwd_int_kernel<4><<<grid, 128, 0, stream>>>(in, out, out2, out2 ? C2 : 1, out2_batch_stride, h, w, scale);
    else
      upsample_fwd_int_kernel<2><<<grid, 128, 0, stream>>>(in, out, out2, out2 ? C2 : 1, out2_batch_stride, h, w, scale);
  } else {
    upsample_fwd_kernel<<<dim3(ceil_div(ceil_div(W, 4), 128), H, planes), 128, 0, stream>>>(
        in, out, out2, out2 ? C2 : 1, out2_batch_stride, h, w, H, W, (float)h / (float)H, (float)w / (float)W, scale);
  }
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  static const bool no_int = getenv("UOF_UPSAMPLE_GENERIC") != nullptr;
  const int R = no_int ? 0 : int_ratio(h, w, H, W);
  const unsigned vec = R == 4 ? 16u : 8u;
  if (R && aligned_to(gout, vec) && aligned_to(g2, vec) && aligned_to(g3, vec) && (!g2 || g2_batch_stride % (vec / 4) == 0)) {
    const dim3 igrid(ceil_div(w, 128), ceil_div(h, kIntRows), planes);
#define UOF_UPB(RR, MM) upsample_bwd_int_kernel<RR, MM><<<igrid, 128, 0, stream>>>(gout, g2, g2 ? C2 : 1, g2_batch_stride, g3, gin, h, w, scale)
    if (R == 4) { if (g2 || g3) UOF_UPB(4, true); else UOF_UPB(4, false); }
    else { if (g2 || g3) UOF_UPB(2, true); else UOF_UPB(2, false); }
#undef UOF_UPB
    count_launch();
    return check_launch("upsample_bilinear_bwd (integer ratio)");
  }
  const dim3 grid(ceil_div(w, 128), h, planes);
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
