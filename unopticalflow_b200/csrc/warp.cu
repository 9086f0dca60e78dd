// a2/a3: bilinear backward warp with fused mesh grid, normalisation and validity mask.
// Replaces warp_flow, /root/reference/core/networks/structures/net_utils.py:16-54
// (CPU mesh grid + blocking H2D copy + ~10 elementwise launches + 1-2 grid_sample + 2 index_put_).
//
// NCHW kernels: one thread per output pixel; sampling coordinates, corner indices and weights are
// computed once and reused for every channel (the per-plane gathers of a warp are neighbouring
// addresses, so they coalesce for smooth flows).  channels_last kernels: one thread per pixel and
// float4 channel group; each corner read is a contiguous 16 B vector (north star: "coalesced,
// float4-vectorised NHWC access").
#include "common.cuh"

namespace uof {
namespace {

constexpr float kMaskThreshold = 0.9999f;   // net_utils.py:50

__device__ __forceinline__ float cover_of(const Bilinear& bl) {
  // grid_sample of a ones image accumulates nw, ne, sw, se in this order
  return ((bl.w00 + bl.w01) + bl.w10) + bl.w11;
}

// ---------------------------------------------------------------------------------- NCHW fwd
__global__ void __launch_bounds__(256)
warp_fwd_nchw_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                     int C, int H, int W, int use_mask, int align_corners) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (px >= W) return;
  const size_t plane = (size_t)H * W;
  const float* fb = flow + (size_t)b * 2 * plane + (size_t)y * W + px;
  const float ix = sample_coord((float)px, __ldg(fb), W, align_corners);
  const float iy = sample_coord((float)y, __ldg(fb + plane), H, align_corners);
  const Bilinear bl = make_bilinear(ix, iy, H, W);
  float m = 1.0f;
  if (use_mask) m = cover_of(bl) < kMaskThreshold ? 0.0f : 1.0f;
  // clamped addresses keep every gather in bounds; out-of-bounds corners have zero weight
  const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
  const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
  const int o00 = ya * W + xa, o01 = ya * W + xb, o10 = yb * W + xa, o11 = yb * W + xb;
  const float* xp = x + (size_t)b * C * plane;
  float* op = out + (size_t)b * C * plane + (size_t)y * W + px;
#pragma unroll 4
  for (int c = 0; c < C; ++c) {
    const float* p = xp + (size_t)c * plane;
    float v = __ldg(p + o00) * bl.w00;
    v = fmaf(__ldg(p + o01), bl.w01, v);
    v = fmaf(__ldg(p + o10), bl.w10, v);
    v = fmaf(__ldg(p + o11), bl.w11, v);
    op[(size_t)c * plane] = v * m;
  }
}

// ---------------------------------------------------------------------------------- NCHW bwd
template <bool NEED_GX>
__global__ void __launch_bounds__(256)
warp_bwd_nchw_kernel(const float* __restrict__ gout, const float* __restrict__ x, const float* __restrict__ flow,
                     float* __restrict__ gx, float* __restrict__ gflow, int C, int H, int W, int use_mask,
                     int align_corners, float sx, float sy) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (px >= W) return;
  const size_t plane = (size_t)H * W;
  const size_t pix = (size_t)y * W + px;
  const float* fb = flow + (size_t)b * 2 * plane + pix;
  const float ix = sample_coord((float)px, __ldg(fb), W, align_corners);
  const float iy = sample_coord((float)y, __ldg(fb + plane), H, align_corners);
  const Bilinear bl = make_bilinear(ix, iy, H, W);
  float m = 1.0f;
  if (use_mask) m = cover_of(bl) < kMaskThreshold ? 0.0f : 1.0f;
  float* gfb = gflow + (size_t)b * 2 * plane + pix;
  if (m == 0.0f && !NEED_GX) {   // masked pixels carry no gradient at all
    gfb[0] = 0.0f;
    gfb[plane] = 0.0f;
    return;
  }
  const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
  const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
  const int o00 = ya * W + xa, o01 = ya * W + xb, o10 = yb * W + xa, o11 = yb * W + xb;
  const float ux = (floorf(ix) + 1.0f) - ix, uy = (floorf(iy) + 1.0f) - iy;
  const float* xp = x + (size_t)b * C * plane;
  const float* gp = gout + (size_t)b * C * plane + pix;
  float* gxp = NEED_GX ? gx + (size_t)b * C * plane : nullptr;
  float gix = 0.0f, giy = 0.0f;
#pragma unroll 2
  for (int c = 0; c < C; ++c) {
    const float g = __ldg(gp + (size_t)c * plane) * m;
    const float* p = xp + (size_t)c * plane;
    const float v00 = bl.in00 ? __ldg(p + o00) : 0.0f;
    const float v01 = bl.in01 ? __ldg(p + o01) : 0.0f;
    const float v10 = bl.in10 ? __ldg(p + o10) : 0.0f;
    const float v11 = bl.in11 ? __ldg(p + o11) : 0.0f;
    // d out / d ix and d out / d iy of the bilinear interpolant
    gix = fmaf(g, (v01 - v00) * uy + (v11 - v10) * bl.ty, gix);
    giy = fmaf(g, (v10 - v00) * ux + (v11 - v01) * bl.tx, giy);
    if (NEED_GX) {
      float* q = gxp + (size_t)c * plane;
      if (bl.in00) atomicAdd(q + o00, g * bl.w00);
      if (bl.in01) atomicAdd(q + o01, g * bl.w01);
      if (bl.in10) atomicAdd(q + o10, g * bl.w10);
      if (bl.in11) atomicAdd(q + o11, g * bl.w11);
    }
  }
  gfb[0] = gix * sx;
  gfb[plane] = giy * sy;
}

// --------------------------------------------------------------------------- channels_last fwd
// x,out stored NHWC; thread = (pixel, 4-channel group); C4 = C/4 groups.
__global__ void __launch_bounds__(256)
warp_fwd_nhwc_kernel(const float4* __restrict__ x, const float* __restrict__ flow, float4* __restrict__ out,
                     int C4, int H, int W, long long npix_total, int use_mask, int align_corners) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix_total * C4) return;
  const int cg = (int)(t % C4);
  const long long pixg = t / C4;           // b*H*W + y*W + x
  const size_t plane = (size_t)H * W;
  const int b = (int)(pixg / plane);
  const int rem = (int)(pixg - (long long)b * plane);
  const int y = rem / W, px = rem - y * W;
  const float* fb = flow + (size_t)b * 2 * plane + rem;
  const float ix = sample_coord((float)px, __ldg(fb), W, align_corners);
  const float iy = sample_coord((float)y, __ldg(fb + plane), H, align_corners);
  const Bilinear bl = make_bilinear(ix, iy, H, W);
  float m = 1.0f;
  if (use_mask) m = cover_of(bl) < kMaskThreshold ? 0.0f : 1.0f;
  const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
  const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
  const float4* xb_ = x + (size_t)b * plane * C4 + cg;
  const float4 a = __ldg(xb_ + (size_t)(ya * W + xa) * C4), bq = __ldg(xb_ + (size_t)(ya * W + xb) * C4);
  const float4 c = __ldg(xb_ + (size_t)(yb * W + xa) * C4), d = __ldg(xb_ + (size_t)(yb * W + xb) * C4);
  float4 r;
  r.x = fmaf(d.x, bl.w11, fmaf(c.x, bl.w10, fmaf(bq.x, bl.w01, a.x * bl.w00))) * m;
  r.y = fmaf(d.y, bl.w11, fmaf(c.y, bl.w10, fmaf(bq.y, bl.w01, a.y * bl.w00))) * m;
  r.z = fmaf(d.z, bl.w11, fmaf(c.z, bl.w10, fmaf(bq.z, bl.w01, a.z * bl.w00))) * m;
  r.w = fmaf(d.w, bl.w11, fmaf(c.w, bl.w10, fmaf(bq.w, bl.w01, a.w * bl.w00))) * m;
  out[(size_t)pixg * C4 + cg] = r;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}

// --------------------------------------------------------------------------- channels_last bwd
// One warp-contiguous group of C4 threads (C4 <= 32 lanes handled by a segmented shuffle) shares a
// pixel; gx uses 16-byte vector atomics (red.global.add.v4.f32, sm_90+), gflow is reduced over the
// channel groups with atomicAdd into a zeroed gflow.
template <bool NEED_GX>
__global__ void __launch_bounds__(256)
warp_bwd_nhwc_kernel(const float4* __restrict__ gout, const float4* __restrict__ x, const float* __restrict__ flow,
                     float4* __restrict__ gx, float* __restrict__ gflow, int C4, int H, int W,
                     long long npix_total, int use_mask, int align_corners, float sx, float sy) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix_total * C4) return;
  const int cg = (int)(t % C4);
  const long long pixg = t / C4;
  const size_t plane = (size_t)H * W;
  const int b = (int)(pixg / plane);
  const int rem = (int)(pixg - (long long)b * plane);
  const int y = rem / W, px = rem - y * W;
  const float* fb = flow + (size_t)b * 2 * plane + rem;
  const float ix = sample_coord((float)px, __ldg(fb), W, align_corners);
  const float iy = sample_coord((float)y, __ldg(fb + plane), H, align_corners);
  const Bilinear bl = make_bilinear(ix, iy, H, W);
  float m = 1.0f;
  if (use_mask) m = cover_of(bl) < kMaskThreshold ? 0.0f : 1.0f;
  const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
  const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
  const size_t o00 = (size_t)(ya * W + xa) * C4, o01 = (size_t)(ya * W + xb) * C4;
  const size_t o10 = (size_t)(yb * W + xa) * C4, o11 = (size_t)(yb * W + xb) * C4;
  const float ux = (floorf(ix) + 1.0f) - ix, uy = (floorf(iy) + 1.0f) - iy;
  const float4 g = scale4(__ldg(gout + (size_t)pixg * C4 + cg), m);
  const float4* xb_ = x + (size_t)b * plane * C4 + cg;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 v00 = bl.in00 ? __ldg(xb_ + o00) : z, v01 = bl.in01 ? __ldg(xb_ + o01) : z;
  const float4 v10 = bl.in10 ? __ldg(xb_ + o10) : z, v11 = bl.in11 ? __ldg(xb_ + o11) : z;
  const float gix = (dot4(g, sub4(v01, v00)) * uy + dot4(g, sub4(v11, v10)) * bl.ty) * sx;
  const float giy = (dot4(g, sub4(v10, v00)) * ux + dot4(g, sub4(v11, v01)) * bl.tx) * sy;
  float* gfb = gflow + (size_t)b * 2 * plane + rem;
  atomicAdd(gfb, gix);
  atomicAdd(gfb + plane, giy);
  if (NEED_GX) {
    float4* q = gx + (size_t)b * plane * C4 + cg;
    if (bl.in00) atomicAdd(q + o00, scale4(g, bl.w00));
    if (bl.in01) atomicAdd(q + o01, scale4(g, bl.w01));
    if (bl.in10) atomicAdd(q + o10, scale4(g, bl.w10));
    if (bl.in11) atomicAdd(q + o11, scale4(g, bl.w11));
  }
}

int check_args(const char* who, const void* a, const void* b, const void* c, int B, int C, int H, int W, int channels_last) {
  UOF_REQUIRE(a && b && c, "%s: null pointer", who);
  UOF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "%s: bad shape B=%d C=%d H=%d W=%d", who, B, C, H, W);
  UOF_REQUIRE(B <= 65535 && H <= 65535, "%s: B and H must be <= 65535", who);
  UOF_REQUIRE((long long)H * W < (1ll << 30), "%s: image too large", who);
  if (channels_last) UOF_REQUIRE(C % 4 == 0, "%s: channels_last needs C %% 4 == 0 (C=%d)", who, C);
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_warp_fwd(const float* x, const float* flow, float* out, int B, int C, int H, int W, int use_mask,
                            int align_corners, int channels_last, uof_stream_t stream_) {
  if (int rc = check_args("warp_fwd", x, flow, out, B, C, H, W, channels_last)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!channels_last) {
    const int bx = W >= 256 ? 256 : (W >= 128 ? 128 : (W >= 64 ? 64 : 32));
    dim3 grid(ceil_div(W, bx), H, B);
    warp_fwd_nchw_kernel<<<grid, bx, 0, stream>>>(x, flow, out, C, H, W, use_mask, align_corners);
  } else {
    const long long npix = (long long)B * H * W, total = npix * (C / 4);
    warp_fwd_nhwc_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(x), flow, reinterpret_cast<float4*>(out), C / 4, H, W, npix, use_mask,
        align_corners);
  }
  count_launch();
  return check_launch("warp_fwd");
}

extern "C" int uof_warp_bwd(const float* gout, const float* x, const float* flow, float* gx, float* gflow, int B, int C,
                            int H, int W, int use_mask, int align_corners, int channels_last, uof_stream_t stream_) {
  if (int rc = check_args("warp_bwd", gout, x, flow, B, C, H, W, channels_last)) return rc;
  UOF_REQUIRE(gflow, "warp_bwd: gflow is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const float sx = coord_scale(W, align_corners), sy = coord_scale(H, align_corners);
  const size_t xbytes = (size_t)B * C * H * W * sizeof(float);
  if (gx) {
    UOF_CUDA(cudaMemsetAsync(gx, 0, xbytes, stream));
  }
  if (!channels_last) {
    const int bx = W >= 256 ? 256 : (W >= 128 ? 128 : (W >= 64 ? 64 : 32));
    dim3 grid(ceil_div(W, bx), H, B);
    if (gx)
      warp_bwd_nchw_kernel<true><<<grid, bx, 0, stream>>>(gout, x, flow, gx, gflow, C, H, W, use_mask, align_corners, sx, sy);
    else
      warp_bwd_nchw_kernel<false><<<grid, bx, 0, stream>>>(gout, x, flow, gx, gflow, C, H, W, use_mask, align_corners, sx, sy);
  } else {
    UOF_CUDA(cudaMemsetAsync(gflow, 0, (size_t)B * 2 * H * W * sizeof(float), stream));
    const long long npix = (long long)B * H * W, total = npix * (C / 4);
    const unsigned nb = (unsigned)ceil_div_ll(total, 256);
    if (gx)
      warp_bwd_nhwc_kernel<true><<<nb, 256, 0, stream>>>(reinterpret_cast<const float4*>(gout),
                                                          reinterpret_cast<const float4*>(x), flow,
                                                          reinterpret_cast<float4*>(gx), gflow, C / 4, H, W, npix,
                                                          use_mask, align_corners, sx, sy);
    else
      warp_bwd_nhwc_kernel<false><<<nb, 256, 0, stream>>>(reinterpret_cast<const float4*>(gout),
                                                           reinterpret_cast<const float4*>(x), flow, nullptr, gflow,
                                                           C / 4, H, W, npix, use_mask, align_corners, sx, sy);
  }
  count_launch();
  return check_launch("warp_bwd");
}
