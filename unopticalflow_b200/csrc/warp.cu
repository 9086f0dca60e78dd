// a2/a3: bilinear backward warp with fused mesh grid, normalisation and validity mask.
// Replaces warp_flow, /root/reference/core/networks/structures/net_utils.py:16-54
// (CPU mesh grid + blocking H2D copy + ~10 elementwise launches + 1-2 grid_sample + 2 index_put_).
//
// NCHW kernels ("fat threads"): a thread owns PXT = 4 consecutive pixels of a row and a chunk of CCH
// channels.  Flow is read with one float4 per component, the four bilinear footprints (indices, weights,
// mask) are computed once, and then all 16 gathers of a channel are issued back to back (up to 64 loads
// in flight per thread) -- the first version (one thread per pixel, one dependent round trip per
// channel) was latency-bound at 10-20 % of HBM bandwidth (profiles/r1_*).  Splitting channels over
// the grid keeps the 148 SMs busy on the small pyramid levels (8x26 ... 32x104).
// channels_last kernels: one thread per pixel and float4 channel group; each corner read is a contiguous
// 16 B vector (north star: "coalesced, float4-vectorised NHWC access").
#include "common.cuh"

namespace uof {
namespace {

constexpr float kMaskThreshold = 0.9999f;   // net_utils.py:50
constexpr int CCH = 8;                      // channels per thread (NCHW kernels)

__device__ __forceinline__ float cover_of(const Bilinear& bl) {
  // grid_sample of a ones image accumulates nw, ne, sw, se in this order
  return ((bl.w00 + bl.w01) + bl.w10) + bl.w11;
}

struct Footprint {
  int o00, o01, o10, o11;        // clamped plane offsets of the four corners
  float w00, w01, w10, w11;      // weights, zero for out-of-bounds corners, already multiplied by the mask
};

template <int PXT>
struct ThreadCoord {
  int b, chunk, y, x0;
  bool live;
};

template <int PXT>
__device__ __forceinline__ ThreadCoord<PXT> locate(long long t, int nchunk, int H, int WQ, int B) {
  ThreadCoord<PXT> tc;
  const int xq = (int)(t % WQ);
  long long r = t / WQ;
  tc.y = (int)(r % H);
  r /= H;
  tc.chunk = (int)(r % nchunk);
  tc.b = (int)(r / nchunk);
  tc.x0 = xq * PXT;
  tc.live = tc.b < B;
  return tc;
}

template <int PXT>
__device__ __forceinline__ void load_flow(const float* __restrict__ fb, size_t plane, int W, int x0, float* fx, float* fy) {
  if (PXT == 4) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(fb)), b = __ldg(reinterpret_cast<const float4*>(fb + plane));
    fx[0] = a.x; fx[1] = a.y; fx[2] = a.z; fx[3] = a.w;
    fy[0] = b.x; fy[1] = b.y; fy[2] = b.z; fy[3] = b.w;
  } else {
#pragma unroll
    for (int p = 0; p < PXT; ++p) {
      fx[p] = __ldg(fb + p);
      fy[p] = __ldg(fb + plane + p);
    }
  }
}

// ---------------------------------------------------------------------------------- NCHW fwd
template <int PXT>
__global__ void __launch_bounds__(128)
warp_fwd_nchw_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out, int B, int C,
                     int H, int W, int nchunk, int use_mask, int align_corners) {
  const int WQ = W / PXT;
  const ThreadCoord<PXT> tc = locate<PXT>((long long)blockIdx.x * blockDim.x + threadIdx.x, nchunk, H, WQ, B);
  if (!tc.live) return;
  const size_t plane = (size_t)H * W;
  const size_t pix = (size_t)tc.y * W + tc.x0;
  float fx[PXT], fy[PXT];
  load_flow<PXT>(flow + (size_t)tc.b * 2 * plane + pix, plane, W, tc.x0, fx, fy);
  Footprint fp[PXT];
#pragma unroll
  for (int p = 0; p < PXT; ++p) {
    const Bilinear bl = make_bilinear(sample_coord((float)(tc.x0 + p), fx[p], W, align_corners),
                                      sample_coord((float)tc.y, fy[p], H, align_corners), H, W);
    const float m = (use_mask && cover_of(bl) < kMaskThreshold) ? 0.0f : 1.0f;
    const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
    const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
    fp[p].o00 = ya * W + xa; fp[p].o01 = ya * W + xb; fp[p].o10 = yb * W + xa; fp[p].o11 = yb * W + xb;
    // (v*w)*m == v*(w*m) exactly for m in {0,1}
    fp[p].w00 = bl.w00 * m; fp[p].w01 = bl.w01 * m; fp[p].w10 = bl.w10 * m; fp[p].w11 = bl.w11 * m;
  }
  const int c0 = tc.chunk * CCH, c1 = min(C, c0 + CCH);
  const float* xp = x + ((size_t)tc.b * C + c0) * plane;
  float* op = out + ((size_t)tc.b * C + c0) * plane + pix;
#pragma unroll 2
  for (int c = c0; c < c1; ++c, xp += plane, op += plane) {
    float v[PXT][4];
#pragma unroll
    for (int p = 0; p < PXT; ++p) {
      v[p][0] = __ldg(xp + fp[p].o00);
      v[p][1] = __ldg(xp + fp[p].o01);
      v[p][2] = __ldg(xp + fp[p].o10);
      v[p][3] = __ldg(xp + fp[p].o11);
    }
    float r[PXT];
#pragma unroll
    for (int p = 0; p < PXT; ++p)
      r[p] = fmaf(v[p][3], fp[p].w11, fmaf(v[p][2], fp[p].w10, fmaf(v[p][1], fp[p].w01, v[p][0] * fp[p].w00)));
    if (PXT == 4) {
      *reinterpret_cast<float4*>(op) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int p = 0; p < PXT; ++p) op[p] = r[p];
    }
  }
}

// ---------------------------------------------------------------------------------- NCHW bwd
// ATOMIC_GFLOW: several channel chunks contribute to the same gflow element (gflow zero-filled by the host).
template <int PXT, bool NEED_GX, bool ATOMIC_GFLOW>
__global__ void __launch_bounds__(128)
warp_bwd_nchw_kernel(const float* __restrict__ gout, const float* __restrict__ x, const float* __restrict__ flow,
                     float* __restrict__ gx, float* __restrict__ gflow, int B, int C, int H, int W, int nchunk,
                     int use_mask, int align_corners, float sx, float sy) {
  const int WQ = W / PXT;
  const ThreadCoord<PXT> tc = locate<PXT>((long long)blockIdx.x * blockDim.x + threadIdx.x, nchunk, H, WQ, B);
  if (!tc.live) return;
  const size_t plane = (size_t)H * W;
  const size_t pix = (size_t)tc.y * W + tc.x0;
  float fx[PXT], fy[PXT];
  load_flow<PXT>(flow + (size_t)tc.b * 2 * plane + pix, plane, W, tc.x0, fx, fy);
  Footprint fp[PXT];
  float ux[PXT], uy[PXT], tx[PXT], ty[PXT], msk[PXT];
  bool in00[PXT], in01[PXT], in10[PXT], in11[PXT];
#pragma unroll
  for (int p = 0; p < PXT; ++p) {
    const float ix = sample_coord((float)(tc.x0 + p), fx[p], W, align_corners);
    const float iy = sample_coord((float)tc.y, fy[p], H, align_corners);
    const Bilinear bl = make_bilinear(ix, iy, H, W);
    msk[p] = (use_mask && cover_of(bl) < kMaskThreshold) ? 0.0f : 1.0f;
    const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
    const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
    fp[p].o00 = ya * W + xa; fp[p].o01 = ya * W + xb; fp[p].o10 = yb * W + xa; fp[p].o11 = yb * W + xb;
    fp[p].w00 = bl.w00; fp[p].w01 = bl.w01; fp[p].w10 = bl.w10; fp[p].w11 = bl.w11;
    in00[p] = bl.in00; in01[p] = bl.in01; in10[p] = bl.in10; in11[p] = bl.in11;
    tx[p] = bl.tx; ty[p] = bl.ty;
    ux[p] = (floorf(ix) + 1.0f) - ix;
    uy[p] = (floorf(iy) + 1.0f) - iy;
  }
  const int c0 = tc.chunk * CCH, c1 = min(C, c0 + CCH);
  const float* xp = x + ((size_t)tc.b * C + c0) * plane;
  const float* gp = gout + ((size_t)tc.b * C + c0) * plane + pix;
  float* gxp = NEED_GX ? gx + ((size_t)tc.b * C + c0) * plane : nullptr;
  float gix[PXT], giy[PXT];
#pragma unroll
  for (int p = 0; p < PXT; ++p) gix[p] = giy[p] = 0.0f;
#pragma unroll 2
  for (int c = c0; c < c1; ++c, xp += plane, gp += plane) {
    float g[PXT];
    if (PXT == 4) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp));
      g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
    } else {
#pragma unroll
      for (int p = 0; p < PXT; ++p) g[p] = __ldg(gp + p);
    }
    float v[PXT][4];
#pragma unroll
    for (int p = 0; p < PXT; ++p) {      // clamped addresses are always valid; zero the out-of-bounds corners after
      v[p][0] = __ldg(xp + fp[p].o00);
      v[p][1] = __ldg(xp + fp[p].o01);
      v[p][2] = __ldg(xp + fp[p].o10);
      v[p][3] = __ldg(xp + fp[p].o11);
    }
#pragma unroll
    for (int p = 0; p < PXT; ++p) {
      const float gm = g[p] * msk[p];
      const float v00 = in00[p] ? v[p][0] : 0.0f, v01 = in01[p] ? v[p][1] : 0.0f;
      const float v10 = in10[p] ? v[p][2] : 0.0f, v11 = in11[p] ? v[p][3] : 0.0f;
      // d out / d ix and d out / d iy of the bilinear interpolant
      gix[p] = fmaf(gm, (v01 - v00) * uy[p] + (v11 - v10) * ty[p], gix[p]);
      giy[p] = fmaf(gm, (v10 - v00) * ux[p] + (v11 - v01) * tx[p], giy[p]);
      if (NEED_GX) {
        float* q = gxp + (size_t)(c - c0) * plane;
        if (in00[p]) atomicAdd(q + fp[p].o00, gm * fp[p].w00);
        if (in01[p]) atomicAdd(q + fp[p].o01, gm * fp[p].w01);
        if (in10[p]) atomicAdd(q + fp[p].o10, gm * fp[p].w10);
        if (in11[p]) atomicAdd(q + fp[p].o11, gm * fp[p].w11);
      }
    }
  }
  float* gfb = gflow + (size_t)tc.b * 2 * plane + pix;
  if (ATOMIC_GFLOW) {
#pragma unroll
    for (int p = 0; p < PXT; ++p) {
      atomicAdd(gfb + p, gix[p] * sx);
      atomicAdd(gfb + plane + p, giy[p] * sy);
    }
  } else if (PXT == 4) {
    *reinterpret_cast<float4*>(gfb) = make_float4(gix[0] * sx, gix[1] * sx, gix[2] * sx, gix[3] * sx);
    *reinterpret_cast<float4*>(gfb + plane) = make_float4(giy[0] * sy, giy[1] * sy, giy[2] * sy, giy[3] * sy);
  } else {
#pragma unroll
    for (int p = 0; p < PXT; ++p) {
      gfb[p] = gix[p] * sx;
      gfb[plane + p] = giy[p] * sy;
    }
  }
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
// gx uses 16-byte vector atomics (red.global.add.v4.f32, sm_90+); the per-group partial flow gradients
// are accumulated with atomicAdd into a zero-filled gflow.
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
  UOF_REQUIRE((long long)H * W < (1ll << 30), "%s: image too large", who);
  UOF_REQUIRE((long long)B * C * H * W < (1ll << 40), "%s: tensor too large", who);
  if (channels_last) UOF_REQUIRE(C % 4 == 0, "%s: channels_last needs C %% 4 == 0 (C=%d)", who, C);
  return UOF_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_warp_fwd(const float* x, const float* flow, float* out, int B, int C, int H, int W, int use_mask,
                            int align_corners, int channels_last, uof_stream_t stream_) {
  if (int rc = check_args("warp_fwd", x, flow, out, B, C, H, W, channels_last)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!channels_last) {
    const int nchunk = ceil_div(C, CCH);
    const bool v4 = (W % 4 == 0) && aligned16(flow) && aligned16(out);
    const long long threads = (long long)B * nchunk * H * (v4 ? W / 4 : W);
    const unsigned blocks = (unsigned)ceil_div_ll(threads, 128);
    if (v4)
      warp_fwd_nchw_kernel<4><<<blocks, 128, 0, stream>>>(x, flow, out, B, C, H, W, nchunk, use_mask, align_corners);
    else
      warp_fwd_nchw_kernel<1><<<blocks, 128, 0, stream>>>(x, flow, out, B, C, H, W, nchunk, use_mask, align_corners);
  } else {
    const long long npix = (long long)B * H * W, total = npix * (C / 4);
    warp_fwd_nhwc_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(x), flow, reinterpret_cast<float4*>(out), C / 4, H, W, npix, use_mask,
        align_corners);
  }
  count_launch();
  return check_launch("warp_fwd");
}

template <int PXT>
static void launch_bwd_nchw(const float* gout, const float* x, const float* flow, float* gx, float* gflow, int B, int C,
                            int H, int W, int nchunk, int use_mask, int align_corners, float sx, float sy,
                            cudaStream_t stream) {
  const long long threads = (long long)B * nchunk * H * (W / PXT);
  const unsigned blocks = (unsigned)ceil_div_ll(threads, 128);
#define UOF_LAUNCH(GX, AT)                                                                                              \
  warp_bwd_nchw_kernel<PXT, GX, AT><<<blocks, 128, 0, stream>>>(gout, x, flow, gx, gflow, B, C, H, W, nchunk, use_mask, \
                                                                align_corners, sx, sy)
  if (gx) {
    if (nchunk > 1) UOF_LAUNCH(true, true); else UOF_LAUNCH(true, false);
  } else {
    if (nchunk > 1) UOF_LAUNCH(false, true); else UOF_LAUNCH(false, false);
  }
#undef UOF_LAUNCH
}

extern "C" int uof_warp_bwd(const float* gout, const float* x, const float* flow, float* gx, float* gflow, int B, int C,
                            int H, int W, int use_mask, int align_corners, int channels_last, uof_stream_t stream_) {
  if (int rc = check_args("warp_bwd", gout, x, flow, B, C, H, W, channels_last)) return rc;
  UOF_REQUIRE(gflow, "warp_bwd: gflow is null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const float sx = coord_scale(W, align_corners), sy = coord_scale(H, align_corners);
  if (gx) UOF_CUDA(cudaMemsetAsync(gx, 0, (size_t)B * C * H * W * sizeof(float), stream));
  if (!channels_last) {
    const int nchunk = ceil_div(C, CCH);
    if (nchunk > 1) UOF_CUDA(cudaMemsetAsync(gflow, 0, (size_t)B * 2 * H * W * sizeof(float), stream));
    const bool v4 = (W % 4 == 0) && aligned16(flow) && aligned16(gout) && aligned16(gflow);
    if (v4)
      launch_bwd_nchw<4>(gout, x, flow, gx, gflow, B, C, H, W, nchunk, use_mask, align_corners, sx, sy, stream);
    else
      launch_bwd_nchw<1>(gout, x, flow, gx, gflow, B, C, H, W, nchunk, use_mask, align_corners, sx, sy, stream);
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
