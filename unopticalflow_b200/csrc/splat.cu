// a12 + a13: forward splat ("transformerFwd" range map) and forward-backward consistency mask.
// These are named by the north star but DO NOT EXIST in the reference: there is
// no reference file:line to cite; the algorithm is the published bilinear forward splat of
// Wang et al., "Occlusion Aware Unsupervised Learning of Optical Flow" restated in oracle/ops.py.
//
// Splat: one thread per source pixel (NHWC), four bilinear corner contributions.  Integer contract:
// tx = x + fx, ty = y + fy in plain fp32 adds (no FMA contraction), floorf, flat index b*H*W+yc*W+xc.
// Atomics are warp-aggregated: along a row the (x1, y) corners of lane i coincide with the (x0, y)
// corners of lane i+1 whenever the flow is locally smooth, so a lane first takes over its left
// neighbour's right-hand contributions (shuffle + index compare) and the neighbour skips its
// atomics -- this halves the RED traffic on smooth flows.  C % 4 == 0 uses 16-byte vector REDs.
#include <type_traits>

#include "common.cuh"

namespace uof {
namespace {

struct SplatGeom {
  int x0, y0;
  float wa, wb, wc, wd;          // (x0,y0) (x0,y1) (x1,y0) (x1,y1), App. D
  long long ia, ib, ic, id;      // flat targets or -1
};

__device__ __forceinline__ SplatGeom splat_geom(float fx, float fy, int b, int y, int x, int H, int W) {
  SplatGeom g;
  const float tx = __fadd_rn((float)x, fx), ty = __fadd_rn((float)y, fy);
  float x0f = floorf(tx), y0f = floorf(ty);
  const float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
  g.wa = (x1f - tx) * (y1f - ty);
  g.wb = (x1f - tx) * (ty - y0f);
  g.wc = (tx - x0f) * (y1f - ty);
  g.wd = (tx - x0f) * (ty - y0f);
  x0f = fminf(fmaxf(x0f, -2.0f), (float)W + 1.0f);   // NaN/inf safe int conversion
  y0f = fminf(fmaxf(y0f, -2.0f), (float)H + 1.0f);
  g.x0 = (int)x0f;
  g.y0 = (int)y0f;
  const bool xi0 = g.x0 >= 0 && g.x0 < W, xi1 = g.x0 + 1 >= 0 && g.x0 + 1 < W;
  const bool yi0 = g.y0 >= 0 && g.y0 < H, yi1 = g.y0 + 1 >= 0 && g.y0 + 1 < H;
  const long long base = (long long)b * H * W;
  g.ia = (xi0 && yi0) ? base + (long long)g.y0 * W + g.x0 : -1;
  g.ib = (xi0 && yi1) ? base + (long long)(g.y0 + 1) * W + g.x0 : -1;
  g.ic = (xi1 && yi0) ? base + (long long)g.y0 * W + g.x0 + 1 : -1;
  g.id = (xi1 && yi1) ? base + (long long)(g.y0 + 1) * W + g.x0 + 1 : -1;
  return g;
}

// ---------------------------------------------------------------------------- range map (C == 1)
// u == nullptr means "splat ones".  grid = (x chunks, H, B): no index decoding, and everything inside the image is
// 32-bit (offsets y*W + x relative to the image base; the first version carried 64-bit flat targets through the
// shuffles and decoded a 64-bit linear thread index with two div/mod pairs).
__global__ void __launch_bounds__(128)
splat1_fwd_kernel(const float* __restrict__ u, const float2* __restrict__ flow, float* __restrict__ out, int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const size_t base = (size_t)blockIdx.z * H * W;
  const bool live = x < W;
  const int lane = threadIdx.x & 31;
  int ia = -1, ib = -1, ic = -1, id = -1;                    // offsets inside the image, -1 = out of bounds
  float ca = 0.0f, cb = 0.0f, cc = 0.0f, cd = 0.0f;
  if (live) {
    const size_t t = base + (size_t)y * W + x;
    const float2 f = __ldg(flow + t);
    const SplatGeom g = splat_geom(f.x, f.y, 0, y, x, H, W);   // b = 0: targets relative to the image
    const float v = u ? __ldg(u + t) : 1.0f;
    ia = (int)g.ia; ib = (int)g.ib; ic = (int)g.ic; id = (int)g.id;
    ca = v * g.wa; cb = v * g.wb; cc = v * g.wc; cd = v * g.wd;
  }
  // warp aggregation: hand my right-hand column (x1) to lane+1 if it is that lane's left-hand column
  const int nic = __shfl_up_sync(kFullMask, ic, 1), nid = __shfl_up_sync(kFullMask, id, 1);
  const float ncc = __shfl_up_sync(kFullMask, cc, 1), ncd = __shfl_up_sync(kFullMask, cd, 1);
  const bool take = lane > 0 && nic == ia && nid == ib && (ia >= 0 || ib >= 0);
  const bool taken = __shfl_down_sync(kFullMask, (int)take, 1) && lane < 31;
  if (take) {
    ca += ncc;
    cb += ncd;
  }
  float* o = out + base;
  if (ia >= 0) atomicAdd(o + ia, ca);
  if (ib >= 0) atomicAdd(o + ib, cb);
  if (!taken) {
    if (ic >= 0) atomicAdd(o + ic, cc);
    if (id >= 0) atomicAdd(o + id, cd);
  }
}

// ------------------------------------------------------------------------------- general C fwd
// NHWC values with C channels (C % 4 == 0: float4 groups and 16-byte vector REDs).  grid = (chunks of W * CG threads, H, B),
// thread = (pixel x, channel group cg) with cg fastest, so a warp touches contiguous channels of neighbouring pixels;
// targets are 32-bit offsets inside the image.  Same warp aggregation as the C == 1 kernel: the thread CG lanes to the
// left owns the same channel group of pixel x - 1, and when its right-hand column is this thread's left-hand column the
// two contributions are added in registers and it skips its two REDs (needs CG < 32).  The first version decoded a
// 64-bit linear index with three div/mod pairs and issued all four REDs per thread.
template <bool VEC4>
__global__ void __launch_bounds__(256)
splat_fwd_kernel(const float* __restrict__ u, const float2* __restrict__ flow, float* __restrict__ out, int H, int W, int C,
                 int CG) {
  typedef typename std::conditional<VEC4, float4, float>::type V;
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, lane = threadIdx.x & 31;
  const size_t base = (size_t)blockIdx.z * H * W;          // pixels before this image
  const unsigned x = t / (unsigned)CG, cg = t - x * (unsigned)CG;
  const bool live = x < (unsigned)W;
  int ia = -1, ib = -1, ic = -1, id = -1;
  float v[4] = {0.f, 0.f, 0.f, 0.f}, wa = 0.f, wb = 0.f, wc = 0.f, wd = 0.f;
  if (live) {
    const size_t pix = base + (size_t)y * W + x;
    const float2 f = __ldg(flow + pix);
    const SplatGeom g = splat_geom(f.x, f.y, 0, y, (int)x, H, W);      // b = 0: targets relative to the image
    ia = (int)g.ia; ib = (int)g.ib; ic = (int)g.ic; id = (int)g.id;
    wa = g.wa; wb = g.wb; wc = g.wc; wd = g.wd;
    if (VEC4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(u) + pix * CG + cg);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      v[0] = __ldg(u + pix * CG + cg);
    }
  }
  constexpr int NV = VEC4 ? 4 : 1;
  float ca[NV], cb[NV], cc[NV], cd[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    ca[k] = v[k] * wa; cb[k] = v[k] * wb; cc[k] = v[k] * wc; cd[k] = v[k] * wd;
  }
  bool take = false, taken = false;
  if (CG < 32) {             // uniform
    const int nic = __shfl_up_sync(kFullMask, ic, CG), nid = __shfl_up_sync(kFullMask, id, CG);
    take = lane >= CG && nic == ia && nid == ib && (ia >= 0 || ib >= 0);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float ncc = __shfl_up_sync(kFullMask, cc[k], CG), ncd = __shfl_up_sync(kFullMask, cd[k], CG);
      if (take) {
        ca[k] += ncc;
        cb[k] += ncd;
      }
    }
    taken = __shfl_down_sync(kFullMask, (int)take, CG) && lane + CG < 32;
  }
  V* o = reinterpret_cast<V*>(out) + base * CG + cg;
  auto put = [&](int idx, const float* c) {
    if (idx < 0) return;
    if (VEC4)
      atomicAdd(reinterpret_cast<float4*>(o + (size_t)idx * CG), make_float4(c[0], c[VEC4 ? 1 : 0], c[VEC4 ? 2 : 0], c[VEC4 ? 3 : 0]));
    else
      atomicAdd(reinterpret_cast<float*>(o + (size_t)idx * CG), c[0]);
  };
  put(ia, ca);
  put(ib, cb);
  if (!taken) {
    put(ic, cc);
    put(id, cd);
  }
}

// ------------------------------------------------------------------------------------ backward
// gu[b,y,x,c] = sum_k w_k * gout[target_k, c];   gflow = sum_c u_c * d(w_k)/d(tx,ty) * gout[target_k,c]
// One thread per source pixel, loops over channels (gather; no atomics).
__global__ void __launch_bounds__(256)
splat_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ u, const float2* __restrict__ flow,
                 float* __restrict__ gu, float2* __restrict__ gflow, int B, int H, int W, int C) {
  const long long n = (long long)B * H * W;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int x = (int)(t % W), y = (int)((t / W) % H), b = (int)(t / ((long long)W * H));
  const float2 f = __ldg(flow + t);
  const SplatGeom g = splat_geom(f.x, f.y, b, y, x, H, W);
  const float tx = __fadd_rn((float)x, f.x), ty = __fadd_rn((float)y, f.y);
  const float x0f = floorf(tx), y0f = floorf(ty);
  const float ax = (x0f + 1.0f) - tx, bx = tx - x0f, ay = (y0f + 1.0f) - ty, by = ty - y0f;
  float gfx = 0.0f, gfy = 0.0f;
  for (int c = 0; c < C; ++c) {
    const float ga = g.ia >= 0 ? __ldg(gout + g.ia * C + c) : 0.0f;
    const float gb = g.ib >= 0 ? __ldg(gout + g.ib * C + c) : 0.0f;
    const float gc = g.ic >= 0 ? __ldg(gout + g.ic * C + c) : 0.0f;
    const float gd = g.id >= 0 ? __ldg(gout + g.id * C + c) : 0.0f;
    if (gu) gu[t * C + c] = ga * g.wa + gb * g.wb + gc * g.wc + gd * g.wd;
    const float uv = u ? __ldg(u + t * C + c) : 1.0f;
    // wa = ax*ay, wb = ax*by, wc = bx*ay, wd = bx*by;  d ax/d tx = -1, d bx/d tx = +1, ...
    gfx = fmaf(uv, (gc - ga) * ay + (gd - gb) * by, gfx);
    gfy = fmaf(uv, (gb - ga) * ax + (gd - gc) * bx, gfy);
  }
  if (gflow) gflow[t] = make_float2(gfx, gfy);
}

__global__ void __launch_bounds__(256)
splat_targets_kernel(const float2* __restrict__ flow, long long* __restrict__ idx, int B, int H, int W) {
  const long long n = (long long)B * H * W;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int x = (int)(t % W), y = (int)((t / W) % H), b = (int)(t / ((long long)W * H));
  const float2 f = __ldg(flow + t);
  const SplatGeom g = splat_geom(f.x, f.y, b, y, x, H, W);
  long long* o = idx + t * 4;
  o[0] = g.ia; o[1] = g.ib; o[2] = g.ic; o[3] = g.id;
}

__global__ void __launch_bounds__(256) clamp01_kernel(float* __restrict__ p, long long n) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) p[t] = fminf(fmaxf(p[t], 0.0f), 1.0f);
}

// --------------------------------------------------------------------- fb-consistency mask (a13)
__global__ void __launch_bounds__(256)
fb_mask_kernel(const float* __restrict__ ffwd, const float* __restrict__ frev, float* __restrict__ mask, int H, int W,
               float alpha, float beta, int align_corners) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (px >= W) return;
  const size_t plane = (size_t)H * W, pix = (size_t)y * W + px;
  const float* fb = ffwd + (size_t)b * 2 * plane + pix;
  const float fx = __ldg(fb), fy = __ldg(fb + plane);
  const Bilinear bl = make_bilinear(sample_coord((float)px, fx, W, align_corners),
                                    sample_coord((float)y, fy, H, align_corners), H, W);
  const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
  const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
  const float* rp = frev + (size_t)b * 2 * plane;
  float back[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float* p = rp + k * plane;
    float v = __ldg(p + ya * W + xa) * bl.w00;
    v = fmaf(__ldg(p + ya * W + xb), bl.w01, v);
    v = fmaf(__ldg(p + yb * W + xa), bl.w10, v);
    v = fmaf(__ldg(p + yb * W + xb), bl.w11, v);
    back[k] = v;
  }
  const float dx = back[0] + fx, dy = back[1] + fy;
  const float diff = sqrtf(dx * dx + dy * dy);
  const float bound = fmaxf(alpha, beta * sqrtf(fx * fx + fy * fy));
  mask[(size_t)b * plane + pix] = diff < bound ? 1.0f : 0.0f;
}

int check_splat(const char* who, const void* flow, int B, int H, int W, int C) {
  UOF_REQUIRE(flow, "%s: null flow", who);
  UOF_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "%s: bad shape B=%d H=%d W=%d C=%d", who, B, H, W, C);
  UOF_REQUIRE((long long)B * H * W * C < (1ll << 40), "%s: problem too large", who);
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_splat_fwd(const float* u, const float* flow, float* out, int B, int H, int W, int C, uof_stream_t stream_) {
  if (int rc = check_splat("splat_fwd", flow, B, H, W, C)) return rc;
  UOF_REQUIRE(out, "splat_fwd: null output");
  UOF_REQUIRE(u || C == 1, "splat_fwd: u == NULL (ones) needs C == 1");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long npix = (long long)B * H * W;
  UOF_REQUIRE(H <= 65535 && B <= 65535 && (long long)H * W < (1ll << 31) && (long long)W * C < (1ll << 31),
              "splat_fwd: image too large for one launch");
  UOF_CUDA(cudaMemsetAsync(out, 0, (size_t)npix * C * sizeof(float), stream));
  const float2* f2 = reinterpret_cast<const float2*>(flow);
  if (C == 1) {
    splat1_fwd_kernel<<<dim3(ceil_div(W, 128), H, B), 128, 0, stream>>>(u, f2, out, H, W);
  } else if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(u) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    splat_fwd_kernel<true><<<dim3((unsigned)ceil_div_ll((long long)W * (C / 4), 256), H, B), 256, 0, stream>>>(u, f2, out, H, W, C, C / 4);
  } else {
    splat_fwd_kernel<false><<<dim3((unsigned)ceil_div_ll((long long)W * C, 256), H, B), 256, 0, stream>>>(u, f2, out, H, W, C, C);
  }
  count_launch(1);
  return check_launch("splat_fwd");
}

extern "C" int uof_splat_bwd(const float* gout, const float* u, const float* flow, float* gu, float* gflow, int B, int H,
                             int W, int C, uof_stream_t stream_) {
  if (int rc = check_splat("splat_bwd", flow, B, H, W, C)) return rc;
  UOF_REQUIRE(gout && (gu || gflow), "splat_bwd: nothing to compute");
  UOF_REQUIRE(u || C == 1, "splat_bwd: u == NULL (ones) needs C == 1");
  const long long npix = (long long)B * H * W;
  splat_bwd_kernel<<<(unsigned)ceil_div_ll(npix, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      gout, u, reinterpret_cast<const float2*>(flow), gu, reinterpret_cast<float2*>(gflow), B, H, W, C);
  count_launch();
  return check_launch("splat_bwd");
}

extern "C" int uof_splat_targets(const float* flow, long long* idx, int B, int H, int W, uof_stream_t stream_) {
  if (int rc = check_splat("splat_targets", flow, B, H, W, 1)) return rc;
  UOF_REQUIRE(idx, "splat_targets: null output");
  const long long npix = (long long)B * H * W;
  splat_targets_kernel<<<(unsigned)ceil_div_ll(npix, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const float2*>(flow), idx, B, H, W);
  count_launch();
  return check_launch("splat_targets");
}

extern "C" int uof_clamp01(float* range, long long n, uof_stream_t stream_) {
  UOF_REQUIRE(range && n > 0, "clamp01: bad arguments");
  clamp01_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(range, n);
  count_launch();
  return check_launch("clamp01");
}

extern "C" int uof_fb_consistency_mask(const float* flow_fwd, const float* flow_rev, float* mask, int B, int H, int W,
                                       float alpha, float beta, int align_corners, uof_stream_t stream_) {
  UOF_REQUIRE(flow_fwd && flow_rev && mask, "fb_consistency_mask: null pointer");
  UOF_REQUIRE(B > 0 && H > 0 && W > 0 && B <= 65535 && H <= 65535, "fb_consistency_mask: bad shape B=%d H=%d W=%d", B, H, W);
  const int bx = W >= 256 ? 256 : (W >= 128 ? 128 : (W >= 64 ? 64 : 32));
  dim3 grid(ceil_div(W, bx), H, B);
  fb_mask_kernel<<<grid, bx, 0, static_cast<cudaStream_t>(stream_)>>>(flow_fwd, flow_rev, mask, H, W, alpha, beta,
                                                                        align_corners);
  count_launch();
  return check_launch("fb_consistency_mask");
}
