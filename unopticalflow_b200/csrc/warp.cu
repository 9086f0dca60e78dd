// a2/a3: bilinear backward warp with fused mesh grid, normalisation and validity mask.
// Replaces warp_flow, /root/reference/core/networks/structures/net_utils.py:16-54
// (CPU mesh grid + blocking H2D copy + ~10 elementwise launches + 1-2 grid_sample + 2 index_put_).
//
// NCHW kernels: a warp owns a run of 32*PXT consecutive pixels of one image row and a chunk of CCH channels;
// lane l handles pixels l, l+32, ..., l+32*(PXT-1) of the run.  Every load/store/RED instruction of the warp
// therefore touches ~32 neighbouring addresses (4-6 sectors), while each thread still has PXT independent
// footprints and 4*PXT gathers in flight per channel.  History (profiles/): v1 (one pixel per thread, one
// dependent round trip per channel) was latency-bound at 10-20 % of HBM peak; v2 gave each thread 4
// *adjacent* pixels, which fixed the latency but made every warp instruction span 512 B (13-25 sectors per
// request, ncu) -- this version keeps the per-thread parallelism and restores coalescing.  Channels are
// split over the grid so that the small pyramid levels (8x26 ... 32x104) still fill the 148 SMs.
// channels_last kernels: one thread per pixel and float4 channel group; each corner read is a contiguous
// 16 B vector (north star: "coalesced, float4-vectorised NHWC access").
#include <stdlib.h>

#include "common.cuh"

namespace uof {
namespace {

constexpr float kMaskThreshold = 0.9999f;   // net_utils.py:50
constexpr int CCH = 8;                      // channels per thread (NCHW kernels)
constexpr int kWarps = 4;                   // warps per block

__device__ __forceinline__ float cover_of(const Bilinear& bl) {
  // grid_sample of a ones image accumulates nw, ne, sw, se in this order
  return ((bl.w00 + bl.w01) + bl.w10) + bl.w11;
}

struct Footprint {
  int o00, o01, o10, o11;        // clamped plane offsets of the four corners
  float w00, w01, w10, w11;      // weights, zero for out-of-bounds corners
};

struct RunCoord {
  int b, chunk, y, x0;           // x0 = first pixel of the warp's run (lane 0, k = 0)
  bool live;
};

// Warp -> (batch, channel chunk, row, run of 32*PXT pixels) without any division: grid = (x, nchunk, B) and within
// the (b, chunk) plane either
//   linear map (run_magic != 0): warp w = blockIdx.x * kWarps + warp id covers run w % runs of row w / runs, so a
//     block reads 4 consecutive runs of one row; w / runs = umulhi(w, run_magic) with run_magic = 2^32 / runs + 1
//     (exact for w * runs < 2^32, checked by the host), or
//   row map (run_magic == 0): grid.x = runs * ceil(H / kWarps), a block covers the same run of 4 consecutive rows
//     (blockIdx.x = run + runs * row group would need a division again, so the row map keeps runs in `runs_per_row`
//     and decodes with the same multiply).
// The first version decoded a linear 64-bit index with three long-long div/mod pairs: ~150 of the ~580 instructions a
// warp executed for an image warp (ncu source page, round 2).
__device__ __forceinline__ RunCoord locate_run(int H, int run_px, int runs_per_row, unsigned run_magic, int row_map) {
  RunCoord rc;
  const unsigned warp = threadIdx.x >> 5;
  unsigned run;
  if (!row_map) {
    const unsigned w = blockIdx.x * kWarps + warp;
    rc.y = (int)(run_magic ? __umulhi(w, run_magic) : w);                 // run_magic == 0: one run per row
    run = w - (unsigned)rc.y * (unsigned)runs_per_row;
  } else {
    const unsigned yg = run_magic ? __umulhi(blockIdx.x, run_magic) : blockIdx.x;       // row group
    run = blockIdx.x - yg * (unsigned)runs_per_row;
    rc.y = (int)(yg * kWarps + warp);
  }
  rc.chunk = (int)blockIdx.y;
  rc.b = (int)blockIdx.z;
  rc.x0 = (int)run * run_px;
  rc.live = rc.y < H;
  return rc;
}

// grid for the mapping above; `row_map` selects the second form
dim3 run_grid(int B, int nchunk, int H, int runs, bool row_map) {
  return row_map ? dim3(runs * ceil_div(H, kWarps), nchunk, B) : dim3(ceil_div(H * runs, kWarps), nchunk, B);
}
inline unsigned magic_of(int d) { return d <= 1 ? 0u : (unsigned)((1ull << 32) / (unsigned)d + 1ull); }

// ---------------------------------------------------------------------------------- NCHW fwd
template <int PXT, int UNR>
__global__ void __launch_bounds__(kWarps * 32)
warp_fwd_nchw_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out, int B, int C,
                     int H, int W, int nchunk, int runs_per_row, unsigned run_magic, int row_map, int use_mask,
                     int align_corners) {
  const int lane = threadIdx.x & 31;
  const RunCoord rc = locate_run(H, 32 * PXT, runs_per_row, run_magic, row_map);
  if (!rc.live) return;
  const size_t plane = (size_t)H * W;
  const float* fb = flow + (size_t)rc.b * 2 * plane + (size_t)rc.y * W;
  float fx[PXT], fy[PXT];
  bool ok[PXT];
#pragma unroll
  for (int k = 0; k < PXT; ++k) {
    const int px = rc.x0 + 32 * k + lane;
    ok[k] = px < W;
    fx[k] = ok[k] ? __ldg(fb + px) : 0.0f;
    fy[k] = ok[k] ? __ldg(fb + plane + px) : 0.0f;
  }
  Footprint fp[PXT];
#pragma unroll
  for (int k = 0; k < PXT; ++k) {
    const int px = rc.x0 + 32 * k + lane;
    const Bilinear bl = make_bilinear(sample_coord((float)px, fx[k], W, align_corners),
                                      sample_coord((float)rc.y, fy[k], H, align_corners), H, W);
    const float m = (use_mask && cover_of(bl) < kMaskThreshold) ? 0.0f : 1.0f;
    const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
    const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
    fp[k].o00 = ya * W + xa; fp[k].o01 = ya * W + xb; fp[k].o10 = yb * W + xa; fp[k].o11 = yb * W + xb;
    // (v*w)*m == v*(w*m) exactly for m in {0,1}
    fp[k].w00 = bl.w00 * m; fp[k].w01 = bl.w01 * m; fp[k].w10 = bl.w10 * m; fp[k].w11 = bl.w11 * m;
  }
  const int cch = (C + nchunk - 1) / nchunk;       // channels per thread
  const int c0 = rc.chunk * cch, c1 = min(C, c0 + cch);
  const float* xp = x + ((size_t)rc.b * C + c0) * plane;
  float* op = out + ((size_t)rc.b * C + c0) * plane + (size_t)rc.y * W + rc.x0 + lane;
  // UNR channels per batch: all 4 * PXT * UNR gathers of a batch are issued before the first use (with a plain
  // `#pragma unroll` over channels ptxas kept the per-channel load -> blend -> store order, 4 * PXT loads in flight)
  for (int c = c0; c < c1; c += UNR, xp += (size_t)UNR * plane, op += (size_t)UNR * plane) {
    float v[UNR][PXT][4];
#pragma unroll
    for (int cb = 0; cb < UNR; ++cb) {
      if (c + cb < c1) {
        const float* xc = xp + (size_t)cb * plane;
#pragma unroll
        for (int k = 0; k < PXT; ++k) {      // clamped offsets are always valid addresses
          v[cb][k][0] = __ldg(xc + fp[k].o00);
          v[cb][k][1] = __ldg(xc + fp[k].o01);
          v[cb][k][2] = __ldg(xc + fp[k].o10);
          v[cb][k][3] = __ldg(xc + fp[k].o11);
        }
      }
    }
#pragma unroll
    for (int cb = 0; cb < UNR; ++cb) {
      if (c + cb < c1) {
#pragma unroll
        for (int k = 0; k < PXT; ++k)
          if (ok[k])
            op[(size_t)cb * plane + 32 * k] =
                fmaf(v[cb][k][3], fp[k].w11, fmaf(v[cb][k][2], fp[k].w10, fmaf(v[cb][k][1], fp[k].w01, v[cb][k][0] * fp[k].w00)));
      }
    }
  }
}

// ---------------------------------------------------------------------------------- NCHW bwd
// ATOMIC_GFLOW: several channel chunks contribute to the same gflow element (gflow zero-filled by the host).
template <int PXT, bool NEED_GX, bool ATOMIC_GFLOW, int UNR>
__global__ void __launch_bounds__(kWarps * 32)
warp_bwd_nchw_kernel(const float* __restrict__ gout, const float* __restrict__ x, const float* __restrict__ flow,
                     float* __restrict__ gx, float* __restrict__ gflow, int B, int C, int H, int W, int nchunk,
                     int runs_per_row, unsigned run_magic, int row_map, int use_mask, int align_corners, float sx,
                     float sy) {
  const int lane = threadIdx.x & 31;
  const RunCoord rc = locate_run(H, 32 * PXT, runs_per_row, run_magic, row_map);
  if (!rc.live) return;
  const size_t plane = (size_t)H * W;
  const size_t row = (size_t)rc.y * W;
  const float* fb = flow + (size_t)rc.b * 2 * plane + row;
  float fx[PXT], fy[PXT];
  bool ok[PXT];
#pragma unroll
  for (int k = 0; k < PXT; ++k) {
    const int px = rc.x0 + 32 * k + lane;
    ok[k] = px < W;
    fx[k] = ok[k] ? __ldg(fb + px) : 0.0f;
    fy[k] = ok[k] ? __ldg(fb + plane + px) : 0.0f;
  }
  Footprint fp[PXT];
  float ux[PXT], uy[PXT], tx[PXT], ty[PXT], msk[PXT];
  unsigned inb[PXT];    // bit0..3: corner 00,01,10,11 in bounds
  int x0u = 0, y0u = 0; // unclamped top-left corner (of the last pixel; only used when PXT == 1)
#pragma unroll
  for (int k = 0; k < PXT; ++k) {
    const int px = rc.x0 + 32 * k + lane;
    const float ix = sample_coord((float)px, fx[k], W, align_corners);
    const float iy = sample_coord((float)rc.y, fy[k], H, align_corners);
    const Bilinear bl = make_bilinear(ix, iy, H, W);
    msk[k] = (!ok[k] || (use_mask && cover_of(bl) < kMaskThreshold)) ? 0.0f : 1.0f;
    const int xa = min(max(bl.x0, 0), W - 1), xb = min(max(bl.x0 + 1, 0), W - 1);
    const int ya = min(max(bl.y0, 0), H - 1), yb = min(max(bl.y0 + 1, 0), H - 1);
    fp[k].o00 = ya * W + xa; fp[k].o01 = ya * W + xb; fp[k].o10 = yb * W + xa; fp[k].o11 = yb * W + xb;
    fp[k].w00 = bl.w00; fp[k].w01 = bl.w01; fp[k].w10 = bl.w10; fp[k].w11 = bl.w11;
    inb[k] = (bl.in00 ? 1u : 0u) | (bl.in01 ? 2u : 0u) | (bl.in10 ? 4u : 0u) | (bl.in11 ? 8u : 0u);
    if (!ok[k]) inb[k] = 0u;
    tx[k] = bl.tx; ty[k] = bl.ty;
    ux[k] = (floorf(ix) + 1.0f) - ix;
    uy[k] = (floorf(iy) + 1.0f) - iy;
    x0u = bl.x0;
    y0u = bl.y0;
  }
  // side-by-side test for the warp-aggregated scatter (PXT == 1 only): lane-1's footprint is one column to the left
  bool take = false, taken = false;
  if (NEED_GX && PXT == 1) {
    const int px0 = __shfl_up_sync(kFullMask, x0u, 1), py0 = __shfl_up_sync(kFullMask, y0u, 1);
    const int pok = __shfl_up_sync(kFullMask, (int)ok[0], 1);
    take = lane > 0 && ok[0] && pok && px0 + 1 == x0u && py0 == y0u;
    taken = __shfl_down_sync(kFullMask, (int)take, 1) && lane < 31;
  }
  const int cch = (C + nchunk - 1) / nchunk;       // channels per thread
  const int c0 = rc.chunk * cch, c1 = min(C, c0 + cch);
  const float* xp = x + ((size_t)rc.b * C + c0) * plane;
  const float* gp = gout + ((size_t)rc.b * C + c0) * plane + row + rc.x0 + lane;
  float* gxp = NEED_GX ? gx + ((size_t)rc.b * C + c0) * plane : nullptr;
  float gix[PXT], giy[PXT];
#pragma unroll
  for (int k = 0; k < PXT; ++k) gix[k] = giy[k] = 0.0f;
  // UNR channels per batch, all loads of a batch issued before the first use (see the forward kernel)
  for (int c = c0; c < c1; c += UNR, xp += (size_t)UNR * plane, gp += (size_t)UNR * plane) {
    float g[UNR][PXT], v[UNR][PXT][4];
#pragma unroll
    for (int cb = 0; cb < UNR; ++cb) {
      if (c + cb < c1) {
        const float* xc = xp + (size_t)cb * plane;
#pragma unroll
        for (int k = 0; k < PXT; ++k) g[cb][k] = ok[k] ? __ldg(gp + (size_t)cb * plane + 32 * k) : 0.0f;
#pragma unroll
        for (int k = 0; k < PXT; ++k) {
          v[cb][k][0] = __ldg(xc + fp[k].o00);
          v[cb][k][1] = __ldg(xc + fp[k].o01);
          v[cb][k][2] = __ldg(xc + fp[k].o10);
          v[cb][k][3] = __ldg(xc + fp[k].o11);
        }
      }
    }
#pragma unroll
    for (int cb = 0; cb < UNR; ++cb) {
      if (c + cb >= c1) continue;
#pragma unroll
      for (int k = 0; k < PXT; ++k) {
        const float gm = g[cb][k] * msk[k];
        const float v00 = (inb[k] & 1u) ? v[cb][k][0] : 0.0f, v01 = (inb[k] & 2u) ? v[cb][k][1] : 0.0f;
        const float v10 = (inb[k] & 4u) ? v[cb][k][2] : 0.0f, v11 = (inb[k] & 8u) ? v[cb][k][3] : 0.0f;
        // d out / d ix and d out / d iy of the bilinear interpolant
        gix[k] = fmaf(gm, (v01 - v00) * uy[k] + (v11 - v10) * ty[k], gix[k]);
        giy[k] = fmaf(gm, (v10 - v00) * ux[k] + (v11 - v01) * tx[k], giy[k]);
        if (NEED_GX) {
          float* q = gxp + (size_t)(c + cb - c0) * plane;
          float c00 = gm * fp[k].w00, c01 = gm * fp[k].w01, c10 = gm * fp[k].w10, c11 = gm * fp[k].w11;
          if (PXT == 1) {
            // warp-aggregated scatter: the right-hand corners of lane-1 are this lane's left-hand corners whenever the
            // two footprints sit side by side (`take`); add them here and let lane-1 skip its two REDs (`taken`)
            const float n01 = __shfl_up_sync(kFullMask, c01, 1), n11 = __shfl_up_sync(kFullMask, c11, 1);
            if (take) {
              c00 += n01;
              c10 += n11;
            }
            if (inb[k] & 1u) atomicAdd(q + fp[k].o00, c00);
            if (inb[k] & 4u) atomicAdd(q + fp[k].o10, c10);
            if (!taken) {
              if (inb[k] & 2u) atomicAdd(q + fp[k].o01, c01);
              if (inb[k] & 8u) atomicAdd(q + fp[k].o11, c11);
            }
          } else {
            if (inb[k] & 1u) atomicAdd(q + fp[k].o00, c00);
            if (inb[k] & 2u) atomicAdd(q + fp[k].o01, c01);
            if (inb[k] & 4u) atomicAdd(q + fp[k].o10, c10);
            if (inb[k] & 8u) atomicAdd(q + fp[k].o11, c11);
          }
        }
      }
    }
  }
  float* gfb = gflow + (size_t)rc.b * 2 * plane + row + rc.x0 + lane;
#pragma unroll
  for (int k = 0; k < PXT; ++k) {
    if (!ok[k]) continue;
    if (ATOMIC_GFLOW) {
      atomicAdd(gfb + 32 * k, gix[k] * sx);
      atomicAdd(gfb + plane + 32 * k, giy[k] * sy);
    } else {
      gfb[32 * k] = gix[k] * sx;
      gfb[plane + 32 * k] = giy[k] * sy;
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
// gx uses 16-byte vector atomics (sm_90+); the per-group partial flow gradients are accumulated with
// atomicAdd into a zero-filled gflow.
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

// pixels per thread: enough to cover short rows with one run, at most 4
int env_int(const char* name) {
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
}

// Pixels per thread.  Measured on B200 (profiles/r1_warp_sweep.txt): the forward kernel is fastest with 2, the
// backward kernel (more registers per pixel, atomics) with 1 -- occupancy beats per-thread ILP there.
int pick_pxt(int W, bool bwd) {
  static const int forced = env_int("UOF_WARP_PXT");
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  if (bwd || W <= 32) return 1;
  return 2;
}

// Channels per thread (the grid splits C into ceil(C / cch) chunks).
int pick_cch() {
  static const int forced = env_int("UOF_WARP_CCH");
  return forced > 0 ? forced : CCH;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_warp_fwd(const float* x, const float* flow, float* out, int B, int C, int H, int W, int use_mask,
                            int align_corners, int channels_last, uof_stream_t stream_) {
  if (int rc = check_args("warp_fwd", x, flow, out, B, C, H, W, channels_last)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!channels_last) {
    const int nchunk = ceil_div(C, pick_cch()), pxt = pick_pxt(W, false), runs = ceil_div(W, 32 * pxt);
    UOF_REQUIRE(B <= 65535 && nchunk <= 65535 && (long long)H * runs * runs < (1ll << 31), "warp_fwd: grid too large (B=%d)", B);
    // measured (profiles/r2_warp_map.txt): the forward kernel is faster when a block reads 4 consecutive runs of one row
    static const int fmap = env_int("UOF_WARP_FWD_ROWMAP");
    const dim3 blocks = run_grid(B, nchunk, H, runs, fmap == 1);
    const int rmap = fmap == 1 ? 1 : 0;
#define UOF_FWD(P, U) warp_fwd_nchw_kernel<P, U><<<blocks, kWarps * 32, 0, stream>>>(x, flow, out, B, C, H, W, nchunk, runs, magic_of(runs), rmap, use_mask, align_corners)
    // channels per load batch: 4 for feature maps (measured 30.4 -> 26.4 us at 16x32x64x208), 2 for 3-channel images
    static const int unr_env = env_int("UOF_WARP_UNROLL");
    const int unr = (unr_env == 2 || unr_env == 4) ? unr_env : (C >= 8 ? 4 : 2);
    if (pxt == 4) UOF_FWD(4, 2); else if (pxt == 2) { if (unr == 4) UOF_FWD(2, 4); else UOF_FWD(2, 2); } else { if (unr == 4) UOF_FWD(1, 4); else UOF_FWD(1, 2); }
#undef UOF_FWD
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
                            int H, int W, int nchunk, int runs, int use_mask, int align_corners, float sx, float sy,
                            cudaStream_t stream) {
  static const int bmap = env_int("UOF_WARP_BWD_LINMAP");
  const dim3 blocks = run_grid(B, nchunk, H, runs, bmap != 1);
  const int rmap = bmap != 1 ? 1 : 0;
  const unsigned magic = magic_of(runs);
  static const int unr_env = env_int("UOF_WARP_UNROLL");
  const int unr = (unr_env == 2 || unr_env == 4) ? unr_env : (C >= 8 ? 4 : 2);
#define UOF_LAUNCH(GX, AT)                                                                                              \
  do {                                                                                                                  \
    if (unr == 4 && PXT == 1)                                                                                           \
      warp_bwd_nchw_kernel<PXT, GX, AT, (PXT == 1 ? 4 : 2)><<<blocks, kWarps * 32, 0, stream>>>(                         \
          gout, x, flow, gx, gflow, B, C, H, W, nchunk, runs, magic, rmap, use_mask, align_corners, sx, sy);            \
    else                                                                                                                \
      warp_bwd_nchw_kernel<PXT, GX, AT, 2><<<blocks, kWarps * 32, 0, stream>>>(                                         \
          gout, x, flow, gx, gflow, B, C, H, W, nchunk, runs, magic, rmap, use_mask, align_corners, sx, sy);            \
  } while (0)
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
  if (!channels_last) {      // argument validation before the first CUDA call
    const int nchunk = ceil_div(C, pick_cch()), runs = ceil_div(W, 32 * pick_pxt(W, true));
    UOF_REQUIRE(B <= 65535 && nchunk <= 65535 && (long long)H * runs * runs < (1ll << 31), "warp_bwd: grid too large (B=%d)", B);
  }
  if (gx) UOF_CUDA(cudaMemsetAsync(gx, 0, (size_t)B * C * H * W * sizeof(float), stream));
  if (!channels_last) {
    const int nchunk = ceil_div(C, pick_cch()), pxt = pick_pxt(W, true), runs = ceil_div(W, 32 * pxt);
    if (nchunk > 1) UOF_CUDA(cudaMemsetAsync(gflow, 0, (size_t)B * 2 * H * W * sizeof(float), stream));
    if (pxt == 4)
      launch_bwd_nchw<4>(gout, x, flow, gx, gflow, B, C, H, W, nchunk, runs, use_mask, align_corners, sx, sy, stream);
    else if (pxt == 2)
      launch_bwd_nchw<2>(gout, x, flow, gx, gflow, B, C, H, W, nchunk, runs, use_mask, align_corners, sx, sy, stream);
    else
      launch_bwd_nchw<1>(gout, x, flow, gx, gflow, B, C, H, W, nchunk, runs, use_mask, align_corners, sx, sy, stream);
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
