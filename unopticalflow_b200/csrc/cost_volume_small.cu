// a1, smallest pyramid levels: cost volume forward for feature maps of at most kMaxQuads pixel quads per image
// (the 4x13 and 8x26 levels of the 256x832 pyramid: 1.6 - 4.5 MB per call).
// Replaces PWC_tf.corr_naive, /root/reference/core/networks/structures/pwc_tf.py:97-106.
//
// The tiled kernels (cost_volume.cu, cost_volume_tma.cu) are built for throughput: 8x32-pixel CTA tiles, channel
// slabs through shared memory.  On these levels they run one or two tiles per image -- a 4x13 map uses 20 % of a
// tile -- so they split the channel loop over CTAs and combine the partial sums with fp32 atomics into a zeroed
// output: 19-22 us per call for a few MFLOP, dominated by the atomics and the staging latency (ncu, round 2).
//
// Here the parallelism comes from the displacement rows instead: one CTA per (image, dy) computes out[b, 9 dy + dx]
// for all nine dx, all pixels and all channels -- 9 x B CTAs (144 at B = 16, one wave of the 148 SMs), no atomics,
// no zero-fill.  A thread owns a quad of 4 adjacent pixels x 9 dx (36 accumulators) and slides a 12-float window of
// the second feature map; the loads go straight through L1 (the whole working set of a CTA is < 100 KB and every
// value is reused by the neighbouring quads and by the other dy CTAs of the image through L2).  When an image has
// few quads the channels are interleaved over G thread groups, reduced through shared memory at the end.
#include <stdlib.h>

#include "cost_volume.h"

namespace uof {
namespace cv {
namespace {

constexpr int kMaxQuads = 64;       // quads per image this kernel accepts: 4x13 (16) and 8x26 (56); measured at 16x52 (208
                                    // quads, 2 channel groups) the 16 scalar loads per 36 FMA make it LSU-bound: 35 us vs 18 us tiled
constexpr int kAcc = ND * PX;       // 36 accumulators per thread

__global__ void __launch_bounds__(512)
cost_volume_fwd_small_kernel(const float* __restrict__ f1, long long f1_bs, const float* __restrict__ f2,
                             float* __restrict__ out, int C, int H, int W, long long out_bs, int qpr /* quads per row */, int nq_pad /* multiple of 32 */,
                             int G, float inv_c) {
  extern __shared__ float red[];                  // [G][kAcc][nq_pad], only used when G > 1
  const int dy = blockIdx.x, b = blockIdx.y;
  const int q = threadIdx.x % nq_pad, g = threadIdx.x / nq_pad;
  const int y = q / qpr, x = (q - y * qpr) * PX;
  const bool live = y < H;                        // quads beyond the image only take part in the barriers
  const size_t plane = (size_t)H * W;
  const int y2 = y + dy - RAD;                    // row of the second feature map
  const bool row_ok = live && y2 >= 0 && y2 < H;

  float acc[ND][PX];
#pragma unroll
  for (int j = 0; j < ND; ++j)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[j][p] = 0.0f;

  if (row_ok) {                                   // a row outside the second map contributes zeros only
    bool aok[PX], wok[PX + ND - 1];
#pragma unroll
    for (int p = 0; p < PX; ++p) aok[p] = x + p < W;
#pragma unroll
    for (int k = 0; k < PX + ND - 1; ++k) wok[k] = x + k - RAD >= 0 && x + k - RAD < W;
    const float* a_ptr = f1 + (size_t)b * f1_bs + (size_t)y * W + x;
    const float* w_ptr = f2 + (size_t)b * C * plane + (size_t)y2 * W + x - RAD;
#pragma unroll 4
    for (int c = g; c < C; c += G) {      // latency-bound: four channels (64 loads) in flight per thread
      const float* ap = a_ptr + (size_t)c * plane;
      const float* wp = w_ptr + (size_t)c * plane;
      float a[PX], win[PX + ND - 1];
#pragma unroll
      for (int p = 0; p < PX; ++p) a[p] = aok[p] ? __ldg(ap + p) : 0.0f;
#pragma unroll
      for (int k = 0; k < PX + ND - 1; ++k) win[k] = wok[k] ? __ldg(wp + k) : 0.0f;
#pragma unroll
      for (int j = 0; j < ND; ++j)
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[j][p] = fmaf(a[p], win[p + j], acc[j][p]);
    }
  }

  float* ob = out + (size_t)b * out_bs + (size_t)(dy * ND) * plane + (size_t)y * W + x;
  if (G == 1) {
    if (!live) return;
#pragma unroll
    for (int j = 0; j < ND; ++j)
#pragma unroll
      for (int p = 0; p < PX; ++p)
        if (x + p < W) ob[(size_t)j * plane + p] = acc[j][p] * inv_c;
    return;
  }
  // combine the G channel groups: group g sums values g, g + G, ... of every quad
#pragma unroll
  for (int j = 0; j < ND; ++j)
#pragma unroll
    for (int p = 0; p < PX; ++p) red[((size_t)g * kAcc + j * PX + p) * nq_pad + q] = acc[j][p];
  __syncthreads();
  if (!live) return;
  for (int v = g; v < kAcc; v += G) {
    float s = 0.0f;
    for (int gg = 0; gg < G; ++gg) s += red[((size_t)gg * kAcc + v) * nq_pad + q];
    const int j = v / PX, p = v - j * PX;
    if (x + p < W) ob[(size_t)j * plane + p] = s * inv_c;
  }
}

// ------------------------------------------------------------------------------------------ backward
// The tiled backward kernels run these levels as one or two 256-pixel tiles per image (a 4x13 map fills 20 % of a tile,
// and W % 4 != 0 rules out TMA), split the channel loop over CTAs to find parallelism and pay a 108-gather prologue, a
// cp.async pipeline fill and two CTA barriers per 8-channel slab for a few MFLOP: 20-25 us per call.
// A first dedicated kernel (round 2) read everything straight through L1 -- a lane's four consecutive pixels made every
// scalar load of a warp span 512 B and the kernel was LSU-bound and slower (31.5 / 39.2 us).  This one stages through
// shared memory instead: a CTA owns (image, role, band of BH output rows, chunk of channels) and first copies
//   Gs[81][BH][Wp]        the coefficients of its pixels: role 0 gout[d] at the pixel itself, role 1 the index-flipped
//                         gout[-d] at the displaced pixel (zero outside the map), so both roles run the same loop;
//   Fp[ch][BH + 8][Wp + 8] its channels of the other feature map with a zero halo of 4,
// with rows padded to whole quads (Wp = 4 ceil(W / 4)), so that every later access is an aligned LDS.128 without bounds
// checks (staging: see the kernel).  A thread then owns a quad of 4 adjacent pixels and CT channels: per displacement row 3 CT window loads + 9
// coefficient loads (LDS.128) feed 36 CT FMAs.  No atomics, every output written once, `gadd` folded in.
template <int CT, int NT, bool VEC>
__global__ void __launch_bounds__(NT)
cost_volume_bwd_staged_kernel(const float* __restrict__ gout, long long gout_bs, const float* __restrict__ f1, long long f1_bs,
                              const float* __restrict__ f2, const float* __restrict__ gadd, long long gadd_bs,
                              float* __restrict__ gf1, float* __restrict__ gf2, int C, int H, int W, int BH, int nbands, int nq_pad,
                              float inv_c) {
  extern __shared__ float4 staged_smem[];
  const int Wq = (W + PX - 1) / PX, Wp = Wq * PX, Wf = Wp + 2 * RAD;
  const int groups = NT / nq_pad, cpc = groups * CT;       // channels per CTA
  float* Gs = reinterpret_cast<float*>(staged_smem);       // [81][BH][Wp]
  float* Fp = Gs + ND * ND * BH * Wp;                      // [cpc][BH + 8][Wf]
  const bool mirror = (int)blockIdx.y >= nbands;
  const int band = (int)blockIdx.y - (mirror ? nbands : 0);
  const int y0 = band * BH, b = blockIdx.z, c0 = (int)blockIdx.x * cpc;
  const size_t plane = (size_t)H * W;

  // ---- staging.  Shared memory is zeroed (halo, padding, out-of-map coefficients), then every SOURCE element the CTA needs
  // is loaded once with plain coalesced loads -- 16-byte ones when the CTA covers whole images whose planes are multiples
  // of 4 floats, because then the 81 coefficient planes / the CTA's channel planes are one contiguous block -- and
  // scattered to its padded (role 1: flipped and shifted) position.  (4-byte cp.async staging was measured at ~2 elements
  // per clock and SM: 12 of the kernel's 18 us at 4x13.)
  {
    float4* z = staged_smem;
    const int total4 = (ND * ND * BH * Wp + cpc * (BH + 2 * RAD) * Wf) / 4;
    for (int i = threadIdx.x; i < total4; i += NT) z[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
  __syncthreads();
  const float* gb = gout + (size_t)b * gout_bs;
  const float* fsrc = (mirror ? f1 + (size_t)b * f1_bs : f2 + (size_t)b * C * plane) + (size_t)c0 * plane;
  const int frows = BH + 2 * RAD;
  const int iplane = H * W;
  // (plane, row, column) of a flat element index are decoded once per load and advanced incrementally for the following
  // elements of a vector: the divisions of a per-element decode were most of the staging instructions
  struct Pos { int p, y, x; };
  auto decode = [&](int e) {
    Pos t;
    t.p = e / iplane;
    const int rem = e - t.p * iplane;
    t.y = rem / W;
    t.x = rem - t.y * W;
    return t;
  };
  auto advance = [&](Pos& t) {
    if (++t.x == W) {
      t.x = 0;
      if (++t.y == H) {
        t.y = 0;
        ++t.p;
      }
    }
  };
  auto put_g = [&](const Pos& t, float v) {    // t = (displacement, row, column) of a gout element of this image
    int dd = t.p, qy = t.y, qx = t.x;
    if (mirror) {                              // coefficient (i', j') at pixel q is gout[(8 - i', 8 - j')] at q + (i' - 4, j' - 4)
      const int is = t.p / ND, js = t.p - is * ND;
      dd = (ND - 1 - is) * ND + (ND - 1 - js);
      qy = t.y - (ND - 1 - is - RAD);
      qx = t.x - (ND - 1 - js - RAD);
    }
    if (qy >= y0 && qy < y0 + BH && qy < H && qx >= 0 && qx < W) Gs[((size_t)dd * BH + (qy - y0)) * Wp + qx] = v;
  };
  auto put_f = [&](const Pos& t, float v) {    // t = (channel of this CTA, row, column)
    const int r = t.y - (y0 - RAD);
    if (r >= 0 && r < frows) Fp[((size_t)t.p * frows + r) * Wf + t.x + RAD] = v;
  };
  const int nch = min(cpc, C - c0);            // channels of this CTA that exist
  const int ng = ND * ND * iplane, nf = nch * iplane;
  if (VEC) {
    const float4* g4 = reinterpret_cast<const float4*>(gb);
    const float4* f4 = reinterpret_cast<const float4*>(fsrc);
    for (int i = threadIdx.x; i < ng / 4; i += NT) {
      const float4 v = __ldg(g4 + i);
      Pos t = decode(4 * i);
      put_g(t, v.x); advance(t); put_g(t, v.y); advance(t); put_g(t, v.z); advance(t); put_g(t, v.w);
    }
    for (int i = threadIdx.x; i < nf / 4; i += NT) {
      const float4 v = __ldg(f4 + i);
      Pos t = decode(4 * i);
      put_f(t, v.x); advance(t); put_f(t, v.y); advance(t); put_f(t, v.z); advance(t); put_f(t, v.w);
    }
  } else {
    for (int i = threadIdx.x; i < ng; i += NT) put_g(decode(i), __ldg(gb + i));
    for (int i = threadIdx.x; i < nf; i += NT) put_f(decode(i), __ldg(fsrc + i));
  }
  __syncthreads();

  const int q = threadIdx.x % nq_pad, g = threadIdx.x / nq_pad;
  const int y = q / Wq, x = (q - y * Wq) * PX;
  if (y >= BH || y0 + y >= H || g >= groups) return;
  float acc[CT][PX];
#pragma unroll
  for (int cc = 0; cc < CT; ++cc)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[cc][p] = 0.0f;
  const float* gq = Gs + (size_t)y * Wp + x;
  const float* fq = Fp + ((size_t)(g * CT) * frows + y) * Wf + x;
#pragma unroll 1
  for (int i = 0; i < ND; ++i) {
    float win[CT][PX + ND - 1];
#pragma unroll
    for (int cc = 0; cc < CT; ++cc) {
      const float4* wp = reinterpret_cast<const float4*>(fq + ((size_t)cc * frows + i) * Wf);
      const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
      win[cc][0] = w0.x; win[cc][1] = w0.y; win[cc][2] = w0.z; win[cc][3] = w0.w;
      win[cc][4] = w1.x; win[cc][5] = w1.y; win[cc][6] = w1.z; win[cc][7] = w1.w;
      win[cc][8] = w2.x; win[cc][9] = w2.y; win[cc][10] = w2.z; win[cc][11] = w2.w;
    }
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const float4 k4 = *reinterpret_cast<const float4*>(gq + (size_t)((i * ND + j) * BH) * Wp);
      const float kc[PX] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
      for (int cc = 0; cc < CT; ++cc)
#pragma unroll
        for (int p = 0; p < PX; ++p) acc[cc][p] = fmaf(kc[p], win[cc][p + j], acc[cc][p]);
    }
  }
  const int cbase = c0 + g * CT;
  float* dst = (mirror ? gf2 : gf1) + ((size_t)b * C + cbase) * plane + (size_t)(y0 + y) * W + x;
  const float* add = (!mirror && gadd) ? gadd + (size_t)b * gadd_bs + (size_t)cbase * plane + (size_t)(y0 + y) * W + x : nullptr;
#pragma unroll
  for (int cc = 0; cc < CT; ++cc) {
    if (cbase + cc >= C) break;
#pragma unroll
    for (int p = 0; p < PX; ++p)
      if (x + p < W) dst[(size_t)cc * plane + p] = fmaf(acc[cc][p], inv_c, add ? __ldg(add + (size_t)cc * plane + p) : 0.0f);
  }
}

}  // namespace

// Returns false when the level is not "small" (or UOF_CV_NO_SMALL is set); otherwise launches and stores the status.
bool fwd_small(const float* f1, long long f1_bs, const float* f2, float* out, int B, int C, int H, int W, long long out_bs,
               cudaStream_t stream, int* rc) {
  static const bool off = getenv("UOF_CV_NO_SMALL") != nullptr;
  const int qpr = ceil_div(W, PX);
  const long long nq = (long long)qpr * H;
  if (off || nq > kMaxQuads || B > 65535) return false;
  const int nq_pad = ceil_div((int)nq, 32) * 32;
  int G = 512 / nq_pad;
  G = G < 1 ? 1 : (G > 8 ? 8 : G);
  if (G > C) G = C;
  const size_t smem = G > 1 ? (size_t)G * kAcc * nq_pad * sizeof(float) : 0;
  auto fail = [&](cudaError_t e) {
    set_error("cost_volume_fwd (small): %s", cudaGetErrorString(e));
    *rc = UOF_ERR_CUDA;
    return true;
  };
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(cost_volume_fwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(e);
  }
  cost_volume_fwd_small_kernel<<<dim3(ND, B), nq_pad * G, smem, stream>>>(f1, f1_bs, f2, out, C, H, W, out_bs, qpr, nq_pad, G,
                                                                         1.0f / (float)C);
  count_launch();
  *rc = check_launch("cost_volume_fwd (small)");
  return true;
}

template <int CT, int NT, bool VEC>
bool launch_staged(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
                   long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, int BH, int nq_pad, cudaStream_t stream,
                   int* rc) {
  const int Wp = ceil_div(W, PX) * PX, groups = NT / nq_pad, cpc = groups * CT, nbands = ceil_div(H, BH);
  const size_t smem = ((size_t)ND * ND * BH * Wp + (size_t)cpc * (BH + 2 * RAD) * (Wp + 2 * RAD)) * sizeof(float);
  if (smem > 200 * 1024) return false;
  static size_t configured[64] = {};   // per instantiation and device (the attribute is per device)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
  if (dev < 0 || smem > configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(cost_volume_bwd_staged_kernel<CT, NT, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cost_volume_bwd (staged): %s", cudaGetErrorString(e));
      *rc = UOF_ERR_CUDA;
      return true;
    }
    if (dev >= 0) configured[dev] = smem;
  }
  cost_volume_bwd_staged_kernel<CT, NT, VEC><<<dim3(ceil_div(C, cpc), 2 * nbands, B), NT, smem, stream>>>(
      gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, C, H, W, BH, nbands, nq_pad, 1.0f / (float)C);
  count_launch();
  *rc = check_launch("cost_volume_bwd (staged)");
  return true;
}

bool bwd_small(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
               long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, cudaStream_t stream, int* rc) {
  // Measured (B200, kernel_bench, round 2): 11.3 us at 16x196x4x13 against 25.2 us for the tiled kernels (18.6 us while the
  // staging still went through 4-byte cp.async), but 21.7 vs 20.0 us at 16x128x8x26 -- there the 81-tap loop itself is
  // shared-memory bound (2.3 B of LDS per FMA at 4 channels per thread against 1.5 in the tiled kernels) -- and far slower at
  // 16x52 in bands of rows (every band CTA re-stages the coefficient planes).  Default: images of at most kStagedQuads
  // quads, i.e. the 4x13 level; UOF_CV_BWD_SMALL_MAXQ raises the limit (the larger forms stay parity-tested),
  // UOF_CV_NO_SMALL_BWD=1 keeps the tiled kernels everywhere.
  constexpr int kStagedQuads = 16;
  static const bool off = getenv("UOF_CV_NO_SMALL_BWD") != nullptr;
  static const int maxq = getenv("UOF_CV_BWD_SMALL_MAXQ") ? atoi(getenv("UOF_CV_BWD_SMALL_MAXQ")) : kStagedQuads;
  const int qpr = ceil_div(W, PX);
  const long long nq_img = (long long)qpr * H;
  if (off || nq_img > maxq || B > 65535 || (long long)ND * ND * H * W >= (1ll << 30) || (long long)C * H * W >= (1ll << 30)) return false;
  // band height: the whole image when it has <= 64 quads, else the most rows that give <= 64 quads
  int BH = H;
  while ((long long)BH * qpr > 64 && BH > 1) BH = (BH + 1) / 2;
  const int nq = BH * qpr;
  int nq_pad = 16;
  while (nq_pad < nq) nq_pad *= 2;
  // Both phases of the kernel (issuing the staging copies, the 81-tap loop) are latency bound with one CTA per SM, so the
  // CTA has 16 warps; channels per thread: 2 when an image has few quads (many thread groups), 4 for a full 64-quad band
  // (UOF_CV_STAGED_WIDE=1: the first form, 8 warps with 4 / 8 channels per thread)
  static const bool wide = getenv("UOF_CV_STAGED_WIDE") != nullptr;
  // 16-byte staging loads: the CTA covers whole images and every plane / batch stride / base pointer is 16-byte aligned
  const uintptr_t bits = reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2);
  const bool vec = BH == H && ((long long)H * W) % 4 == 0 && gout_bs % 4 == 0 && f1_bs % 4 == 0 && (bits & 15u) == 0;
#define UOF_STAGED(CT_, NT_) (vec ? launch_staged<CT_, NT_, true>(gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, B, C, H, W, BH, nq_pad, stream, rc) \
                                  : launch_staged<CT_, NT_, false>(gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, B, C, H, W, BH, nq_pad, stream, rc))
  if (wide) return nq_pad >= 64 ? UOF_STAGED(8, 256) : UOF_STAGED(4, 256);
  return nq_pad >= 64 ? UOF_STAGED(4, 512) : UOF_STAGED(2, 512);
#undef UOF_STAGED
}

}  // namespace cv
}  // namespace uof
