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
// checks.  The copies are 4-byte zero-filling cp.async (W % 4 != 0 at these levels), all in flight at once: with plain
// loads every (displacement, row) iteration of a warp waited for its own load and the staging alone took 40 us.  A thread then owns a quad of 4 adjacent pixels and CT channels: per displacement row 3 CT window loads + 9
// coefficient loads (LDS.128) feed 36 CT FMAs.  No atomics, every output written once, `gadd` folded in.
constexpr int kColIters = 3;          // staged rows are at most 96 floats wide (W <= 64)

template <int CT>
__global__ void __launch_bounds__(256)
cost_volume_bwd_staged_kernel(const float* __restrict__ gout, long long gout_bs, const float* __restrict__ f1, long long f1_bs,
                              const float* __restrict__ f2, const float* __restrict__ gadd, long long gadd_bs,
                              float* __restrict__ gf1, float* __restrict__ gf2, int C, int H, int W, int BH, int nbands, int nq_pad,
                              float inv_c) {
  extern __shared__ float4 staged_smem[];
  const int Wq = (W + PX - 1) / PX, Wp = Wq * PX, Wf = Wp + 2 * RAD;
  const int groups = 256 / nq_pad, cpc = groups * CT;      // channels per CTA
  float* Gs = reinterpret_cast<float*>(staged_smem);       // [81][BH][Wp]
  float* Fp = Gs + ND * ND * BH * Wp;                      // [cpc][BH + 8][Wf]
  const bool mirror = (int)blockIdx.y >= nbands;
  const int band = (int)blockIdx.y - (mirror ? nbands : 0);
  const int y0 = band * BH, b = blockIdx.z, c0 = (int)blockIdx.x * cpc;
  const size_t plane = (size_t)H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- stage the coefficients: a warp takes whole displacements, lanes run along x (no runtime division in these loops:
  // the first version decoded a flat (displacement, row) index and spent 14 k instructions per warp on the staging alone)
  const float* gb = gout + (size_t)b * gout_bs;
  for (int d = warp; d < ND * ND; d += 8) {
    const int i = d / ND, j = d - i * ND;
    // role 1: coefficient of displacement (i, j) at pixel q is gout[(8 - i, 8 - j)] at q + (i - 4, j - 4)
    const int xs = mirror ? j - RAD : 0, yo = mirror ? i - RAD : 0;
    const float* dsrc = gb + (size_t)(mirror ? (ND - 1 - i) * ND + (ND - 1 - j) : d) * plane + xs;
    float* dst = Gs + (size_t)d * BH * Wp;
    for (int y = 0; y < BH; ++y, dst += Wp) {
      const int ys = y0 + y + yo;
      const bool row_ok = ys >= 0 && ys < H && y0 + y < H;
      const float* src = dsrc + (size_t)min(max(ys, 0), H - 1) * W;
#pragma unroll
      for (int t = 0; t < kColIters; ++t) {                  // a lane owns columns lane, lane + 32, lane + 64 (Wp <= 64 + 8)
        const int x = lane + 32 * t;
        const bool ok = row_ok && x < W && x + xs >= 0 && x + xs < W;
        if (x < Wp) cp_async_4(dst + x, ok ? src + x : gb, ok);   // asynchronous, zero-filled: all copies of a CTA in flight at once
      }
    }
  }
  // ---- stage the feature rows with their zero halo: a warp takes whole channels
  const float* fsrc = mirror ? f1 + (size_t)b * f1_bs : f2 + (size_t)b * C * plane;
  const int frows = BH + 2 * RAD;
  for (int cc = warp; cc < cpc; cc += 8) {
    const bool c_ok = c0 + cc < C;
    const float* csrc = fsrc + (size_t)min(c0 + cc, C - 1) * plane - RAD;
    float* dst = Fp + (size_t)cc * frows * Wf;
    for (int r = 0; r < frows; ++r, dst += Wf) {
      const int yy = y0 - RAD + r;
      const bool row_ok = c_ok && yy >= 0 && yy < H;
      const float* src = csrc + (size_t)min(max(yy, 0), H - 1) * W;
#pragma unroll
      for (int t = 0; t < kColIters; ++t) {
        const int k = lane + 32 * t;
        const bool ok = row_ok && k >= RAD && k < W + RAD;
        if (k < Wf) cp_async_4(dst + k, ok ? src + k : fsrc, ok);
      }
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
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

template <int CT>
bool launch_staged(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
                   long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, int BH, int nq_pad, cudaStream_t stream,
                   int* rc) {
  const int Wp = ceil_div(W, PX) * PX, groups = 256 / nq_pad, cpc = groups * CT, nbands = ceil_div(H, BH);
  const size_t smem = ((size_t)ND * ND * BH * Wp + (size_t)cpc * (BH + 2 * RAD) * (Wp + 2 * RAD)) * sizeof(float);
  if (smem > 200 * 1024) return false;
  static size_t configured[64] = {};   // per instantiation and device (the attribute is per device)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
  if (dev < 0 || smem > configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(cost_volume_bwd_staged_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cost_volume_bwd (staged): %s", cudaGetErrorString(e));
      *rc = UOF_ERR_CUDA;
      return true;
    }
    if (dev >= 0) configured[dev] = smem;
  }
  cost_volume_bwd_staged_kernel<CT><<<dim3(ceil_div(C, cpc), 2 * nbands, B), 256, smem, stream>>>(
      gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, C, H, W, BH, nbands, nq_pad, 1.0f / (float)C);
  count_launch();
  *rc = check_launch("cost_volume_bwd (staged)");
  return true;
}

bool bwd_small(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
               long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, cudaStream_t stream, int* rc) {
  // Measured (B200, kernel_bench, round 2): 18.6 us at 16x196x4x13 against 25.2 us for the tiled kernels, but 24.1 vs 20.3 us
  // at 16x128x8x26 (128 CTAs each staging 146 KB through 4-byte cp.async: the staging, not the 81-tap loop, is the
  // time) and 51 vs 24 us at 16x52 in bands of 4 rows (gout re-staged once per channel chunk).  Default: images of at most
  // kStagedQuads quads, i.e. the 4x13 level; UOF_CV_BWD_SMALL_MAXQ raises the limit (the larger forms stay parity-tested),
  // UOF_CV_NO_SMALL_BWD=1 keeps the tiled kernels everywhere.
  constexpr int kStagedQuads = 16;
  static const bool off = getenv("UOF_CV_NO_SMALL_BWD") != nullptr;
  static const int maxq = getenv("UOF_CV_BWD_SMALL_MAXQ") ? atoi(getenv("UOF_CV_BWD_SMALL_MAXQ")) : kStagedQuads;
  const int qpr = ceil_div(W, PX);
  const long long nq_img = (long long)qpr * H;
  if (off || nq_img > maxq || qpr > 16 || B > 65535) return false;      // W <= 64: kColIters column passes per staged row
  // band height: the whole image when it has <= 64 quads, else the most rows that give <= 64 quads
  int BH = H;
  while ((long long)BH * qpr > 64 && BH > 1) BH = (BH + 1) / 2;
  const int nq = BH * qpr;
  int nq_pad = 16;
  while (nq_pad < nq) nq_pad *= 2;
  // few quads (many thread groups): 4 channels per thread; a full 64-quad band: 8, which halves the re-reads of gout
  if (nq_pad >= 64) return launch_staged<8>(gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, B, C, H, W, BH, nq_pad, stream, rc);
  return launch_staged<4>(gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1, gf2, B, C, H, W, BH, nq_pad, stream, rc);
}

}  // namespace cv
}  // namespace uof
