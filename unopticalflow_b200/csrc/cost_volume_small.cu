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
// cp.async pipeline fill and two CTA barriers per 8-channel slab for a few MFLOP: 20-25 us per call (round 1).
// Here the parallelism comes from the channels: one CTA per (image, role, chunk of G * CC channels), a thread owns a quad
// of 4 adjacent pixels and CC channels.  For each displacement row it keeps the 9 x 4 coefficients of its quad in
// registers (role 0: gout at the pixel itself; role 1: the index-flipped gout at the displaced pixel, see
// cost_volume.cu) and slides the 12-float window of each of its channels over them: 36 + 12 CC independent loads
// (L1/L2 hits: an image's whole working set is < 100 KB) and 36 CC FMAs per row, no shared memory, no barrier, no
// atomics; every output element is written once, with the concat-slice gradient `gadd` folded in.
template <int CC>
__global__ void __launch_bounds__(256)
cost_volume_bwd_small_kernel(const float* __restrict__ gout, long long gout_bs, const float* __restrict__ f1, long long f1_bs,
                             const float* __restrict__ f2, const float* __restrict__ gadd, long long gadd_bs,
                             float* __restrict__ gf1, float* __restrict__ gf2, int C, int H, int W, int qpr, int nq_pad,
                             int G, float inv_c) {
  const bool mirror = blockIdx.y != 0;
  const int b = blockIdx.z;
  const int q = threadIdx.x % nq_pad, g = threadIdx.x / nq_pad;
  const int y = q / qpr, x = (q - y * qpr) * PX;
  const int c0 = ((int)blockIdx.x * G + g) * CC;
  if (y >= H || c0 >= C) return;
  const size_t plane = (size_t)H * W;
  const float* gb = gout + (size_t)b * gout_bs;
  const float* src = (mirror ? f1 + (size_t)b * f1_bs : f2 + (size_t)b * C * plane) + (size_t)c0 * plane;

  bool pok[PX], wok[PX + ND - 1];
#pragma unroll
  for (int p = 0; p < PX; ++p) pok[p] = x + p < W;
#pragma unroll
  for (int k = 0; k < PX + ND - 1; ++k) wok[k] = x + k - RAD >= 0 && x + k - RAD < W;

  float acc[CC][PX];
#pragma unroll
  for (int cc = 0; cc < CC; ++cc)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[cc][p] = 0.0f;

  for (int i = 0; i < ND; ++i) {
    const int ys = y + i - RAD;                    // row of the other feature map (and, mirrored, of gout)
    if (ys < 0 || ys >= H) continue;
    float kc[ND][PX];
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      const float* gp = mirror ? gb + (size_t)((ND - 1 - i) * ND + (ND - 1 - j)) * plane + (size_t)ys * W + x + j - RAD
                               : gb + (size_t)(i * ND + j) * plane + (size_t)y * W + x;
#pragma unroll
      for (int p = 0; p < PX; ++p) kc[j][p] = (pok[p] && (!mirror || wok[p + j])) ? __ldg(gp + p) : 0.0f;
    }
    const float* wrow = src + (size_t)ys * W + x - RAD;
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      if (c0 + cc < C) {
        float win[PX + ND - 1];
#pragma unroll
        for (int k = 0; k < PX + ND - 1; ++k) win[k] = wok[k] ? __ldg(wrow + (size_t)cc * plane + k) : 0.0f;
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
          for (int p = 0; p < PX; ++p) acc[cc][p] = fmaf(kc[j][p], win[p + j], acc[cc][p]);
      }
    }
  }

  float* dst = (mirror ? gf2 : gf1) + ((size_t)b * C + c0) * plane + (size_t)y * W + x;
  const float* add = (!mirror && gadd) ? gadd + (size_t)b * gadd_bs + (size_t)c0 * plane + (size_t)y * W + x : nullptr;
#pragma unroll
  for (int cc = 0; cc < CC; ++cc) {
    if (c0 + cc >= C) break;
#pragma unroll
    for (int p = 0; p < PX; ++p)
      if (pok[p]) dst[(size_t)cc * plane + p] = fmaf(acc[cc][p], inv_c, add ? __ldg(add + (size_t)cc * plane + p) : 0.0f);
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

bool bwd_small(const float* gout, long long gout_bs, const float* f1, long long f1_bs, const float* f2, const float* gadd,
               long long gadd_bs, float* gf1, float* gf2, int B, int C, int H, int W, cudaStream_t stream, int* rc) {
  // Measured (B200, gpurun_out/r2c): 31.5 us at 16x196x4x13 and 39.2 us at 16x128x8x26 against 25.5 / 20.6 us for the tiled
  // kernels -- a lane's four consecutive pixels make every scalar load of the warp span 512 B (16 sectors per request,
  // W % 4 != 0 rules out vector loads), so the kernel is LSU-bound exactly like the forward variant at 16x52.  Kept
  // opt-in (UOF_CV_SMALL_BWD=1) and parity-tested; the tiled kernels stay the default for these levels.
  static const bool off = getenv("UOF_CV_SMALL_BWD") == nullptr;
  static const int maxq = getenv("UOF_CV_BWD_SMALL_MAXQ") ? atoi(getenv("UOF_CV_BWD_SMALL_MAXQ")) : kMaxQuads;
  const int qpr = ceil_div(W, PX);
  const long long nq = (long long)qpr * H;
  if (off || nq > maxq || nq > 256 || B > 65535) return false;
  constexpr int CC = 4;
  const int nq_pad = ceil_div((int)nq, 32) * 32;
  const int G = 256 / nq_pad < 1 ? 1 : 256 / nq_pad;
  const int chunks = ceil_div(C, G * CC);
  cost_volume_bwd_small_kernel<CC><<<dim3(chunks, 2, B), nq_pad * G, 0, stream>>>(gout, gout_bs, f1, f1_bs, f2, gadd, gadd_bs, gf1,
                                                                               gf2, C, H, W, qpr, nq_pad, G, 1.0f / (float)C);
  count_launch();
  *rc = check_launch("cost_volume_bwd (small)");
  return true;
}

}  // namespace cv
}  // namespace uof
