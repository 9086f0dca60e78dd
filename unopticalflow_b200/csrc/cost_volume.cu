// a1: PWC-Net 81-displacement cost volume, forward and backward (fp32, NCHW).
// Replaces PWC_tf.corr_naive, /root/reference/core/networks/structures/pwc_tf.py:97-106
// (164 ATen launches per call forward, ~4x that backward) with one launch each way.
//
// Design (see DESIGN.md "cost volume"):
//   CTA tile = 8x32 pixels; the matching (8+8)x(32+8) halo tile of the second feature map is staged
//   through shared memory in chunks of CK channels with a 2-stage cp.async (LDGSTS, zero-fill)
//   pipeline, so every feature element is read from HBM once instead of 81 times.
//   192 threads = 64 pixel quads x 3 displacement-row groups; each thread keeps a 4-pixel x
//   (3 dy x 9 dx) block of accumulators in registers (108 FFMA per 40 floats read from smem)
//   and slides a 12-float register window along dx.
//   Small pyramid levels split the channel loop over extra CTAs (fp32 atomics into a zeroed
//   output) so that the grid still covers the 148 SMs.
#include "cost_volume.h"

namespace uof {
namespace {

constexpr int RAD = 4;                 // max displacement (pwc_tf.py:16 md=4)
constexpr int ND = 2 * RAD + 1;        // 9
constexpr int TH = 8, TW = 32;         // pixel tile
constexpr int HTH = TH + 2 * RAD;      // 16 halo rows
constexpr int HTW = TW + 2 * RAD;      // 40 halo cols
constexpr int CK = 8;                  // channels per pipeline stage
constexpr int PX = 4;                  // pixels per thread along x
constexpr int DYG = 3;                 // displacement rows per thread
constexpr int NGROUP = ND / DYG;       // 3 thread groups over dy
constexpr int QUADS = (TW / PX) * TH;  // 64
constexpr int NT = QUADS * NGROUP;     // 192 threads
constexpr int S1 = CK * TH * TW;       // floats of the f1 tile per stage
constexpr int S2 = CK * HTH * HTW;     // floats of the f2 halo tile per stage

// Stage a CK-channel slab of a [rows x cols] window whose top-left image coordinate is (ys, xs);
// everything outside the image or beyond channel C is zero-filled by the async copy itself.
template <bool VEC4, int ROWS, int COLS>
__device__ __forceinline__ void stage_tile(float* __restrict__ dst, const float* __restrict__ src_b,
                                           int c0, int C, int H, int W, int ys, int xs, int tid) {
  if (VEC4) {
    constexpr int CV = COLS / 4;
    for (int e = tid; e < CK * ROWS * CV; e += NT) {
      int cv = e % CV, r = (e / CV) % ROWS, cc = e / (CV * ROWS);
      int c = c0 + cc, y = ys + r, x = xs + 4 * cv;
      bool ok = (c < C) && (y >= 0) && (y < H) && (x >= 0) && (x < W);
      const float* g = ok ? src_b + ((size_t)c * H + y) * W + x : src_b;
      cp_async_16(dst + (cc * ROWS + r) * COLS + 4 * cv, g, ok);
    }
  } else {
    for (int e = tid; e < CK * ROWS * COLS; e += NT) {
      int col = e % COLS, r = (e / COLS) % ROWS, cc = e / (COLS * ROWS);
      int c = c0 + cc, y = ys + r, x = xs + col;
      bool ok = (c < C) && (y >= 0) && (y < H) && (x >= 0) && (x < W);
      const float* g = ok ? src_b + ((size_t)c * H + y) * W + x : src_b;
      cp_async_4(dst + (cc * ROWS + r) * COLS + col, g, ok);
    }
  }
}

// ------------------------------------------------------------------------------------ forward
template <bool VEC4>
__global__ void __launch_bounds__(NT, 2)
cost_volume_fwd_kernel(const float* __restrict__ f1, long long f1_bs, const float* __restrict__ f2,
                       float* __restrict__ out, int C, int H, int W, long long out_bs, int ksplit, float inv_c) {
  extern __shared__ __align__(16) float smem[];
  constexpr int kStage = S1 + S2;      // stage k: f1 tile at smem + k*kStage, f2 halo tile S1 floats later

  const int tid = threadIdx.x;
  const int gx = tid % (TW / PX), ty = (tid / (TW / PX)) % TH, dg = tid / QUADS;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int b = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  const int nchunks = (C + CK - 1) / CK;
  const int per = (nchunks + ksplit - 1) / ksplit;
  const int k_begin = ks * per, k_end = min(nchunks, k_begin + per);
  if (k_begin >= k_end) return;

  const float* f1b = f1 + (size_t)b * f1_bs;
  const float* f2b = f2 + (size_t)b * C * H * W;

  float acc[DYG][ND][PX];
#pragma unroll
  for (int r = 0; r < DYG; ++r)
#pragma unroll
    for (int j = 0; j < ND; ++j)
#pragma unroll
      for (int p = 0; p < PX; ++p) acc[r][j][p] = 0.0f;

  stage_tile<VEC4, TH, TW>(smem, f1b, k_begin * CK, C, H, W, y0, x0, tid);
  stage_tile<VEC4, HTH, HTW>(smem + S1, f2b, k_begin * CK, C, H, W, y0 - RAD, x0 - RAD, tid);
  cp_async_commit();

  for (int k = k_begin; k < k_end; ++k) {
    const int cur = (k - k_begin) & 1;
    if (k + 1 < k_end) {
      stage_tile<VEC4, TH, TW>(smem + (cur ^ 1) * kStage, f1b, (k + 1) * CK, C, H, W, y0, x0, tid);
      stage_tile<VEC4, HTH, HTW>(smem + (cur ^ 1) * kStage + S1, f2b, (k + 1) * CK, C, H, W, y0 - RAD, x0 - RAD, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* a_base = smem + cur * kStage + ty * TW + PX * gx;
    const float* w_base = smem + cur * kStage + S1 + (ty + dg * DYG) * HTW + PX * gx;
#pragma unroll 2
    for (int cc = 0; cc < CK; ++cc) {
      const float4 a4 = *reinterpret_cast<const float4*>(a_base + cc * TH * TW);
      const float a[PX] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int r = 0; r < DYG; ++r) {
        const float4* wp = reinterpret_cast<const float4*>(w_base + cc * HTH * HTW + r * HTW);
        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
          for (int p = 0; p < PX; ++p) acc[r][j][p] = fmaf(a[p], win[p + j], acc[r][j][p]);
      }
    }
    __syncthreads();
  }

  const int y = y0 + ty, x = x0 + PX * gx;
  if (y >= H || x >= W) return;
  float* ob = out + (size_t)b * out_bs + (size_t)y * W + x;
  const size_t plane = (size_t)H * W;
#pragma unroll
  for (int r = 0; r < DYG; ++r) {
#pragma unroll
    for (int j = 0; j < ND; ++j) {
      float* o = ob + (size_t)((dg * DYG + r) * ND + j) * plane;
      if (ksplit == 1) {
        if (VEC4) {
          *reinterpret_cast<float4*>(o) =
              make_float4(acc[r][j][0] * inv_c, acc[r][j][1] * inv_c, acc[r][j][2] * inv_c, acc[r][j][3] * inv_c);
        } else {
#pragma unroll
          for (int p = 0; p < PX; ++p)
            if (x + p < W) o[p] = acc[r][j][p] * inv_c;
        }
      } else {
#pragma unroll
        for (int p = 0; p < PX; ++p)
          if (x + p < W) atomicAdd(o + p, acc[r][j][p] * inv_c);
      }
    }
  }
}

// ----------------------------------------------------------------------------------- backward
// One kernel, two roles (blockIdx.y selects):
//   role 0: gf1[c,p] = 1/C * sum_d gout[d,p]          * f2[c,p+d]
//   role 1: gf2[c,q] = 1/C * sum_d gout[flip(d),q+d]  * f1[c,q+d]     (d -> -d re-indexed as flip)
// Per thread the 4 x 27 coefficients live in registers for the whole channel loop; the three
// displacement-row groups are reduced through shared memory once per CK-channel chunk.
template <bool VEC4>
__global__ void __launch_bounds__(NT, 2)
cost_volume_bwd_kernel(const float* __restrict__ gout, long long gout_bs, const float* __restrict__ f1, long long f1_bs,
                       const float* __restrict__ f2, const float* __restrict__ gadd, long long gadd_bs,
                       float* __restrict__ gf1, float* __restrict__ gf2, int C, int H, int W, int tiles_x, int csplit,
                       float inv_c) {
  extern __shared__ __align__(16) float smem[];
  float* red = smem + 2 * S2;          // [NGROUP][CK][TH][TW]; halo stage k lives at smem + k*S2

  const int tid = threadIdx.x;
  const int gx = tid % (TW / PX), ty = (tid / (TW / PX)) % TH, dg = tid / QUADS;
  const int x0 = (blockIdx.x % tiles_x) * TW, y0 = (blockIdx.x / tiles_x) * TH;
  const bool mirror = blockIdx.y != 0;
  const int b = blockIdx.z / csplit, cs = blockIdx.z % csplit;
  const int nchunks = (C + CK - 1) / CK;
  const int per = (nchunks + csplit - 1) / csplit;
  const int k_begin = cs * per, k_end = min(nchunks, k_begin + per);
  if (k_begin >= k_end) return;

  const float* src_b = mirror ? f1 + (size_t)b * f1_bs : f2 + (size_t)b * C * H * W;
  float* dst_b = (mirror ? gf2 : gf1) + (size_t)b * C * H * W;
  const float* gb = gout + (size_t)b * gout_bs;
  const float* add_b = (!mirror && gadd) ? gadd + (size_t)b * gadd_bs : nullptr;   // concat-slice gradient folded into gf1
  const size_t plane = (size_t)H * W;
  const int y = y0 + ty, x = x0 + PX * gx;

  // prefetch the first chunk while the coefficients are gathered
  stage_tile<VEC4, HTH, HTW>(smem, src_b, k_begin * CK, C, H, W, y0 - RAD, x0 - RAD, tid);
  cp_async_commit();

  float kc[DYG][ND][PX];
#pragma unroll
  for (int r = 0; r < DYG; ++r) {
    const int i = dg * DYG + r;
#pragma unroll
    for (int j = 0; j < ND; ++j) {
#pragma unroll
      for (int p = 0; p < PX; ++p) {
        float v = 0.0f;
        if (!mirror) {
          if (y < H && x + p < W) v = __ldg(gb + (size_t)(i * ND + j) * plane + (size_t)y * W + x + p);
        } else {
          const int ys = y + i - RAD, xs = x + p + j - RAD;
          if (y < H && x + p < W && ys >= 0 && ys < H && xs >= 0 && xs < W)
            v = __ldg(gb + (size_t)((ND - 1 - i) * ND + (ND - 1 - j)) * plane + (size_t)ys * W + xs);
        }
        kc[r][j][p] = v;
      }
    }
  }

  for (int k = k_begin; k < k_end; ++k) {
    const int cur = (k - k_begin) & 1;
    if (k + 1 < k_end) {
      stage_tile<VEC4, HTH, HTW>(smem + (cur ^ 1) * S2, src_b, (k + 1) * CK, C, H, W, y0 - RAD, x0 - RAD, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // stage `cur` landed; previous iteration's reads of `red` are complete

    const float* w_base = smem + cur * S2 + (ty + dg * DYG) * HTW + PX * gx;
#pragma unroll
    for (int cc = 0; cc < CK; ++cc) {
      float part[PX] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int r = 0; r < DYG; ++r) {
        const float4* wp = reinterpret_cast<const float4*>(w_base + cc * HTH * HTW + r * HTW);
        const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const float win[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
          for (int p = 0; p < PX; ++p) part[p] = fmaf(kc[r][j][p], win[p + j], part[p]);
      }
      *reinterpret_cast<float4*>(red + ((dg * CK + cc) * TH + ty) * TW + PX * gx) =
          make_float4(part[0], part[1], part[2], part[3]);
    }
    __syncthreads();

    // cross-group reduction + store: CK*TH*TW/4 float4 results per chunk
    for (int e = tid; e < CK * TH * (TW / 4); e += NT) {
      const int q = e % (TW / 4), r = (e / (TW / 4)) % TH, cc = e / ((TW / 4) * TH);
      const int c = k * CK + cc, yy = y0 + r, xx = x0 + 4 * q;
      if (c >= C || yy >= H || xx >= W) continue;
      const float* rp = red + (cc * TH + r) * TW + 4 * q;
      float4 s = *reinterpret_cast<const float4*>(rp);
#pragma unroll
      for (int g = 1; g < NGROUP; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(rp + g * CK * TH * TW);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      float* o = dst_b + ((size_t)c * H + yy) * W + xx;
      float v[4] = {s.x * inv_c, s.y * inv_c, s.z * inv_c, s.w * inv_c};
      if (add_b) {
        const float* ap = add_b + ((size_t)c * H + yy) * W + xx;
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (xx + p < W) v[p] += __ldg(ap + p);
      }
      if (VEC4) {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (xx + p < W) o[p] = v[p];
      }
    }
    // next iteration's first __syncthreads orders these reads of `red` before it is rewritten
  }
}

// The TMA + mbarrier variants of both kernels live in cost_volume_tma.cu; the cp.async kernels above are the
// fallback for W % 4 != 0 (levels 8x26 and 4x13 of the 256x832 pyramid) and for UOF_DISABLE_TMA=1.
constexpr size_t kFwdSmem = 2 * (S1 + S2) * sizeof(float);                       // 57344
constexpr size_t kBwdSmem = (2 * S2 + NGROUP * CK * TH * TW) * sizeof(float);    // 65536

bool vec4_ok(const void* a, const void* b, const void* c, int W, long long bs) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return (W % 4 == 0) && (bs % 4 == 0) && al(a) && al(b) && al(c);
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_cost_volume_fwd_ex(const float* f1, long long f1_batch_stride, const float* f2, float* out, int B, int C,
                                      int H, int W, long long out_batch_stride, uof_stream_t stream_) {
  UOF_REQUIRE(f1 && f2 && out, "cost_volume_fwd: null pointer");
  UOF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cost_volume_fwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  UOF_REQUIRE(out_batch_stride >= (long long)UOF_NUM_DISPLACEMENTS * H * W, "cost_volume_fwd: out_batch_stride too small");
  UOF_REQUIRE(f1_batch_stride >= (long long)C * H * W, "cost_volume_fwd: f1_batch_stride too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = 0;
  if (cv::fwd_small(f1, f1_batch_stride, f2, out, B, C, H, W, out_batch_stride, stream, &rc)) return rc;
  if (cv::fwd_tma(f1, f1_batch_stride, f2, out, B, C, H, W, out_batch_stride, stream, &rc)) return rc;
  const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
  const int nchunks = ceil_div(C, CK);
  const int ksplit = cv::pick_ksplit_atomic((long long)tx * ty * B, nchunks);
  UOF_REQUIRE((long long)B * ksplit <= 65535, "cost_volume_fwd: batch too large for one launch");
  if (ksplit > 1) {
    UOF_CUDA(cudaMemset2DAsync(out, out_batch_stride * sizeof(float), 0,
                               (size_t)UOF_NUM_DISPLACEMENTS * H * W * sizeof(float), B, stream));
  }
  dim3 grid(tx, ty, B * ksplit);
  const float inv_c = 1.0f / (float)C;
  const bool v4 = vec4_ok(f1, f2, out, W, out_batch_stride) && f1_batch_stride % 4 == 0;
  auto kern = v4 ? cost_volume_fwd_kernel<true> : cost_volume_fwd_kernel<false>;
  UOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem));
  kern<<<grid, NT, kFwdSmem, stream>>>(f1, f1_batch_stride, f2, out, C, H, W, out_batch_stride, ksplit, inv_c);
  count_launch();
  return check_launch("cost_volume_fwd");
}

extern "C" int uof_cost_volume_fwd(const float* f1, const float* f2, float* out, int B, int C, int H, int W,
                                   long long out_batch_stride, uof_stream_t stream) {
  return uof_cost_volume_fwd_ex(f1, (long long)C * H * W, f2, out, B, C, H, W, out_batch_stride, stream);
}

extern "C" int uof_cost_volume_bwd_ex(const float* gout, long long gout_batch_stride, const float* f1, long long f1_batch_stride,
                                      const float* f2, const float* gadd, long long gadd_batch_stride, float* gf1, float* gf2,
                                      int B, int C, int H, int W, uof_stream_t stream_) {
  UOF_REQUIRE(gout && f1 && f2 && gf1 && gf2, "cost_volume_bwd: null pointer");
  UOF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cost_volume_bwd: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  UOF_REQUIRE(gout_batch_stride >= (long long)UOF_NUM_DISPLACEMENTS * H * W, "cost_volume_bwd: gout_batch_stride too small");
  UOF_REQUIRE(f1_batch_stride >= (long long)C * H * W, "cost_volume_bwd: f1_batch_stride too small");
  UOF_REQUIRE(!gadd || gadd_batch_stride >= (long long)C * H * W, "cost_volume_bwd: gadd_batch_stride too small");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = 0;
  const bool plain = f1_batch_stride == (long long)C * H * W && !gadd;
  if (plain && cv::bwd_tc(gout, gout_batch_stride, f1, f2, gf1, gf2, B, C, H, W, stream, &rc)) return rc;      // tcgen05 banded GEMM (opt-in)
  if (cv::bwd_small(gout, gout_batch_stride, f1, f1_batch_stride, f2, gadd, gadd_batch_stride, gf1, gf2, B, C, H, W, stream, &rc))
    return rc;
  if (cv::bwd_tma(gout, gout_batch_stride, f1, f1_batch_stride, f2, gadd, gadd_batch_stride, gf1, gf2, B, C, H, W, stream, &rc))
    return rc;
  const int tx = ceil_div(W, TW), ty = ceil_div(H, TH);
  const int nchunks = ceil_div(C, CK);
  const int csplit = cv::pick_split((long long)tx * ty * B * 2, nchunks);
  UOF_REQUIRE((long long)B * csplit <= 65535, "cost_volume_bwd: batch too large for one launch");
  dim3 grid(tx * ty, 2, B * csplit);
  const float inv_c = 1.0f / (float)C;
  const bool v4 = vec4_ok(f1, f2, gf1, W, gout_batch_stride) && vec4_ok(gf2, gout, gf1, W, f1_batch_stride);
  auto kern = v4 ? cost_volume_bwd_kernel<true> : cost_volume_bwd_kernel<false>;
  UOF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
  kern<<<grid, NT, kBwdSmem, stream>>>(gout, gout_batch_stride, f1, f1_batch_stride, f2, gadd, gadd_batch_stride, gf1, gf2, C, H, W,
                                       tx, csplit, inv_c);
  count_launch();
  return check_launch("cost_volume_bwd");
}

extern "C" int uof_cost_volume_bwd(const float* gout, long long gout_batch_stride, const float* f1, const float* f2,
                                   float* gf1, float* gf2, int B, int C, int H, int W, uof_stream_t stream) {
  return uof_cost_volume_bwd_ex(gout, gout_batch_stride, f1, (long long)C * H * W, f2, nullptr, 0, gf1, gf2, B, C, H, W, stream);
}
