// the data formats either side of the hot path -- input pre-processing and flow evaluation.
// Replaces, in /root/reference:
//   core/dataset/kitti_prepared.py:63-91,146-153   resize_img (3 x cv2.resize) + random_flip_img (cv2.flip) + /255.0 +
//                                                  transpose(2,0,1) + .float(): per sample ~15 ms of CPU work and a 7.7 MB
//                                                  fp32 host->device copy; here the decoded uint8 strip (4.2 MB at KITTI size)
//                                                  is copied and ONE kernel produces the (B,3,3H,W) fp32 batch;
//   core/evaluation/flowlib.py:107-138             KITTI 16-bit flow PNG arithmetic (decode / encode);
//   core/evaluation/evaluate_flow.py:85-160        rescale + cv2.resize of the prediction, EPE map, masked means and the
//                                                  3 px / 5 % outlier rate -- one kernel per image, no HBM temporaries.
// cv2.resize(INTER_LINEAR) is OpenCV's algorithm (resize.cpp), restated in oracle/io_ops.py and pinned bit-exactly against
// the installed cv2: 8-bit images use 11-bit fixed-point coefficients and the rounding
// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2; float images two unfused multiply-adds.
// All of it is integer / streaming work: HBM-bound by design, nothing here wants tensor cores.
#include "common.cuh"

namespace uof {
namespace {

// Source index pair and fraction of cv2's bilinear resize along one axis: f = (float)((d + 0.5) * src/dst - 0.5).
// Along x an index outside the row zeroes the fraction; along y the fraction is kept and the row indices are clamped.
struct Tap {
  int i0, i1;
  float f;
};
__device__ __forceinline__ Tap cv_tap(int d, double scale, int src, bool vertical) {
  Tap t;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (!vertical) {
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= src - 1) { f = 0.0f; s = src - 1; }
  }
  t.i0 = min(max(s, 0), src - 1);
  t.i1 = min(max(s + 1, 0), src - 1);
  t.f = f;
  return t;
}

__constant__ float kInv255[256];        // (float)(v / 255.0): the reference divides in float64 and rounds once

// grid = (x chunks, nimg * H, B); a thread produces one output pixel (3 channels) of one image of the strip.
__global__ void __launch_bounds__(128)
preprocess_u8_kernel(const unsigned char* __restrict__ src, long long src_bs, const unsigned char* __restrict__ flip,
                     float* __restrict__ out, int nimg, int H0, int W0, int H, int W, double sy, double sx) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const int row = blockIdx.y, b = blockIdx.z;
  const int k = row / H, y = row - k * H;                     // image of the strip, row inside it
  const Tap ty = cv_tap(y, sy, H0, true), tx = cv_tap(x, sx, W0, false);
  const int ax1 = __float2int_rn(tx.f * 2048.0f), ax0 = __float2int_rn((1.0f - tx.f) * 2048.0f);     // saturate_cast<short>
  const int by1 = __float2int_rn(ty.f * 2048.0f), by0 = __float2int_rn((1.0f - ty.f) * 2048.0f);
  const unsigned char* s = src + (size_t)b * src_bs + (size_t)k * H0 * W0 * 3;
  const unsigned char* r0 = s + (size_t)ty.i0 * W0 * 3;
  const unsigned char* r1 = s + (size_t)ty.i1 * W0 * 3;
  const int xo = (flip && flip[b]) ? W - 1 - x : x;           // cv2.flip(img, 1) of the re-stacked strip
  const size_t plane = (size_t)nimg * H * W;
  float* o = out + (size_t)b * 3 * plane + (size_t)row * W + xo;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = (int)__ldg(r0 + tx.i0 * 3 + c) * ax0 + (int)__ldg(r0 + tx.i1 * 3 + c) * ax1;       // HResizeLinear, x 2048
    const int h1 = (int)__ldg(r1 + tx.i0 * 3 + c) * ax0 + (int)__ldg(r1 + tx.i1 * 3 + c) * ax1;
    const int v = ((((h0 >> 4) * by0) >> 16) + (((h1 >> 4) * by1) >> 16) + 2) >> 2;                   // VResizeLinear<uchar>
    o[(size_t)c * plane] = kInv255[min(max(v, 0), 255)];
  }
}

// flowlib.py:119-127 on the decoded (H,W,3) uint16 array
__global__ void __launch_bounds__(256)
flow_png_decode_kernel(const unsigned short* __restrict__ raw, float* __restrict__ flow, long long npix) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix) return;
  const unsigned short u = raw[3 * t], v = raw[3 * t + 1], ok = raw[3 * t + 2];
  flow[3 * t] = ok ? ((float)u - 32768.0f) * (1.0f / 64.0f) : 0.0f;       // exact in fp32
  flow[3 * t + 1] = ok ? ((float)v - 32768.0f) * (1.0f / 64.0f) : 0.0f;
  flow[3 * t + 2] = (float)ok;
}

// flowlib.py:131-138: fp32 clip(flo * 64 + 2^15, 0, 2^16 - 1) truncated to uint16; third channel 1
__global__ void __launch_bounds__(256)
flow_png_encode_kernel(const float* __restrict__ flo, int cstride, unsigned short* __restrict__ raw, long long npix) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npix) return;
  const float u = fmaxf(fminf(__fadd_rn(__fmul_rn(flo[(size_t)cstride * t], 64.0f), 32768.0f), 65535.0f), 0.0f);
  const float v = fmaxf(fminf(__fadd_rn(__fmul_rn(flo[(size_t)cstride * t + 1], 64.0f), 32768.0f), 65535.0f), 0.0f);
  raw[3 * t] = (unsigned short)u;
  raw[3 * t + 1] = (unsigned short)v;
  raw[3 * t + 2] = 1;
}

constexpr int kEvalSums = 13;

// evaluate_flow.py:104-160 for one image.  pred: (2,h,w) planar fp32 at network resolution; gt: (H,W,3) [u,v,valid];
// noc / move: (H,W).  A thread owns one ground-truth pixel; the 13 masked sums are reduced warp -> block -> one fp64
// atomic per block and quantity.
__global__ void __launch_bounds__(256)
flow_eval_kernel(const float* __restrict__ pred, int h, int w, const float* __restrict__ gt, const float* __restrict__ noc,
                 const float* __restrict__ move, int H, int W, float img_h, float img_w, double sy, double sx,
                 double* __restrict__ sums) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  double acc[kEvalSums];
#pragma unroll
  for (int k = 0; k < kEvalSums; ++k) acc[k] = 0.0;
  if (x < W) {
    const Tap ty = cv_tap(y, sy, h, true), tx = cv_tap(x, sx, w, false);
    const float ux = 1.0f - tx.f, uy = 1.0f - ty.f;
    float flo[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float* p = pred + (size_t)c * h * w;
      const float den = c ? img_h : img_w, num = c ? (float)H : (float)W;
      auto at = [&](int yy, int xx) { return __fmul_rn(__fdiv_rn(__ldg(p + (size_t)yy * w + xx), den), num); };   // :107-108
      const float r0 = __fadd_rn(__fmul_rn(at(ty.i0, tx.i0), ux), __fmul_rn(at(ty.i0, tx.i1), tx.f));             // HResizeLinear<float>
      const float r1 = __fadd_rn(__fmul_rn(at(ty.i1, tx.i0), ux), __fmul_rn(at(ty.i1, tx.i1), tx.f));
      flo[c] = __fadd_rn(__fmul_rn(r0, uy), __fmul_rn(r1, ty.f));                                                 // VResizeLinear<float>
    }
    const size_t pix = (size_t)y * W + x;
    const double gu = gt[3 * pix], gv = gt[3 * pix + 1], valid = gt[3 * pix + 2];
    const double du = (double)flo[0] - gu, dv = (double)flo[1] - gv;
    const double epe = sqrt(du * du + dv * dv);                                                                   // :132-134
    const double nc = noc[pix];
    const double mag = fmax(sqrt(gu * gu + gv * gv), 1e-10);
    auto bad = [&](double m) { return (epe * m > 3.0 && epe * m / mag > 0.05) ? 1.0 : 0.0; };                      // :85-90
    acc[0] = epe * valid;
    acc[1] = valid;
    acc[2] = epe * nc;
    acc[3] = nc;
    acc[4] = epe * (valid - nc);
    acc[5] = valid - nc;
    acc[6] = bad(valid);
    if (move) {
      const double mv = move[pix];
      acc[7] = bad(valid * mv);
      acc[8] = valid * mv;
      acc[9] = bad(valid * (1.0 - mv));
      acc[10] = valid * (1.0 - mv);
      acc[11] = epe * valid * mv;
      acc[12] = epe * valid * (1.0 - mv);
    }
  }
  __shared__ double part[8][kEvalSums];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kEvalSums; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    if (lane == 0) part[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kEvalSums) {
    double v = 0.0;
    for (int j = 0; j < (int)(blockDim.x >> 5); ++j) v += part[j][threadIdx.x];
    if (v != 0.0) atomicAdd(sums + threadIdx.x, v);
  }
}

bool g_lut_ready = false;

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_preprocess_u8(const unsigned char* src, long long src_batch_stride, const unsigned char* flip, float* out,
                                 int B, int nimg, int H0, int W0, int H, int W, uof_stream_t stream_) {
  UOF_REQUIRE(src && out, "preprocess_u8: null pointer");
  UOF_REQUIRE(B > 0 && nimg > 0 && H0 > 0 && W0 > 0 && H > 0 && W > 0, "preprocess_u8: bad shape B=%d nimg=%d %dx%d -> %dx%d", B,
              nimg, H0, W0, H, W);
  UOF_REQUIRE(src_batch_stride >= (long long)nimg * H0 * W0 * 3, "preprocess_u8: src_batch_stride too small");
  UOF_REQUIRE(B <= 65535 && (long long)nimg * H <= 65535, "preprocess_u8: too many rows / samples for one launch");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!g_lut_ready) {       // one-time, outside any capture in practice (the first call of a process)
    float lut[256];
    for (int v = 0; v < 256; ++v) lut[v] = (float)((double)v / 255.0);
    UOF_CUDA(cudaMemcpyToSymbolAsync(kInv255, lut, sizeof(lut), 0, cudaMemcpyHostToDevice, stream));
    UOF_CUDA(cudaStreamSynchronize(stream));
    g_lut_ready = true;
  }
  preprocess_u8_kernel<<<dim3(ceil_div(W, 128), nimg * H, B), 128, 0, stream>>>(src, src_batch_stride, flip, out, nimg, H0, W0, H, W,
                                                                                (double)H0 / (double)H, (double)W0 / (double)W);
  count_launch();
  return check_launch("preprocess_u8");
}

extern "C" int uof_flow_png_decode(const unsigned short* raw, float* flow, long long npix, uof_stream_t stream_) {
  UOF_REQUIRE(raw && flow && npix > 0, "flow_png_decode: bad arguments");
  flow_png_decode_kernel<<<(unsigned)ceil_div_ll(npix, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(raw, flow, npix);
  count_launch();
  return check_launch("flow_png_decode");
}

extern "C" int uof_flow_png_encode(const float* flow, int channel_stride, unsigned short* raw, long long npix,
                                   uof_stream_t stream_) {
  UOF_REQUIRE(flow && raw && npix > 0 && channel_stride >= 2, "flow_png_encode: bad arguments");
  flow_png_encode_kernel<<<(unsigned)ceil_div_ll(npix, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(flow, channel_stride, raw,
                                                                                                          npix);
  count_launch();
  return check_launch("flow_png_encode");
}

extern "C" int uof_flow_eval(const float* pred, int h, int w, const float* gt, const float* noc_mask, const float* moving_mask,
                             int H, int W, int img_h, int img_w, double* sums, uof_stream_t stream_) {
  UOF_REQUIRE(pred && gt && noc_mask && sums, "flow_eval: null pointer");
  UOF_REQUIRE(h > 0 && w > 0 && H > 0 && W > 0 && img_h > 0 && img_w > 0 && H <= 65535, "flow_eval: bad shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UOF_CUDA(cudaMemsetAsync(sums, 0, kEvalSums * sizeof(double), stream));
  flow_eval_kernel<<<dim3(ceil_div(W, 256), H), 256, 0, stream>>>(pred, h, w, gt, noc_mask, moving_mask, H, W, (float)img_h,
                                                                  (float)img_w, (double)h / (double)H, (double)w / (double)W, sums);
  count_launch();
  return check_launch("flow_eval");
}
