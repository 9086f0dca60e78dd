// a9: loss image pyramid.  Replaces Model_flow.generate_img_pyramid,
// /root/reference/core/networks/model_flow_paper.py:54-60 (adaptive_avg_pool2d per level, no grad).
// One launch produces every level >= 1 of `nimg` vertically stacked images (the triplet (B,3,3H,W) is read
// in place through explicit strides, no split copy); level 0 equals the input.
//
// Fast path (H % 4 == 0, W % 4 == 0, levels <= 3): a thread owns a 4x4 input block, reads it with four
// coalesced float4 loads and emits the 2x2 level-1 means and the level-2 mean hierarchically -- the input is
// read exactly once.  General path: adaptive_avg_pool2d bin rule [floor(i*H/h), ceil((i+1)*H/h)), one thread
// per output element.
#include "common.cuh"

namespace uof {
namespace {

struct PyrParams {
  float* out[UOF_MAX_LEVELS];          // out[l] is (nimg, B, C, h_l, w_l) contiguous
  int h[UOF_MAX_LEVELS], w[UOF_MAX_LEVELS];
  long long begin[UOF_MAX_LEVELS + 1];
  int nout, nimg, B, C, H, W;
  long long si, sb, sc, sh;            // element strides: image, batch, channel, row
  float* out0;                         // optional dense (nimg, B, C, H, W) copy of level 0 (pow2 kernel only)
  int slot[4];                         // output slot of input image i (a permutation of 0..nimg-1)
};

__global__ void __launch_bounds__(256) pyramid_generic_kernel(const __grid_constant__ PyrParams P, const float* __restrict__ img) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.begin[P.nout]) return;
  int l = 0;
  while (l + 1 < P.nout && t >= P.begin[l + 1]) ++l;
  long long e = t - P.begin[l];
  const int h = P.h[l], w = P.w[l];
  const int j = (int)(e % w), i = (int)((e / w) % h);
  long long r = e / ((long long)w * h);
  const int c = (int)(r % P.C);
  r /= P.C;
  const int b = (int)(r % P.B), im = (int)(r / P.B);
  const int ys = (int)(((long long)i * P.H) / h), ye = (int)((((long long)i + 1) * P.H + h - 1) / h);
  const int xs = (int)(((long long)j * P.W) / w), xe = (int)((((long long)j + 1) * P.W + w - 1) / w);
  const float* src = img + im * P.si + b * P.sb + c * P.sc;
  float s = 0.0f;
  for (int y = ys; y < ye; ++y)
    for (int x = xs; x < xe; ++x) s += __ldg(src + y * P.sh + x);
  P.out[l][e] = s / (float)((ye - ys) * (xe - xs));
}

// thread = one 4x4 input block of one (image, batch, channel) plane; block = 32 x 4 such blocks, grid = (W/128, H/16,
// planes): only the plane index is decoded (block-uniform), the first version decoded a 64-bit linear thread index
__global__ void __launch_bounds__(128) pyramid_pow2_kernel(const __grid_constant__ PyrParams P, const float* __restrict__ img) {
  const int W4 = P.W >> 2, H4 = P.H >> 2;
  const int bx = blockIdx.x * 32 + (threadIdx.x & 31), by = blockIdx.y * 4 + (threadIdx.x >> 5);
  if (bx >= W4 || by >= H4) return;
  long long plane_id = blockIdx.z;                              // (im*B + b)*C + c
  const int c = (int)(blockIdx.z % (unsigned)P.C);
  const int b = (int)((blockIdx.z / (unsigned)P.C) % (unsigned)P.B), im = (int)(blockIdx.z / (unsigned)(P.C * P.B));
  const float* src = img + im * P.si + b * P.sb + c * P.sc + (long long)(4 * by) * P.sh + 4 * bx;
  float4 r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) r[k] = __ldg(reinterpret_cast<const float4*>(src + k * P.sh));
  // outputs are stacked by slot, not by input image: lets the caller order the triplet as [left; right; centre]
  plane_id = ((long long)P.slot[im] * P.B + b) * P.C + c;
  if (P.out0) {        // dense copy of level 0 in the same stacked order (the encoder's 3B batch, the [l;r] warp sources)
    float* o0 = P.out0 + plane_id * ((long long)P.H * P.W) + (long long)(4 * by) * P.W + 4 * bx;
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<float4*>(o0 + (long long)k * P.W) = r[k];
  }
  // level 1: 2x2 means, accumulated row-major like adaptive_avg_pool2d
  const float a00 = ((r[0].x + r[0].y) + r[1].x + r[1].y) * 0.25f, a01 = ((r[0].z + r[0].w) + r[1].z + r[1].w) * 0.25f;
  const float a10 = ((r[2].x + r[2].y) + r[3].x + r[3].y) * 0.25f, a11 = ((r[2].z + r[2].w) + r[3].z + r[3].w) * 0.25f;
  const int w1 = P.W >> 1, h1 = P.H >> 1;
  float* o1 = P.out[0] + plane_id * ((long long)h1 * w1) + (long long)(2 * by) * w1 + 2 * bx;
  *reinterpret_cast<float2*>(o1) = make_float2(a00, a01);
  *reinterpret_cast<float2*>(o1 + w1) = make_float2(a10, a11);
  if (P.nout > 1) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += ((r[k].x + r[k].y) + r[k].z) + r[k].w;
    P.out[1][plane_id * ((long long)H4 * W4) + (long long)by * W4 + bx] = s * 0.0625f;
  }
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_img_pyramid_stacked(const float* img, long long stride_img, long long stride_b, long long stride_c,
                                       long long stride_h, float* out0, const int* slot, float* const* outs, int nlevels,
                                       int nimg, int B, int C, int H, int W, uof_stream_t stream_) {
  UOF_REQUIRE(img && outs, "img_pyramid: null pointer");
  UOF_REQUIRE(nlevels >= 2 && nlevels <= UOF_MAX_LEVELS + 1, "img_pyramid: nlevels must be 2..%d", UOF_MAX_LEVELS + 1);
  UOF_REQUIRE(nimg > 0 && B > 0 && C > 0 && H > 0 && W > 0, "img_pyramid: bad shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PyrParams P;
  P.nout = nlevels - 1;
  P.nimg = nimg; P.B = B; P.C = C; P.H = H; P.W = W;
  P.si = stride_img; P.sb = stride_b; P.sc = stride_c; P.sh = stride_h;
  P.out0 = out0;
  UOF_REQUIRE(!slot || nimg <= 4, "img_pyramid: a slot permutation supports at most 4 images");
  for (int i = 0; i < 4; ++i) P.slot[i] = (slot && i < nimg) ? slot[i] : i;
  for (int i = 0; slot && i < nimg; ++i) UOF_REQUIRE(slot[i] >= 0 && slot[i] < nimg, "img_pyramid: bad slot %d", slot[i]);
  long long total = 0;
  for (int l = 0; l < P.nout; ++l) {
    const int s = l + 1;
    P.h[l] = (int)((double)H / (double)(1 << s));   // int(img_h / 2**s), model_flow_paper.py:58
    P.w[l] = (int)((double)W / (double)(1 << s));
    UOF_REQUIRE(P.h[l] > 0 && P.w[l] > 0 && outs[l], "img_pyramid: level %d is empty", s);
    P.out[l] = outs[l];
    P.begin[l] = total;
    total += (long long)nimg * B * C * P.h[l] * P.w[l];
  }
  P.begin[P.nout] = total;
  const bool strides_ok = (stride_img % 4 == 0) && (stride_b % 4 == 0) && (stride_c % 4 == 0) && (stride_h % 4 == 0) &&
                          (reinterpret_cast<uintptr_t>(img) & 15u) == 0 && (reinterpret_cast<uintptr_t>(outs[0]) & 7u) == 0;
  const bool fast = H % 4 == 0 && W % 4 == 0 && P.nout <= 2 && strides_ok && (!out0 || (reinterpret_cast<uintptr_t>(out0) & 15u) == 0);
  UOF_REQUIRE(fast || (!out0 && !slot), "img_pyramid: the stacked level-0 copy / slot order need H, W %% 4 == 0, <= 3 levels and 16-byte aligned strides");
  if (fast) {
    UOF_REQUIRE((long long)nimg * B * C <= 65535 && H / 16 + 1 <= 65535, "img_pyramid: too many planes for one launch");
    pyramid_pow2_kernel<<<dim3(ceil_div(W / 4, 32), ceil_div(H / 4, 4), nimg * B * C), 128, 0, stream>>>(P, img);
  } else {
    pyramid_generic_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(P, img);
  }
  count_launch();
  return check_launch("img_pyramid");
}

extern "C" int uof_img_pyramid(const float* img, long long stride_img, long long stride_b, long long stride_c,
                               long long stride_h, float* const* outs, int nlevels, int nimg, int B, int C, int H, int W,
                               uof_stream_t stream_) {
  return uof_img_pyramid_stacked(img, stride_img, stride_b, stride_c, stride_h, nullptr, nullptr, outs, nlevels, nimg, B, C, H,
                                 W, stream_);
}
