// a9: loss image pyramid.  Replaces Model_flow.generate_img_pyramid,
// /root/reference/core/networks/model_flow_paper.py:54-60 (adaptive_avg_pool2d per level, no grad).
// One launch produces every level >= 1; level 0 equals the input.  Bin rule of adaptive_avg_pool2d:
// rows [floor(i*H/h), ceil((i+1)*H/h)), accumulated row-major then divided by the bin size.  The
// input is addressed with explicit strides so that each image of the vertically stacked triplet
// (B,3,3H,W) is read in place (no split copy).
#include "common.cuh"

namespace uof {
namespace {

struct PyrParams {
  float* out[UOF_MAX_LEVELS];
  int h[UOF_MAX_LEVELS], w[UOF_MAX_LEVELS];
  long long begin[UOF_MAX_LEVELS + 1];
  int nout, B, C, H, W;
  long long sb, sc, sh;
};

__global__ void __launch_bounds__(256) pyramid_kernel(const __grid_constant__ PyrParams P, const float* __restrict__ img) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= P.begin[P.nout]) return;
  int l = 0;
  while (l + 1 < P.nout && t >= P.begin[l + 1]) ++l;
  long long e = t - P.begin[l];
  const int h = P.h[l], w = P.w[l];
  const int j = (int)(e % w), i = (int)((e / w) % h);
  const int c = (int)((e / ((long long)w * h)) % P.C), b = (int)(e / ((long long)w * h * P.C));
  const int ys = (int)(((long long)i * P.H) / h), ye = (int)((((long long)i + 1) * P.H + h - 1) / h);
  const int xs = (int)(((long long)j * P.W) / w), xe = (int)((((long long)j + 1) * P.W + w - 1) / w);
  const float* src = img + b * P.sb + c * P.sc;
  float s = 0.0f;
  for (int y = ys; y < ye; ++y)
    for (int x = xs; x < xe; ++x) s += __ldg(src + y * P.sh + x);
  P.out[l][e] = s / (float)((ye - ys) * (xe - xs));
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_img_pyramid(const float* img, long long stride_b, long long stride_c, long long stride_h,
                               float* const* outs, int nlevels, int B, int C, int H, int W, uof_stream_t stream_) {
  UOF_REQUIRE(img && outs, "img_pyramid: null pointer");
  UOF_REQUIRE(nlevels >= 2 && nlevels <= UOF_MAX_LEVELS + 1, "img_pyramid: nlevels must be 2..%d", UOF_MAX_LEVELS + 1);
  UOF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "img_pyramid: bad shape");
  PyrParams P;
  P.nout = nlevels - 1;
  P.B = B; P.C = C; P.H = H; P.W = W;
  P.sb = stride_b; P.sc = stride_c; P.sh = stride_h;
  long long total = 0;
  for (int l = 0; l < P.nout; ++l) {
    const int s = l + 1;
    P.h[l] = (int)((double)H / (double)(1 << s));   // int(img_h / 2**s), model_flow_paper.py:58
    P.w[l] = (int)((double)W / (double)(1 << s));
    UOF_REQUIRE(P.h[l] > 0 && P.w[l] > 0 && outs[l], "img_pyramid: level %d is empty", s);
    P.out[l] = outs[l];
    P.begin[l] = total;
    total += (long long)B * C * P.h[l] * P.w[l];
  }
  P.begin[P.nout] = total;
  pyramid_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(P, img);
  count_launch();
  return check_launch("img_pyramid");
}
