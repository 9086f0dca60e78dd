// Decoder/encoder glue: fused bias + LeakyReLU(0.1) after every 3x3 convolution.
// The reference builds every layer as nn.Sequential(Conv2d(bias=True), LeakyReLU(0.1))
// (/root/reference/core/networks/structures/net_utils.py:7-11).  In PyTorch eager that costs, per layer and step,
// a cuDNN bias/scale epilogue kernel + a LeakyReLU kernel forward and a LeakyReLU-backward kernel + a bias-gradient
// reduction backward (ncu launch list, profiles/: ~170 launches, ~3 ms of the 54.5 ms step).  Here the convolution runs
// bias-free in cuDNN and ONE kernel each way does the rest:
//   forward : y[b,c,:] = lrelu(y[b,c,:] + bias[c])                       (in place on the convolution output)
//   backward: gx = gout * (y > 0 ? 1 : slope);  gbias[c] = sum_{b,hw} gx  (one pass, block reduction + one atomic per block)
#include "common.cuh"

namespace uof {
namespace {

constexpr int kThreads = 256;

// grid: (chunks of a plane, C, B); a block handles up to kThreads*4*ITER elements of one (b, c) plane
constexpr int ITER = 4;

// `in` is the dense convolution output; the activation goes to d1 (batch stride bs1; d1 == in, bs1 == C*plane is the
// in-place form) and, when d2 != nullptr, also to d2 (batch stride bs2).  The strided destinations are channel slices of
// the decoder's pre-allocated concat buffers (pwc_tf.py:113-118: every activation is concatenated twice), so no torch.cat
// copy kernel ever runs: 7 -> 4 passes over each activation.
template <bool VEC4, bool TWO>
__global__ void __launch_bounds__(kThreads)
bias_lrelu_fwd_kernel(const float* in /* may alias d1 */, const float* __restrict__ bias, float* d1, long long bs1,
                      float* __restrict__ d2, long long bs2, int C, int plane, float slope) {
  const int c = blockIdx.y, b = blockIdx.z;
  const float bv = __ldg(bias + c);
  const float* p = in + ((size_t)b * C + c) * plane;
  float* o1 = d1 + (size_t)b * bs1 + (size_t)c * plane;
  float* o2 = TWO ? d2 + (size_t)b * bs2 + (size_t)c * plane : nullptr;
  if (VEC4) {
    const int n4 = plane >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = (blockIdx.x * ITER + it) * kThreads + threadIdx.x;
      if (i < n4) {
        float4 v = p4[i];
        v.x += bv; v.y += bv; v.z += bv; v.w += bv;
        v.x = v.x > 0.0f ? v.x : v.x * slope;
        v.y = v.y > 0.0f ? v.y : v.y * slope;
        v.z = v.z > 0.0f ? v.z : v.z * slope;
        v.w = v.w > 0.0f ? v.w : v.w * slope;
        reinterpret_cast<float4*>(o1)[i] = v;
        if (TWO) reinterpret_cast<float4*>(o2)[i] = v;
      }
    }
  } else {
#pragma unroll
    for (int it = 0; it < ITER * 4; ++it) {
      const int i = (blockIdx.x * ITER * 4 + it) * kThreads + threadIdx.x;
      if (i < plane) {
        float v = p[i] + bv;
        v = v > 0.0f ? v : v * slope;
        o1[i] = v;
        if (TWO) o2[i] = v;
      }
    }
  }
}

// gout = g1 (+ g2): the incoming gradient may arrive as TWO tensors.  In the decoder every activation feeds two consumers
// (the next convolution and a torch.cat, pwc_tf.py:119-131); autograd would add the two gradients with a strided
// elementwise kernel (they are channel slices of cat gradients: dense inside a sample, batch stride of the cat) and hand
// the sum over -- 36 launches / 0.55 ms per step plus 14 .contiguous() copies (ncu launch list, round 2).  Here both
// arrive as they are (`bs1`, `bs2` = batch strides in floats) and are summed on the fly.
template <bool VEC4, bool TWO>
__global__ void __launch_bounds__(kThreads)
bias_lrelu_bwd_kernel(const float* __restrict__ g1, long long bs1, const float* __restrict__ g2, long long bs2,
                      const float* __restrict__ y_, long long y_bs, float* __restrict__ gx, float* __restrict__ gbias, int C,
                      int plane, float slope) {
  const int c = blockIdx.y, b = blockIdx.z;
  const size_t base = ((size_t)b * C + c) * plane;
  const float* y = y_ + (size_t)b * y_bs + (size_t)c * plane - base;      // so that y + base is this (b, c) plane
  const float* g1p = g1 + (size_t)b * bs1 + (size_t)c * plane;
  const float* g2p = TWO ? g2 + (size_t)b * bs2 + (size_t)c * plane : nullptr;
  float acc = 0.0f;
  if (VEC4) {
    const int n4 = plane >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g1p);
    const float4* h4 = reinterpret_cast<const float4*>(g2p);
    const float4* y4 = reinterpret_cast<const float4*>(y + base);
    float4* o4 = reinterpret_cast<float4*>(gx + base);
    float4 g[ITER], h[ITER], v[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {      // all loads first
      const int i = (blockIdx.x * ITER + it) * kThreads + threadIdx.x;
      if (i < n4) {
        g[it] = __ldg(g4 + i);
        if (TWO) h[it] = __ldg(h4 + i);
        v[it] = __ldg(y4 + i);
      }
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = (blockIdx.x * ITER + it) * kThreads + threadIdx.x;
      if (i < n4) {
        float4 t = g[it];
        if (TWO) { t.x += h[it].x; t.y += h[it].y; t.z += h[it].z; t.w += h[it].w; }
        float4 r;
        r.x = v[it].x > 0.0f ? t.x : t.x * slope;
        r.y = v[it].y > 0.0f ? t.y : t.y * slope;
        r.z = v[it].z > 0.0f ? t.z : t.z * slope;
        r.w = v[it].w > 0.0f ? t.w : t.w * slope;
        o4[i] = r;
        acc += (r.x + r.y) + (r.z + r.w);
      }
    }
  } else {
#pragma unroll
    for (int it = 0; it < ITER * 4; ++it) {
      const int i = (blockIdx.x * ITER * 4 + it) * kThreads + threadIdx.x;
      if (i < plane) {
        float gv = __ldg(g1p + i);
        if (TWO) gv += __ldg(g2p + i);
        const float r = __ldg(y + base + i) > 0.0f ? gv : gv * slope;
        gx[base + i] = r;
        acc += r;
      }
    }
  }
  __shared__ float part[kThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = threadIdx.x < kThreads / 32 ? part[threadIdx.x] : 0.0f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(gbias + c, s);
  }
}

int check(const char* who, const void* a, const void* b, int B, int C, int H, int W) {
  UOF_REQUIRE(a && b, "%s: null pointer", who);
  UOF_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && B <= 65535 && C <= 65535, "%s: bad shape B=%d C=%d H=%d W=%d", who, B, C, H, W);
  UOF_REQUIRE((long long)H * W < (1ll << 30), "%s: plane too large", who);
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_bias_lrelu_fwd2(const float* in, const float* bias, float* dst1, long long dst1_batch_stride, float* dst2,
                                   long long dst2_batch_stride, int B, int C, int H, int W, float slope, uof_stream_t stream_) {
  if (int rc = check("bias_lrelu_fwd", in, bias, B, C, H, W)) return rc;
  UOF_REQUIRE(dst1, "bias_lrelu_fwd: null destination");
  const int plane = H * W;
  UOF_REQUIRE(dst1_batch_stride >= (long long)C * plane && (!dst2 || dst2_batch_stride >= (long long)C * plane),
              "bias_lrelu_fwd: destination batch stride smaller than a sample");
  const uintptr_t bits = reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(dst1) | reinterpret_cast<uintptr_t>(dst2);
  const bool v4 = (plane % 4 == 0) && (bits & 15u) == 0 && dst1_batch_stride % 4 == 0 && (!dst2 || dst2_batch_stride % 4 == 0);
  dim3 grid(ceil_div(v4 ? plane / 4 : ceil_div(plane, 4), kThreads * ITER), C, B);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
#define UOF_FWD(V, T) bias_lrelu_fwd_kernel<V, T><<<grid, kThreads, 0, stream>>>(in, bias, dst1, dst1_batch_stride, dst2, dst2_batch_stride, C, plane, slope)
  if (v4) { if (dst2) UOF_FWD(true, true); else UOF_FWD(true, false); }
  else { if (dst2) UOF_FWD(false, true); else UOF_FWD(false, false); }
#undef UOF_FWD
  count_launch();
  return check_launch("bias_lrelu_fwd");
}

extern "C" int uof_bias_lrelu_fwd(float* y, const float* bias, int B, int C, int H, int W, float slope, uof_stream_t stream) {
  return uof_bias_lrelu_fwd2(y, bias, y, (long long)C * H * W, nullptr, 0, B, C, H, W, slope, stream);
}

extern "C" int uof_bias_lrelu_bwd3(const float* g1, long long g1_batch_stride, const float* g2, long long g2_batch_stride,
                                   const float* y, long long y_batch_stride, float* gx, float* gbias, int B, int C, int H, int W,
                                   float slope, uof_stream_t stream_) {
  if (int rc = check("bias_lrelu_bwd", g1, y, B, C, H, W)) return rc;
  UOF_REQUIRE(gx && gbias, "bias_lrelu_bwd: null output");
  const int plane = H * W;
  UOF_REQUIRE(g1_batch_stride >= (long long)C * plane && (!g2 || g2_batch_stride >= (long long)C * plane) &&
                  y_batch_stride >= (long long)C * plane,
              "bias_lrelu_bwd: batch stride smaller than a sample");
  const uintptr_t bits = reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(g2) | reinterpret_cast<uintptr_t>(y) |
                         reinterpret_cast<uintptr_t>(gx);
  const bool v4 = (plane % 4 == 0) && (bits & 15u) == 0 && g1_batch_stride % 4 == 0 && (!g2 || g2_batch_stride % 4 == 0) &&
                  y_batch_stride % 4 == 0;
  dim3 grid(ceil_div(v4 ? plane / 4 : ceil_div(plane, 4), kThreads * ITER), C, B);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UOF_CUDA(cudaMemsetAsync(gbias, 0, (size_t)C * sizeof(float), stream));
#define UOF_BWD(V, T) bias_lrelu_bwd_kernel<V, T><<<grid, kThreads, 0, stream>>>(g1, g1_batch_stride, g2, g2_batch_stride, y, y_batch_stride, gx, gbias, C, plane, slope)
  if (v4) { if (g2) UOF_BWD(true, true); else UOF_BWD(true, false); }
  else { if (g2) UOF_BWD(false, true); else UOF_BWD(false, false); }
#undef UOF_BWD
  count_launch();
  return check_launch("bias_lrelu_bwd");
}

extern "C" int uof_bias_lrelu_bwd2(const float* g1, long long g1_batch_stride, const float* g2, long long g2_batch_stride,
                                   const float* y, float* gx, float* gbias, int B, int C, int H, int W, float slope,
                                   uof_stream_t stream) {
  return uof_bias_lrelu_bwd3(g1, g1_batch_stride, g2, g2_batch_stride, y, (long long)C * H * W, gx, gbias, B, C, H, W, slope, stream);
}

extern "C" int uof_bias_lrelu_bwd(const float* gout, const float* y, float* gx, float* gbias, int B, int C, int H, int W,
                                  float slope, uof_stream_t stream_) {
  return uof_bias_lrelu_bwd2(gout, (long long)C * H * W, nullptr, 0, y, gx, gbias, B, C, H, W, slope, stream_);
}
