// Stand-alone kernels for the reference's individual loss methods (the fused kernels in photo_loss.cu serve
// Model_flow.forward; these serve callers that use the methods one by one, SURVEY 8b "losses" seam):
//   uof_diff_weight_fwd/bwd   <- Model_flow.compute_diff_weight     (model_flow_paper.py:101-134)
//   uof_masked_mean_fwd/bwd   <- Model_flow.compute_loss_with_mask  (model_flow_paper.py:90-99)
// Pure streaming kernels: float4 along x when the plane size allows, warp-shuffle + one atomic per warp for the
// per-sample sums.
#include "common.cuh"

namespace uof {
namespace {

constexpr float kThird = 1.0f / 3.0f;
constexpr float kInvSigma2 = 1.0f / 0.03f;
constexpr float kEps = 1e-12f;

// ---------------------------------------------------------------------------------------- diff + weight
// one thread per pixel; planes are (B,3,H,W) / (B,1,H,W)
__global__ void __launch_bounds__(256)
diff_weight_fwd_kernel(const float* __restrict__ img, const float* __restrict__ wl, const float* __restrict__ wr,
                       float* __restrict__ diff_l, float* __restrict__ diff_r, float* __restrict__ weight_l,
                       float* __restrict__ weight_r, int B, int plane) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * plane) return;
  const int b = (int)(t / plane), p = (int)(t - (long long)b * plane);
  const size_t o3 = (size_t)b * 3 * plane + p;
  float I[3], L[3], R[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    I[c] = __ldg(img + o3 + (size_t)c * plane);
    L[c] = __ldg(wl + o3 + (size_t)c * plane);
    R[c] = __ldg(wr + o3 + (size_t)c * plane);
  }
  const float dl = (fabsf(I[0] - L[0]) + fabsf(I[1] - L[1]) + fabsf(I[2] - L[2])) * kThird;   // :117
  const float dr = (fabsf(I[0] - R[0]) + fabsf(I[1] - R[1]) + fabsf(I[2] - R[2])) * kThird;   // :118
  const float vl = (L[0] == 0.0f && L[1] == 0.0f && L[2] == 0.0f) ? 0.0f : 1.0f;              // :112
  const float vr = (R[0] == 0.0f && R[1] == 0.0f && R[2] == 0.0f) ? 0.0f : 1.0f;              // :111
  const float tt = __expf(-fabsf(dl - dr));                                                   // :120-121, see photo_loss.cu
  const float h = __fdividef(1.0f, 1.0f + tt) - 0.5f;
  const float g = 2.0f * __expf(-(h * h) * kInvSigma2);                                       // :126
  diff_l[t] = dl;
  diff_r[t] = dr;
  weight_l[t] = g * vl;                                                                       // :128
  weight_r[t] = g * vr;                                                                       // :129
}

// d diff / d warped_c = -sign(I_c - W_c) / 3   (weights are detached in the reference, :122)
__global__ void __launch_bounds__(256)
diff_weight_bwd_kernel(const float* __restrict__ img, const float* __restrict__ wl, const float* __restrict__ wr,
                       const float* __restrict__ g_diff_l, const float* __restrict__ g_diff_r, float* __restrict__ gwl,
                       float* __restrict__ gwr, int B, int plane) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * plane) return;
  const int b = (int)(t / plane), p = (int)(t - (long long)b * plane);
  const size_t o3 = (size_t)b * 3 * plane + p;
  const float gl = g_diff_l ? __ldg(g_diff_l + t) * kThird : 0.0f, gr = g_diff_r ? __ldg(g_diff_r + t) * kThird : 0.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const size_t o = o3 + (size_t)c * plane;
    const float I = __ldg(img + o);
    const float a = I - __ldg(wl + o), r = I - __ldg(wr + o);
    gwl[o] = a > 0.0f ? -gl : (a < 0.0f ? gl : 0.0f);
    gwr[o] = r > 0.0f ? -gr : (r < 0.0f ? gr : 0.0f);
  }
}

// ------------------------------------------------------------------------------------------ masked mean
struct MaskedParams {
  const float* diff[UOF_MAX_LEVELS];
  const float* mask[UOF_MAX_LEVELS];
  float* gdiff[UOF_MAX_LEVELS];
  int plane[UOF_MAX_LEVELS];
  int warp_begin[UOF_MAX_LEVELS + 1];
  int warps_per_sample[UOF_MAX_LEVELS];
  int nlevels, B, C;       // diff has C channels (1 in the reference), the mask one
};
constexpr int kPxPerWarp = 32 * 8;

__device__ __forceinline__ bool locate(const MaskedParams& P, int gw, int& l, int& b, int& px0) {
  if (gw >= P.warp_begin[P.nlevels]) return false;
  l = 0;
  while (l + 1 < P.nlevels && gw >= P.warp_begin[l + 1]) ++l;
  const int local = gw - P.warp_begin[l];
  b = local / P.warps_per_sample[l];
  px0 = (local % P.warps_per_sample[l]) * kPxPerWarp;
  return true;
}

__global__ void __launch_bounds__(128) masked_mean_fwd_kernel(const __grid_constant__ MaskedParams P, float* __restrict__ sums) {
  const int lane = threadIdx.x & 31, gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  int l, b, px0;
  if (!locate(P, gw, l, b, px0)) return;
  const int plane = P.plane[l];
  const float* m = P.mask[l] + (size_t)b * plane;
  const float* d = P.diff[l] + (size_t)b * P.C * plane;
  float num = 0.0f, den = 0.0f;
#pragma unroll
  for (int it = 0; it < kPxPerWarp / 32; ++it) {
    const int p = px0 + it * 32 + lane;
    if (p < plane) {
      const float w = __ldg(m + p);
      float s = 0.0f;
      for (int c = 0; c < P.C; ++c) s += __ldg(d + (size_t)c * plane + p);
      num = fmaf(s, w, num);
      den += w;
    }
  }
  num = warp_sum(num);
  den = warp_sum(den);
  if (lane == 0) {
    atomicAdd(sums + ((size_t)l * P.B + b) * 2, num);
    atomicAdd(sums + ((size_t)l * P.B + b) * 2 + 1, den);
  }
}

// loss[b] = sum_l mean(diff * mask_repeated) / (mean(mask) + eps): the reference repeats the mask to 3 channels and
// broadcasts a 1-channel diff against it, so the numerator mean runs over max(C,3)/C copies -- the value is unchanged.
__global__ void masked_mean_finalize_kernel(const __grid_constant__ MaskedParams P, const float* __restrict__ sums,
                                            float* __restrict__ loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  float acc = 0.0f;
  for (int l = 0; l < P.nlevels; ++l) {
    const float n = (float)P.plane[l];
    const float* s = sums + ((size_t)l * P.B + b) * 2;
    acc += (s[0] / (n * (float)P.C)) / (s[1] / n + kEps);
  }
  loss[b] = acc;
}

__global__ void __launch_bounds__(128)
masked_mean_bwd_kernel(const __grid_constant__ MaskedParams P, const float* __restrict__ sums, const float* __restrict__ g_loss) {
  const int lane = threadIdx.x & 31, gw = blockIdx.x * 4 + (threadIdx.x >> 5);
  int l, b, px0;
  if (!locate(P, gw, l, b, px0)) return;
  const int plane = P.plane[l];
  const float n = (float)plane;
  const float coef = __ldg(g_loss + b) / (n * (float)P.C) / (sums[((size_t)l * P.B + b) * 2 + 1] / n + kEps);
  const float* m = P.mask[l] + (size_t)b * plane;
  float* g = P.gdiff[l] + (size_t)b * P.C * plane;
#pragma unroll
  for (int it = 0; it < kPxPerWarp / 32; ++it) {
    const int p = px0 + it * 32 + lane;
    if (p < plane) {
      const float v = coef * __ldg(m + p);
      for (int c = 0; c < P.C; ++c) g[(size_t)c * plane + p] = v;
    }
  }
}

int fill_masked(MaskedParams& P, const float* const* diff, const float* const* mask, float* const* gdiff, const int* H,
                const int* W, int nlevels, int B, int C, bool bwd) {
  UOF_REQUIRE(diff && mask && H && W && nlevels >= 1 && nlevels <= UOF_MAX_LEVELS, "masked_mean: nlevels must be 1..%d", UOF_MAX_LEVELS);
  UOF_REQUIRE(B > 0 && C > 0, "masked_mean: bad batch/channels");
  long long total = 0;
  for (int l = 0; l < nlevels; ++l) {
    UOF_REQUIRE(diff[l] && mask[l] && H[l] > 0 && W[l] > 0, "masked_mean: level %d incomplete", l);
    if (bwd) UOF_REQUIRE(gdiff && gdiff[l], "masked_mean_bwd: level %d has no gradient buffer", l);
    UOF_REQUIRE((long long)H[l] * W[l] < (1ll << 30), "masked_mean: level %d too large", l);
    P.diff[l] = diff[l];
    P.mask[l] = mask[l];
    P.gdiff[l] = bwd ? gdiff[l] : nullptr;
    P.plane[l] = H[l] * W[l];
    P.warps_per_sample[l] = ceil_div(P.plane[l], kPxPerWarp);
    P.warp_begin[l] = (int)total;
    total += (long long)P.warps_per_sample[l] * B;
    UOF_REQUIRE(total < (1ll << 30), "masked_mean: problem too large");
  }
  P.warp_begin[nlevels] = (int)total;
  P.nlevels = nlevels;
  P.B = B;
  P.C = C;
  return UOF_OK;
}

// ---- objective of the training step (train.py:147-150): loss = sum_k w_k * mean_B(loss_k) -------------------------------
// One block; replaces 4 means + 4 multiplies + a stack + a sum forward and their ~16 element-wise backward launches
// (each ~2 us inside the CUDA graph) by one launch each way.
constexpr int kMaxTerms = 8;
struct CombineParams {
  const float* src[kMaxTerms];
  float* dst[kMaxTerms];
  float w[kMaxTerms];
  int n[kMaxTerms];
  int K;
};

__global__ void __launch_bounds__(128) weighted_mean_sum_fwd_kernel(const __grid_constant__ CombineParams P, float* __restrict__ out) {
  __shared__ float part[4];
  float total = 0.0f;
  for (int k = 0; k < P.K; ++k) {
    float s = 0.0f;
    for (int i = threadIdx.x; i < P.n[k]; i += blockDim.x) s += __ldg(P.src[k] + i);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) total = fmaf(P.w[k], (part[0] + part[1] + part[2] + part[3]) / (float)P.n[k], total);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = total;
}

__global__ void __launch_bounds__(128) weighted_mean_sum_bwd_kernel(const __grid_constant__ CombineParams P, const float* __restrict__ g) {
  const float gv = __ldg(g);
  for (int k = 0; k < P.K; ++k) {
    const float v = gv * P.w[k] / (float)P.n[k];
    for (int i = threadIdx.x; i < P.n[k]; i += blockDim.x) P.dst[k][i] = v;
  }
}

int fill_combine(CombineParams& P, const float* const* src, float* const* dst, const float* w, const int* n, int K) {
  UOF_REQUIRE(w && n && K >= 1 && K <= kMaxTerms, "weighted_mean_sum: 1..%d terms", kMaxTerms);
  P.K = K;
  for (int k = 0; k < K; ++k) {
    UOF_REQUIRE(n[k] > 0 && (src ? src[k] != nullptr : dst[k] != nullptr), "weighted_mean_sum: term %d incomplete", k);
    P.src[k] = src ? src[k] : nullptr;
    P.dst[k] = dst ? dst[k] : nullptr;
    P.w[k] = w[k];
    P.n[k] = n[k];
  }
  return UOF_OK;
}

}  // namespace
}  // namespace uof

using namespace uof;

extern "C" int uof_diff_weight_fwd(const float* img, const float* warped_l, const float* warped_r, float* diff_l,
                                   float* diff_r, float* weight_l, float* weight_r, int B, int H, int W, uof_stream_t stream_) {
  UOF_REQUIRE(img && warped_l && warped_r && diff_l && diff_r && weight_l && weight_r, "diff_weight_fwd: null pointer");
  UOF_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "diff_weight_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  const long long n = (long long)B * H * W;
  diff_weight_fwd_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      img, warped_l, warped_r, diff_l, diff_r, weight_l, weight_r, B, H * W);
  count_launch();
  return check_launch("diff_weight_fwd");
}

extern "C" int uof_diff_weight_bwd(const float* img, const float* warped_l, const float* warped_r, const float* g_diff_l,
                                   const float* g_diff_r, float* gwarped_l, float* gwarped_r, int B, int H, int W,
                                   uof_stream_t stream_) {
  UOF_REQUIRE(img && warped_l && warped_r && gwarped_l && gwarped_r, "diff_weight_bwd: null pointer");
  UOF_REQUIRE(B > 0 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), "diff_weight_bwd: bad shape B=%d H=%d W=%d", B, H, W);
  const long long n = (long long)B * H * W;
  diff_weight_bwd_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      img, warped_l, warped_r, g_diff_l, g_diff_r, gwarped_l, gwarped_r, B, H * W);
  count_launch();
  return check_launch("diff_weight_bwd");
}

extern "C" int uof_masked_mean_fwd(const float* const* diff, const float* const* mask, const int* H, const int* W, int nlevels,
                                   int B, int C, float* sums, float* loss, uof_stream_t stream_) {
  UOF_REQUIRE(sums && loss, "masked_mean_fwd: null output");
  MaskedParams P;
  if (int rc = fill_masked(P, diff, mask, nullptr, H, W, nlevels, B, C, false)) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  UOF_CUDA(cudaMemsetAsync(sums, 0, (size_t)nlevels * B * 2 * sizeof(float), stream));
  masked_mean_fwd_kernel<<<ceil_div(P.warp_begin[nlevels], 4), 128, 0, stream>>>(P, sums);
  masked_mean_finalize_kernel<<<ceil_div(B, 64), 64, 0, stream>>>(P, sums, loss);
  count_launch(2);
  return check_launch("masked_mean_fwd");
}

extern "C" int uof_masked_mean_bwd(const float* const* diff, const float* const* mask, float* const* gdiff, const int* H,
                                   const int* W, int nlevels, int B, int C, const float* sums, const float* g_loss,
                                   uof_stream_t stream_) {
  UOF_REQUIRE(sums && g_loss, "masked_mean_bwd: null input");
  MaskedParams P;
  if (int rc = fill_masked(P, diff, mask, gdiff, H, W, nlevels, B, C, true)) return rc;
  masked_mean_bwd_kernel<<<ceil_div(P.warp_begin[nlevels], 4), 128, 0, static_cast<cudaStream_t>(stream_)>>>(P, sums, g_loss);
  count_launch();
  return check_launch("masked_mean_bwd");
}

extern "C" int uof_weighted_mean_sum_fwd(const float* const* terms, const float* weights, const int* n, int K, float* out,
                                         uof_stream_t stream_) {
  UOF_REQUIRE(terms && out, "weighted_mean_sum_fwd: null pointer");
  CombineParams P;
  if (int rc = fill_combine(P, terms, nullptr, weights, n, K)) return rc;
  weighted_mean_sum_fwd_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream_)>>>(P, out);
  count_launch();
  return check_launch("weighted_mean_sum_fwd");
}

extern "C" int uof_weighted_mean_sum_bwd(const float* g_out, const float* weights, const int* n, int K, float* const* g_terms,
                                         uof_stream_t stream_) {
  UOF_REQUIRE(g_out && g_terms, "weighted_mean_sum_bwd: null pointer");
  CombineParams P;
  if (int rc = fill_combine(P, nullptr, g_terms, weights, n, K)) return rc;
  weighted_mean_sum_bwd_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream_)>>>(P, g_out);
  count_launch();
  return check_launch("weighted_mean_sum_bwd");
}
