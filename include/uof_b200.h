/*
 * uof_b200.h — C ABI of the B200-native Model_flow hot path (libuof_b200.so).
 *
 * Drop-in boundary for jianfenglihg/UnOpticalFlow's operator seams under core/networks
 * (SURVEY.md section 8b).  The reference has no FFI: its seams are Python callables
 * (`PWC_tf.corr`, `warp_flow`, `SSIM`, `Model_flow.compute_*`).  Each entry point below
 * replaces the ATen op chain behind one of them; the reference file:line it replaces is
 * cited per function.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless stated otherwise;
 *   - tensors are dense NCHW (the reference's layout) unless a `channels_last` flag or an
 *     explicit stride argument says otherwise;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising; it is safe to capture the calls into a CUDA graph;
 *   - return value: 0 = ok, UOF_ERR_* otherwise; uof_last_error() gives a message;
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef UOF_B200_H
#define UOF_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define UOF_OK 0
#define UOF_ERR_INVALID_ARGUMENT 1
#define UOF_ERR_CUDA 2
#define UOF_ERR_UNSUPPORTED 3

#define UOF_MAX_LEVELS 4
#define UOF_NUM_DISPLACEMENTS 81   /* (2*4+1)^2, pwc_tf.py:16,97 */

typedef void* uof_stream_t;        /* cudaStream_t */

/* library / diagnostics --------------------------------------------------------------- */
int uof_abi_version(void);
const char* uof_last_error(void);
/* number of kernels (memsets not counted) this library has enqueued since load */
long long uof_launch_count(void);

/* a1: cost volume.  Replaces PWC_tf.corr_naive (pwc_tf.py:97-106):
 *   out[b, 9*i+j, y, x] = (1/C) * sum_c f1[b,c,y,x] * f2[b,c,y+i-4,x+j-4]   (zero outside)
 * f1,f2: (B,C,H,W).  out: (B,81,H,W) with batch stride `out_batch_stride` elements
 * (81*H*W when dense; larger when writing into a pre-allocated concat buffer). */
int uof_cost_volume_fwd(const float* f1, const float* f2, float* out,
                        int B, int C, int H, int W, long long out_batch_stride,
                        uof_stream_t stream);
/* backward of the above: gf1,gf2 (B,C,H,W) are fully overwritten. */
int uof_cost_volume_bwd(const float* gout, long long gout_batch_stride,
                        const float* f1, const float* f2, float* gf1, float* gf2,
                        int B, int C, int H, int W, uof_stream_t stream);

/* a1 + decoder glue (pwc_tf.py:119-131): the same cost volume with
 *   - f1 read in place with an explicit batch stride (floats), e.g. the `c1` slice of the decoder's concat buffer
 *     cat((corr, c1, up_flow), 1), so c1 is stored once;
 *   - backward: `gadd` (nullable, (B,C,H,W) view with batch stride `gadd_batch_stride`) is added to gf1 in the kernel's
 *     epilogue -- the gradient of that concat slice -- so d/d c1 needs no separate add. */
int uof_cost_volume_fwd_ex(const float* f1, long long f1_batch_stride, const float* f2, float* out,
                           int B, int C, int H, int W, long long out_batch_stride, uof_stream_t stream);
int uof_cost_volume_bwd_ex(const float* gout, long long gout_batch_stride,
                           const float* f1, long long f1_batch_stride, const float* f2,
                           const float* gadd, long long gadd_batch_stride, float* gf1, float* gf2,
                           int B, int C, int H, int W, uof_stream_t stream);

/* a2/a3: bilinear backward warp.  Replaces warp_flow (net_utils.py:16-54): mesh grid,
 * normalisation, grid_sample(zeros padding) and, with use_mask, the validity mask
 * (sum of in-bounds corner weights >= 0.9999, net_utils.py:47-52) in ONE kernel.
 * x,out: (B,C,H,W) NCHW, or NHWC storage when channels_last != 0 (C % 4 == 0 required);
 * flow: (B,2,H,W) NCHW always.  align_corners is a flags word: bit 0 selects the grid_sample
 * convention (0 = installed torch default, 1 = torch-1.2 behaviour); bit 1
 * (UOF_COORD_HOST) selects the fp32 rounding of the coordinate chain: clear = ATen's CUDA kernels
 * (reciprocal multiply + FMA, what the reference computes on a GPU), set = ATen's CPU kernels
 * (true division, no contraction: the CPU oracle's arithmetic).  The two differ by one ulp of the
 * normalised coordinate, i.e. up to 1e-4 relative in the flow gradient at W ~ 800. */
#define UOF_ALIGN_CORNERS 1
#define UOF_COORD_HOST 2
int uof_warp_fwd(const float* x, const float* flow, float* out,
                 int B, int C, int H, int W, int use_mask, int align_corners,
                 int channels_last, uof_stream_t stream);
/* gx may be NULL (image warps: x carries no gradient).  gx is zero-filled here and then
 * accumulated with fp32 atomics; gflow (B,2,H,W) is overwritten. */
int uof_warp_bwd(const float* gout, const float* x, const float* flow, float* gx, float* gflow,
                 int B, int C, int H, int W, int use_mask, int align_corners,
                 int channels_last, uof_stream_t stream);

/* a4+a5+a6 fused: photometric weights + masked L1 + masked SSIM for BOTH directions, all
 * pyramid levels in one launch.  Replaces Model_flow.compute_diff_weight (:101-134),
 * compute_loss_with_mask x2 (:90-99, :241-242), compute_loss_ssim x2 (:137-148, :244-245)
 * and SSIM (pytorch_ssim/ssim.py:4-19).   "l" = warped from the left image (reference
 * "bwd"), "r" = from the right image (reference "fwd"). */
typedef struct {
  const float* img;        /* (B,3,H,W) target image at this level            */
  const float* warped_l;   /* (B,3,H,W) masked warp of the left image          */
  const float* warped_r;   /* (B,3,H,W) masked warp of the right image         */
  float* weight_l;         /* (B,1,H,W) out: soft weight map, or NULL          */
  float* weight_r;
  float* diff_l;           /* (B,1,H,W) out: mean_c |img - warped|, or NULL    */
  float* diff_r;
  float* gwarped_l;        /* (B,3,H,W) out of the backward pass               */
  float* gwarped_r;
  int H, W;
} uof_photo_level;
/* Reduction workspaces of the fused loss entry points (photo, smooth, consis): `sums` holds nlevels*B*K partial sums
 * followed by UOF_SUMS_EXTRA floats reserved for the kernels (smooth_loss_fwd keeps a block counter there: the last
 * block to finish forms the (B) losses, so it needs no separate finalize launch).  Allocate nlevels*B*K + UOF_SUMS_EXTRA floats; the backward entry
 * points read only the first nlevels*B*K. */
#define UOF_SUMS_EXTRA 1
/* sums: (nlevels,B,6) + UOF_SUMS_EXTRA workspace, zero-filled here, then
 *   [0]=sum d_l*w_l [1]=sum w_l [2]=sum d_r*w_r [3]=sum w_r [4]=sum ssim_term_l [5]=sum ssim_term_r
 * loss_pixel, loss_ssim: (B) out, summed over levels and both directions. */
int uof_photo_loss_fwd(const uof_photo_level* levels, int nlevels, int B,
                       float* sums, float* loss_pixel, float* loss_ssim, uof_stream_t stream);
/* backward: writes gwarped_l/gwarped_r.  If every level carries weight_l AND weight_r they must hold the maps written
 * by uof_photo_loss_fwd on the same inputs (the faster "split" kernel reads them); pass NULL to have them recomputed. */
int uof_photo_loss_bwd(const uof_photo_level* levels, int nlevels, int B, const float* sums,
                       const float* g_loss_pixel, const float* g_loss_ssim, uof_stream_t stream);

/* a3+a4+a5+a6 fused (round 2): the image warps of Model_flow.forward (warp_flow_pyramid, model_flow_paper.py:62-66,
 * 236-237 -> net_utils.py:16-54 with use_mask=True) evaluated INSIDE the photometric kernels above.  The forward kernel
 * gathers the left / right image at (pixel + flow) itself -- same coordinate arithmetic, same validity mask, same blend
 * order as uof_warp_fwd, so the warped values are bit-identical -- and the backward kernel turns d loss / d warped into
 * d loss / d flow in its epilogue (uof_warp_bwd without gradient w.r.t. the image).  One launch each way for all pyramid
 * levels and both directions instead of 3 + 1 (forward) and 1 + 3 (backward).
 * Requires even W and 8-byte aligned planes at every level (else UOF_ERR_UNSUPPORTED: use the separate entry points). */
typedef struct {
  const float* img;        /* (B,3,H,W) target image at this level                                     */
  const float* src_l;      /* (B,3,H,W) left image at this level, sampled with flow_l                  */
  const float* src_r;      /* (B,3,H,W) right image, sampled with flow_r                               */
  const float* flow_l;     /* (B,2,H,W) flow target -> left  (reference "bwd")                         */
  const float* flow_r;     /* (B,2,H,W) flow target -> right (reference "fwd")                         */
  float* warped_l;         /* (B,3,H,W) masked warp: written by fwd (NULL: not stored), read by bwd    */
  float* warped_r;
  float* weight_l;         /* (B,1,H,W) weight maps: written by fwd (NULL: not stored), read by bwd    */
  float* weight_r;
  float* diff_l;           /* (B,1,H,W) out of fwd: mean_c |img - warped|, or NULL                     */
  float* diff_r;
  float* gflow_l;          /* (B,2,H,W) out of bwd: d loss / d flow_l                                  */
  float* gflow_r;
  int H, W;
} uof_photo_warp_level;
/* coord_flags: bit 0 align_corners, bit 1 UOF_COORD_HOST (see uof_warp_fwd).  sums / loss_*: as uof_photo_loss_fwd. */
int uof_photo_warp_loss_fwd(const uof_photo_warp_level* levels, int nlevels, int B, int coord_flags,
                            float* sums, float* loss_pixel, float* loss_ssim, uof_stream_t stream);
/* backward: needs warped_*, weight_* as written by the forward call on the same inputs; writes gflow_*. */
int uof_photo_warp_loss_bwd(const uof_photo_warp_level* levels, int nlevels, int B, int coord_flags, const float* sums,
                            const float* g_loss_pixel, const float* g_loss_ssim, uof_stream_t stream);

/* a4 seam: Model_flow.compute_diff_weight (model_flow_paper.py:101-134) on one pyramid level.
 * img, warped_*: (B,3,H,W); diff_*, weight_*: (B,1,H,W).  Weights carry no gradient (detached in the reference). */
int uof_diff_weight_fwd(const float* img, const float* warped_l, const float* warped_r, float* diff_l, float* diff_r,
                        float* weight_l, float* weight_r, int B, int H, int W, uof_stream_t stream);
/* g_diff_* (B,1,H,W) may be NULL (treated as zero); gwarped_* (B,3,H,W) are overwritten. */
int uof_diff_weight_bwd(const float* img, const float* warped_l, const float* warped_r, const float* g_diff_l,
                        const float* g_diff_r, float* gwarped_l, float* gwarped_r, int B, int H, int W,
                        uof_stream_t stream);

/* a5 seam: Model_flow.compute_loss_with_mask (model_flow_paper.py:90-99):
 *   loss[b] = sum_l mean(diff_l * mask_l) / (mean(mask_l) + 1e-12).
 * diff, mask, H, W: HOST arrays of nlevels entries; diff[l] is (B,C,H,W), mask[l] (B,1,H,W).
 * sums: (nlevels,B,2) workspace (zero-filled here).  Gradient flows to diff only. */
int uof_masked_mean_fwd(const float* const* diff, const float* const* mask, const int* H, const int* W, int nlevels,
                        int B, int C, float* sums, float* loss, uof_stream_t stream);
int uof_masked_mean_bwd(const float* const* diff, const float* const* mask, float* const* gdiff, const int* H,
                        const int* W, int nlevels, int B, int C, const float* sums, const float* g_loss,
                        uof_stream_t stream);

/* Objective of the training step (train.py:147-150): out[0] = sum_k weights[k] * mean(terms[k][0..n[k])).
 * terms / g_terms, weights, n: HOST arrays of K <= 8 entries (terms[k] device pointers); out, g_out: device scalars.
 * Backward: g_terms[k][i] = g_out[0] * weights[k] / n[k]. */
int uof_weighted_mean_sum_fwd(const float* const* terms, const float* weights, const int* n, int K, float* out,
                              uof_stream_t stream);
int uof_weighted_mean_sum_bwd(const float* g_out, const float* weights, const int* n, int K, float* const* g_terms,
                              uof_stream_t stream);

/* a6 seam: SSIM map.  Replaces SSIM(x,y) (pytorch_ssim/ssim.py:4-19) on N = B*C planes. */
int uof_ssim_fwd(const float* x, const float* y, float* out, int N, int H, int W, uof_stream_t stream);
int uof_ssim_bwd(const float* gout, const float* x, const float* y, float* gx, float* gy,
                 int N, int H, int W, uof_stream_t stream);

/* a7: edge-aware second-order smoothness.  Replaces Model_flow.gradients/cal_grad2_error/
 * compute_loss_flow_smooth (model_flow_paper.py:152-177).  flow: (B,2,H,W) in pixels (the /20
 * is applied inside); img: (Bimg,3,H,W), sample b reads image b % Bimg so both flow
 * directions can share one launch. */
typedef struct {
  const float* flow;
  const float* img;
  float* gflow;            /* (B,2,H,W) out of the backward pass */
  int H, W;
} uof_smooth_level;
int uof_smooth_loss_fwd(const uof_smooth_level* levels, int nlevels, int B, int Bimg,
                        float* sums /* (nlevels,B,2) + UOF_SUMS_EXTRA */, float* loss /* (B) */, uof_stream_t stream);
int uof_smooth_loss_bwd(const uof_smooth_level* levels, int nlevels, int B, int Bimg,
                        const float* g_loss, uof_stream_t stream);
/* same, `accumulate` != 0: gflow += (the buffer already holds another loss's gradient w.r.t. the same flows, e.g. the one
 * written by uof_photo_warp_loss_bwd) -- saves autograd's add kernels when several losses read one flow pyramid. */
int uof_smooth_loss_bwd_acc(const uof_smooth_level* levels, int nlevels, int B, int Bimg,
                            const float* g_loss, int accumulate, uof_stream_t stream);

/* a8: flow-direction consistency.  Replaces get_flow_normalization + compute_loss_flow_consis
 * (model_flow_paper.py:44-51,180-195).  Gradient flows to flow_fwd only. */
typedef struct {
  const float* flow_fwd;   /* (B,2,H,W) */
  const float* flow_bwd;   /* (B,2,H,W) */
  const float* weight_fwd; /* (B,1,H,W) */
  float* gflow_fwd;        /* (B,2,H,W) out of the backward pass */
  int H, W;
} uof_consis_level;
int uof_consis_loss_fwd(const uof_consis_level* levels, int nlevels, int B,
                        float* sums /* (nlevels,B,2) + UOF_SUMS_EXTRA */, float* loss /* (B) */, uof_stream_t stream);
int uof_consis_loss_bwd(const uof_consis_level* levels, int nlevels, int B, const float* sums,
                        const float* g_loss, uof_stream_t stream);
int uof_consis_loss_bwd_acc(const uof_consis_level* levels, int nlevels, int B, const float* sums,
                            const float* g_loss, int accumulate, uof_stream_t stream);

/* a9: image pyramid.  Replaces Model_flow.generate_img_pyramid (model_flow_paper.py:54-60) for
 * levels 1..nlevels-1 (level 0 is the input itself); adaptive_avg_pool2d bin rule
 * [floor(i*H/h), ceil((i+1)*H/h)).  `nimg` images of shape (B,C,H,W) are addressed with explicit element
 * strides (image, batch, channel, row; columns dense), so the three images of the vertically stacked
 * triplet (B,3,3H,W) are processed in place by one launch (stride_img = H*W_row_stride).
 * outs[l] receives level l+1 as a dense (nimg,B,C,h,w) tensor. */
int uof_img_pyramid(const float* img, long long stride_img, long long stride_b, long long stride_c,
                    long long stride_h, float* const* outs /* host array of nlevels-1 device pointers */,
                    int nlevels, int nimg, int B, int C, int H, int W, uof_stream_t stream);
/* Same, with the outputs stacked by SLOT: input image i goes to slot[i] (host array, a permutation of 0..nimg-1; NULL =
 * identity), and an optional dense (nimg,B,C,H,W) copy `out0` of level 0 in the same order.  Model_flow.forward orders
 * the triplet as [left; right; centre], so that the encoder's 3B batch and the [left; right] sources of the image warps
 * (model_flow_paper.py:206-209,233-235) are plain views instead of torch.cat copies.  Needs H, W % 4 == 0, nlevels <= 3. */
int uof_img_pyramid_stacked(const float* img, long long stride_img, long long stride_b, long long stride_c,
                            long long stride_h, float* out0, const int* slot, float* const* outs, int nlevels, int nimg,
                            int B, int C, int H, int W, uof_stream_t stream);

/* a11 glue: fused bias + LeakyReLU after a bias-free convolution.  Replaces the bias epilogue + LeakyReLU
 * of conv() (net_utils.py:7-11) forward, and LeakyReLU-backward + bias-gradient reduction backward.
 * y: (B,C,H,W) convolution output, updated IN PLACE to lrelu(y + bias[c]). */
int uof_bias_lrelu_fwd(float* y, const float* bias, int B, int C, int H, int W, float slope, uof_stream_t stream);
/* y = the post-activation tensor of the forward pass; gx (B,C,H,W) overwritten, gbias (C) zero-filled then accumulated. */
int uof_bias_lrelu_bwd(const float* gout, const float* y, float* gx, float* gbias, int B, int C, int H, int W,
                       float slope, uof_stream_t stream);
/* Same with the incoming gradient given as g1 (+ g2, may be NULL): (B,C,H,W) views that are dense inside a sample and
 * have an arbitrary batch stride (in floats) -- e.g. channel slices of torch.cat gradients -- summed on the fly, so the
 * two consumers of a decoder activation (next convolution and concat, pwc_tf.py:119-131) need no separate add/copy. */
int uof_bias_lrelu_bwd2(const float* g1, long long g1_batch_stride, const float* g2, long long g2_batch_stride,
                        const float* y, float* gx, float* gbias, int B, int C, int H, int W, float slope,
                        uof_stream_t stream);

/* Concat-free dense block (pwc_tf.py:113-118: x = cat((conv(x), x), 1) five times per level): the activation of the dense
 * convolution output `in` is written to dst1 and, when dst2 != NULL, also to dst2 -- (B,C,H,W) views with explicit batch
 * strides (floats), i.e. channel slices of pre-allocated concat buffers (dst1 == in with a dense stride is the in-place
 * form).  Every activation of the decoder is concatenated twice; writing it where the concatenations will be read removes
 * the torch.cat copy kernels (7 -> 4 passes over each activation).  bwd3 = bwd2 with the saved activation `y` read
 * through a batch stride as well (it lives in one of those buffers). */
int uof_bias_lrelu_fwd2(const float* in, const float* bias, float* dst1, long long dst1_batch_stride, float* dst2,
                        long long dst2_batch_stride, int B, int C, int H, int W, float slope, uof_stream_t stream);
int uof_bias_lrelu_bwd3(const float* g1, long long g1_batch_stride, const float* g2, long long g2_batch_stride,
                        const float* y, long long y_batch_stride, float* gx, float* gbias, int B, int C, int H, int W,
                        float slope, uof_stream_t stream);

/* a11 glue: bilinear up-sampling (align_corners = False, ATen upsample_bilinear2d semantics) fused with a scale factor:
 * out = scale * interpolate(in).  Replaces `F.interpolate(flow, scale_factor=2.0, mode='bilinear') * 2.0`
 * (pwc_tf.py:119,132,144,157) and `F.interpolate(flow * 4.0, [h, w], mode='bilinear')` (pwc_tf.py:174-177).
 * in: (planes,h,w), out: (planes,H,W), H >= h, W >= w.  The backward pass is a gather (no atomics). */
int uof_upsample_bilinear_fwd(const float* in, float* out, int planes, int h, int w, int H, int W, float scale,
                              uof_stream_t stream);
int uof_upsample_bilinear_bwd(const float* gout, float* gin, int planes, int h, int w, int H, int W, float scale,
                              uof_stream_t stream);

/* Decoder glue (pwc_tf.py:119-125): the up-sampled flow has three consumers (warp, concat, residual add).
 * fwd2: the values are also written to `out2` (nullable), a channel slice of a wider buffer: plane p goes to sample
 * p / C2, channel p % C2, batch stride `out2_batch_stride` floats (the `up_flow` slice of the concat buffer).
 * bwd3: the incoming gradient is gout + g2 + g3 (g2, g3 nullable): g2 addressed like out2 (the concat-slice gradient),
 * g3 dense like gout (the residual branch); summed on the fly. */
int uof_upsample_bilinear_fwd2(const float* in, float* out, float* out2, int C2, long long out2_batch_stride,
                               int planes, int h, int w, int H, int W, float scale, uof_stream_t stream);
int uof_upsample_bilinear_bwd3(const float* gout, const float* g2, int C2, long long g2_batch_stride, const float* g3,
                               float* gin, int planes, int h, int w, int H, int W, float scale, uof_stream_t stream);

/* a12: forward splat ("transformerFwd").  NOT in the reference.
 * u: (B,H,W,C) NHWC or NULL for a range map of ones (then C must be 1); flow: (B,H,W,2) in pixels;
 * out: (B,H,W,C), zero-filled here then accumulated with warp-aggregated fp32 atomics. */
int uof_splat_fwd(const float* u, const float* flow, float* out,
                  int B, int H, int W, int C, uof_stream_t stream);
/* gu (nullable) and gflow (nullable) are gathers from gout: no atomics. */
int uof_splat_bwd(const float* gout, const float* u, const float* flow, float* gu, float* gflow,
                  int B, int H, int W, int C, uof_stream_t stream);
/* integer contract: idx (B,H,W,4) int64 flat target indices b*H*W + yc*W + xc in corner order
 * (x0,y0),(x0,y1),(x1,y0),(x1,y1); -1 where the corner is out of bounds. */
int uof_splat_targets(const float* flow, long long* idx, int B, int H, int W, uof_stream_t stream);
/* visibility = clamp(range_map, 0, 1), in place on `range` (B*H*W elements). */
int uof_clamp01(float* range, long long n, uof_stream_t stream);

/* a13: forward-backward consistency mask.  NOT in the reference (App. D).
 * flows (B,2,H,W); mask (B,1,H,W) = |f_fwd + warp(f_rev, f_fwd)|_2 < max(alpha, beta*|f_fwd|_2). */
int uof_fb_consistency_mask(const float* flow_fwd, const float* flow_rev, float* mask,
                            int B, int H, int W, float alpha, float beta, int align_corners,
                            uof_stream_t stream);

/* ---- the data formats either side of the hot path ------------------------------------------------ */

/* Input pipeline.  Replaces KITTI_Prepared.resize_img / random_flip_img / preprocess_img and the transpose + .float() of
 * __getitem__ (core/dataset/kitti_prepared.py:63-91,146-153): the decoded uint8 strip of `nimg` vertically stacked BGR
 * images, (nimg*H0, W0, 3) per sample with `src_batch_stride` BYTES between samples, is cut into its images; each is resized
 * to (H, W) exactly as cv2.resize(INTER_LINEAR) does on 8-bit data (bit-exact: 11-bit fixed-point coefficients, OpenCV's
 * rounding), re-stacked, mirrored horizontally where flip[b] != 0 (flip: B device bytes, NULL = never), divided by 255 and
 * written as the (B, 3, nimg*H, W) fp32 batch Model_flow.forward takes. */
int uof_preprocess_u8(const unsigned char* src, long long src_batch_stride, const unsigned char* flip, float* out,
                      int B, int nimg, int H0, int W0, int H, int W, uof_stream_t stream);

/* KITTI 16-bit flow PNG arithmetic (core/evaluation/flowlib.py:107-138) on the decoded (H,W,3) uint16 array [u, v, valid]:
 * decode: flow[...,0:2] = (raw - 2^15) / 64 where valid, else 0; flow[...,2] = valid      -> (H,W,3) fp32
 * encode: raw = [clip(u*64 + 2^15, 0, 65535), clip(v*64 + 2^15, 0, 65535), 1] truncated; `flow` has `channel_stride` >= 2
 *         floats per pixel (2 for (H,W,2), 3 for a decoded (H,W,3) array). */
int uof_flow_png_decode(const unsigned short* raw, float* flow, long long npix, uof_stream_t stream);
int uof_flow_png_encode(const float* flow, int channel_stride, unsigned short* raw, long long npix, uof_stream_t stream);

/* Flow evaluation of one image.  Replaces the loop body of eval_flow_avg and calculate_error_rate
 * (core/evaluation/evaluate_flow.py:85-160): pred (2,h,w) planar fp32 in network-resolution pixels is rescaled by
 * (W/img_w, H/img_h), resized to the ground-truth size like cv2.resize(float32, INTER_LINEAR), and compared with
 * gt (H,W,3) fp32 [u, v, valid]; noc_mask (H,W); moving_mask (H,W) or NULL.  sums: 13 DEVICE doubles, zero-filled here:
 *   [0] sum epe*valid  [1] sum valid  [2] sum epe*noc  [3] sum noc  [4] sum epe*(valid-noc)  [5] sum (valid-noc)
 *   [6] outliers(valid)  [7] outliers(valid*move)  [8] sum valid*move  [9] outliers(valid*(1-move))
 *   [10] sum valid*(1-move)  [11] sum epe*valid*move  [12] sum epe*valid*(1-move)
 * with outlier = epe*m > 3 and epe*m / max(|gt|, 1e-10) > 0.05. */
#define UOF_FLOW_EVAL_SUMS 13
int uof_flow_eval(const float* pred, int h, int w, const float* gt, const float* noc_mask, const float* moving_mask,
                  int H, int W, int img_h, int img_w, double* sums, uof_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UOF_B200_H */
