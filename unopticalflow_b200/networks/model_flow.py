"""`Model_flow` with the reference's interface (model_flow_paper.py:14-255) on the CUDA hot path.

`forward(inputs) -> {'loss_pixel','loss_ssim','loss_flow_smooth','loss_flow_consis'}` each `(B,)`,
`inference_flow(img1, img2) -> (B,2,H,W)`, the `compute_*` / pyramid helper methods, and the same
98 state-dict keys.  What changes is how the step is executed:

* the three encoder passes run as one 3B batch and the two decoder passes as one 2B batch
  ([centre;centre] vs [left;right]) -- the per-sample math is unchanged;
* cost volume, warps, image pyramid and every loss run as hand-written sm_100a kernels
  (`unopticalflow_b200.ops`), with all pyramid levels and both directions of a loss in one launch;
* scale 3 is not warped in `forward` because no loss reads it.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .structures import FeaturePyramid, PWC_tf


class Model_flow(nn.Module):
    def __init__(self, cfg, align_corners=None):
        super().__init__()
        self.align_corners = ops.DEFAULT_ALIGN_CORNERS if align_corners is None else bool(align_corners)
        self.fpyramid = FeaturePyramid()
        self.pwc_model = PWC_tf(align_corners=self.align_corners)
        if cfg.mode == 'depth' or cfg.mode == 'flowposenet':        # model_flow_paper.py:19-24
            for p in self.parameters():
                p.requires_grad = False
        self.dataset = cfg.dataset
        self.num_scales = cfg.num_scales
        self.flow_consist_alpha = cfg.h_flow_consist_alpha
        self.flow_consist_beta = cfg.h_flow_consist_beta

    # ---- reference helper API (list-of-tensors in, (B,) out) ------------------------------------
    def get_flow_norm(self, flow, p=2):
        return torch.norm(flow, p=p, dim=1).unsqueeze(1) + 1e-12

    def get_flow_normalization(self, flow, p=2):
        return flow / self.get_flow_norm(flow, p).repeat(1, 2, 1, 1)

    def generate_img_pyramid(self, img, num_pyramid):
        return ops.img_pyramid(img, num_pyramid)

    def warp_flow_pyramid(self, img_pyramid, flow_pyramid):
        return [ops.warp_flow(i, f, use_mask=True, align_corners=self.align_corners)
                for i, f in zip(img_pyramid, flow_pyramid)]

    def compute_diff_weight(self, img_pyramid_from_l, img_pyramid, img_pyramid_from_r):
        """-> diff_bwd, diff_fwd, weight_bwd, weight_fwd (model_flow_paper.py:101-134); weights detached as in the
        reference, diffs differentiable w.r.t. the warped images."""
        return ops.diff_weight(img_pyramid_from_l, img_pyramid, img_pyramid_from_r, self.num_scales)

    def compute_loss_pixel(self, img_pyramid, img_warped_pyramid, occ_mask_list):
        """model_flow_paper.py:68-77 (not called by the reference's forward): masked mean of |img - warped| over 3 channels."""
        diffs = [torch.abs(img_pyramid[s] - img_warped_pyramid[s]) for s in range(self.num_scales)]
        return ops.loss_with_mask(diffs, occ_mask_list, self.num_scales)

    def compute_loss_pixel_without_mask(self, img_pyramid, img_warped_pyramid):
        """model_flow_paper.py:79-87 (not called by the reference's forward)."""
        diffs = [torch.abs(img_pyramid[s] - img_warped_pyramid[s]) for s in range(self.num_scales)]
        ones = [torch.ones_like(d[:, :1]) for d in diffs]
        return ops.loss_with_mask(diffs, ones, self.num_scales) * (1.0 + 1e-12)      # the masked form divides by mean(1) + 1e-12

    def compute_loss_with_mask(self, diff_list, occ_mask_list):
        return ops.loss_with_mask(diff_list, occ_mask_list, self.num_scales)

    def compute_loss_ssim(self, img_pyramid, img_warped_pyramid, occ_mask_list):
        total = 0
        for s in range(self.num_scales):
            img, wp, m = img_pyramid[s], img_warped_pyramid[s], occ_mask_list[s]
            s_map = ops.SSIM(img * m, wp * m)
            total = total + torch.clamp((1.0 - s_map) / 2.0, 0, 1).mean((1, 2, 3)) / (m.mean((1, 2, 3)) + 1e-12)
        return total

    def gradients(self, img):
        """model_flow_paper.py:152-155: forward differences (dx, dy)."""
        return img[:, :, :, 1:] - img[:, :, :, :-1], img[:, :, 1:, :] - img[:, :, :-1, :]

    def cal_grad2_error(self, flow, img):
        """model_flow_paper.py:157-167: edge-aware second-order smoothness of ONE level, (B,).  The fused kernel computes the
        reference's `cal_grad2_error(flow / 20, img)` (compute_loss_flow_smooth); the factor is undone on the way in."""
        return ops.flow_smooth_loss([flow * 20.0], [img], 1)

    def compute_loss_flow_smooth(self, optical_flows, img_pyramid):
        return ops.flow_smooth_loss(optical_flows, img_pyramid, self.num_scales)

    def compute_loss_flow_consis(self, fwd_flow_pyramid, bwd_flow_pyramid, occ_mask_list):
        return ops.flow_consis_loss(fwd_flow_pyramid, bwd_flow_pyramid, occ_mask_list, self.num_scales)

    # ---- north-star extras (absent from the reference) -----------------------------
    def get_occlusion_mask_from_flow(self, flow):
        """(B,2,H,W) flow -> (B,1,H,W) soft visibility clamp(range_map, 0, 1)."""
        return ops.occlusion_mask(flow.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)

    def get_consistent_mask(self, flow_fwd, flow_rev):
        return ops.fb_consistency_mask(flow_fwd, flow_rev, self.flow_consist_alpha, self.flow_consist_beta,
                                       self.align_corners)

    # ---- inference -----------------------------------------------------------------------------------
    def inference_flow(self, img1, img2):
        B = img1.shape[0]
        feats = self.fpyramid(torch.cat((img1, img2), 0))
        return self.pwc_model([f[:B] for f in feats], [f[B:] for f in feats], [img1.shape[2], img1.shape[3]])[0]

    # ---- training step ----------------------------------------------------------------------------------
    def forward(self, inputs, output_flow=False, use_flow_loss=True, is_second_phase=False):
        """model_flow_paper.py:205-255.  `use_flow_loss` / `is_second_phase` are accepted and ignored, as the reference's
        forward does (it never reads them); `output_flow=True` additionally returns the forward and backward flow pyramids
        (the reference accepts the flag and returns the loss pack only)."""
        assert inputs.shape[1] == 3
        B, H, W = inputs.shape[0], int(inputs.shape[2] / 3), inputs.shape[3]
        S = self.num_scales
        stacked = ops.img_pyramid_triplet_stacked(inputs, S)
        if stacked is not None:
            # one launch: every pyramid level (level 0 as a dense copy) stacked as [left; right; centre], so the encoder's
            # 3B batch and the [left; right] warp sources are views -- no torch.cat of the images anywhere
            feats = self.fpyramid(stacked[0].view(3 * B, 3, H, W))                   # one 3B encoder pass
            # [left ; right] is the leading 2B block of the stacked batch and the centre features the trailing B: two views
            # (no concatenation forward, one concatenation of the two gradients backward)
            parts = [ops.split_at(f, 2 * B) for f in feats]
            f1 = [c for _, c in parts]                                               # centre, paired with both (PWC_tf.forward repeats it)
            f2 = [lr for lr, _ in parts]                                             # [left   ; right ]
            pyr_c = [t[2] for t in stacked]
            sources = [t[:2].reshape(2 * B, 3, t.shape[3], t.shape[4]) for t in stacked]
        else:
            imgl, img, imgr = inputs[:, :, :H], inputs[:, :, H:2 * H], inputs[:, :, 2 * H:3 * H]
            feats = self.fpyramid(torch.cat((imgl, img, imgr), 0))
            # split (not three slices): its backward is ONE concatenation of the three gradients instead of three
            # zero-filled full-size tensors plus adds
            parts = [f.split(B, 0) for f in feats]                                   # (left, centre, right)
            f1 = [c for _, c, _ in parts]
            f2 = [torch.cat((l, r), 0) for l, _, r in parts]
            pyr_l, pyr_c, pyr_r, _ = ops.img_pyramid_triplet(inputs, S)              # one launch, triplet read in place
            sources = [torch.cat((pyr_l[s], pyr_r[s]), 0) for s in range(S)]
        flows = self.pwc_model(f1, f2, [H, W])                                       # (2B,2,h,w): [bwd ; fwd]
        if ops.FUSE_IMAGE_WARP:
            # one autograd node for the four losses: image warps evaluated inside the photometric kernels (one launch each
            # way instead of 3 + 1) and the three flow gradients accumulated into one buffer by the backward kernels
            loss_pixel, loss_ssim, smooth, consis, w_bwd, w_fwd = ops.flow_loss_pack(pyr_c, sources, flows, S, self.align_corners)
        else:
            warped = [ops.warp_flow(sources[s], flows[s], use_mask=True, align_corners=self.align_corners)
                      for s in range(S)]                                              # [from_l ; from_r]
            loss_pixel, loss_ssim, w_bwd, w_fwd = ops.photometric_losses_stacked(pyr_c, warped, S)
            smooth2 = ops.flow_smooth_loss(flows, pyr_c, S)                          # (2B,): [bwd ; fwd]
            smooth = smooth2[B:] + smooth2[:B]
            consis = ops.flow_consis_loss([f[B:] for f in flows[:S]], [f[:B] for f in flows[:S]], w_fwd, S)
        loss_pack = {'loss_pixel': loss_pixel, 'loss_ssim': loss_ssim, 'loss_flow_smooth': smooth, 'loss_flow_consis': consis}
        if output_flow:
            return loss_pack, [f[B:] for f in flows], [f[:B] for f in flows]
        return loss_pack
