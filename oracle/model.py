"""Model-level oracle (TEST INFRASTRUCTURE — see oracle/__init__.py).

A CPU/fp32 restatement of the reference's `Model_flow` training step built only from
`oracle.ops` and stock `torch.nn` convolutions.  State-dict keys equal the reference's
(98 tensors, SURVEY section 5 "checkpoint"), so weights can be exchanged with both the
reference and the product model.  Table-driven on purpose: it is a restatement of what
the reference computes, not a transcription of its source.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

ENCODER_CHANNELS = (16, 32, 64, 96, 128, 196)      # feature_pyramid.py:10-21
DECODER_WIDTHS = (128, 128, 96, 64, 32)            # pwc_tf.py:26 (dd)
LEVEL_FEATURES = {6: 0, 5: 128, 4: 96, 3: 64, 2: 32}   # c1k channels concatenated at level k (pwc_tf.py:38-70)
CONTEXT_NET = ((128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1))   # pwc_tf.py:76-81 (out, dilation)


def conv_lrelu(cin, cout, stride=1, dilation=1):
    """3x3 conv + LeakyReLU(0.1) as an nn.Sequential so keys read '<name>.0.weight' (net_utils.py:7-11)."""
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, dilation, dilation, bias=True), nn.LeakyReLU(0.1))


class FeaturePyramid(nn.Module):
    """12 convs, strides 2/1 alternating; returns the six stride-1 outputs.  feature_pyramid.py:7-36."""

    def __init__(self):
        super().__init__()
        cin = 3
        for k, cout in enumerate(ENCODER_CHANNELS):
            setattr(self, 'conv%d' % (2 * k + 1), conv_lrelu(cin, cout, stride=2))
            setattr(self, 'conv%d' % (2 * k + 2), conv_lrelu(cout, cout, stride=1))
            cin = cout

    def forward(self, img):
        feats, t = [], img
        for k in range(len(ENCODER_CHANNELS)):
            t = getattr(self, 'conv%d' % (2 * k + 2))(getattr(self, 'conv%d' % (2 * k + 1))(t))
            feats.append(t)
        return tuple(feats)


class PWC_tf(nn.Module):
    """PWC-Net decoder.  pwc_tf.py:16-179.  `corr` / `warp` are the operator seams."""

    def __init__(self, md=ops.MAX_DISP, align_corners=False):
        super().__init__()
        self.align_corners = align_corners
        nd = (2 * md + 1) ** 2
        w = DECODER_WIDTHS
        for lvl in (6, 5, 4, 3, 2):
            cin = nd + (LEVEL_FEATURES[lvl] + 2 if lvl < 6 else 0)
            ins = (cin, w[0], w[0] + w[1], w[1] + w[2], w[2] + w[3])
            for i in range(5):
                setattr(self, 'conv%d_%d' % (lvl, i), conv_lrelu(ins[i], w[i]))
            setattr(self, 'predict_flow%d' % lvl, nn.Conv2d(w[3] + w[4], 2, 3, 1, 1, bias=True))
        cin = w[4] + 2
        for i, (cout, dil) in enumerate(CONTEXT_NET):
            setattr(self, 'dc_conv%d' % (i + 1), conv_lrelu(cin, cout, dilation=dil))
            cin = cout
        self.dc_conv7 = nn.Conv2d(cin, 2, 3, 1, 1, bias=True)

    def corr(self, a, b):
        return ops.cost_volume(a, b)

    def warp(self, x, flow):
        return ops.warp_flow(x, flow, use_mask=False, align_corners=self.align_corners)

    def _estimate(self, lvl, x):
        """The five 'dense-ish' convs of one level: each conv sees the previous two outputs."""
        x0 = getattr(self, 'conv%d_0' % lvl)(x)
        x1 = getattr(self, 'conv%d_1' % lvl)(x0)
        x2 = getattr(self, 'conv%d_2' % lvl)(torch.cat((x0, x1), 1))
        x3 = getattr(self, 'conv%d_3' % lvl)(torch.cat((x1, x2), 1))
        x4 = getattr(self, 'conv%d_4' % lvl)(torch.cat((x2, x3), 1))
        return getattr(self, 'predict_flow%d' % lvl)(torch.cat((x3, x4), 1)), x4

    def forward(self, feats1, feats2, img_hw):
        flows, up, x4 = {}, None, None
        for lvl in (6, 5, 4, 3, 2):
            c1, c2 = feats1[lvl - 1], feats2[lvl - 1]
            if up is None:
                flow, x4 = self._estimate(lvl, self.corr(c1, c2))                    # pwc_tf.py:112-118
            else:
                cv = self.corr(c1, self.warp(c2, up))                                 # :121-122
                res, x4 = self._estimate(lvl, torch.cat((cv, c1, up), 1))             # :123-129
                flow = res + up                                                       # :130
            flows[lvl] = flow
            if lvl > 2:
                up = F.interpolate(flow, scale_factor=2.0, mode='bilinear') * 2.0     # :119,131,144,156
        t = torch.cat((flows[2], x4), 1)
        for i in range(1, 7):
            t = getattr(self, 'dc_conv%d' % i)(t)
        flows[2] = flows[2] + self.dc_conv7(t)                                        # :170-171
        H, W = img_hw[0], img_hw[1]
        return [F.interpolate(flows[2 + s] * 4.0, [H // 2 ** s, W // 2 ** s], mode='bilinear')
                for s in range(4)]                                                    # :173-177


class Model_flow(nn.Module):
    """model_flow_paper.py:14-255."""

    def __init__(self, cfg, align_corners=False):
        super().__init__()
        self.align_corners = align_corners
        self.fpyramid = FeaturePyramid()
        self.pwc_model = PWC_tf(align_corners=align_corners)
        self.dataset = cfg.dataset
        self.num_scales = cfg.num_scales
        self.flow_consist_alpha = cfg.h_flow_consist_alpha
        self.flow_consist_beta = cfg.h_flow_consist_beta

    def inference_flow(self, img1, img2):
        hw = [img1.shape[2], img1.shape[3]]
        return self.pwc_model(self.fpyramid(img1), self.fpyramid(img2), hw)[0]       # :198-202

    def forward(self, inputs):
        assert inputs.shape[1] == 3
        H = int(inputs.shape[2] / 3)
        W = inputs.shape[3]
        imgl, img, imgr = inputs[:, :, :H], inputs[:, :, H:2 * H], inputs[:, :, 2 * H:3 * H]   # :206-209
        fl, fc, fr = self.fpyramid(imgl), self.fpyramid(img), self.fpyramid(imgr)               # :214
        flows_bwd = self.pwc_model(fc, fl, [H, W])                                              # :216
        flows_fwd = self.pwc_model(fc, fr, [H, W])                                              # :218
        n = len(flows_fwd)
        pl, pc, pr = ops.img_pyramid(imgl, n), ops.img_pyramid(img, n), ops.img_pyramid(imgr, n)   # :229-231
        from_l = ops.warp_pyramid(pl, flows_bwd, self.align_corners)                            # :233
        from_r = ops.warp_pyramid(pr, flows_fwd, self.align_corners)                            # :235
        S = self.num_scales
        d_b, d_f, w_b, w_f = ops.diff_weight(from_l, pc, from_r, S)                             # :240
        return {
            'loss_pixel': ops.loss_with_mask(d_f, w_f, S) + ops.loss_with_mask(d_b, w_b, S),    # :241-242
            'loss_ssim': ops.loss_ssim(pc, from_r, w_f, S) + ops.loss_ssim(pc, from_l, w_b, S),  # :244-245
            'loss_flow_smooth': ops.loss_flow_smooth(flows_fwd, pc, S) + ops.loss_flow_smooth(flows_bwd, pc, S),  # :248-249
            'loss_flow_consis': ops.loss_flow_consis(flows_fwd, flows_bwd, w_f, S),             # :251
        }


class Cfg:
    """The five fields Model_flow reads (model_flow_paper.py:19-30) with kitti.yaml values."""
    mode = 'flow'
    dataset = 'kitti_depth'
    num_scales = 3
    h_flow_consist_alpha = 3.0
    h_flow_consist_beta = 0.05


def train_step(model, optimizer, inputs, weights=(0.85, 10.0, 0.01)):
    """train.py:137-152: zero_grad, forward, weighted sum of batch means, backward, Adam step."""
    optimizer.zero_grad()
    pack = model(inputs)
    loss = ops.total_loss(pack, *weights)
    loss.backward()
    optimizer.step()
    return loss.detach(), {k: v.detach() for k, v in pack.items()}


def make_optimizer(model, lr=1e-4):
    """train.py:39 (Adam, defaults otherwise)."""
    return torch.optim.Adam([{'params': [p for p in model.parameters() if p.requires_grad], 'lr': lr}])
