"""Encoder / decoder skeleton around the CUDA operators.

Mirrors the reference's `core/networks/structures` API: `conv`, `FeaturePyramid`
(feature_pyramid.py:7-36), `PWC_tf` (pwc_tf.py:16-179) with its `corr` / `warp` seams, and
`warp_flow`.  By the north star the convolutions stay in PyTorch/cuDNN; `corr` and `warp` run the
hand-written sm_100a kernels.  State-dict keys equal the reference's 98 keys, so reference
checkpoints load unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..ops import warp_flow  # noqa: F401  (re-exported like structures/__init__.py:5)

_ENCODER = (16, 32, 64, 96, 128, 196)
_DENSE = (128, 128, 96, 64, 32)
_LEVEL_C = {6: 196, 5: 128, 4: 96, 3: 64, 2: 32}
_CONTEXT = ((128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1))


class ConvLReLU(nn.Sequential):
    """nn.Sequential(Conv2d(bias=True), LeakyReLU(0.1)) -- the reference's `conv()` block (net_utils.py:7-11), same
    modules and state-dict keys ('0.weight', '0.bias') -- executed as a bias-free cuDNN convolution followed by ONE fused
    bias + LeakyReLU kernel (in place), whose backward also produces the bias gradient.  Like every other operator of
    the package it has no CPU path."""

    def forward(self, x, fork=False, dsts=None):
        """`fork=True`: return the activation twice, one tensor per consumer (ops._BiasLeakyReLU).
        `dsts`: write the activation to one or two `(buffer, channel_offset)` destinations instead (`(None, 0)` = in place)
        and return one tensor per destination (ops.bias_leaky_relu_to) -- the concat-free dense block of PWC_tf._level."""
        cv, act = self[0], self[1]
        y = F.conv2d(x, cv.weight, None, cv.stride, cv.padding, cv.dilation, cv.groups)
        if dsts is not None:
            return ops.bias_leaky_relu_to(y, cv.bias, dsts, act.negative_slope)
        return ops.bias_leaky_relu_(y, cv.bias, act.negative_slope, fork)


def conv(in_planes, out_planes, kernel_size=3, stride=1, padding=1, dilation=1):
    """Conv2d + LeakyReLU(0.1) (net_utils.py:7-11)."""
    return ConvLReLU(nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                               dilation=dilation, bias=True), nn.LeakyReLU(0.1))


class FeaturePyramid(nn.Module):
    def __init__(self):
        super().__init__()
        cin = 3
        for k, cout in enumerate(_ENCODER):
            setattr(self, 'conv%d' % (2 * k + 1), conv(cin, cout, stride=2))
            setattr(self, 'conv%d' % (2 * k + 2), conv(cout, cout, stride=1))
            cin = cout

    def forward(self, img):
        feats, t = [], img
        last = len(_ENCODER) - 1
        for k in range(len(_ENCODER)):
            # every level but the last feeds the next level AND the decoder: fork (see ConvLReLU.forward)
            fork = k < last
            out = getattr(self, 'conv%d' % (2 * k + 2))(getattr(self, 'conv%d' % (2 * k + 1))(t), fork=fork)
            t, f = out if fork else (out, out)
            feats.append(f)
        return tuple(feats)


class PWC_tf(nn.Module):
    def __init__(self, md=4, align_corners=None):
        super().__init__()
        if md != 4:
            raise ValueError('the cost-volume kernel is built for md=4 (81 displacements)')
        self.align_corners = align_corners
        self.corr = self.corr_cuda          # seam, like `self.corr = self.corr_naive` (pwc_tf.py:19)
        nd = (2 * md + 1) ** 2
        w = _DENSE
        for lvl in (6, 5, 4, 3, 2):
            cin = nd if lvl == 6 else nd + _LEVEL_C[lvl] + 2
            ins = (cin, w[0], w[0] + w[1], w[1] + w[2], w[2] + w[3])
            for i in range(5):
                setattr(self, 'conv%d_%d' % (lvl, i), conv(ins[i], w[i]))
            setattr(self, 'predict_flow%d' % lvl, self.predict_flow(w[3] + w[4]))
        cin = w[4] + 2
        for i, (cout, dil) in enumerate(_CONTEXT):
            setattr(self, 'dc_conv%d' % (i + 1), conv(cin, cout, padding=dil, dilation=dil))
            cin = cout
        self.dc_conv7 = self.predict_flow(cin)

    def predict_flow(self, in_planes):
        return nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1, bias=True)

    @staticmethod
    def _flow_head(m, x):
        """A `predict_flow` convolution (no activation, pwc_tf.py:84-85) as a bias-free cuDNN convolution + the fused bias
        kernel with slope 1 (identity): its backward also produces the bias gradient -- ATen's sum over (N, H, W) for a
        2-channel tensor runs on two blocks (12-30 us per flow head)."""
        y = F.conv2d(x, m.weight, None, m.stride, m.padding, m.dilation, m.groups)
        return ops.bias_leaky_relu_(y, m.bias, 1.0, False)

    def corr_cuda(self, input1, input2):
        return ops.corr(input1, input2)

    corr_naive = corr_cuda                  # reference name (pwc_tf.py:97)

    def warp(self, x, flow):
        return ops.warp_flow(x, flow, use_mask=False, align_corners=self.align_corners)

    def _level(self, lvl, x):
        """pwc_tf.py:113-118 etc.: the five-convolution block whose inputs are concatenations of the two previous
        activations.  Every activation has two consumers; it is written once to each place it is read from -- the channel
        slices of the pre-allocated concat buffers (and, for the first one, in place for the next convolution) -- so that
        `torch.cat` never copies (ops.bias_leaky_relu_to / ops.cat_alias), and the fused LeakyReLU backward receives the two
        gradients separately and sums them itself."""
        w = _DENSE
        B, _, H, W = x.shape
        new = lambda c: torch.empty((B, c, H, W), device=x.device, dtype=torch.float32)
        b2, b3, b4, b5 = new(w[0] + w[1]), new(w[1] + w[2]), new(w[2] + w[3]), new(w[3] + w[4])
        conv = lambda i: getattr(self, 'conv%d_%d' % (lvl, i))
        x0a, x0b = conv(0)(x, dsts=[(None, 0), (b2, 0)])
        x1a, x1b = conv(1)(x0a, dsts=[(b2, w[0]), (b3, 0)])
        x2a, x2b = conv(2)(ops.cat_alias(b2, (x0b, x1a)), dsts=[(b3, w[1]), (b4, 0)])
        x3a, x3b = conv(3)(ops.cat_alias(b3, (x1b, x2a)), dsts=[(b4, w[2]), (b5, 0)])
        if lvl == 2:        # x4 also feeds the context network
            x4a, x4b = conv(4)(ops.cat_alias(b4, (x2b, x3a)), dsts=[(b5, w[3]), (None, 0)])
        else:
            (x4a,) = conv(4)(ops.cat_alias(b4, (x2b, x3a)), dsts=[(b5, w[3])])
            x4b = None
        return self._flow_head(getattr(self, 'predict_flow%d' % lvl), ops.cat_alias(b5, (x3b, x4a))), x4b

    def forward(self, feature_list_1, feature_list_2, img_hw):
        """pwc_tf.py:108-179.  The first feature list may carry 1/rep of the second one's batch: it is then taken as
        repeated along the batch (Model_flow.forward pairs [centre; centre] with [left; right]) without being copied."""
        flows, flow, x4 = {}, None, None
        fused = self.corr == self.corr_cuda          # default seam: the fused decoder-input node (ops.decoder_input)
        for lvl in (6, 5, 4, 3, 2):
            c1, c2 = feature_list_1[lvl - 1], feature_list_2[lvl - 1]
            rep = c2.shape[0] // c1.shape[0]
            if flow is None:
                flow, x4 = self._level(lvl, self.corr(c1.repeat(rep, 1, 1, 1) if rep > 1 else c1, c2))
            else:
                if fused and c2.shape[2] == 2 * flow.shape[2] and c2.shape[3] == 2 * flow.shape[3]:
                    # up-sampling, warp, cost volume and the concatenation as one node: the cost volume is written straight
                    # into the concat buffer and reads c1 from it, `up` is written into it by the up-sampling kernel
                    x, up = ops.decoder_input(c1, c2, flow, self.align_corners)
                else:                                # a user-installed corr keeps the reference's three-step form
                    up = ops.upsample_bilinear_scaled(flow, (2 * flow.shape[2], 2 * flow.shape[3]), 2.0)      # pwc_tf.py:119
                    c1r = c1.repeat(rep, 1, 1, 1) if rep > 1 else c1
                    x = torch.cat((self.corr(c1r, self.warp(c2, up)), c1r, up), 1)
                res, x4 = self._level(lvl, x)
                flow = res + up
            flows[lvl] = flow
        t = torch.cat((flows[2], x4), 1)
        for i in range(1, 7):
            t = getattr(self, 'dc_conv%d' % i)(t)
        flows[2] = flows[2] + self._flow_head(self.dc_conv7, t)
        img_h, img_w = img_hw[0], img_hw[1]
        # pwc_tf.py:174-177: F.interpolate(flow * 4.0, size) -- the power-of-two scale commutes exactly with the interpolation
        return [ops.upsample_bilinear_scaled(flows[2 + s], (img_h // 2 ** s, img_w // 2 ** s), 4.0) for s in range(4)]
