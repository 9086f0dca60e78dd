"""Training-step semantics of the reference driver, restated for the B200 path.

train.py:35-39 (model + Adam), :137-152 (zero_grad / forward / weighted loss / backward / step) and
core/config/config_utils.py:3-9 (loss-weight dict).  The reference's DataParallel (train.py:36-37) becomes
one process per GPU with DistributedDataParallel over NCCL.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

KITTI_CFG = SimpleNamespace(mode='flow', dataset='kitti_depth', num_scales=3, h_flow_consist_alpha=3.0,
                            h_flow_consist_beta=0.05, w_ssim=0.85, w_flow_smooth=10.0, w_flow_consis=0.01,
                            img_hw=(256, 832), lr=1e-4, batch_size=8)          # config/kitti.yaml:11-38, train.py:168,170
SINTEL_CFG = SimpleNamespace(mode='flow', dataset='sintel', num_scales=3, h_flow_consist_alpha=3.0,
                             h_flow_consist_beta=0.05, w_ssim=0.85, w_flow_smooth=6.0, w_flow_consis=0.01,
                             img_hw=(384, 832), lr=1e-4, batch_size=8)         # config/sintel.yaml


def generate_loss_weights_dict(cfg):
    """core/config/config_utils.py:3-9."""
    return {'loss_pixel': 1 - cfg.w_ssim, 'loss_ssim': cfg.w_ssim, 'loss_flow_smooth': cfg.w_flow_smooth,
            'loss_flow_consis': cfg.w_flow_consis}


def total_loss(loss_pack, weights):
    """train.py:147-150: sum_k w_k * mean_B(loss_k)."""
    return torch.stack([weights[k] * loss_pack[k].mean() for k in loss_pack]).sum()


def make_optimizer(model, lr=1e-4, fused=True):
    """train.py:39 -- Adam with default betas/eps; `fused` uses the single-launch multi-tensor CUDA Adam."""
    params = [p for p in model.parameters() if p.requires_grad]
    return torch.optim.Adam([{'params': params, 'lr': lr}], fused=fused and params[0].is_cuda)


def train_step(model, optimizer, inputs, weights):
    """One iteration of train.py:137-152.  Returns the scalar loss tensor (no host sync)."""
    optimizer.zero_grad(set_to_none=True)
    loss_pack = model(inputs)
    loss = total_loss(loss_pack, weights)
    loss.backward()
    optimizer.step()
    return loss.detach()
