"""Flow evaluation on the GPU.

Mirrors `core/evaluation/evaluate_flow.py` (calculate_error_rate :85-91, eval_flow_avg :93-174, load_gt_flow_kitti :59-83)
and the KITTI flow-PNG arithmetic of `core/evaluation/flowlib.py` (read_flow_png :107-128, write_flow_png :131-145) with
the same names, argument meaning and report string.  The reference rescales and `cv2.resize`s every prediction on the
CPU, materialises the EPE map and half a dozen masked temporaries in float64 numpy and sums them; here one
`uof_flow_eval` launch per image does the rescale, the bilinear resize (OpenCV's float algorithm), the EPE, the masked
sums and the 3 px / 5 % outlier counts with fp64 accumulators, and only 13 doubles per image come back to the host.
PNG container decoding / encoding stays on the CPU (cv2); the 16-bit arithmetic runs on the device.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib

N_SUMS = 13      # include/uof_b200.h UOF_FLOW_EVAL_SUMS


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dev(device=None):
    dev = torch.device(device if device is not None else 'cuda')
    if dev.type != 'cuda':
        raise RuntimeError('unopticalflow_b200.evaluation runs on a CUDA device; there is no CPU fallback')
    return dev


def _f32_cuda(a, dev):
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a
    return t.to(device=dev, dtype=torch.float32).contiguous()


# ------------------------------------------------------------------------------------------ flow PNG
def decode_flow_png(raw, device=None):
    """(H,W,3) uint16 [u, v, valid] (numpy, or an int16/uint16 tensor holding the same bits) -> (H,W,3) fp32 CUDA tensor
    [(u - 2^15)/64, (v - 2^15)/64, valid] with invalid pixels zeroed (flowlib.py:119-127)."""
    dev = _dev(device)
    if isinstance(raw, np.ndarray):
        if raw.dtype != np.uint16:
            raise TypeError('decode_flow_png takes the decoded uint16 array (got %s)' % raw.dtype)
        raw = torch.from_numpy(np.ascontiguousarray(raw).view(np.int16))
    raw = raw.to(dev).contiguous()
    if raw.dim() != 3 or raw.shape[2] != 3 or raw.element_size() != 2:
        raise ValueError('decode_flow_png: expected (H,W,3) 16-bit data, got %r' % (tuple(raw.shape),))
    H, W, _ = raw.shape
    flow = torch.empty((H, W, 3), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.call('uof_flow_png_decode', ctypes.c_void_p(raw.data_ptr()), ctypes.c_void_p(flow.data_ptr()), H * W, _stream(dev))
    return flow


def encode_flow_png(flo, device=None):
    """(H,W,>=2) float flow -> (H,W,3) uint16 numpy array as write_flow_png stores it (flowlib.py:131-138)."""
    dev = _dev(device)
    f = _f32_cuda(flo, dev)
    if f.dim() != 3 or f.shape[2] < 2:
        raise ValueError('encode_flow_png: expected (H,W,>=2), got %r' % (tuple(f.shape),))
    H, W, C = f.shape
    raw = torch.empty((H, W, 3), device=dev, dtype=torch.int16)
    with torch.cuda.device(dev):
        _lib.call('uof_flow_png_encode', ctypes.c_void_p(f.data_ptr()), C, ctypes.c_void_p(raw.data_ptr()), H * W, _stream(dev))
    return raw.cpu().numpy().view(np.uint16)


def read_flow_png(flow_file, device=None):
    """flowlib.py:107-128 -> (H,W,3) fp32 CUDA tensor.  The file is decoded with cv2 (BGR order, hence the reversal)."""
    import cv2
    raw = cv2.imread(flow_file, cv2.IMREAD_UNCHANGED)
    if raw is None or raw.dtype != np.uint16 or raw.ndim != 3:
        raise ValueError('%s is not a 16-bit 3-channel PNG' % flow_file)
    return decode_flow_png(np.ascontiguousarray(raw[:, :, ::-1]), device)


def write_flow_png(flo, flow_file, device=None):
    """flowlib.py:131-145."""
    import cv2
    raw = encode_flow_png(flo, device)
    cv2.imwrite(flow_file, np.ascontiguousarray(raw[:, :, ::-1]))


# ------------------------------------------------------------------------------------------ evaluation
def flow_eval_sums(pred_flow, gt_flow, noc_mask, img_hw, moving_mask=None, device=None):
    """The 13 masked sums of one image (include/uof_b200.h, uof_flow_eval) as a CUDA fp64 tensor.
    pred_flow: (h,w,2) numpy / tensor as the reference passes it, or a (2,h,w) CUDA tensor (Model_flow.inference_flow's
    output for one sample); gt_flow (H,W,3); noc_mask, moving_mask (H,W)."""
    dev = _dev(device)
    p = _f32_cuda(pred_flow, dev)
    if p.dim() != 3:
        raise ValueError('flow_eval: pred_flow must be (h,w,2) or (2,h,w)')
    if p.shape[2] == 2 and p.shape[0] != 2:
        p = p.permute(2, 0, 1).contiguous()
    elif p.shape[0] != 2:
        raise ValueError('flow_eval: pred_flow must be (h,w,2) or (2,h,w), got %r' % (tuple(p.shape),))
    gt = _f32_cuda(gt_flow, dev)
    noc = _f32_cuda(noc_mask, dev)
    mv = _f32_cuda(moving_mask, dev) if moving_mask is not None else None
    H, W = gt.shape[0], gt.shape[1]
    if gt.shape[2] != 3 or tuple(noc.shape) != (H, W) or (mv is not None and tuple(mv.shape) != (H, W)):
        raise ValueError('flow_eval: gt_flow must be (H,W,3) and the masks (H,W)')
    sums = torch.empty(N_SUMS, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        _lib.call('uof_flow_eval', ctypes.c_void_p(p.data_ptr()), int(p.shape[1]), int(p.shape[2]), ctypes.c_void_p(gt.data_ptr()),
                  ctypes.c_void_p(noc.data_ptr()), ctypes.c_void_p(mv.data_ptr() if mv is not None else 0), H, W,
                  int(img_hw[0]), int(img_hw[1]), ctypes.c_void_p(sums.data_ptr()), _stream(dev))
    return sums


def calculate_error_rate(epe_map, gt_flow, mask):
    """evaluate_flow.py:85-91 on tensors (any device): fraction of `mask` pixels with EPE > 3 px and > 5 % of |gt|."""
    epe_map, gt_flow, mask = (torch.as_tensor(t).double() for t in (epe_map, gt_flow, mask))
    mag = torch.clamp(torch.sqrt((gt_flow ** 2).sum(2)), min=1e-10)
    bad = (epe_map * mask > 3) & (epe_map * mask / mag > 0.05)
    return float(bad.sum() / mask.sum())


def eval_flow_avg(gt_flows, noc_masks, pred_flows, cfg, moving_masks=None, write_img=False):
    """Drop-in for evaluate_flow.py:93-174: same arguments (cfg.img_hw is read), same two-line report string.
    `write_img` (the colour-coded dumps) is not supported."""
    if write_img:
        raise NotImplementedError('eval_flow_avg(write_img=True): flow visualisation is outside the hot path (DESIGN.md section 7)')
    num = len(gt_flows)
    all_sums = [flow_eval_sums(pred_flows[i], gt_flows[i], noc_masks[i], cfg.img_hw, moving_masks[i] if moving_masks else None)
                for i in range(num)]
    S = torch.stack(all_sums).cpu().numpy()                                   # one read-back for the whole set
    error = error_noc = error_occ = error_move = error_static = error_rate = 0.0
    error_move_rate = error_static_rate = 0.0
    for s in S:
        error += s[0] / s[1]                                                   # :135
        error_noc += s[2] / s[3]                                               # :137
        error_occ += s[4] / max(s[5], 1.0)                                     # :139-140
        error_rate += s[6] / s[1]                                              # :142-143
        if moving_masks:
            error_move_rate += s[7] / s[8]                                     # :148-149
            error_static_rate += s[9] / s[10]
            error_move += s[11] / s[8]                                         # :154-156
            error_static += s[12] / s[10]
    if moving_masks:
        result = "{:>10}, {:>10}, {:>10}, {:>10}, {:>10}, {:>10}, {:>10}, {:>10} \n".format(
            'epe', 'epe_noc', 'epe_occ', 'epe_move', 'epe_static', 'move_err_rate', 'static_err_rate', 'err_rate')
        result += "{:10.4f}, {:10.4f}, {:10.4f}, {:10.4f}, {:10.4f}, {:10.4f}, {:10.4f}, {:10.4f} \n".format(
            error / num, error_noc / num, error_occ / num, error_move / num, error_static / num, error_move_rate / num,
            error_static_rate / num, error_rate / num)
        return result
    result = "{:>10}, {:>10}, {:>10}, {:>10} \n".format('epe', 'epe_noc', 'epe_occ', 'err_rate')
    result += "{:10.4f}, {:10.4f}, {:10.4f}, {:10.4f} \n".format(error / num, error_noc / num, error_occ / num, error_rate / num)
    return result


def load_gt_flow_kitti(gt_dataset_dir, mode, device=None):
    """evaluate_flow.py:59-83: -> (gt_flows, noc_masks) as lists of CUDA tensors ((H,W,3) and (H,W))."""
    if mode == 'kitti_2012':
        num_gt = 194
    elif mode == 'kitti_2015':
        num_gt = 200
    else:
        raise ValueError('Mode {} not found.'.format(mode))
    gt_flows, noc_masks = [], []
    for i in range(num_gt):
        gt_flows.append(read_flow_png(os.path.join(gt_dataset_dir, 'flow_occ', str(i).zfill(6) + '_10.png'), device))
        noc_masks.append(read_flow_png(os.path.join(gt_dataset_dir, 'flow_noc', str(i).zfill(6) + '_10.png'), device)[:, :, 2].contiguous())
    return gt_flows, noc_masks
