"""Input pipeline of the training step on the GPU.

Mirrors `core/dataset/kitti_prepared.py` of the reference: `KITTI_Prepared(data_dir, num_scales, img_hw, num_iterations)`
with `get_data_list`, `rand_num`, `__len__`, `read_cam_intrinsic`, `rescale_intrinsics`, `get_multiscale_intrinsics`, and the
pre-processing chain `resize_img` -> `random_flip_img` -> `/ 255.0` -> `transpose(2,0,1)` -> `.float()` (:63-91, :146-153).
What changes is where that chain runs: the reference does three `cv2.resize` calls, a flip, a float64 division and a
transpose per sample on the CPU (~15 ms) and ships 7.7 MB of fp32 per triplet to the GPU; here `__getitem__` returns the
DECODED uint8 strip (4.2 MB at KITTI size) and `collate` / `preprocess_batch` turn a batch of strips into the
(B,3,3H,W) fp32 tensor with one `uof_preprocess_u8` launch per sample -- bit-identical to the reference's result
(tests/golden/io_preprocess_*.npz, recorded from the unmodified class).  File decoding stays on the CPU (cv2.imread).
There is no CPU fallback for the pre-processing itself.
"""
from __future__ import annotations

import copy
import ctypes
import os

import numpy as np
import torch
import torch.utils.data

from . import _lib


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _as_cuda_u8(img, device):
    """numpy / CPU uint8 (nimg*H0, W0, 3) -> contiguous CUDA uint8 tensor (pinned staging for host arrays)."""
    if isinstance(img, np.ndarray):
        if img.dtype != np.uint8:
            raise TypeError('preprocess_img takes the decoded uint8 image (got %s)' % img.dtype)
        img = torch.from_numpy(np.ascontiguousarray(img))
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3:
        raise ValueError('preprocess_img: expected a (nimg*H, W, 3) uint8 image, got %s %r' % (img.dtype, tuple(img.shape)))
    if not img.is_cuda:
        img = img.contiguous()
        if torch.cuda.is_available():
            img = img.pin_memory()
        img = img.to(device, non_blocking=True)
    return img.contiguous()


def preprocess_batch(imgs, img_hw, flips=None, nimg=3, device=None, out=None):
    """Batch form of KITTI_Prepared.preprocess_img + the tail of __getitem__ (kitti_prepared.py:83-91,146-153).

    imgs: sequence of decoded strips, each (nimg*H0_i, W0_i, 3) uint8 (numpy, CPU or CUDA tensor; sizes may differ per
    sample, as KITTI's do).  flips: per-sample booleans (the reference's coin, `random_flip_img`), None = no flip.
    -> (B, 3, nimg*H, W) fp32 CUDA tensor, equal to torch.stack of the reference's per-sample results."""
    dev = torch.device(device if device is not None else 'cuda')
    if dev.type != 'cuda':
        raise RuntimeError('unopticalflow_b200.dataset.preprocess_batch runs on a CUDA device; there is no CPU fallback')
    B = len(imgs)
    H, W = int(img_hw[0]), int(img_hw[1])
    if out is None:
        out = torch.empty((B, 3, nimg * H, W), device=dev, dtype=torch.float32)
    if tuple(out.shape) != (B, 3, nimg * H, W) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError('preprocess_batch: `out` must be a contiguous (B,3,%d,%d) fp32 tensor' % (nimg * H, W))
    flip_t = None
    if flips is not None:
        flip_t = torch.tensor([1 if f else 0 for f in flips], dtype=torch.uint8).to(dev)
    per = 3 * nimg * H * W * 4
    keep = []
    with torch.cuda.device(dev):
        for b, im in enumerate(imgs):
            src = _as_cuda_u8(im, dev)
            keep.append(src)
            H0 = int(src.shape[0] / nimg)                                   # kitti_prepared.py:70
            W0 = int(src.shape[1])
            fp = ctypes.c_void_p(flip_t.data_ptr() + b) if flip_t is not None else ctypes.c_void_p(0)
            _lib.call('uof_preprocess_u8', ctypes.c_void_p(src.data_ptr()), src.numel(), fp,
                      ctypes.c_void_p(out.data_ptr() + b * per), 1, nimg, H0, W0, H, W, _stream(dev))
        for t in keep:                                                          # the staging buffers outlive this call's launches
            t.record_stream(torch.cuda.current_stream(dev))
    return out


def preprocess_img(img, img_hw, flip=False, nimg=3, device=None):
    """One sample: == torch.from_numpy(KITTI_Prepared.preprocess_img(img, img_hw).transpose(2,0,1)).float() with the
    reference's coin replaced by `flip`; returns a (3, nimg*H, W) fp32 CUDA tensor."""
    return preprocess_batch([img], img_hw, [flip], nimg, device)[0]


class KITTI_Prepared(torch.utils.data.Dataset):
    """Drop-in for core/dataset/kitti_prepared.py:10-153 with the pre-processing deferred to the GPU.

    `__getitem__` returns `(strip_uint8, flip)`: the decoded (3*H0, W0, 3) image as a CPU uint8 tensor and the
    reference's flip coin.  Use `KITTI_Prepared.collate` as the DataLoader's `collate_fn` (it keeps the strips as a list)
    and `to_batch` on the training process to obtain the (B,3,3H,W) fp32 CUDA batch `Model_flow.forward` takes."""

    def __init__(self, data_dir, num_scales=3, img_hw=(256, 832), num_iterations=None):
        super().__init__()
        self.data_dir = data_dir
        self.num_scales = num_scales
        self.img_hw = img_hw
        self.num_iterations = num_iterations
        self.data_list = self.get_data_list(os.path.join(self.data_dir, 'train.txt'))

    def get_data_list(self, info_file):                                       # kitti_prepared.py:22-33
        with open(info_file, 'r') as f:
            lines = f.readlines()
        data_list = []
        for line in lines:
            k = line.strip('\n').split()
            data_list.append({'image_file': os.path.join(self.data_dir, k[0]),
                              'cam_intrinsic_file': os.path.join(self.data_dir, k[1])})
        return data_list

    def count(self):
        return len(self.data_list)

    def rand_num(self, idx):                                                  # kitti_prepared.py:38-41
        num_total = self.count()
        np.random.seed(idx)
        return np.random.randint(num_total)

    def __len__(self):
        return self.count() if self.num_iterations is None else self.num_iterations

    def random_flip(self):
        return bool(np.random.rand() > 0.5)                                    # kitti_prepared.py:78

    def read_cam_intrinsic(self, fname):                                      # kitti_prepared.py:103-110
        with open(fname, 'r') as f:
            lines = f.readlines()
        data = [float(k) for k in lines[-1].strip('\n').split(' ')[1:]]
        return np.array(data).reshape(3, 4)[:3, :3]

    def rescale_intrinsics(self, K, img_hw_orig, img_hw_new):                # kitti_prepared.py:112-115 (as written there)
        K[0, :] = K[0, :] * img_hw_new[0] / img_hw_orig[0]
        K[1, :] = K[1, :] * img_hw_new[1] / img_hw_orig[1]
        return K

    def get_intrinsics_per_scale(self, K, scale):
        K_new = copy.deepcopy(K)
        K_new[0, :] = K_new[0, :] / (2 ** scale)
        K_new[1, :] = K_new[1, :] / (2 ** scale)
        return K_new, np.linalg.inv(K_new)

    def get_multiscale_intrinsics(self, K, num_scales):
        K_ms, K_inv_ms = [], []
        for s in range(num_scales):
            K_new, K_new_inv = self.get_intrinsics_per_scale(K, s)
            K_ms.append(K_new[None, :, :])
            K_inv_ms.append(K_new_inv[None, :, :])
        return np.concatenate(K_ms, 0), np.concatenate(K_inv_ms, 0)

    def preprocess_img(self, img, img_hw=None, is_test=False):
        """Reference name and meaning (kitti_prepared.py:83-91) on the GPU: (3*H0, W0, 3) uint8 -> (3, 3H, W) fp32 CUDA."""
        return preprocess_img(img, self.img_hw if img_hw is None else img_hw, flip=(not is_test) and self.random_flip())

    def __getitem__(self, idx):
        if self.num_iterations is not None:
            idx = self.rand_num(idx)
        import cv2                                                             # decode stays on the CPU
        img = cv2.imread(self.data_list[idx]['image_file'])
        if img is None:
            raise FileNotFoundError(self.data_list[idx]['image_file'])
        return torch.from_numpy(img), self.random_flip()

    @staticmethod
    def collate(samples):
        """collate_fn: keeps the variable-size strips as a list -> (list of uint8 tensors, list of flips)."""
        return [s[0] for s in samples], [s[1] for s in samples]

    def to_batch(self, collated, device=None):
        """(strips, flips) from `collate` -> (B,3,3H,W) fp32 CUDA batch."""
        strips, flips = collated
        return preprocess_batch(strips, self.img_hw, flips, 3, device)
