"""CPU oracle for SURVEY 8(f) rank 4 -- the input pipeline and the flow evaluation either side of the hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, `__graft_entry__.smoke()` and bench.py's CPU legs, never by the product
(unopticalflow_b200/), which has no CPU fallback.

numpy restatement of
  * core/dataset/kitti_prepared.py:63-91  (resize_img / random_flip_img / preprocess_img) and :133-153 (__getitem__),
  * core/evaluation/flowlib.py:107-145   (read_flow_png / write_flow_png, the KITTI 16-bit flow encoding),
  * core/evaluation/evaluate_flow.py:85-174 (calculate_error_rate / eval_flow_avg).

Third-party arithmetic: the resizes are OpenCV's `cv2.resize(..., INTER_LINEAR)` (the reference pins
opencv-python==4.1.0.25, requirements.txt; installed here: 4.13.0).  OpenCV's algorithm is restated below
(modules/imgproc/src/resize.cpp: 11-bit fixed-point coefficients for 8-bit images, the `VResizeLinear<uchar, ...>`
rounding `(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`, border handling by zeroing the horizontal
fraction but clamping the vertical row index) and PINNED against the installed cv2: bit-exact on 60 random shapes and the
KITTI / Sintel shapes (oracle/make_golden_io.py prints the check; tests/test_oracle_golden.py repeats it when cv2 is
importable and otherwise checks the committed fixtures).  The PNG container itself (pypng in the reference, absent here)
is not restated: the functions take / return the decoded (H,W,3) uint16 array -- "parity unpinned" for the file decode,
pinned for the arithmetic.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11                      # INTER_RESIZE_COEF_BITS
COEF_SCALE = 1 << COEF_BITS         # 2048


def synthetic_strip(h0: int, w0: int, nimg: int = 3, seed: int = 0):
    """Deterministic uint8 test strip (nimg*h0, w0, 3) built from integer arithmetic only (no RNG state, no cv2), so the
    golden script and the tests construct the same bytes anywhere: smooth ramps, a checker texture and hashed noise."""
    y = np.arange(nimg * h0, dtype=np.int64)[:, None, None]
    x = np.arange(w0, dtype=np.int64)[None, :, None]
    c = np.arange(3, dtype=np.int64)[None, None, :]
    ramp = (y * (3 + c) + x * (5 - c)) // 4
    checker = (((y // 6) + (x // 9)) % 2) * 70
    hashed = ((y * 73856093 + x * 19349663 + (c + seed) * 83492791) ^ ((y * x + seed) >> 3)) % 61
    return ((ramp + checker + hashed) % 256).astype(np.uint8)


def linear_coeffs(dst: int, src: int, vertical: bool):
    """Source indices and fractions of cv2's bilinear resize along one axis (resize.cpp, `resizeGeneric_` set-up):
    f = (float)((d + 0.5) * (double)src / dst - 0.5); s = floor(f); f -= s.  Along x a source index outside the row sets
    f = 0 and clamps s; along y the fraction is kept and the two ROW indices are clamped instead."""
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), f


def resize_linear_u8(img: np.ndarray, size_wh):
    """== cv2.resize(img_uint8, (W, H))  (INTER_LINEAR), bit-exact.  img: (h, w) or (h, w, C) uint8."""
    W, H = int(size_wh[0]), int(size_wh[1])
    squeeze = img.ndim == 2
    im = img.reshape(img.shape[0], img.shape[1], -1).astype(np.int64)
    h, w = im.shape[:2]
    x0, x1, fx = linear_coeffs(W, w, False)
    y0, y1, fy = linear_coeffs(H, h, True)
    one = np.float32(1.0)
    sc = np.float32(COEF_SCALE)
    ax0, ax1 = np.rint((one - fx) * sc).astype(np.int64), np.rint(fx * sc).astype(np.int64)      # saturate_cast<short> = cvRound
    by0, by1 = np.rint((one - fy) * sc).astype(np.int64), np.rint(fy * sc).astype(np.int64)
    rows = im[:, x0] * ax0[None, :, None] + im[:, x1] * ax1[None, :, None]                       # HResizeLinear, x 2048
    s0, s1 = rows[y0], rows[y1]
    out = ((((s0 >> 4) * by0[:, None, None]) >> 16) + (((s1 >> 4) * by1[:, None, None]) >> 16) + 2) >> 2
    out = out.astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def resize_linear_f32(img: np.ndarray, size_wh):
    """== cv2.resize(img_float32, (W, H), interpolation=cv2.INTER_LINEAR) up to the last bit of the two multiply-adds
    (HResizeLinear / VResizeLinear<float>: D = S[x0]*(1-fx) + S[x1]*fx, then dst = S0*(1-fy) + S1*fy, in fp32)."""
    W, H = int(size_wh[0]), int(size_wh[1])
    squeeze = img.ndim == 2
    im = img.reshape(img.shape[0], img.shape[1], -1).astype(np.float32)
    h, w = im.shape[:2]
    x0, x1, fx = linear_coeffs(W, w, False)
    y0, y1, fy = linear_coeffs(H, h, True)
    one = np.float32(1.0)
    rows = im[:, x0] * (one - fx)[None, :, None] + im[:, x1] * fx[None, :, None]
    out = rows[y0] * (one - fy)[:, None, None] + rows[y1] * fy[:, None, None]
    out = out.astype(np.float32)
    return out[:, :, 0] if squeeze else out


def preprocess_img(img: np.ndarray, img_hw, flip: bool = False, nimg: int = 3):
    """kitti_prepared.py:63-91 + :146-153: the (nimg*H0, W0, 3) uint8 strip is cut into its images, each resized to
    img_hw, re-stacked, optionally mirrored (cv2.flip(img, 1)), divided by 255.0 and returned as float32 CHW
    (`torch.from_numpy(img.transpose(2,0,1)).float()`; the division is done in float64 and rounded once)."""
    H0 = img.shape[0] // nimg                                                # :70 int(img_h / 3)
    parts = [resize_linear_u8(img[k * H0:(k + 1) * H0], (img_hw[1], img_hw[0])) for k in range(nimg)]
    out = np.concatenate(parts, 0)
    if flip:
        out = out[:, ::-1]
    return (out / 255.0).transpose(2, 0, 1).astype(np.float32)


def decode_flow_png(raw: np.ndarray):
    """flowlib.py:107-128 on the decoded (H,W,3) uint16 array [u, v, valid]: flow = (raw - 2^15) / 64, zero where invalid."""
    flow = raw.astype(np.float64)
    invalid = flow[:, :, 2] == 0
    flow[:, :, 0:2] = (flow[:, :, 0:2] - 2 ** 15) / 64.0
    flow[invalid, 0] = 0
    flow[invalid, 1] = 0
    return flow


def encode_flow_png(flo: np.ndarray):
    """flowlib.py:131-138: (H,W,>=2) float flow -> (H,W,3) uint16 [clip(u*64 + 2^15), clip(v*64 + 2^15), 1] (fp32, truncated)."""
    h, w = flo.shape[:2]
    out = np.ones((h, w, 3), dtype=np.float32)
    out[:, :, 0] = np.maximum(np.minimum(flo[:, :, 0] * 64.0 + 2 ** 15, 2 ** 16 - 1), 0)
    out[:, :, 1] = np.maximum(np.minimum(flo[:, :, 1] * 64.0 + 2 ** 15, 2 ** 16 - 1), 0)
    return out.astype(np.uint16)


def calculate_error_rate(epe_map, gt_flow, mask):
    """evaluate_flow.py:85-91."""
    bad = np.logical_and(epe_map * mask > 3,
                         epe_map * mask / np.maximum(np.sqrt(np.sum(np.square(gt_flow), axis=2)), 1e-10) > 0.05)
    return bad.sum() / mask.sum()


def eval_flow_terms(gt_flow, noc_mask, pred_flow, img_hw, moving_mask=None):
    """One image of evaluate_flow.py:104-160 -> dict of the per-image terms that eval_flow_avg averages.
    gt_flow (H,W,3) float64 [u,v,valid], noc_mask (H,W), pred_flow (h,w,2) float32 at network resolution."""
    H, W = gt_flow.shape[0:2]
    pred = np.copy(pred_flow)
    pred[:, :, 0] = pred[:, :, 0] / img_hw[1] * W                                         # :107
    pred[:, :, 1] = pred[:, :, 1] / img_hw[0] * H                                         # :108
    flo = resize_linear_f32(pred, (W, H))                                                 # :110-111
    epe = np.sqrt(np.sum(np.square(flo[:, :, 0:2] - gt_flow[:, :, 0:2]), axis=2))       # :132-134
    valid = gt_flow[:, :, 2]
    t = {'epe': np.sum(epe * valid) / np.sum(valid),                                      # :135
         'epe_noc': np.sum(epe * noc_mask) / np.sum(noc_mask),                            # :137
         'epe_occ': np.sum(epe * (valid - noc_mask)) / max(np.sum(valid - noc_mask), 1.0),  # :139-140
         'err_rate': calculate_error_rate(epe, gt_flow[:, :, 0:2], valid)}                # :142-143
    if moving_mask is not None:
        mv = moving_mask
        t['move_err_rate'] = calculate_error_rate(epe, gt_flow[:, :, 0:2], valid * mv)    # :148-149
        t['static_err_rate'] = calculate_error_rate(epe, gt_flow[:, :, 0:2], valid * (1.0 - mv))
        t['epe_move'] = np.sum(epe * valid * mv) / np.sum(valid * mv)                     # :154-156
        t['epe_static'] = np.sum(epe * valid * (1.0 - mv)) / np.sum(valid * (1.0 - mv))
    return t


MOVING_KEYS = ('epe', 'epe_noc', 'epe_occ', 'epe_move', 'epe_static', 'move_err_rate', 'static_err_rate', 'err_rate')
PLAIN_KEYS = ('epe', 'epe_noc', 'epe_occ', 'err_rate')


def format_result(means: dict, moving: bool) -> str:
    """The two-line report string of evaluate_flow.py:162-174."""
    keys = MOVING_KEYS if moving else PLAIN_KEYS
    head = ', '.join('{:>10}'.format(k) for k in keys) + ' \n'
    body = ', '.join('{:10.4f}'.format(float(means[k])) for k in keys) + ' \n'
    return head + body


def eval_flow_avg(gt_flows, noc_masks, pred_flows, img_hw, moving_masks=None):
    """evaluate_flow.py:93-174 without the image dumps: -> (report string, dict of means)."""
    n = len(gt_flows)
    acc = {}
    for i in range(n):
        t = eval_flow_terms(gt_flows[i], noc_masks[i], pred_flows[i], img_hw, moving_masks[i] if moving_masks else None)
        for k, v in t.items():
            acc[k] = acc.get(k, 0.0) + v
    means = {k: v / n for k, v in acc.items()}
    return format_result(means, bool(moving_masks)), means
