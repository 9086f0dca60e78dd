"""GPU parity tests of the I/O row: uof_preprocess_u8, uof_flow_png_decode/encode, uof_flow_eval and
their host mirrors (unopticalflow_b200.dataset / .evaluation) against the CPU oracle (oracle/io_ops.py) and the fixtures
recorded from the unmodified reference.  Bit-exact for the 8-bit resize + normalisation and the PNG arithmetic; 1e-9
relative for the fp64 evaluation sums (summation order); identical report strings."""
import os
import types
import zlib

import numpy as np
import pytest
import torch

from oracle import io_ops as IO
from util import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def D(cuda):
    from unopticalflow_b200 import _lib, dataset, evaluation
    _lib.load()
    return types.SimpleNamespace(ds=dataset, ev=evaluation)


def gold(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize('name', ['io_preprocess_down', 'io_preprocess_up', 'io_preprocess_mixed'])
def test_preprocess_matches_reference_fixture(D, name):
    g = gold(name + '.npz')
    hw = tuple(int(v) for v in g['hw'])
    for flip in (0, 1):
        out = D.ds.preprocess_img(g['img'], hw, bool(flip))
        assert out.is_cuda and out.dtype == torch.float32
        assert np.array_equal(out.cpu().numpy(), g['out_flip%d' % flip])


@pytest.mark.parametrize('tag', ['kitti', 'sintel'])
def test_preprocess_full_size_crc(D, tag):
    """BASELINE configs' image sizes (375x1242 -> 256x832, 436x1024 -> 448x1024): CRC32 of the fp32 batch == the CRC recorded
    from KITTI_Prepared.preprocess_img of the unmodified reference."""
    g = gold('io_preprocess_%s.npz' % tag)
    h0, w0, H, W = (int(v) for v in g['shape'])
    img = IO.synthetic_strip(h0, w0, 3, seed=h0)
    for flip in (0, 1):
        out = D.ds.preprocess_img(img, (H, W), bool(flip)).cpu().numpy()
        assert zlib.crc32(np.ascontiguousarray(out).tobytes()) == int(g['crc'][flip])


def test_preprocess_batch_ragged_sizes_and_two_image_strips(D):
    """A batch whose samples have different raw sizes (as KITTI's do), per-sample flips, CUDA and numpy inputs mixed; and the
    two-image strips of resize_img_origin (kitti_prepared.py:50-61)."""
    rng = np.random.default_rng(3)
    shapes = [(375, 1242), (370, 1226), (94, 311), (20, 33), (300, 500)]
    imgs = [rng.integers(0, 256, (3 * h, w, 3), dtype=np.uint8) for h, w in shapes]
    flips = [False, True, True, False, True]
    mixed = [torch.from_numpy(a).cuda() if i % 2 else a for i, a in enumerate(imgs)]
    out = D.ds.preprocess_batch(mixed, (128, 416), flips)
    assert tuple(out.shape) == (5, 3, 384, 416)
    for b in range(5):
        assert np.array_equal(out[b].cpu().numpy(), IO.preprocess_img(imgs[b], (128, 416), flips[b])), b
    two = rng.integers(0, 256, (2 * 61, 77, 3), dtype=np.uint8)
    o2 = D.ds.preprocess_batch([two], (40, 96), None, nimg=2)
    assert np.array_equal(o2[0].cpu().numpy(), IO.preprocess_img(two, (40, 96), False, nimg=2))
    with pytest.raises(TypeError):
        D.ds.preprocess_img(two.astype(np.float32), (40, 96))


def test_flow_png_roundtrip(D):
    g = gold('io_flow_png.npz')
    dec = D.ev.decode_flow_png(g['raw'])
    assert np.array_equal(dec.cpu().numpy().astype(np.float64), g['decoded'])
    assert np.array_equal(D.ev.encode_flow_png(g['flo']), g['encoded'])
    f = (np.arange(-500, 500, dtype=np.float32).reshape(10, 100, 1) / 64.0).repeat(2, 2)
    assert np.array_equal(D.ev.decode_flow_png(D.ev.encode_flow_png(f)).cpu().numpy()[:, :, :2], f)


def test_eval_flow_matches_reference_fixture(D):
    g = gold('io_eval_flow.npz')
    n = int(g['n'][0])
    cfg = types.SimpleNamespace(img_hw=tuple(int(v) for v in g['img_hw']))
    gts, nocs, preds, moves = ([g['%s%d' % (k, i)] for i in range(n)] for k in ('gt', 'noc', 'pred', 'move'))
    assert D.ev.eval_flow_avg(gts, nocs, preds, cfg) == bytes(g['report_plain']).decode()
    assert D.ev.eval_flow_avg(gts, nocs, preds, cfg, moving_masks=moves) == bytes(g['report_move']).decode()
    for i in range(n):
        s = D.ev.flow_eval_sums(torch.from_numpy(preds[i]).permute(2, 0, 1).cuda(), gts[i], nocs[i], cfg.img_hw, moves[i]).cpu().numpy()
        t = g['terms%d' % i]          # MOVING_KEYS order: epe, epe_noc, epe_occ, epe_move, epe_static, move_rate, static_rate, err_rate
        mine = [s[0] / s[1], s[2] / s[3], s[4] / max(s[5], 1.0), s[11] / s[8], s[12] / s[10], s[7] / s[8], s[9] / s[10], s[6] / s[1]]
        assert np.allclose(mine, t, rtol=1e-9, atol=0), (mine, t)


def test_eval_flow_kitti_size_against_oracle(D):
    """375x1242 ground truth, 256x832 prediction (KITTI-2015 eval shape, BASELINE configs[2]) against the numpy oracle."""
    rng = np.random.default_rng(11)
    H, W, h, w = 375, 1242, 256, 832
    gt = np.zeros((H, W, 3))
    gt[:, :, :2] = np.rint(rng.normal(0, 20, (H, W, 2)) * 64) / 64
    gt[:, :, 2] = rng.random((H, W)) > 0.4
    gt[:, :, :2] *= gt[:, :, 2:3]
    noc = gt[:, :, 2] * (rng.random((H, W)) > 0.1)
    pred = rng.normal(0, 15, (h, w, 2)).astype(np.float32)
    t = IO.eval_flow_terms(gt, noc, pred, (h, w))
    s = D.ev.flow_eval_sums(pred, gt, noc, (h, w)).cpu().numpy()
    mine = [s[0] / s[1], s[2] / s[3], s[4] / max(s[5], 1.0), s[6] / s[1]]
    assert np.allclose(mine, [t[k] for k in IO.PLAIN_KEYS], rtol=1e-9, atol=0), (mine, t)
