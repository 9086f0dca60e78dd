"""CPU tests of the I/O oracle (oracle/io_ops.py) against the fixtures recorded from the unmodified
reference by oracle/make_golden_io.py (KITTI_Prepared.preprocess_img, eval_flow_avg) and, when cv2 is importable, against
cv2.resize itself.  Bit-exact for the 8-bit resize, the flow-PNG arithmetic and the report strings."""
import zlib

import numpy as np
import pytest

from oracle import io_ops as IO
from util import GOLDEN
import os


def gold(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize('name', ['io_preprocess_down', 'io_preprocess_up', 'io_preprocess_mixed'])
def test_preprocess_matches_reference_fixture(name):
    g = gold(name + '.npz')
    for flip in (0, 1):
        out = IO.preprocess_img(g['img'], tuple(int(v) for v in g['hw']), bool(flip))
        assert out.dtype == np.float32 and np.array_equal(out, g['out_flip%d' % flip])


@pytest.mark.parametrize('tag', ['kitti', 'sintel'])
def test_preprocess_full_size_crc(tag):
    g = gold('io_preprocess_%s.npz' % tag)
    h0, w0, H, W = (int(v) for v in g['shape'])
    img = IO.synthetic_strip(h0, w0, 3, seed=h0)
    assert zlib.crc32(img.tobytes()) == int(g['img_crc'][0])
    for flip in (0, 1):
        out = IO.preprocess_img(img, (H, W), bool(flip))
        assert out.shape == (3, 3 * H, W)
        assert zlib.crc32(np.ascontiguousarray(out).tobytes()) == int(g['crc'][flip])


def test_resize_against_cv2_when_available():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(7)
    for _ in range(25):
        h, w, H, W = [int(v) for v in rng.integers(3, 200, 4)]
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(cv2.resize(a, (W, H)), IO.resize_linear_u8(a, (W, H))), (h, w, H, W)
        f = rng.normal(0, 5, (h, w, 2)).astype(np.float32)
        d = np.abs(cv2.resize(f, (W, H), interpolation=cv2.INTER_LINEAR) - IO.resize_linear_f32(f, (W, H))).max()
        assert d <= 1e-5, (h, w, H, W, d)
    one = rng.integers(0, 256, (7, 9), dtype=np.uint8)                     # single-channel, identity and 1-pixel edge cases
    assert np.array_equal(cv2.resize(one, (9, 7)), IO.resize_linear_u8(one, (9, 7)))
    assert np.array_equal(cv2.resize(one[:1, :1], (5, 4)), IO.resize_linear_u8(one[:1, :1], (5, 4)))


def test_eval_flow_matches_reference_fixture():
    g = gold('io_eval_flow.npz')
    n = int(g['n'][0])
    hw = tuple(int(v) for v in g['img_hw'])
    gts, nocs, preds, moves = ([g['%s%d' % (k, i)] for i in range(n)] for k in ('gt', 'noc', 'pred', 'move'))
    rep_plain, _ = IO.eval_flow_avg(gts, nocs, preds, hw)
    rep_move, _ = IO.eval_flow_avg(gts, nocs, preds, hw, moves)
    assert rep_plain == bytes(g['report_plain']).decode() and rep_move == bytes(g['report_move']).decode()
    for i in range(n):
        t = IO.eval_flow_terms(gts[i], nocs[i], preds[i], hw, moves[i])
        assert np.allclose([t[k] for k in IO.MOVING_KEYS], g['terms%d' % i], rtol=1e-12, atol=0)
        assert abs(t['err_rate'] - g['err_rate_ref'][i]) < 1e-15       # calculate_error_rate of the reference itself
    assert 0.2 < t['err_rate'] < 0.8                                        # both sides of the outlier test are exercised


def test_flow_png_arithmetic():
    g = gold('io_flow_png.npz')
    dec = IO.decode_flow_png(g['raw'])
    assert np.array_equal(dec, g['decoded'])
    assert np.all(dec[g['raw'][:, :, 2] == 0][:, :2] == 0)
    assert np.array_equal(IO.encode_flow_png(g['flo']), g['encoded'])
    # round trip inside the representable range: decode(encode(f)) == f quantised to 1/64 (truncation toward zero of f*64+2^15)
    f = (np.arange(-500, 500, dtype=np.float32).reshape(10, 100, 1) / 64.0).repeat(2, 2)
    back = IO.decode_flow_png(IO.encode_flow_png(f))
    assert np.array_equal(back[:, :, :2], f.astype(np.float64))
    big = np.full((1, 2, 2), 1e6, dtype=np.float32)
    big[0, 1] = -1e6
    e = IO.encode_flow_png(big)
    assert e[0, 0, 0] == 65535 and e[0, 1, 0] == 0 and e[0, 0, 2] == 1
