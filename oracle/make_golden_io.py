"""Golden fixtures for the I/O rows, recorded from the UNMODIFIED reference in the build container:

    python oracle/make_golden_io.py [--ref /root/reference] [--out tests/golden]

* core/dataset/kitti_prepared.py  KITTI_Prepared.preprocess_img  (cv2.resize of each third, flip, /255)
* core/evaluation/evaluate_flow.py  eval_flow_avg / calculate_error_rate  (cv2.resize of the prediction, EPE, Fl)

Harness-side shims (the reference files are not touched): `core.evaluation` imports `png` (pypng) and matplotlib, which are
not installed here and are only needed to open PNG files / colour-code flow images -- empty stub modules are registered
under those names; the dataset class is
instantiated with `object.__new__` because its constructor reads a file list.  The oracle (oracle/io_ops.py) is checked
against every recorded value and, separately, against cv2 itself on random shapes; the script prints both.
Fixtures are small (seeded synthetic images at reduced sizes); the KITTI- and Sintel-size resizes are recorded as
CRC32s of the output for the integer test strip of oracle.io_ops.synthetic_strip.
"""
from __future__ import annotations

import argparse
import os
import sys
import types
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import io_ops as IO  # noqa: E402


def load_reference(ref):
    # shims: pypng and matplotlib are not installed here; flowlib.py needs them only to open PNG files and to colour-code
    # flow images, neither of which this script calls
    for name in ('png', 'matplotlib', 'matplotlib.colors', 'matplotlib.pyplot', 'PIL', 'PIL.Image'):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, ref)
    sys.path.insert(0, os.path.join(ref, 'core', 'dataset'))
    sys.path.insert(0, os.path.join(ref, 'core', 'evaluation'))
    import kitti_prepared
    import evaluate_flow
    return kitti_prepared.KITTI_Prepared, evaluate_flow


def synth_triplet(rng, h0, w0):
    """A smooth-ish uint8 BGR strip of three images (random low-frequency field + noise)."""
    base = rng.integers(0, 256, (3 * h0 // 8 + 2, w0 // 8 + 2, 3)).astype(np.float32)
    import cv2
    img = cv2.resize(base, (w0, 3 * h0), interpolation=cv2.INTER_CUBIC) + rng.normal(0, 12, (3 * h0, w0, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden'))
    args = ap.parse_args()
    import cv2
    KP, EF = load_reference(args.ref)
    ds = object.__new__(KP)
    rng = np.random.default_rng(20261017)
    worst = 0.0

    # ---- preprocess_img: small cases stored in full --------------------------------------------------------------
    for name, (h0, w0, hw) in {'io_preprocess_down': (47, 155, (32, 104)), 'io_preprocess_up': (30, 52, (32, 64)),
                               'io_preprocess_mixed': (40, 100, (48, 64))}.items():
        img = synth_triplet(rng, h0, w0)
        out = {}
        for flip in (0, 1):
            ds.random_flip_img = (lambda im, f=flip: cv2.flip(im, 1) if f else im)       # make the coin deterministic
            ref = ds.preprocess_img(img, hw).transpose(2, 0, 1)                          # __getitem__ :146-147
            ref32 = ref.astype(np.float32)                                                # torch.from_numpy(img).float()
            mine = IO.preprocess_img(img, hw, bool(flip))
            assert np.array_equal(ref32, mine), (name, flip, np.abs(ref32 - mine).max())
            out['out_flip%d' % flip] = ref32
        np.savez_compressed(os.path.join(args.out, name + '.npz'), img=img, hw=np.array(hw), **out)
        print('%-22s oracle == reference (bit-exact), %s -> %s' % (name, img.shape, out['out_flip0'].shape))

    # ---- KITTI-size resize: CRC + sample rows ----------------------------------------------------------------------
    for tag, (h0, w0, hw) in {'kitti': (375, 1242, (256, 832)), 'sintel': (436, 1024, (448, 1024))}.items():
        img = IO.synthetic_strip(h0, w0, 3, seed=h0)
        crcs = []
        for flip in (0, 1):
            ds.random_flip_img = (lambda im, f=flip: cv2.flip(im, 1) if f else im)
            ref = ds.preprocess_img(img, hw).transpose(2, 0, 1).astype(np.float32)
            assert np.array_equal(ref, IO.preprocess_img(img, hw, bool(flip)))
            crcs.append(zlib.crc32(np.ascontiguousarray(ref).tobytes()))
        np.savez_compressed(os.path.join(args.out, 'io_preprocess_%s.npz' % tag), shape=np.array([h0, w0, hw[0], hw[1]]),
                            crc=np.array(crcs, dtype=np.int64), img_crc=np.array([zlib.crc32(img.tobytes())], dtype=np.int64))
        print('io_preprocess_%-9s oracle == reference (bit-exact) at 3x%dx%d -> 3x%dx%d, crc %08x / %08x (flipped)'
              % (tag, h0, w0, hw[0], hw[1], crcs[0], crcs[1]))

    # ---- cv2 itself, random shapes ---------------------------------------------------------------------------------
    bad = 0
    for _ in range(60):
        h, w, H, W = [int(v) for v in rng.integers(5, 300, 4)]
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        bad += int((cv2.resize(a, (W, H)) != IO.resize_linear_u8(a, (W, H))).sum())
        f = rng.normal(0, 5, (h, w, 2)).astype(np.float32)
        d = np.abs(cv2.resize(f, (W, H), interpolation=cv2.INTER_LINEAR) - IO.resize_linear_f32(f, (W, H))).max()
        worst = max(worst, float(d))
    print('cv2.resize uint8 vs oracle on 60 random shapes: %d differing bytes; float32: max abs diff %.2e' % (bad, worst))
    assert bad == 0 and worst < 1e-5

    # ---- eval_flow_avg ---------------------------------------------------------------------------------------------
    cfg = types.SimpleNamespace(img_hw=(32, 104), model_dir='/tmp')
    gts, nocs, preds, moves = [], [], [], []
    for i in range(3):
        H, W = 47 + i, 155 - i
        gt = np.zeros((H, W, 3), dtype=np.float64)
        coarse = rng.normal(0, 25, (6, 12, 2)).astype(np.float32)               # a smooth field, quantised like a KITTI PNG
        raw = np.rint(cv2.resize(coarse, (W, H), interpolation=cv2.INTER_CUBIC) * 64.0) + 2 ** 15
        valid = (rng.random((H, W)) > 0.3)
        gt[:, :, 0:2] = (raw - 2 ** 15) / 64.0 * valid[:, :, None]
        gt[:, :, 2] = valid
        noc = (valid & (rng.random((H, W)) > 0.2)).astype(np.float64)
        pred = (rng.normal(0, 6, (32, 104, 2))).astype(np.float32)
        # make the prediction close to the ground truth on part of the image so that both sides of the 3 px / 5 % test occur
        small = cv2.resize(((raw - 2 ** 15) / 64.0).astype(np.float32), (104, 32), interpolation=cv2.INTER_LINEAR)
        small[:, :, 0] *= 104.0 / W
        small[:, :, 1] *= 32.0 / H
        pred[:, :52] = small[:, :52] + rng.normal(0, 1.2, (32, 52, 2)).astype(np.float32)
        gts.append(gt); nocs.append(noc); preds.append(pred)
        moves.append((rng.random((H, W)) > 0.6).astype(np.float64))
    ref_plain = EF.eval_flow_avg(gts, nocs, preds, cfg)
    ref_move = EF.eval_flow_avg(gts, nocs, preds, cfg, moving_masks=moves)
    my_plain, m1 = IO.eval_flow_avg(gts, nocs, preds, cfg.img_hw)
    my_move, m2 = IO.eval_flow_avg(gts, nocs, preds, cfg.img_hw, moves)
    assert ref_plain == my_plain and ref_move == my_move, (ref_plain, my_plain, ref_move, my_move)
    terms = [IO.eval_flow_terms(gts[i], nocs[i], preds[i], cfg.img_hw, moves[i]) for i in range(3)]
    ref_rate = [EF.calculate_error_rate(np.sqrt(np.sum(np.square(IO.resize_linear_f32(
        preds[i] / np.array([cfg.img_hw[1], cfg.img_hw[0]], dtype=np.float32) * np.array([gts[i].shape[1], gts[i].shape[0]], dtype=np.float32),
        (gts[i].shape[1], gts[i].shape[0])) - gts[i][:, :, :2]), axis=2)), gts[i][:, :, 0:2], gts[i][:, :, 2]) for i in range(3)]
    print('eval_flow_avg          oracle == reference report strings:\n' + ref_move.rstrip())
    save = {'img_hw': np.array(cfg.img_hw), 'n': np.array([3])}
    for i in range(3):
        save.update({'gt%d' % i: gts[i], 'noc%d' % i: nocs[i], 'pred%d' % i: preds[i], 'move%d' % i: moves[i],
                     'terms%d' % i: np.array([terms[i][k] for k in IO.MOVING_KEYS])})
    save['report_plain'] = np.frombuffer(ref_plain.encode(), dtype=np.uint8)
    save['report_move'] = np.frombuffer(ref_move.encode(), dtype=np.uint8)
    save['err_rate_ref'] = np.array(ref_rate)
    np.savez_compressed(os.path.join(args.out, 'io_eval_flow.npz'), **save)

    # ---- KITTI flow PNG arithmetic -----------------------------------------------------------------------------------
    raw = rng.integers(0, 2 ** 16, (20, 33, 3)).astype(np.uint16)
    raw[:, :, 2] = rng.integers(0, 2, (20, 33))
    dec = IO.decode_flow_png(raw)
    flo = rng.normal(0, 300, (20, 33, 2)).astype(np.float32)       # large: exercises both clamps
    enc = IO.encode_flow_png(flo)
    np.savez_compressed(os.path.join(args.out, 'io_flow_png.npz'), raw=raw, decoded=dec, flo=flo, encoded=enc)
    print('io_flow_png            recorded from the oracle (flowlib.py:107-138 restated; pypng absent: file decode unpinned)')


if __name__ == '__main__':
    main()
