"""Per-kernel device timing of the hand-written operators at the benchmark shapes.

Every case captures `reps` launches of one C-ABI entry point into a CUDA graph, each launch on a different set of
buffers (the sets together exceed the 126 MB L2, so no launch finds its inputs in cache), replays the graph and
times it with CUDA events: the result is pure device time per launch without host launch latency.
Used by bench.py for the roofline table and by `python -m unopticalflow_b200.kernel_bench` while tuning.
"""
from __future__ import annotations

import argparse
import json
import re

import torch

from . import _lib, ops

L2_BYTES = 126e6
FP32_PEAK_TFLOPS = 69.0      # measured on this pool's B200 with a pure FFMA kernel (profiles/microbench/ffma2_result.txt)
LEVELS = ((196, 4, 13), (128, 8, 26), (96, 16, 52), (64, 32, 104), (32, 64, 208))   # (C,h,w) at 256x832, SURVEY 8


def _time_graph(fn_list, iters=5):
    """fn_list: closures, one per buffer set.  Returns mean device microseconds per closure call."""
    for f in fn_list:          # warm-up (also triggers cudaFuncSetAttribute etc. outside capture)
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fn_list:
            f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float('inf')
    for _ in range(iters):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 1e3 * best / len(fn_list)


def _nsets(bytes_per_set):
    return int(max(3, min(24, 2 * L2_BYTES // max(bytes_per_set, 1) + 1)))


def cases(B=8, H=256, W=832):
    """Yield (name, algorithmic_bytes, [closures]) for the kernels of one B-triplet training step (2B pairs)."""
    dev = 'cuda'
    B2 = 2 * B
    S = 3
    r = lambda *s: torch.randn(*s, device=dev)
    u = lambda *s: torch.rand(*s, device=dev)

    for (C, h, w) in LEVELS:
        n = B2 * h * w
        nb = (2 * C + 81) * 4 * n
        sets = [(r(B2, C, h, w), r(B2, C, h, w)) for _ in range(_nsets(nb))]
        outs = [torch.empty(B2, 81, h, w, device=dev) for _ in sets]
        yield ('cost_volume_fwd[%dx%dx%dx%d]' % (B2, C, h, w), nb,
               [lambda a=a, b=b, o=o: _lib.call('uof_cost_volume_fwd', ops._p(a), ops._p(b), ops._p(o), B2, C, h, w, 81 * h * w,
                                                ops._stream(a)) for (a, b), o in zip(sets, outs)])
        nbb = (4 * C + 81) * 4 * n
        gouts = [r(B2, 81, h, w) for _ in sets]
        g1 = [torch.empty_like(a) for a, _ in sets]
        g2 = [torch.empty_like(a) for a, _ in sets]
        yield ('cost_volume_bwd[%dx%dx%dx%d]' % (B2, C, h, w), nbb,
               [lambda a=a, b=b, go=go, x1=x1, x2=x2: _lib.call('uof_cost_volume_bwd', ops._p(go), 81 * h * w, ops._p(a), ops._p(b),
                                                                ops._p(x1), ops._p(x2), B2, C, h, w, ops._stream(a))
                for (a, b), go, x1, x2 in zip(sets, gouts, g1, g2)])

    def warp_case(C, h, w, use_mask, need_gx, tag):
        n = B2 * h * w
        nb = (2 * C + 2) * 4 * n
        k = _nsets(nb)
        xs = [u(B2, C, h, w) for _ in range(k)]
        # decoder flows are x2-upsampled, i.e. smooth: low-resolution noise, bilinearly upsampled (sigma ~2.5 px)
        fl = [torch.nn.functional.interpolate(r(B2, 2, max(h // 8, 2), max(w // 8, 2)) * 2.5, size=(h, w), mode='bilinear',
                                              align_corners=False).contiguous() for _ in range(k)]
        outs = [torch.empty_like(x) for x in xs]
        yield ('warp_fwd[%s %dx%dx%dx%d]' % (tag, B2, C, h, w), nb,
               [lambda x=x, f=f, o=o: _lib.call('uof_warp_fwd', ops._p(x), ops._p(f), ops._p(o), B2, C, h, w, use_mask, 0, 0,
                                                ops._stream(x)) for x, f, o in zip(xs, fl, outs)])
        nbb = ((3 * C + 4) if need_gx else (2 * C + 4)) * 4 * n
        go = [r(B2, C, h, w) for _ in range(k)]
        gx = [torch.empty_like(x) if need_gx else None for x in xs]
        gf = [torch.empty_like(f) for f in fl]
        yield ('warp_bwd[%s %dx%dx%dx%d]' % (tag, B2, C, h, w), nbb,
               [lambda x=x, f=f, g=g, a=a, b=b: _lib.call('uof_warp_bwd', ops._p(g), ops._p(x), ops._p(f), ops._p(a), ops._p(b),
                                                          B2, C, h, w, use_mask, 0, 0, ops._stream(x))
                for x, f, g, a, b in zip(xs, fl, go, gx, gf)])

    for (C, h, w) in LEVELS[1:]:
        yield from warp_case(C, h, w, 0, True, 'feat')
    for s in range(S):
        yield from warp_case(3, H >> s, W >> s, 1, False, 'img')

    # loss kernels: all three scales in one launch, called through the C ABI on pre-allocated buffers
    from ._lib import ConsisLevel, PhotoLevel, SmoothLevel
    px = sum(B * (H >> s) * (W >> s) for s in range(S))
    k = _nsets(36 * px)
    keep = []            # keeps every tensor / ctypes array alive while the closures exist

    def photo_set():
        imgs = [u(B, 3, H >> s, W >> s) for s in range(S)]
        both = [u(B2, 3, H >> s, W >> s) * (torch.rand(B2, 1, H >> s, W >> s, device=dev) > 0.02) for s in range(S)]
        wl = [torch.empty(B, 1, H >> s, W >> s, device=dev) for s in range(S)]
        wr = [torch.empty(B, 1, H >> s, W >> s, device=dev) for s in range(S)]
        gb = [torch.empty_like(t) for t in both]
        lv = (PhotoLevel * S)()
        for s in range(S):
            lv[s] = PhotoLevel(imgs[s].data_ptr(), both[s][:B].data_ptr(), both[s][B:].data_ptr(), wl[s].data_ptr(),
                               wr[s].data_ptr(), None, None, gb[s][:B].data_ptr(), gb[s][B:].data_ptr(), H >> s, W >> s)
        sums, lp, ls = torch.zeros(S * B * 6 + _lib.SUMS_EXTRA, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev)
        g = torch.ones(B, device=dev)
        keep.append((imgs, both, wl, wr, gb, lv, sums, lp, ls, g))
        return imgs, wr, lv, sums, lp, ls, g

    psets = [photo_set() for _ in range(k)]
    anchor = psets[0][3]      # the stream must be looked up at call time (graph capture runs on a side stream)
    yield ('photo_loss_fwd[3 scales, 2 dirs]', 36 * px,
           [lambda q=q: _lib.call('uof_photo_loss_fwd', q[2], S, B, ops._p(q[3]), ops._p(q[4]), ops._p(q[5]), ops._stream(anchor)) for q in psets])
    yield ('photo_loss_bwd[3 scales, 2 dirs]', 60 * px,
           [lambda q=q: _lib.call('uof_photo_loss_bwd', q[2], S, B, ops._p(q[3]), ops._p(q[6]), ops._p(q[6]), ops._stream(anchor)) for q in psets])

    # image warps fused into the photometric kernels (what Model_flow.forward launches): decoder-like flows, [l ; r] stacked
    from ._lib import PhotoWarpLevel

    def photo_warp_set():
        imgs = [u(B, 3, H >> s, W >> s) for s in range(S)]
        src = [u(B2, 3, H >> s, W >> s) for s in range(S)]
        lo = r(B2, 2, H // 32, W // 32) * 2.5
        fl = [(torch.nn.functional.interpolate(lo, size=(H >> s, W >> s), mode='bilinear', align_corners=False) / (1 << s)).contiguous()
              for s in range(S)]
        wpd = [torch.empty(B2, 3, H >> s, W >> s, device=dev) for s in range(S)]
        wl = [torch.empty(B, 1, H >> s, W >> s, device=dev) for s in range(S)]
        wr = [torch.empty(B, 1, H >> s, W >> s, device=dev) for s in range(S)]
        gf = [torch.empty_like(t) for t in fl]
        lv = (PhotoWarpLevel * S)()
        for s in range(S):
            lv[s] = PhotoWarpLevel(imgs[s].data_ptr(), src[s][:B].data_ptr(), src[s][B:].data_ptr(), fl[s][:B].data_ptr(),
                                   fl[s][B:].data_ptr(), wpd[s][:B].data_ptr(), wpd[s][B:].data_ptr(), wl[s].data_ptr(),
                                   wr[s].data_ptr(), None, None, gf[s][:B].data_ptr(), gf[s][B:].data_ptr(), H >> s, W >> s)
        sums, lp, ls = torch.zeros(S * B * 6 + _lib.SUMS_EXTRA, device=dev), torch.empty(B, device=dev), torch.empty(B, device=dev)
        g = torch.ones(B, device=dev)
        keep.append((imgs, src, fl, wpd, wl, wr, gf, lv, sums, lp, ls, g))
        return lv, sums, lp, ls, g

    # compulsory traffic of the fused op per target pixel: forward reads img 12 + two sources 24 + two flows 16 and writes
    # two weight maps 8 = 60 B (the 24 B of warped values it also stores for its backward pass are not counted); backward
    # reads img 12 + two sources 24 + two flows 16 + weights 8 and writes two flow gradients 16 = 76 B (the 24 B of saved
    # warped values it reads are not counted)
    wsets = [photo_warp_set() for _ in range(_nsets(84 * px))]
    yield ('photo_warp_loss_fwd[3 scales, 2 dirs, image warps fused]', 60 * px,
           [lambda q=q: _lib.call('uof_photo_warp_loss_fwd', q[0], S, B, 0, ops._p(q[1]), ops._p(q[2]), ops._p(q[3]), ops._stream(anchor)) for q in wsets])
    yield ('photo_warp_loss_bwd[3 scales, 2 dirs, image warps fused]', 76 * px,
           [lambda q=q: _lib.call('uof_photo_warp_loss_bwd', q[0], S, B, 0, ops._p(q[1]), ops._p(q[4]), ops._p(q[4]), ops._stream(anchor)) for q in wsets])

    def flow_set(i):
        fl = [r(B2, 2, H >> s, W >> s) for s in range(S)]
        gf = [torch.empty_like(t) for t in fl]
        imgs, wr = psets[i][0], psets[i][1]
        sl, cl = (SmoothLevel * S)(), (ConsisLevel * S)()
        for s in range(S):
            sl[s] = SmoothLevel(fl[s].data_ptr(), imgs[s].data_ptr(), gf[s].data_ptr(), H >> s, W >> s)
            cl[s] = ConsisLevel(fl[s][B:].data_ptr(), fl[s][:B].data_ptr(), wr[s].data_ptr(), gf[s][B:].data_ptr(), H >> s, W >> s)
        ssum, sloss = torch.zeros(S * B2 * 2 + _lib.SUMS_EXTRA, device=dev), torch.empty(B2, device=dev)
        csum, closs = torch.zeros(S * B * 2 + _lib.SUMS_EXTRA, device=dev), torch.empty(B, device=dev)
        g2 = torch.ones(B2, device=dev)
        keep.append((fl, gf, sl, cl, ssum, sloss, csum, closs, g2))
        return sl, cl, ssum, sloss, csum, closs, g2

    fsets = [flow_set(i) for i in range(k)]
    yield ('smooth_loss_fwd[3 scales x 2B]', 20 * 2 * px,
           [lambda q=q: _lib.call('uof_smooth_loss_fwd', q[0], S, B2, B, ops._p(q[2]), ops._p(q[3]), ops._stream(anchor)) for q in fsets])
    yield ('smooth_loss_bwd[3 scales x 2B]', 28 * 2 * px,
           [lambda q=q: _lib.call('uof_smooth_loss_bwd', q[0], S, B2, B, ops._p(q[6]), ops._stream(anchor)) for q in fsets])
    yield ('consis_loss_fwd[3 scales]', 20 * px,
           [lambda q=q: _lib.call('uof_consis_loss_fwd', q[1], S, B, ops._p(q[4]), ops._p(q[5]), ops._stream(anchor)) for q in fsets])
    yield ('consis_loss_bwd[3 scales]', 28 * px,
           [lambda q=q: _lib.call('uof_consis_loss_bwd', q[1], S, B, ops._p(q[4]), ops._p(q[6]), ops._stream(anchor)) for q in fsets])

    import ctypes
    trip = [u(B, 3, 3 * H, W) for _ in range(k)]
    pouts = [[torch.empty(3, B, 3, H >> s, W >> s, device=dev) for s in range(1, S)] for _ in range(k)]
    pptr = [(ctypes.c_void_p * (S - 1))(*[t.data_ptr() for t in o]) for o in pouts]
    keep.append((trip, pouts, pptr))
    yield ('img_pyramid[levels 1..2 of the 3 images, one launch]', int(3 * B * 3 * H * W * 4 * (1 + 1 / 4 + 1 / 16)),
           [lambda t=t, pp=pp: _lib.call('uof_img_pyramid', ops._p(t), H * t.stride(2), t.stride(0), t.stride(1), t.stride(2), pp, S,
                                         3, B, 3, H, W, ops._stream(anchor)) for t, pp in zip(trip, pptr)])
    # stacked triplet pyramid with the dense level-0 copy (what Model_flow.forward calls)
    lv0 = [torch.empty(3, B, 3, H, W, device=dev) for _ in range(k)]
    slots = (ctypes.c_int * 3)(0, 2, 1)
    keep.append((lv0, slots))
    yield ('img_pyramid_stacked[3 images, levels 0..2, one launch]', int(3 * B * 3 * H * W * 4 * (2 + 1 / 4 + 1 / 16)),
           [lambda t=t, pp=pp, o0=o0: _lib.call('uof_img_pyramid_stacked', ops._p(t), H * t.stride(2), t.stride(0), t.stride(1),
                                                t.stride(2), ops._p(o0), slots, pp, S, 3, B, 3, H, W, ops._stream(anchor))
            for t, pp, o0 in zip(trip, pptr, lv0)])
    # fused x4 bilinear up-sampling + scale of the finest flow (2B x 2 x H/4 x W/4 -> H x W)
    fins = [r(B2, 2, H // 4, W // 4) for _ in range(k)]
    fouts = [torch.empty(B2, 2, H, W, device=dev) for _ in range(k)]
    keep.append((fins, fouts))
    nup = 4 * B2 * 2 * (H * W + (H // 4) * (W // 4))
    yield ('upsample_x4_fwd[%dx2x%dx%d]' % (B2, H // 4, W // 4), nup,
           [lambda a=a, o=o: _lib.call('uof_upsample_bilinear_fwd', ops._p(a), ops._p(o), B2 * 2, H // 4, W // 4, H, W, 4.0,
                                       ops._stream(anchor)) for a, o in zip(fins, fouts)])
    yield ('upsample_x4_bwd[%dx2x%dx%d]' % (B2, H // 4, W // 4), nup,
           [lambda a=a, o=o: _lib.call('uof_upsample_bilinear_bwd', ops._p(o), ops._p(a), B2 * 2, H // 4, W // 4, H, W, 4.0,
                                       ops._stream(anchor)) for a, o in zip(fins, fouts)])
    # fused bias + LeakyReLU of the largest decoder activation (2B x 128 x H/4 x W/4)
    ah, aw = H // 4, W // 4
    ka = 3
    acts = [r(B2, 128, ah, aw) for _ in range(ka)]
    gacts = [r(B2, 128, ah, aw) for _ in range(ka)]
    gxs = [torch.empty_like(a) for a in acts]
    bias, gbias = r(128), torch.empty(128, device=dev)
    keep.append((acts, gacts, gxs, bias, gbias))
    nact = B2 * 128 * ah * aw
    yield ('bias_lrelu_fwd[%dx128x%dx%d]' % (B2, ah, aw), 8 * nact,
           [lambda a=a: _lib.call('uof_bias_lrelu_fwd', ops._p(a), ops._p(bias), B2, 128, ah, aw, 0.1, ops._stream(anchor)) for a in acts])
    yield ('bias_lrelu_bwd[%dx128x%dx%d]' % (B2, ah, aw), 12 * nact,
           [lambda a=a, g=g, o=o: _lib.call('uof_bias_lrelu_bwd', ops._p(g), ops._p(a), ops._p(o), ops._p(gbias), B2, 128, ah, aw, 0.1,
                                            ops._stream(anchor)) for a, g, o in zip(acts, gacts, gxs)])
    # full-resolution flows as the decoder produces them (x4 bilinear up-sampling of a coarse field, sigma ~2.5 px), NHWC;
    # per-pixel random flows (no coinciding corners between neighbours, nothing to aggregate) cost 47 us instead
    fl_nhwc = [torch.nn.functional.interpolate(r(B2, 2, H // 4, W // 4) * 2.5, size=(H, W), mode='bilinear', align_corners=False)
               .permute(0, 2, 3, 1).contiguous() for _ in range(k)]
    rm = [torch.empty(B2, H, W, 1, device=dev) for _ in range(k)]
    keep.append((fl_nhwc, rm))
    yield ('range_map (splat ones) [%dx%dx%d]' % (B2, H, W), 16 * B2 * H * W,
           [lambda f=f, o=o: _lib.call('uof_splat_fwd', None, ops._p(f), ops._p(o), B2, H, W, 1, ops._stream(anchor)) for f, o in zip(fl_nhwc, rm)])


def run(peak_gbs, B=8, H=256, W=832, only=None, once=False):
    _lib.load()
    rows = []
    for name, nbytes, fns in cases(B, H, W):
        if only and only not in name:
            continue
        if once:       # one plain launch per case on cold buffers: the mode `ncu --set full` is run on
            fns[-1]()
            torch.cuda.synchronize()
            print('once', name)
            continue
        us = _time_graph(fns)
        gbs = nbytes / (us * 1e-6) / 1e9
        row = {'kernel': name, 'avg_us': round(us, 2), 'alg_mb': round(nbytes / 1e6, 3), 'achieved_gbs': round(gbs, 1),
               'frac': round(gbs / peak_gbs, 4), 'buffer_sets': len(fns)}
        m = re.match(r'cost_volume_(fwd|bwd)\[(\d+)x(\d+)x(\d+)x(\d+)\]', name)
        if m:      # the cost volume is FP32-FMA bound: also report it against the measured CUDA-core peak (SURVEY 8d FLOPs)
            b_, c_, h_, w_ = (int(v) for v in m.groups()[1:])
            flops = (2 if m.group(1) == 'fwd' else 4) * 81 * c_ * b_ * h_ * w_
            row['tflops'] = round(flops / (us * 1e-6) / 1e12, 2)
            row['fp32_frac'] = round(row['tflops'] / FP32_PEAK_TFLOPS, 3)
        rows.append(row)
        del fns
        torch.cuda.empty_cache()
    return rows


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    ap.add_argument('--peak', type=float, default=6531.9)
    ap.add_argument('--json', default=None)
    ap.add_argument('--once', action='store_true', help='launch every case once, untimed (for ncu --set full)')
    a = ap.parse_args()
    rows = run(a.peak, only=a.only, once=a.once)
    for r_ in rows:
        print('%-58s %9.2f us %9.2f MB %8.1f GB/s  frac %.3f' % (r_['kernel'], r_['avg_us'], r_['alg_mb'], r_['achieved_gbs'], r_['frac']))
    if a.json:
        json.dump(rows, open(a.json, 'w'), indent=1)
