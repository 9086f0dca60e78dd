"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/ and prints oracle-vs-reference errors

The reference has no tests or golden vectors of its own, so the fixtures
are outputs of the reference's own `core.networks` code on seeded synthetic tensors, under
the installed torch (grid_sample align_corners=False) and, for the warp, also with
grid_sample forced to align_corners=True (torch-1.2 semantics).  The script also
asserts that the oracle restatement agrees with the reference on every case it writes.
"""
from __future__ import annotations

import functools
import os
import sys

import numpy as np
import torch

REF = os.environ.get('UOF_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def load_reference():
    """Import the reference's core.networks with the CPU shim for net_utils.py:48."""
    orig = torch.Tensor.get_device
    torch.Tensor.get_device = lambda s: s.device if not s.is_cuda else orig(s)
    sys.path.insert(0, REF)
    import core.networks as ref_networks       # noqa: F401  (registers top-level module names)
    return sys.modules['model_flow_paper'], sys.modules['pwc_tf'], sys.modules['net_utils'], sys.modules['ssim']


def rel(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def npz(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                                    for k, v in arrs.items()})


def flows_like(g, B, H, W, sigma, oob_col=True):
    f = torch.randn(B, 2, H, W, generator=g) * sigma
    if oob_col:
        f[:, 0, :, W - 2] += W          # push one column out of bounds
        f[:, 1, 1, :] -= H              # and one row
    return f


def main():
    from oracle import ops, model as omodel
    mfp, pwc_tf, net_utils, ssim_mod = load_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(1234)
    report = {}

    # ---- App. C smoke vector (net_utils.py:56-60) --------------------------------------
    x = torch.ones(1, 1, 10, 10)
    fl = torch.stack([torch.full((1, 10, 10), 3.0), torch.zeros(1, 10, 10)], 1)
    y_false = net_utils.warp_flow(x, fl)
    gs = net_utils.nn.functional.grid_sample
    net_utils.nn.functional.grid_sample = functools.partial(gs, align_corners=True)
    y_true = net_utils.warp_flow(x, fl)
    net_utils.nn.functional.grid_sample = gs
    report['smoke_ac0'] = rel(ops.warp_flow(x, fl), y_false)
    report['smoke_ac1'] = rel(ops.warp_flow(x, fl, align_corners=True), y_true)
    npz('warp_smoke.npz', y_ac0=y_false, y_ac1=y_true)

    # ---- a1 cost volume ----------------------------------------------------------------
    dec = pwc_tf.PWC_tf()
    for tag, (B, C, H, W) in {'small': (2, 8, 6, 9), 'odd': (1, 5, 11, 7)}.items():
        f1 = torch.randn(B, C, H, W, generator=g, requires_grad=True)
        f2 = torch.randn(B, C, H, W, generator=g, requires_grad=True)
        ct = torch.randn(B, 81, H, W, generator=g)
        out = dec.corr_naive(f1, f2)
        g1, g2 = torch.autograd.grad((out * ct).sum(), (f1, f2))
        o_out = ops.cost_volume(f1, f2)
        o1, o2 = torch.autograd.grad((o_out * ct).sum(), (f1, f2))
        report['corr_' + tag] = max(rel(o_out, out), rel(o1, g1), rel(o2, g2))
        npz('corr_%s.npz' % tag, f1=f1, f2=f2, ct=ct, out=out, g1=g1, g2=g2)

    # ---- a2/a3 warp ---------------------------------------------------------------------
    for tag, (B, C, H, W, sigma) in {'feat': (2, 4, 12, 16, 3.0), 'img': (2, 3, 10, 14, 0.7), 'wild': (1, 3, 9, 13, 20.0)}.items():
        for ac in (False, True):
            for use_mask in (False, True):
                x = torch.rand(B, C, H, W, generator=g, requires_grad=True)
                fl = flows_like(g, B, H, W, sigma).requires_grad_(True)
                ct = torch.randn(B, C, H, W, generator=g)
                if ac:
                    net_utils.nn.functional.grid_sample = functools.partial(gs, align_corners=True)
                out = net_utils.warp_flow(x, fl, use_mask=use_mask)
                gx, gf = torch.autograd.grad((out * ct).sum(), (x, fl))
                net_utils.nn.functional.grid_sample = gs
                o = ops.warp_flow(x, fl, use_mask=use_mask, align_corners=ac)
                ox, of = torch.autograd.grad((o * ct).sum(), (x, fl))
                c, _ = ops.warp_closed_form(x, fl, use_mask=use_mask, align_corners=ac)
                name = 'warp_%s_ac%d_m%d' % (tag, ac, use_mask)
                report[name] = max(rel(o, out), rel(ox, gx), rel(of, gf), rel(c, out))
                npz(name + '.npz', x=x, flow=fl, ct=ct, out=out, gx=gx, gflow=gf)

    # ---- a6 SSIM map --------------------------------------------------------------------
    x = torch.rand(2, 3, 12, 16, generator=g, requires_grad=True)
    y = torch.rand(2, 3, 12, 16, generator=g, requires_grad=True)
    ct = torch.randn(2, 3, 12, 16, generator=g)
    out = ssim_mod.SSIM(x, y)
    gx, gy = torch.autograd.grad((out * ct).sum(), (x, y))
    o = ops.ssim(x, y)
    ox, oy = torch.autograd.grad((o * ct).sum(), (x, y))
    report['ssim'] = max(rel(o, out), rel(ox, gx), rel(oy, gy))
    npz('ssim.npz', x=x, y=y, ct=ct, out=out, gx=gx, gy=gy)

    # ---- a4..a9 loss methods on a 3-scale pyramid ---------------------------------------
    torch.manual_seed(0)
    ref_model = mfp.Model_flow(omodel.Cfg)
    B, H, W = 2, 32, 48
    imgs3 = [torch.rand(B, 3, H, W, generator=g) for _ in range(3)]
    pyr_ref = [ref_model.generate_img_pyramid(i, 4) for i in imgs3]
    pyr_or = [ops.img_pyramid(i, 4) for i in imgs3]
    report['pyramid'] = max(rel(a, b) for pa, pb in zip(pyr_or, pyr_ref) for a, b in zip(pa, pb))
    fb = [flows_like(g, B, H >> s, W >> s, 1.5 / (s + 1), oob_col=(s == 0)).requires_grad_(True) for s in range(4)]
    ff = [flows_like(g, B, H >> s, W >> s, 1.5 / (s + 1), oob_col=False).requires_grad_(True) for s in range(4)]
    cts = torch.randn(4, B, generator=g)

    def run(mod_warp_pyr, dw, lwm, lssim, lsm, lcon, pyr):
        from_l = mod_warp_pyr(pyr[0], fb)
        from_r = mod_warp_pyr(pyr[2], ff)
        d_b, d_f, w_b, w_f = dw(from_l, pyr[1], from_r)
        pack = [lwm(d_f, w_f) + lwm(d_b, w_b),
                lssim(pyr[1], from_r, w_f) + lssim(pyr[1], from_l, w_b),
                lsm(ff, pyr[1]) + lsm(fb, pyr[1]),
                lcon(ff, fb, w_f)]
        total = sum((p * c).sum() for p, c in zip(pack, cts))
        grads = torch.autograd.grad(total, fb[:3] + ff[:3])
        return pack, grads, (from_l, from_r, d_b, d_f, w_b, w_f)

    r_pack, r_grads, r_aux = run(ref_model.warp_flow_pyramid, ref_model.compute_diff_weight,
                                 ref_model.compute_loss_with_mask, ref_model.compute_loss_ssim,
                                 ref_model.compute_loss_flow_smooth, ref_model.compute_loss_flow_consis, pyr_ref)
    o_pack, o_grads, o_aux = run(ops.warp_pyramid, ops.diff_weight, ops.loss_with_mask, ops.loss_ssim,
                                 ops.loss_flow_smooth, ops.loss_flow_consis, pyr_or)
    report['losses'] = max(rel(a, b) for a, b in zip(o_pack, r_pack))
    report['loss_grads'] = max(rel(a, b) for a, b in zip(o_grads, r_grads))
    report['weights'] = max(rel(a, b) for a, b in zip(o_aux[4] + o_aux[5], r_aux[4] + r_aux[5]))
    arrs = {'imgl': imgs3[0], 'img': imgs3[1], 'imgr': imgs3[2], 'cts': cts}
    for s in range(4):
        arrs['fb%d' % s], arrs['ff%d' % s] = fb[s], ff[s]
    for s in range(3):
        arrs['gfb%d' % s], arrs['gff%d' % s] = r_grads[s], r_grads[3 + s]
        arrs['wb%d' % s], arrs['wf%d' % s] = r_aux[4][s], r_aux[5][s]
        arrs['db%d' % s], arrs['df%d' % s] = r_aux[2][s], r_aux[3][s]
    for k, name in enumerate(('loss_pixel', 'loss_ssim', 'loss_flow_smooth', 'loss_flow_consis')):
        arrs[name] = r_pack[k]
    npz('losses.npz', **arrs)

    # ---- a10/a11 full training step, B=1 64x128 and B=2 64x64, and BASELINE.json configs[0] (B=1 256x832) -----------
    for tag, (B, H, W) in {'b1_64x128': (1, 64, 128), 'b2_64x64': (2, 64, 64), 'b1_256x832': (1, 256, 832)}.items():
        torch.manual_seed(0)
        ref_model = mfp.Model_flow(omodel.Cfg)
        torch.manual_seed(0)
        or_model = omodel.Model_flow(omodel.Cfg)
        sd_r, sd_o = ref_model.state_dict(), or_model.state_dict()
        assert list(sd_r.keys()) == list(sd_o.keys()) and len(sd_r) == 98
        assert all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r), 'same seed must give same init'
        inputs = torch.rand(B, 3, 3 * H, W, generator=torch.Generator().manual_seed(1234))
        packs, grads = [], []
        for m in (ref_model, or_model):
            pack = m(inputs)
            loss = ops.total_loss(pack)
            loss.backward()
            packs.append({k: v.detach() for k, v in pack.items()})
            grads.append({k: p.grad.clone() for k, p in m.named_parameters()})
        report['step_' + tag] = max(rel(packs[1][k], packs[0][k]) for k in packs[0])
        gnorm_r = torch.stack([grads[0][k].norm() for k in grads[0]])
        gnorm_o = torch.stack([grads[1][k].norm() for k in grads[1]])
        report['step_grads_' + tag] = max(rel(grads[1][k], grads[0][k]) for k in grads[0] if grads[0][k].abs().max() > 0)
        flows = ref_model.inference_flow(inputs[:, :, H:2 * H], inputs[:, :, 2 * H:])
        npz('step_%s.npz' % tag, grad_norms=gnorm_r, flow_sum=flows.detach().double().sum(),
            flow_abs_mean=flows.detach().abs().mean(),
            predict_flow2_wgrad=grads[0]['pwc_model.predict_flow2.weight'],
            conv1_bgrad=grads[0]['fpyramid.conv1.0.bias'],
            param_keys=np.array(list(sd_r.keys())), **packs[0])
        report['step_gradnorm_' + tag] = rel(gnorm_o, gnorm_r)

    width = max(map(len, report))
    for k, v in report.items():
        print('%-*s  oracle-vs-reference max rel err %.3e' % (width, k, v))
    bad = {k: v for k, v in report.items() if v > 2e-5}
    assert not bad, 'oracle disagrees with the reference: %r' % bad
    print('golden fixtures written to', OUT)


if __name__ == '__main__':
    main()
