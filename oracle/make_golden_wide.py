"""Generate tests/golden/losses_wide.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_wide          # build container only (needs /root/reference)

Same loss chain as the `losses.npz` fixture of oracle/make_golden.py (warp_flow_pyramid with the validity mask ->
compute_diff_weight -> compute_loss_with_mask / compute_loss_ssim / compute_loss_flow_smooth / compute_loss_flow_consis,
model_flow_paper.py:62-66,90-195) on a WIDE frame (16 x 200: several 60-column strips of the CUDA kernels with a partial
last one) with decoder-like flows: a smooth low-resolution field, bilinearly up-sampled, plus noise, with one column and
one row pushed out of bounds.  A separate script so that the existing fixtures stay byte-identical.
"""
from __future__ import annotations

import torch

from oracle.make_golden import load_reference, npz, rel


def main():
    from oracle import ops, model as omodel
    mfp, _, _, _ = load_reference()
    g = torch.Generator().manual_seed(4321)
    torch.manual_seed(0)
    ref_model = mfp.Model_flow(omodel.Cfg)
    B, H, W, S = 1, 16, 200, 3
    imgs3 = [torch.rand(B, 3, H, W, generator=g) for _ in range(3)]
    pyr_ref = [ref_model.generate_img_pyramid(i, 4) for i in imgs3]
    pyr_or = [ops.img_pyramid(i, 4) for i in imgs3]

    def flows():
        lo = (torch.rand(B, 2, 2, 25, generator=g) - 0.5) * 8.0
        out = []
        for s in range(4):
            h, w = H >> s, W >> s
            f = torch.nn.functional.interpolate(lo, size=(h, w), mode='bilinear', align_corners=False) / (1 << s)
            f = f + (torch.rand(B, 2, h, w, generator=g) - 0.5) * 0.5
            if s == 0:
                f[:, 0, :, w - 2] += w
                f[:, 1, 1, :] -= h
            out.append(f.requires_grad_(True))
        return out

    fb, ff = flows(), flows()
    cts = torch.randn(4, B, generator=g)

    def run(warp_pyr, dw, lwm, lssim, lsm, lcon, pyr):
        from_l, from_r = warp_pyr(pyr[0], fb), warp_pyr(pyr[2], ff)
        d_b, d_f, w_b, w_f = dw(from_l, pyr[1], from_r)
        pack = [lwm(d_f, w_f) + lwm(d_b, w_b), lssim(pyr[1], from_r, w_f) + lssim(pyr[1], from_l, w_b),
                lsm(ff, pyr[1]) + lsm(fb, pyr[1]), lcon(ff, fb, w_f)]
        total = sum((p * c).sum() for p, c in zip(pack, cts))
        return pack, torch.autograd.grad(total, fb[:S] + ff[:S]), (w_b, w_f)

    r_pack, r_grads, r_w = run(ref_model.warp_flow_pyramid, ref_model.compute_diff_weight, ref_model.compute_loss_with_mask,
                               ref_model.compute_loss_ssim, ref_model.compute_loss_flow_smooth,
                               ref_model.compute_loss_flow_consis, pyr_ref)
    o_pack, o_grads, o_w = run(ops.warp_pyramid, ops.diff_weight, ops.loss_with_mask, ops.loss_ssim, ops.loss_flow_smooth,
                               ops.loss_flow_consis, pyr_or)
    report = {'losses': max(rel(a, b) for a, b in zip(o_pack, r_pack)),
              'loss_grads': max(rel(a, b) for a, b in zip(o_grads, r_grads)),
              'weights': max(rel(a, b) for a, b in zip(o_w[0] + o_w[1], r_w[0] + r_w[1]))}
    arrs = {'imgl': imgs3[0], 'img': imgs3[1], 'imgr': imgs3[2], 'cts': cts}
    for s in range(4):
        arrs['fb%d' % s], arrs['ff%d' % s] = fb[s], ff[s]
    for s in range(S):
        arrs['gfb%d' % s], arrs['gff%d' % s] = r_grads[s], r_grads[S + s]
        arrs['wb%d' % s], arrs['wf%d' % s] = r_w[0][s], r_w[1][s]
    for k, name in enumerate(('loss_pixel', 'loss_ssim', 'loss_flow_smooth', 'loss_flow_consis')):
        arrs[name] = r_pack[k]
    npz('losses_wide.npz', **arrs)
    for k, v in report.items():
        print('%-12s oracle-vs-reference max rel err %.3e' % (k, v))
    assert all(v <= 2e-5 for v in report.values()), report
    print('wrote tests/golden/losses_wide.npz')


if __name__ == '__main__':
    main()
