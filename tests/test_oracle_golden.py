"""CPU: the oracle restatement against the fixtures produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle."""
import torch

from oracle import model as omodel
from oracle import ops as O
from util import assert_close, load_golden

TOL = 2e-5


def test_smoke_vector_app_c():
    g = load_golden('warp_smoke.npz')
    x = torch.ones(1, 1, 10, 10)
    fl = torch.stack([torch.full((1, 10, 10), 3.0), torch.zeros(1, 10, 10)], 1)
    y0 = O.warp_flow(x, fl)
    assert torch.equal(y0, g['y_ac0'])
    # SURVEY App. C, recorded by hand from the reference's net_utils.py:56-60 smoke input
    assert y0[0, 0, 0].tolist() == [.5, .5, .5, .5, .5, .5, .25, 0, 0, 0]
    assert y0[0, 0, 4].tolist() == [1, 1, 1, 1, 1, 1, .5, 0, 0, 0]
    y1 = O.warp_flow(x, fl, align_corners=True)
    assert torch.equal(y1, g['y_ac1'])
    assert y1[0, 0, 3].tolist() == [1, 1, 1, 1, 1, 1, 1, 0, 0, 0]


def test_cost_volume_golden():
    for tag in ('small', 'odd'):
        g = load_golden('corr_%s.npz' % tag)
        f1, f2 = g['f1'].requires_grad_(True), g['f2'].requires_grad_(True)
        out = O.cost_volume(f1, f2)
        g1, g2 = torch.autograd.grad((out * g['ct']).sum(), (f1, f2))
        assert_close(out, g['out'], TOL, 'corr ' + tag)
        assert_close(g1, g['g1'], TOL)
        assert_close(g2, g['g2'], TOL)


def test_cost_volume_scalar_loop():
    g = torch.Generator().manual_seed(3)
    f1, f2 = torch.randn(1, 3, 5, 6, generator=g), torch.randn(1, 3, 5, 6, generator=g)
    out = O.cost_volume(f1, f2)
    for (i, j, y, x) in [(0, 0, 4, 5), (4, 4, 2, 3), (8, 8, 0, 0), (3, 7, 1, 2), (8, 0, 4, 0)]:
        yy, xx = y + i - 4, x + j - 4
        ref = 0.0 if not (0 <= yy < 5 and 0 <= xx < 6) else float((f1[0, :, y, x] * f2[0, :, yy, xx]).sum() / 3)
        assert abs(float(out[0, 9 * i + j, y, x]) - ref) < 1e-6


def test_warp_golden_and_closed_form():
    for tag in ('feat', 'img', 'wild'):
        for ac in (0, 1):
            for m in (0, 1):
                g = load_golden('warp_%s_ac%d_m%d.npz' % (tag, ac, m))
                x, fl = g['x'].requires_grad_(True), g['flow'].requires_grad_(True)
                out = O.warp_flow(x, fl, use_mask=bool(m), align_corners=bool(ac))
                gx, gf = torch.autograd.grad((out * g['ct']).sum(), (x, fl))
                assert_close(out, g['out'], TOL)
                assert_close(gx, g['gx'], TOL)
                assert_close(gf, g['gflow'], TOL)
                cf, _ = O.warp_closed_form(x, fl, use_mask=bool(m), align_corners=bool(ac))
                assert_close(cf, g['out'], TOL)


def test_ssim_golden():
    g = load_golden('ssim.npz')
    x, y = g['x'].requires_grad_(True), g['y'].requires_grad_(True)
    out = O.ssim(x, y)
    gx, gy = torch.autograd.grad((out * g['ct']).sum(), (x, y))
    assert_close(out, g['out'], TOL)
    assert_close(gx, g['gx'], TOL)
    assert_close(gy, g['gy'], TOL)


def oracle_losses(g):
    """Run the oracle loss chain on the inputs stored in losses.npz."""
    pyr = [O.img_pyramid(g[k], 4) for k in ('imgl', 'img', 'imgr')]
    fb = [g['fb%d' % s].clone().requires_grad_(True) for s in range(4)]
    ff = [g['ff%d' % s].clone().requires_grad_(True) for s in range(4)]
    from_l, from_r = O.warp_pyramid(pyr[0], fb), O.warp_pyramid(pyr[2], ff)
    d_b, d_f, w_b, w_f = O.diff_weight(from_l, pyr[1], from_r)
    pack = [O.loss_with_mask(d_f, w_f) + O.loss_with_mask(d_b, w_b),
            O.loss_ssim(pyr[1], from_r, w_f) + O.loss_ssim(pyr[1], from_l, w_b),
            O.loss_flow_smooth(ff, pyr[1]) + O.loss_flow_smooth(fb, pyr[1]),
            O.loss_flow_consis(ff, fb, w_f)]
    total = sum((p * c).sum() for p, c in zip(pack, g['cts']))
    grads = torch.autograd.grad(total, fb[:3] + ff[:3])
    return pack, grads, w_b, w_f


def test_losses_golden():
    g = load_golden('losses.npz')
    pack, grads, w_b, w_f = oracle_losses(g)
    for k, name in enumerate(('loss_pixel', 'loss_ssim', 'loss_flow_smooth', 'loss_flow_consis')):
        assert_close(pack[k], g[name], TOL, name)
    for s in range(3):
        assert_close(grads[s], g['gfb%d' % s], TOL)
        assert_close(grads[3 + s], g['gff%d' % s], TOL)
        assert_close(w_b[s], g['wb%d' % s], TOL)
        assert_close(w_f[s], g['wf%d' % s], TOL)


def test_losses_wide_golden():
    """The same chain on the wide-frame fixture (oracle/make_golden_wide.py: 16 x 200, decoder-like flows)."""
    g = load_golden('losses_wide.npz')
    pack, grads, w_b, w_f = oracle_losses(g)
    for k, name in enumerate(('loss_pixel', 'loss_ssim', 'loss_flow_smooth', 'loss_flow_consis')):
        assert_close(pack[k], g[name], TOL, name)
    for s in range(3):
        assert_close(grads[s], g['gfb%d' % s], TOL)
        assert_close(grads[3 + s], g['gff%d' % s], TOL)
        assert_close(w_b[s], g['wb%d' % s], TOL)
        assert_close(w_f[s], g['wf%d' % s], TOL)


def test_step_golden():
    for tag, (B, H, W) in {'b1_64x128': (1, 64, 128), 'b2_64x64': (2, 64, 64), 'b1_256x832': (1, 256, 832)}.items():
        g = load_golden('step_%s.npz' % tag)
        torch.manual_seed(0)
        m = omodel.Model_flow(omodel.Cfg)
        assert list(m.state_dict().keys()) == [str(k) for k in g['param_keys']]
        x = torch.rand(B, 3, 3 * H, W, generator=torch.Generator().manual_seed(1234))
        pack = m(x)
        O.total_loss(pack).backward()
        for k in pack:
            assert_close(pack[k], g[k], TOL, k)
        gn = torch.stack([p.grad.norm() for p in m.parameters()])
        assert_close(gn, g['grad_norms'], 1e-4)
        assert_close(m.pwc_model.predict_flow2.weight.grad, g['predict_flow2_wgrad'], 1e-4)


def test_splat_oracle_properties():
    """a12 is unpinned by the reference; check the oracle against first principles instead."""
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 9, 11
    flow = torch.randn(B, H, W, 2, generator=g) * 2
    r = O.range_map(flow)
    idx, inb, wts = O.splat_targets(flow)
    # mass conservation: splatted mass == sum of in-bounds weights
    assert abs(float(r.sum()) - float((wts * inb).sum())) < 1e-3
    # integer flow = pure shift
    flow0 = torch.zeros(B, H, W, 2)
    flow0[..., 0] = 2.0
    u = torch.rand(B, H, W, 3, generator=g)
    out = O.splat(u, flow0)
    assert torch.allclose(out[:, :, 2:], u[:, :, :-2], atol=1e-6) and float(out[:, :, :2].abs().max()) == 0.0
