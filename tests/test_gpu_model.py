"""GPU parity at step level: Model_flow forward + train.py's weighted loss + backward, CUDA path vs the CPU
oracle with identical weights and inputs, and vs the golden values recorded from the unmodified reference."""
import pytest
import torch

from oracle import model as omodel
from oracle import ops as O
from util import REL_TOL, assert_close, load_golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def U(cuda):
    import unopticalflow_b200 as u
    return u


def build_pair(U):
    torch.manual_seed(0)
    ref = omodel.Model_flow(omodel.Cfg)
    m = U.Model_flow(omodel.Cfg)
    m.load_state_dict(ref.state_dict())
    return ref, m.cuda()


@pytest.mark.parametrize('tag,B,H,W', [('b1_64x128', 1, 64, 128), ('b2_64x64', 2, 64, 64)])
def test_step_matches_oracle_and_reference_golden(U, tag, B, H, W):
    ref, m = build_pair(U)
    x = torch.rand(B, 3, 3 * H, W, generator=torch.Generator().manual_seed(1234))
    rp = ref(x)
    O.total_loss(rp).backward()
    gp = m(x.cuda())
    O.total_loss(gp).backward()
    g = load_golden('step_%s.npz' % tag)
    for k in rp:
        assert gp[k].shape == (B,)
        assert_close(gp[k], rp[k], REL_TOL, k + ' vs oracle')
        assert_close(gp[k], g[k], REL_TOL, k + ' vs reference golden')
    # parameter gradients: every tensor within 1e-4 of the largest gradient entry of that tensor ...
    worst = 0.0
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        if q.grad.abs().max() == 0:
            assert p.grad.abs().max() == 0
            continue
        worst = max(worst, rel_err(p.grad, q.grad))
    assert worst <= 5e-4, 'worst per-tensor parameter-gradient rel err %.3e' % worst
    # ... and the global gradient vector within 1e-4
    gv = torch.cat([p.grad.flatten().cpu() for p in m.parameters()])
    rv = torch.cat([q.grad.flatten() for q in ref.parameters()])
    assert float((gv - rv).norm() / rv.norm()) <= REL_TOL
    gn = torch.stack([p.grad.norm() for p in m.parameters()]).cpu()
    assert_close(gn, g['grad_norms'], 2e-4, 'grad norms vs reference golden')


def test_inference_flow_matches_oracle(U):
    ref, m = build_pair(U)
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(2, 3, 64, 128, generator=g), torch.rand(2, 3, 64, 128, generator=g)
    with torch.no_grad():
        fr = ref.inference_flow(a, b)
        fg = m.inference_flow(a.cuda(), b.cuda())
    assert fg.shape == (2, 2, 64, 128)
    assert_close(fg, fr, REL_TOL)


def test_align_corners_true_mode(U):
    """torch-1.2 convention end to end."""
    torch.manual_seed(0)
    ref = omodel.Model_flow(omodel.Cfg, align_corners=True)
    m = U.Model_flow(omodel.Cfg, align_corners=True)
    m.load_state_dict(ref.state_dict())
    m = m.cuda()
    x = torch.rand(1, 3, 192, 128, generator=torch.Generator().manual_seed(1234))
    rp, gp = ref(x), m(x.cuda())
    for k in rp:
        assert_close(gp[k], rp[k], REL_TOL, k)


def test_train_steps_track_oracle(U):
    """Three Adam steps (train.py:137-152): losses stay within tolerance of the CPU oracle's trajectory."""
    ref, m = build_pair(U)
    opt_r, opt_g = omodel.make_optimizer(ref), omodel.make_optimizer(m)
    gen = torch.Generator().manual_seed(1234)
    for step in range(3):
        x = torch.rand(1, 3, 192, 128, generator=gen)
        lr, _ = omodel.train_step(ref, opt_r, x)
        lg, _ = omodel.train_step(m, opt_g, x.cuda())
        assert abs(float(lg) - float(lr)) <= 2e-4 * abs(float(lr)), (step, float(lg), float(lr))


def test_occlusion_extras(U):
    _, m = build_pair(U)
    f = torch.randn(2, 2, 32, 48, device='cuda') * 2
    occ = m.get_occlusion_mask_from_flow(f)
    assert occ.shape == (2, 1, 32, 48) and float(occ.min()) >= 0 and float(occ.max()) <= 1
    ref = O.occlusion_mask(f.cpu().permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    assert_close(occ, ref, REL_TOL)
    noc = m.get_consistent_mask(f, -f)
    assert noc.shape == (2, 1, 32, 48)
