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
    err, noise = global_grad_err_vs_same_gpu_oracle(m, ref, x)
    assert noise <= NOISE_CAP, 'run-to-run spread %.2e: a noisy kernel may not widen its own tolerance' % noise
    assert err <= REL_TOL + noise, (err, noise)


# The run-to-run spread that may be added to the 1e-4 bar is itself capped.
NOISE_CAP = 5e-5


def global_grad_err_vs_same_gpu_oracle(m, ref, x, **loss_kw):
    """The CPU oracle and the CUDA path run different convolution back ends (mkldnn vs cuDNN), which alone moves parameter
    gradients by up to ~1e-4.  Running the oracle's op chain on the SAME GPU (same cuDNN convolutions, TF32 off) isolates the
    hand-written kernels: relative L2 error of the whole gradient vector.

    Returns (err, noise).  `noise` is the run-to-run spread of the two implementations themselves (each run twice on the
    same input): both scatter with fp32 atomics (ATen's grid_sampler backward, our warp / cost-volume kernels), and the
    smoothness loss back-propagates the SIGN of second differences that are pure rounding noise on the bilinearly
    up-sampled flows, so a last-bit difference upstream flips whole gradient entries.  At 448x1024 that spread reaches
    ~1e-4 by itself; the parity bar is applied to the error in excess of it."""
    import copy
    ref_gpu = copy.deepcopy(ref).cuda()
    xc = x.cuda()

    def grad_of(model):
        model.zero_grad(set_to_none=True)
        O.total_loss(model(xc), **loss_kw).backward()
        return torch.cat([p.grad.flatten() for p in model.parameters()]).clone()

    gv, gv2 = grad_of(m), grad_of(m)
    rv, rv2 = grad_of(ref_gpu), grad_of(ref_gpu)
    n = float(rv.norm())
    noise = float((gv - gv2).norm()) / n + float((rv - rv2).norm()) / n
    return float((gv - rv).norm()) / n, noise


def test_step_headline_config_b8_256x832(U):
    """BASELINE.json configs[1] -- the shape bench.py times: B=8 triplets of 256x832.  Loss pack and the global parameter
    gradient of the CUDA path against the oracle's op chain on the same GPU (same cuDNN convolutions, TF32 off; the CPU
    oracle needs ~1 min for this batch), every sample of the batch.  This is the only place the TMA-persistent cost volume,
    the +gx feature-warp backward and the 3B/2B batching are compared end to end at the benchmark shape."""
    import copy
    ref, m = build_pair(U)
    x = torch.rand(8, 3, 3 * 256, 832, generator=torch.Generator().manual_seed(1234))
    ref_gpu = copy.deepcopy(ref).cuda()
    xc = x.cuda()
    with torch.no_grad():
        rp = ref_gpu(xc)
    gp = m(xc)
    for k in rp:
        assert gp[k].shape == (8,)
        assert_close(gp[k], rp[k], REL_TOL, k + ' vs same-GPU oracle chain, B=8 256x832')
    del rp, gp, ref_gpu
    err, noise = global_grad_err_vs_same_gpu_oracle(m, ref, x)
    print('headline-config gradient: err vs same-GPU oracle %.2e, run-to-run spread %.2e' % (err, noise))
    assert noise <= NOISE_CAP, noise
    assert err <= REL_TOL + noise, (err, noise)


def test_step_config0_b1_256x832_vs_cpu_oracle(U):
    """BASELINE.json configs[0]: one 256x832 triplet, CUDA path against the CPU oracle (the reference's CPU-runnable case)
    and against the step-level known answer recorded from the unmodified reference (SURVEY App. C recipe)."""
    ref, m = build_pair(U)
    x = torch.rand(1, 3, 3 * 256, 832, generator=torch.Generator().manual_seed(1234))
    rp = ref(x)
    O.total_loss(rp).backward()
    gp = m(x.cuda())
    O.total_loss(gp).backward()
    for k in rp:
        assert_close(gp[k], rp[k], REL_TOL, k + ' vs CPU oracle, B=1 256x832')
    g = load_golden('step_b1_256x832.npz')
    for k in rp:
        assert_close(gp[k], g[k], REL_TOL, k + ' vs reference golden, B=1 256x832')
    gv = torch.cat([p.grad.flatten().cpu() for p in m.parameters()])
    rv = torch.cat([q.grad.flatten() for q in ref.parameters()])
    assert float((gv - rv).norm() / rv.norm()) <= 3e-4          # mkldnn vs cuDNN convolutions
    err, noise = global_grad_err_vs_same_gpu_oracle(m, ref, x)
    assert noise <= NOISE_CAP, noise
    assert err <= REL_TOL + noise, (err, noise)


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


def test_graphed_train_step_matches_eager(U):
    """The CUDA-graph replay of the whole iteration (train.GraphedTrainStep) follows the eager trajectory from the SAME
    starting point: the constructor's warm-up iterations are undone (parameters and Adam state restored) before capture."""
    from unopticalflow_b200 import train as T
    torch.manual_seed(0)
    m1, m2 = U.Model_flow(T.KITTI_CFG).cuda(), U.Model_flow(T.KITTI_CFG).cuda()
    m2.load_state_dict(m1.state_dict())
    w = T.generate_loss_weights_dict(T.KITTI_CFG)
    gen = torch.Generator().manual_seed(1234)
    xs = [torch.rand(2, 3, 192, 128, generator=gen).cuda() for _ in range(3)]
    opt = T.make_optimizer(m1)
    graphed = T.GraphedTrainStep(m2, xs[0], w, warmup=3)
    for p, q in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(p, q), 'GraphedTrainStep construction moved the parameters'
    assert all(float(st['step']) == 0 and float(st['exp_avg'].abs().max()) == 0 for st in graphed.optimizer.state.values())
    for i in range(4):
        le = float(T.train_step(m1, opt, xs[i % 3], w))
        lg = float(graphed(xs[i % 3]))
        assert abs(le - lg) <= 2e-4 * abs(le), (i, le, lg)


def test_checkpoint_roundtrip_reference_layout(U, tmp_path):
    """train.py:23-31: save_model / load_model with the reference's dict layout; resuming a graphed step from a checkpoint
    (model + Adam state) continues the trajectory of the run that wrote it."""
    from unopticalflow_b200 import train as T
    torch.manual_seed(0)
    w = T.generate_loss_weights_dict(T.KITTI_CFG)
    gen = torch.Generator().manual_seed(7)
    xs = [torch.rand(1, 3, 192, 128, generator=gen).cuda() for _ in range(4)]
    m1 = U.Model_flow(T.KITTI_CFG).cuda()
    run1 = T.GraphedTrainStep(m1, xs[0], w, warmup=2)
    for i in range(2):
        run1(xs[i])
    T.save_model(1, str(tmp_path), 'last.pth', m1, run1)
    data = torch.load(str(tmp_path / 'last.pth'))
    assert set(data) == {'iteration', 'model_state_dict', 'optimizer_state_dict'} and len(data['model_state_dict']) == 98
    expect = [float(run1(xs[i])) for i in (2, 3)]
    m2 = U.Model_flow(T.KITTI_CFG).cuda()
    run2 = T.GraphedTrainStep(m2, xs[0], w, warmup=2)
    it, _, _ = T.load_model(str(tmp_path), 'last.pth', m2, run2)
    assert it == 1
    got = [float(run2(xs[i])) for i in (2, 3)]
    for a, b in zip(got, expect):
        assert abs(a - b) <= 2e-4 * abs(b), (got, expect)
    # a DataParallel-style checkpoint ('module.' prefix, train.py:36-37) loads into the bare model
    data['model_state_dict'] = {'module.' + k: v for k, v in data['model_state_dict'].items()}
    torch.save(data, str(tmp_path / 'dp.pth'))
    T.load_model(str(tmp_path), 'dp.pth', U.Model_flow(T.KITTI_CFG).cuda(), None)


def test_graphed_step_staged_inputs_and_flat_gradients(U):
    """The two additions of the data-parallel step at world size 1: (a) gradients living in one flat buffer
    (train.FlatGradAllReduce, the all-reduce is a no-op without a process group) and (b) the staged input pipeline
    (stage(next) overlapping step()) must give the same loss trajectory as the plain graphed step fed directly."""
    from unopticalflow_b200 import train as T
    torch.manual_seed(0)
    m1, m2 = U.Model_flow(T.KITTI_CFG).cuda(), U.Model_flow(T.KITTI_CFG).cuda()
    m2.load_state_dict(m1.state_dict())
    w = T.generate_loss_weights_dict(T.KITTI_CFG)
    gen = torch.Generator().manual_seed(99)
    host = [torch.rand(2, 3, 192, 128, generator=gen).pin_memory() for _ in range(3)]
    direct = T.GraphedTrainStep(m1, host[0].cuda(), w, warmup=3)
    staged = T.GraphedTrainStep(m2, host[0].cuda(), w, warmup=3, allreduce=True)
    assert all(p.grad.data_ptr() >= staged.exchange.flat.data_ptr() for p in m2.parameters())
    staged.stage(host[0])
    for i in range(5):
        ld = float(direct(host[i % 3]))
        loss = staged()
        staged.stage(host[(i + 1) % 3])
        ls = float(loss)
        assert abs(ld - ls) <= 2e-4 * abs(ld), (i, ld, ls)


def test_occlusion_extras(U):
    _, m = build_pair(U)
    f = torch.randn(2, 2, 32, 48, device='cuda') * 2
    occ = m.get_occlusion_mask_from_flow(f)
    assert occ.shape == (2, 1, 32, 48) and float(occ.min()) >= 0 and float(occ.max()) <= 1
    ref = O.occlusion_mask(f.cpu().permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    assert_close(occ, ref, REL_TOL)
    noc = m.get_consistent_mask(f, -f)
    assert noc.shape == (2, 1, 32, 48)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs 3 and 4 (375x1242 and 436x1024 are rejected by the network itself, so config 3 is an
# op-level test on the encoder-map shapes of a 375x1242 frame and config 4 runs at the padded 448x1024 sintel shape)
EVAL_SHAPES = [(16, 188, 621), (32, 94, 311), (64, 47, 156), (96, 24, 78), (128, 12, 39), (196, 6, 20)]


@pytest.mark.parametrize('C,h,w', EVAL_SHAPES)
def test_config3_eval_shape_ops(U, C, h, w):
    """cost volume + feature warp + occlusion masks at ceil(375/2^k) x ceil(1242/2^k), B=1 (inference, forward only)."""
    g = torch.Generator().manual_seed(C + h)
    f1, f2 = torch.randn(1, C, h, w, generator=g), torch.randn(1, C, h, w, generator=g)
    flow = torch.randn(1, 2, h, w, generator=g) * 2.0
    with torch.no_grad():
        assert_close(U.corr(f1.cuda(), f2.cuda()), O.cost_volume(f1, f2), REL_TOL, 'corr')
        assert_close(U.warp_flow(f2.cuda(), flow.cuda()), O.warp_flow(f2, flow), REL_TOL, 'warp')
        nhwc = flow.permute(0, 2, 3, 1).contiguous()
        assert torch.equal(U.ops.splat_targets(nhwc.cuda()).cpu(), O.splat_targets(nhwc)[0])
        occ_ref = O.occlusion_mask(nhwc)
        occ = U.ops.occlusion_mask(nhwc.cuda()).cpu()
        assert float(((occ > 0.5) == (occ_ref > 0.5)).float().mean()) >= 0.999
        noc_ref = O.fb_consistency_mask(flow, -flow + 0.5, 3.0, 0.05)
        noc = U.ops.fb_consistency_mask(flow.cuda(), (-flow + 0.5).cuda(), 3.0, 0.05).cpu()
        assert float((noc == noc_ref).float().mean()) >= 0.999


def test_config3_padded_inference(U):
    """Full inference_flow on the 384x1280 padded KITTI-2015 frame (the reference rejects 375x1242 itself)."""
    ref, m = build_pair(U)
    g = torch.Generator().manual_seed(8)
    a, b = torch.rand(1, 3, 384, 1280, generator=g), torch.rand(1, 3, 384, 1280, generator=g)
    with torch.no_grad():
        fg = m.inference_flow(a.cuda(), b.cuda())
        fr = ref.inference_flow(a, b)
    assert fg.shape == (1, 2, 384, 1280)
    assert_close(fg, fr, REL_TOL)
    with pytest.raises(ValueError, match='the shape of grid'):
        m.inference_flow(torch.rand(1, 3, 375, 1242).cuda(), torch.rand(1, 3, 375, 1242).cuda())


def test_config4_sintel_shape_step(U):
    """sintel 448x1024: one triplet against the CPU oracle (values + global gradient), then the full batch of 16 for
    finiteness, per-sample independence and run-to-run agreement (atomics make the last bits order dependent)."""
    ref, m = build_pair(U)
    gen = torch.Generator().manual_seed(1234)
    x = torch.rand(16, 3, 3 * 448, 1024, generator=gen)
    rp = ref(x[:1])
    O.total_loss(rp, w_smooth=6.0).backward()
    gp = m(x[:1].cuda())
    O.total_loss(gp, w_smooth=6.0).backward()
    for k in rp:
        assert_close(gp[k], rp[k], REL_TOL, k)
    gv = torch.cat([p.grad.flatten().cpu() for p in m.parameters()])
    rv = torch.cat([q.grad.flatten() for q in ref.parameters()])
    assert float((gv - rv).norm() / rv.norm()) <= 3e-4          # CPU (mkldnn) vs GPU (cuDNN) convolutions differ by ~1e-4
    err, noise = global_grad_err_vs_same_gpu_oracle(m, ref, x[:1], w_smooth=6.0)
    assert noise <= 2 * NOISE_CAP, 'run-to-run spread %.2e' % noise        # 459k pixels of sign-of-noise smoothness gradients
    assert err <= REL_TOL + noise, (err, noise)
    print('sintel-shape gradient: err vs same-GPU oracle %.2e, run-to-run spread %.2e' % (err, noise))
    m.zero_grad(set_to_none=True)
    xc = x.cuda()
    p1 = m(xc)
    O.total_loss(p1, w_smooth=6.0).backward()
    p2 = m(xc)
    for k in p1:
        assert p1[k].shape == (16,) and bool(torch.isfinite(p1[k]).all())
        assert_close(p2[k], p1[k], 1e-5, k + ' run-to-run')
        assert_close(p1[k][:1], gp[k].detach(), REL_TOL, k + ' sample 0 is independent of the rest of the batch')
    assert all(bool(torch.isfinite(p.grad).all()) for p in m.parameters())
