"""CPU: the C-ABI library loads and exports every symbol include/uof_b200.h declares (no compute calls),
and the host-side mirror of the reference interface behaves (names, keys, error behaviour)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'uof_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(uof_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    from unopticalflow_b200 import _lib, build
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from unopticalflow_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'libuof_b200.so does not export %s' % s
        assert s in _lib.SIGNATURES or s in _lib.DIAGNOSTICS, 'no ctypes signature for %s' % s
    assert set(_lib.SIGNATURES) | set(_lib.DIAGNOSTICS) == set(syms)
    assert lib.uof_abi_version() == 1


def test_argument_validation_without_gpu(lib):
    """Invalid arguments are rejected before any CUDA call, with a message."""
    from unopticalflow_b200 import _lib
    with pytest.raises(ValueError, match='null pointer'):
        _lib.call('uof_cost_volume_fwd', None, None, None, 1, 1, 1, 1, 81, None)
    with pytest.raises(ValueError, match='bad shape'):
        _lib.call('uof_warp_fwd', ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 0, 3, 4, 4, 0, 0, 0, None)
    with pytest.raises(ValueError, match='channels_last'):
        _lib.call('uof_warp_fwd', ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 3, 4, 4, 0, 0, 1, None)
    with pytest.raises(ValueError, match='nlevels'):
        _lib.call('uof_photo_loss_fwd', None, 0, 1, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), None)
    P = ctypes.c_void_p(16)
    with pytest.raises(ValueError, match='up-sampling only'):
        _lib.call('uof_upsample_bilinear_fwd', P, P, 2, 8, 8, 4, 16, 1.0, None)
    with pytest.raises(ValueError, match='null pointer'):
        _lib.call('uof_upsample_bilinear_bwd', None, P, 2, 4, 4, 8, 8, 1.0, None)
    with pytest.raises(ValueError, match='batch stride'):
        _lib.call('uof_bias_lrelu_bwd2', P, 10, None, 0, P, P, P, 2, 4, 8, 8, 0.1, None)
    with pytest.raises(ValueError, match='bad slot'):
        slots = (ctypes.c_int * 3)(0, 5, 1)
        outs = (ctypes.c_void_p * 2)(16, 16)
        _lib.call('uof_img_pyramid_stacked', P, 64, 192, 64, 8, P, slots, outs, 3, 3, 1, 3, 8, 8, None)
    with pytest.raises(ValueError, match='too large for one launch'):
        _lib.call('uof_splat_fwd', None, P, P, 1, 70000, 4, 1, None)
    with pytest.raises(ValueError, match='nlevels'):
        _lib.call('uof_photo_warp_loss_fwd', None, 0, 1, 0, P, P, P, None)
    with pytest.raises(ValueError, match='coord_flags'):
        _lib.call('uof_photo_warp_loss_bwd', None, 1, 1, 7, P, P, P, None)
    lv = (_lib.PhotoWarpLevel * 1)()
    lv[0] = _lib.PhotoWarpLevel(16, 16, 16, 16, 16, None, None, None, None, None, None, None, None, 8, 7)
    with pytest.raises(RuntimeError, match='even W'):             # UOF_ERR_UNSUPPORTED: the caller takes the separate kernels
        _lib.call('uof_photo_warp_loss_fwd', lv, 1, 1, 0, P, P, P, None)
    lv[0] = _lib.PhotoWarpLevel(16, 16, 16, 16, 16, None, None, None, None, None, None, None, None, 8, 8)
    with pytest.raises(ValueError, match='needs the warped images'):
        _lib.call('uof_photo_warp_loss_bwd', lv, 1, 1, 0, P, P, P, None)


def test_ctypes_signatures_have_the_header_arity():
    """Every ctypes prototype in _lib.SIGNATURES lists as many arguments as the declaration in include/uof_b200.h, and
    pointer / integer / float kinds agree position by position."""
    from unopticalflow_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'uof_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    decls = dict(re.findall(r'\bint\s+(uof_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;', text, flags=re.S))
    assert len(decls) >= 40

    def kind_of_c(param):
        p = ' '.join(param.split())
        if '*' in p or 'uof_stream_t' in p:
            return 'ptr'
        if p.startswith('float') or ' float ' in ' ' + p + ' ':
            return 'float'
        return 'int'

    def kind_of_ctypes(t):
        if t in (ctypes.c_float, ctypes.c_double):
            return 'float'
        if t in (ctypes.c_int, ctypes.c_longlong):
            return 'int'
        return 'ptr'

    for name, argtypes in _lib.SIGNATURES.items():
        params = [p for p in decls[name].split(',') if p.strip() and p.strip() != 'void']
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for i, (p, t) in enumerate(zip(params, argtypes)):
            assert kind_of_c(p) == kind_of_ctypes(t), (name, i, p.strip(), t)


def test_level_structs_match_the_header(tmp_path):
    """The ctypes mirrors of the per-level structs (_lib.PhotoLevel, PhotoWarpLevel, SmoothLevel, ConsisLevel) have the
    size and field offsets a C compiler gives the structs of include/uof_b200.h (plain C: the header must also compile as C)."""
    import shutil
    import subprocess
    from unopticalflow_b200 import _lib
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no C compiler')
    pairs = {'uof_photo_level': _lib.PhotoLevel, 'uof_photo_warp_level': _lib.PhotoWarpLevel,
             'uof_smooth_level': _lib.SmoothLevel, 'uof_consis_level': _lib.ConsisLevel}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "uof_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run([gcc, '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split('\n')
    got = {l.split()[0]: [int(v) for v in l.split()[1:]] for l in out if l.strip()}
    for cname, cls in pairs.items():
        want = [ctypes.sizeof(cls)] + [getattr(cls, f).offset for f, _ in cls._fields_]
        assert got[cname] == want, (cname, got[cname], want)


def test_library_missing_fails_loudly(tmp_path, monkeypatch):
    from unopticalflow_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    with pytest.raises(_lib.LibraryMissing, match='no CPU or PyTorch fallback'):
        _lib.load(str(tmp_path / 'nope.so'))


def test_no_cpu_fallback():
    import unopticalflow_b200 as u
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        u.corr(x, x)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        u.warp_flow(x, torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        u.SSIM(x, x)


def test_round2_host_ops_without_gpu():
    """The round-2 host entry points refuse CPU tensors like every other operator (no fallback), and train.total_loss keeps
    its torch form for CPU loss packs (the gloo tests and the oracle use it)."""
    import unopticalflow_b200 as u
    from unopticalflow_b200 import train
    img = [torch.zeros(1, 3, 8, 8)]
    src = [torch.zeros(2, 3, 8, 8)]
    flo = [torch.zeros(2, 2, 8, 8)]
    for fn in (u.ops.photometric_losses_warped, u.ops.flow_loss_pack):
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            fn(img, src, flo, 1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        u.ops.weighted_mean_sum([torch.zeros(4)], [1.0])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        u.ops.upsample_bilinear_scaled(torch.zeros(1, 2, 4, 4), (8, 8), 2.0)
    pack = {'loss_pixel': torch.tensor([1.0, 3.0]), 'loss_ssim': torch.tensor([2.0, 2.0])}
    assert float(train.total_loss(pack, {'loss_pixel': 0.5, 'loss_ssim': 2.0})) == pytest.approx(0.5 * 2.0 + 2.0 * 2.0)
    a, b = u.ops.split_at(torch.arange(12.0).view(6, 2).requires_grad_(True), 4)
    assert a.shape == (4, 2) and b.shape == (2, 2)
    x = torch.arange(12.0).view(6, 2).requires_grad_(True)
    a, b = u.ops.split_at(x, 4)
    (a.sum() * 2.0 + b.sum() * 3.0).backward()
    assert torch.equal(x.grad, torch.tensor([[2.0, 2.0]] * 4 + [[3.0, 3.0]] * 2))
    x = torch.arange(12.0).view(6, 2).requires_grad_(True)
    a, _ = u.ops.split_at(x, 4)
    a.sum().backward()                                  # the unused part gets a zero gradient
    assert torch.equal(x.grad, torch.tensor([[1.0, 1.0]] * 4 + [[0.0, 0.0]] * 2))


def test_warp_flow_shape_error_matches_reference():
    import unopticalflow_b200 as u
    with pytest.raises(ValueError, match='the shape of grid .* is not equal to the shape of flow'):
        u.warp_flow(torch.zeros(1, 3, 12, 39), torch.zeros(1, 2, 12, 40))      # SURVEY F6 case


def test_model_surface_and_state_dict_keys():
    import unopticalflow_b200 as u
    from oracle import model as omodel
    from util import load_golden
    torch.manual_seed(0)
    m = u.get_model('flow')(omodel.Cfg)
    torch.manual_seed(0)
    o = omodel.Model_flow(omodel.Cfg)
    sd, so = m.state_dict(), o.state_dict()
    assert list(sd.keys()) == list(so.keys()) == [str(k) for k in load_golden('step_b1_64x128.npz')['param_keys']]
    assert len(sd) == 98 and sum(p.numel() for p in m.parameters()) == 5134324
    assert all(torch.equal(sd[k], so[k]) for k in sd), 'same seed -> same init as the reference'
    for name in ('forward', 'inference_flow', 'generate_img_pyramid', 'warp_flow_pyramid', 'compute_diff_weight',
                 'compute_loss_with_mask', 'compute_loss_ssim', 'compute_loss_flow_smooth', 'compute_loss_flow_consis',
                 'get_flow_normalization'):
        assert callable(getattr(m, name))
    assert m.pwc_model.corr is not None and callable(m.pwc_model.warp)
    with pytest.raises(ValueError, match='Mode depth not found'):
        u.get_model('depth')


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under unopticalflow_b200/ may import it."""
    pkg = os.path.join(ROOT, 'unopticalflow_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(dirpath, f)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path: the unmodified reference from baseline/_ref when installed,
    else the oracle port) prints ONE JSON line with the contract's keys and the same `config` shape as the b200 arm;
    under a multi-rank launch every rank but 0 exits without work."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--hw', '64', '128']
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frame-pairs/s' and d['higher_is_better'] is True and d['value'] > 0
    have_ref = os.path.exists(os.path.join(root, 'baseline', '_ref', 'MANIFEST.json'))
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['config']['batch_per_gpu'] == 8 and d['config']['img_hw'] == [64, 128] and 'workload' in d['config']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert other.returncode == 0 and other.stdout.strip() == ''


def test_api_crumbs_and_argument_checks():
    """Host-side behaviour that needs no GPU: the reference's extra forward kwargs are accepted (model_flow_paper.py:205), the
    helper methods exist (:68-87,152-167), loss operators validate shapes before any pointer reaches the C ABI
    and an all-frozen model is reported clearly (train.py:39)."""
    import inspect
    from types import SimpleNamespace
    import torch
    import unopticalflow_b200 as u
    from unopticalflow_b200 import ops, train as T
    sig = inspect.signature(u.Model_flow.forward)
    assert list(sig.parameters)[1:] == ['inputs', 'output_flow', 'use_flow_loss', 'is_second_phase']
    m = u.Model_flow(T.KITTI_CFG)
    for name in ('gradients', 'cal_grad2_error', 'compute_loss_pixel', 'compute_loss_pixel_without_mask'):
        assert callable(getattr(m, name))
    dx, dy = m.gradients(torch.arange(24.).view(1, 1, 4, 6))
    assert dx.shape == (1, 1, 4, 5) and dy.shape == (1, 1, 3, 6) and float(dx.min()) == 1 and float(dy.max()) == 6
    frozen = u.Model_flow(SimpleNamespace(**{**vars(T.KITTI_CFG), 'mode': 'depth'}))
    with pytest.raises(ValueError, match='no trainable parameters'):
        T.make_optimizer(frozen)

    class FakeCuda(torch.Tensor):          # shape checks run before the library is touched: fake the device test only
        is_cuda = True
    fake = lambda *shape: torch.zeros(*shape).as_subclass(FakeCuda)
    with pytest.raises(ValueError, match='flow_smooth_loss level 0'):
        ops._SmoothLoss.forward(None, 1, fake(2, 2, 8, 8), fake(2, 1, 8, 8))          # 1-channel image
    with pytest.raises(ValueError, match='multiple of the image batch'):
        ops._SmoothLoss.forward(None, 1, fake(3, 2, 8, 8), fake(2, 3, 8, 8))
    with pytest.raises(ValueError, match='flow_consis_loss level 0'):
        ops._ConsisLoss.forward(None, 1, fake(2, 2, 8, 8), fake(2, 2, 8, 8), fake(2, 1, 4, 4))   # weight map of another level
    with pytest.raises(ValueError, match='flow_consis_loss level 0'):
        ops._ConsisLoss.forward(None, 1, fake(2, 2, 8, 8), fake(1, 2, 8, 8), fake(2, 1, 8, 8))   # bwd flows of another batch


def test_checkpoint_prefix_handling(tmp_path):
    """train.py:23-31,50-60: the reference's checkpoint layout; `module.` / `model_flow.` prefixes are stripped on load."""
    import torch
    import unopticalflow_b200 as u
    from unopticalflow_b200 import train as T
    torch.manual_seed(1)
    m = u.Model_flow(T.KITTI_CFG)
    opt = torch.optim.Adam([{'params': T.trainable_parameters(m), 'lr': 1e-4}])
    T.save_model(41, str(tmp_path), 'iter_41.pth', m, opt)
    data = torch.load(str(tmp_path / 'iter_41.pth'))
    assert set(data) == {'iteration', 'model_state_dict', 'optimizer_state_dict'}
    data['model_state_dict'] = {'module.model_flow.' + k: v for k, v in data['model_state_dict'].items()}
    torch.save(data, str(tmp_path / 'last.pth'))
    m2 = u.Model_flow(T.KITTI_CFG)
    it, _, _ = T.load_model(str(tmp_path), 'last.pth', m2, torch.optim.Adam([{'params': T.trainable_parameters(m2), 'lr': 1e-4}]))
    assert it == 41 and all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
