"""CPU, world_size 2, gloo: the N>1 host path.  The step shards by sample: DDP-averaged gradients of
two ranks holding B/2 samples each must equal the single-process gradients of the concatenated batch, using the
product's own train-step host code (unopticalflow_b200.train) around a CPU stand-in model (the oracle)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import model as omodel
    from unopticalflow_b200 import train as T
    torch.manual_seed(0)
    model = omodel.Model_flow(omodel.Cfg)
    ddp = torch.nn.parallel.DistributedDataParallel(model, gradient_as_bucket_view=True, bucket_cap_mb=8)
    w = T.generate_loss_weights_dict(T.KITTI_CFG)
    x = torch.rand(2, 3, 192, 64, generator=torch.Generator().manual_seed(1234))     # global batch of 2
    pack = ddp(x[rank:rank + 1])
    T.total_loss(pack, w).backward()
    grads = torch.cat([p.grad.flatten() for p in model.parameters()])
    # the graph-capturable exchange (train.FlatGradAllReduce): same model, gradients in one flat buffer, one all-reduce
    torch.manual_seed(0)
    model2 = omodel.Model_flow(omodel.Cfg)
    ex = T.FlatGradAllReduce(list(model2.parameters()))
    for _ in range(2):                       # twice: zero() must clear what the previous step accumulated
        ex.zero()
        T.total_loss(model2(x[rank:rank + 1]), w).backward()
        ex.allreduce()
    assert all(p.grad.data_ptr() >= ex.flat.data_ptr() for p in model2.parameters()), 'grads must stay views of the flat buffer'
    flat = torch.cat([p.grad.flatten() for p in model2.parameters()])
    # the overlapped form: observe one backward, re-lay the buffer in completion order, early chunk reduced from a hook
    torch.manual_seed(0)
    model3 = omodel.Model_flow(omodel.Cfg)
    ex3 = T.FlatGradAllReduce(list(model3.parameters()))
    ex3.observe()
    ex3.zero()
    T.total_loss(model3(x[rank:rank + 1]), w).backward()
    n_early = ex3.plan_overlap()
    assert 0 < n_early < len(ex3.params) and 0 < ex3.n_early < ex3.flat.numel()
    for _ in range(2):
        ex3.zero()
        T.total_loss(model3(x[rank:rank + 1]), w).backward()
        assert ex3._early_done, 'the early chunk must have been launched from the gradient hooks during backward'
        ex3.allreduce()
    flat_overlap = torch.cat([p.grad.flatten() for p in model3.parameters()])
    # max-over-ranks reduction used by bench.py for timing
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    torch.save({'grads': grads, 'flat': flat, 'flat_overlap': flat_overlap, 'tmax': t}, os.path.join(out_dir, 'rank%d.pt' % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_ddp_gradients_equal_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / 'rank0.pt'), torch.load(tmp_path / 'rank1.pt')
    assert torch.equal(r0['grads'], r1['grads']), 'DDP must leave identical averaged gradients on every rank'
    assert float(r0['tmax']) == 2.0 and float(r1['tmax']) == 2.0
    from oracle import model as omodel
    from unopticalflow_b200 import train as T
    torch.manual_seed(0)
    model = omodel.Model_flow(omodel.Cfg)
    x = torch.rand(2, 3, 192, 64, generator=torch.Generator().manual_seed(1234))
    T.total_loss(model(x), T.generate_loss_weights_dict(T.KITTI_CFG)).backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()])
    err = float((r0['grads'] - ref).norm() / ref.norm())
    assert err < 1e-5, err
    # flat-buffer exchange == DDP == single process
    assert torch.equal(r0['flat'], r1['flat'])
    err = float((r0['flat'] - ref).norm() / ref.norm())
    assert err < 1e-5, err
    # ... == the two-chunk overlapped exchange
    assert torch.equal(r0['flat_overlap'], r1['flat_overlap'])
    err = float((r0['flat_overlap'] - r0['flat']).norm() / ref.norm())
    assert err < 1e-6, err


def test_loss_weights_and_total_loss():
    from unopticalflow_b200 import train as T
    w = T.generate_loss_weights_dict(T.KITTI_CFG)
    assert w == {'loss_pixel': pytest.approx(0.15), 'loss_ssim': 0.85, 'loss_flow_smooth': 10.0, 'loss_flow_consis': 0.01}
    pack = {k: torch.full((4,), float(i + 1)) for i, k in enumerate(w)}
    assert float(T.total_loss(pack, w)) == pytest.approx(0.15 * 1 + 0.85 * 2 + 10 * 3 + 0.01 * 4)
    assert T.generate_loss_weights_dict(T.SINTEL_CFG)['loss_flow_smooth'] == 6.0


def test_install_rebinds_reference_seams():
    """install() rebinds exactly the names the reference resolves at call time."""
    import types
    from unopticalflow_b200 import ops
    from unopticalflow_b200.install import install
    mods = {}
    for name in ('net_utils', 'pwc_tf', 'model_flow_paper', 'ssim'):
        mods[name] = types.ModuleType(name)
    mods['net_utils'].warp_flow = mods['pwc_tf'].warp_flow = mods['model_flow_paper'].warp_flow = object()
    mods['ssim'].SSIM = mods['model_flow_paper'].SSIM = object()

    class PWC_tf:
        def corr_naive(self, a, b, d=4):
            raise AssertionError('reference op chain must not run')

        def __init__(self):
            self.corr = self.corr_naive

    class Model_flow:
        pass
    mods['pwc_tf'].PWC_tf, mods['model_flow_paper'].Model_flow = PWC_tf, Model_flow
    done = install(mods)
    assert mods['pwc_tf'].warp_flow is ops.warp_flow and mods['model_flow_paper'].warp_flow is ops.warp_flow
    assert mods['net_utils'].warp_flow is ops.warp_flow and mods['model_flow_paper'].SSIM is ops.SSIM
    assert ('pwc_tf', 'PWC_tf.corr_naive') in done
    with pytest.raises(RuntimeError, match='no CPU fallback'):       # new instances route to the CUDA op
        PWC_tf().corr(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    assert Model_flow.compute_loss_flow_smooth is not None
    from unopticalflow_b200.install import uninstall
    assert uninstall() == len(done)
    assert mods['pwc_tf'].warp_flow is not ops.warp_flow and not hasattr(Model_flow, 'compute_loss_flow_smooth')
    with pytest.raises(AssertionError, match='reference op chain'):
        PWC_tf().corr(None, None)


@pytest.mark.skipif(not os.path.isdir('/root/reference/core/networks'), reason='reference checkout not present')
def test_install_on_real_reference():
    import sys
    sys.path.insert(0, '/root/reference')
    import core.networks  # noqa: F401
    from unopticalflow_b200 import ops
    from unopticalflow_b200.install import install
    done = install()
    for mod in ('net_utils', 'pwc_tf', 'model_flow_paper'):
        assert (mod, 'warp_flow') in done and sys.modules[mod].warp_flow is ops.warp_flow
    assert sys.modules['model_flow_paper'].SSIM is ops.SSIM
    from unopticalflow_b200.install import uninstall
    uninstall()
    assert sys.modules['net_utils'].warp_flow is not ops.warp_flow
