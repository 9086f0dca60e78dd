"""Training-step semantics of the reference driver, restated for the B200 path.

train.py:35-39 (model + Adam), :137-152 (zero_grad / forward / weighted loss / backward / step) and
core/config/config_utils.py:3-9 (loss-weight dict).  The reference's DataParallel (train.py:36-37) becomes
one process per GPU with DistributedDataParallel over NCCL.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops

KITTI_CFG = SimpleNamespace(mode='flow', dataset='kitti_depth', num_scales=3, h_flow_consist_alpha=3.0,
                            h_flow_consist_beta=0.05, w_ssim=0.85, w_flow_smooth=10.0, w_flow_consis=0.01,
                            img_hw=(256, 832), lr=1e-4, batch_size=8)          # config/kitti.yaml:11-38, train.py:168,170
SINTEL_CFG = SimpleNamespace(mode='flow', dataset='sintel', num_scales=3, h_flow_consist_alpha=3.0,
                             h_flow_consist_beta=0.05, w_ssim=0.85, w_flow_smooth=6.0, w_flow_consis=0.01,
                             img_hw=(384, 832), lr=1e-4, batch_size=8)         # config/sintel.yaml


def generate_loss_weights_dict(cfg):
    """core/config/config_utils.py:3-9."""
    return {'loss_pixel': 1 - cfg.w_ssim, 'loss_ssim': cfg.w_ssim, 'loss_flow_smooth': cfg.w_flow_smooth,
            'loss_flow_consis': cfg.w_flow_consis}


def total_loss(loss_pack, weights):
    """train.py:147-150: sum_k w_k * mean_B(loss_k)."""
    keys = list(loss_pack)
    if all(loss_pack[k].is_cuda and loss_pack[k].dtype == torch.float32 for k in keys):
        return ops.weighted_mean_sum([loss_pack[k] for k in keys], [float(weights[k]) for k in keys])     # one launch each way
    return torch.stack([weights[k] * loss_pack[k].mean() for k in keys]).sum()


def trainable_parameters(model):
    """The parameters train.py:39 hands to Adam.  `Model_flow` freezes everything in 'depth' / 'flowposenet' mode
    (model_flow_paper.py:19-24); an empty list is reported as such instead of an IndexError further down."""
    params = [p for p in model.parameters() if p.requires_grad]
    if not params:
        raise ValueError('the model has no trainable parameters (cfg.mode freezes Model_flow in depth / flowposenet mode): '
                         'nothing to optimise')
    return params


def make_optimizer(model, lr=1e-4, fused=True):
    """train.py:39 -- Adam with default betas/eps; `fused` uses the single-launch multi-tensor CUDA Adam."""
    params = trainable_parameters(model)
    return torch.optim.Adam([{'params': params, 'lr': lr}], fused=fused and params[0].is_cuda)


def _strip_prefix(state_dict):
    """Checkpoint keys without the wrappers the reference may have saved them under: `module.` (DataParallel / DDP,
    train.py:36-37) and `model_flow.` / `model_pose.model_flow.` (the upstream multi-task models, train.py:52-58)."""
    out = {}
    for k, v in state_dict.items():
        for pre in ('module.', 'model_pose.', 'model_flow.'):
            if k.startswith(pre):
                k = k[len(pre):]
        out[k] = v
    return out


def save_model(iter_, model_dir, filename, model, optimizer):
    """train.py:23-24: {'iteration', 'model_state_dict', 'optimizer_state_dict'} -- the reference's checkpoint layout, so its
    `iter_N.pth` / `last.pth` files and ours are interchangeable.  `optimizer` may be a torch optimizer or a
    `GraphedTrainStep` (both have `state_dict()`)."""
    import os
    inner = getattr(model, 'module', model)
    torch.save({'iteration': iter_, 'model_state_dict': inner.state_dict(), 'optimizer_state_dict': optimizer.state_dict()},
               os.path.join(model_dir, filename))


def load_model(model_dir, filename, model, optimizer=None, map_location=None):
    """train.py:26-31 (+ the prefix renaming of :50-60, in reverse): returns (iteration, model, optimizer)."""
    import os
    data = torch.load(os.path.join(model_dir, filename), map_location=map_location)
    inner = getattr(model, 'module', model)
    inner.load_state_dict(_strip_prefix(data['model_state_dict']))
    if optimizer is not None and data.get('optimizer_state_dict') is not None:
        optimizer.load_state_dict(data['optimizer_state_dict'])
    return data['iteration'], model, optimizer


class FlatGradAllReduce:
    """Data-parallel gradient exchange of SURVEY 8e: every parameter's `.grad` is a view into one flat fp32 buffer
    (5 134 324 floats = 20.5 MB for Model_flow), `zero()` clears it, backward accumulates into the views in place, and
    `allreduce()` averages it over the ranks with NCCL (NVLink 5 / NVSwitch).  Unlike DistributedDataParallel's bucketed
    hooks this is a plain stream-ordered sequence, so the whole iteration -- collective included -- can be captured into a
    CUDA graph (`GraphedTrainStep(allreduce=True)`).

    Overlap (`plan_overlap()` after one observed backward): the buffer is re-laid out in the order in which backward
    finishes the gradients and cut in two.  The EARLY chunk (the decoder's parameters: backward reaches them first) is
    all-reduced asynchronously as soon as its last gradient has been accumulated -- a post-accumulate hook counts them
    down -- so the collective and the wait for the slowest rank run behind the encoder's backward; only the small LATE
    chunk is exchanged after backward.  Round 1 measured the single un-overlapped all-reduce plus rank skew at 0.55 ms of
    a 50.8 ms step on 8 GPUs.  Captured into a CUDA graph the early all-reduce becomes a parallel branch of the graph."""

    def __init__(self, params, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('FlatGradAllReduce: no trainable parameters')
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, dtype=p0.dtype, device=p0.device)
        self._layout(self.params)
        self.n_early = 0                  # floats of the early chunk (0: no overlap, one all-reduce after backward)
        self._early_left = self._early_total = 0
        self._work = None
        self._hooks = []

    def _layout(self, ordered):
        off = 0
        for p in ordered:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def _multi(self):
        return self.dist.is_available() and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1

    def _reduce(self, t, async_op=False):
        cuda = t.is_cuda
        w = self.dist.all_reduce(t, op=self.dist.ReduceOp.AVG if cuda else self.dist.ReduceOp.SUM, group=self.group,
                                 async_op=async_op)
        if not cuda:                       # gloo has no AVG
            if async_op:
                w.wait()
                w = None
            t.div_(self.dist.get_world_size(self.group))
        return w

    # ---- overlap planning ------------------------------------------------------------------------------------------
    def observe(self):
        """Record, during the NEXT backward, the order (and, on CUDA, the device time) at which each gradient is ready."""
        self._order = []
        cuda = self.flat.is_cuda

        def make(i):
            def hook(_p):
                ev = None
                if cuda:
                    ev = torch.cuda.Event(enable_timing=True)
                    ev.record()
                self._order.append((i, ev))
            return hook
        self._obs = [p.register_post_accumulate_grad_hook(make(i)) for i, p in enumerate(self.params)]

    def plan_overlap(self, tail_ms=1.5, early_fraction=0.75):
        """Use the observation to split the buffer: the early chunk ends at the last gradient that is ready at least
        `tail_ms` of device time before backward ends (without timing -- CPU tensors -- after `early_fraction` of the
        bytes in completion order).  Returns the number of parameters in the early chunk."""
        for h in getattr(self, '_obs', []):
            h.remove()
        self._obs = []
        order = getattr(self, '_order', [])
        if len(order) != len(self.params):
            return 0                                       # some gradient never arrived: keep the single all-reduce
        idx = [i for i, _ in order]
        if order[0][1] is not None:
            torch.cuda.synchronize()
            end = order[-1][1]
            k = 0
            for j, (_, ev) in enumerate(order):
                if ev.elapsed_time(end) >= tail_ms:
                    k = j + 1
        else:
            total, acc, k = self.flat.numel(), 0, 0
            for j, i in enumerate(idx):
                acc += self.params[i].numel()
                if acc <= early_fraction * total:
                    k = j + 1
        if k == 0 or k == len(idx):
            return 0
        ordered = [self.params[i] for i in idx]
        self.flat.zero_()
        self._layout(ordered)
        self.n_early = sum(p.numel() for p in ordered[:k])
        self._early_total = k
        for h in self._hooks:
            h.remove()
        self._hooks = [p.register_post_accumulate_grad_hook(self._early_hook) for p in ordered[:k]]
        self._early_left = k
        return k

    def _early_hook(self, _p):
        self._early_left -= 1
        if self._early_left == 0 and self._multi():
            self._work = self._reduce(self.flat[:self.n_early], async_op=True)
            self._early_done = True

    def zero(self):
        self.flat.zero_()
        self._early_left = self._early_total
        self._work, self._early_done = None, False

    def allreduce(self):
        if not self._multi():
            return
        if self.n_early and getattr(self, '_early_done', False):
            self._reduce(self.flat[self.n_early:])         # the late chunk (queued behind the early one on NCCL's stream)
            if self._work is not None:
                self._work.wait()                          # join: the optimiser must see the averaged early chunk
                self._work = None
        else:
            self._reduce(self.flat)


class GraphedTrainStep:
    """The whole iteration (zero_grad, forward, weighted loss, backward, Adam) captured once into a CUDA graph and
    replayed per step.  Every C-ABI entry point is capturable (no allocation, no sync), the
    reference's blocking CPU mesh-grid copies are gone, and Adam runs as the capturable fused multi-tensor kernel.
    Inputs are copied into a static device buffer; the returned loss tensor is overwritten by the next replay.
    `allreduce=True` (one process per GPU, torch.distributed initialised with NCCL): gradients live in one flat buffer
    and are averaged over the ranks by a single captured NCCL all-reduce between backward and Adam
    (`FlatGradAllReduce`) -- the data-parallel step of SURVEY 8e without DDP's host-side hooks.
    With `overlap` the exchange is cut in two and the decoder's chunk is all-reduced behind the encoder's backward (see
    `FlatGradAllReduce.plan_overlap`).  `forward_module` lets the captured forward go through a wrapper of `model`.

    The eager warm-up iterations (cuDNN autotune, lazy initialisation) are REAL optimiser steps on `inputs_like`; the
    parameters and the Adam state are snapshotted before them and restored before capture, so constructing the step does
    not move the training trajectory, and `optimizer_state` (an Adam `state_dict()`, e.g. the `optimizer_state_dict` of a
    reference checkpoint) is loaded after the warm-up, i.e. it is what the first replay starts from."""

    def __init__(self, model, inputs_like, weights, lr=1e-4, warmup=3, forward_module=None, allreduce=False, group=None,
                 optimizer_state=None, overlap=False, overlap_tail_ms=5.0):
        params = trainable_parameters(model)
        self.model, self.weights = (forward_module if forward_module is not None else model), weights
        self.optimizer = torch.optim.Adam([{'params': params, 'lr': lr}], fused=True, capturable=True)
        self.exchange = FlatGradAllReduce(params, group) if allreduce else None
        self.static_in = torch.empty_like(inputs_like)
        self.static_in.copy_(inputs_like)
        saved_params = [p.detach().clone() for p in params]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # eager warm-up on a side stream (cuDNN autotune, lazy inits)
            plan_at = max(warmup - 2, 0)                   # not the first iteration: cuDNN autotunes inside its backward
            for it in range(warmup):
                if it == plan_at and self.exchange is not None and overlap:
                    self.exchange.observe()                # gradient completion order / times of one eager backward
                self._eager()
                if it == plan_at and self.exchange is not None and overlap:
                    self.overlap_params = self.exchange.plan_overlap(tail_ms=overlap_tail_ms)
            # undo the warm-up: parameters back to their values, Adam moments and step counters back to zero -- in place,
            # the graph is captured on these very tensors
            with torch.no_grad():
                for p, q in zip(params, saved_params):
                    p.copy_(q)
                for st in self.optimizer.state.values():
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        torch.cuda.current_stream().wait_stream(side)
        del saved_params
        if optimizer_state is not None:
            self.load_state_dict(optimizer_state)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager()

    def state_dict(self):
        """The Adam state (`optimizer_state_dict` of the reference's checkpoints, train.py:24)."""
        return self.optimizer.state_dict()

    def load_state_dict(self, state):
        """Load an Adam state IN PLACE: the captured graph reads the moment / step tensors it was captured on, so the
        values are copied into them (`step` may arrive as a Python number or a CPU tensor from a non-capturable Adam)."""
        if not self.optimizer.state:                       # no warm-up ran: let torch build the state, capturable layout
            self.optimizer.load_state_dict(state)
            return
        params = self.optimizer.param_groups[0]['params']
        packed = state['state']
        with torch.no_grad():
            for i, p in enumerate(params):
                src = packed.get(i, packed.get(str(i)))
                if src is None:
                    continue
                dst = self.optimizer.state[p]
                for k, v in src.items():
                    if torch.is_tensor(dst.get(k)):
                        dst[k].copy_(torch.as_tensor(v, dtype=dst[k].dtype))
                    else:
                        dst[k] = v

    def _eager(self):
        if self.exchange is not None:
            self.exchange.zero()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss = total_loss(self.model(self.static_in), self.weights)
        loss.backward()
        if self.exchange is not None:
            self.exchange.allreduce()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, inputs=None):
        """Run one iteration.  `inputs` (host-pinned or device tensor) is copied into the static input first; with
        `inputs=None` the batch staged by `stage()` is used."""
        if inputs is not None:
            self.static_in.copy_(inputs, non_blocking=True)
        else:
            assert self._staged is not None, 'call stage(batch) before step()'
            torch.cuda.current_stream().wait_event(self._staged)
            self.static_in.copy_(self._staging, non_blocking=True)        # device-to-device, ~20 us for 61 MB
            self._consumed = torch.cuda.Event()
            self._consumed.record()                                      # the staging buffer may be refilled from here on
            self._staged = None
        self.graph.replay()
        return self.static_loss

    _staged = None

    def stage(self, inputs):
        """Input pipeline: start the host-to-device copy of the NEXT batch on a side stream into a staging buffer, so
        that it overlaps the iteration that is running (the reference's DataLoader + `.cuda()` copies are blocking,
        train.py:139-141).  The following `step()` (no argument) consumes it."""
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream()
            self._staging = torch.empty_like(self.static_in)
            self._consumed = None
        # wait only for the device-to-device read of the staging buffer by the previous step(), NOT for the iteration
        # that is running: that is what lets the copy overlap it
        if self._consumed is not None:
            self._copy_stream.wait_event(self._consumed)
        with torch.cuda.stream(self._copy_stream):
            self._staging.copy_(inputs, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged = ev


def train_step(model, optimizer, inputs, weights):
    """One iteration of train.py:137-152.  Returns the scalar loss tensor (no host sync)."""
    optimizer.zero_grad(set_to_none=True)
    loss_pack = model(inputs)
    loss = total_loss(loss_pack, weights)
    loss.backward()
    optimizer.step()
    return loss.detach()
