#!/usr/bin/env python
"""bench.py — train frame-pairs/s of the Model_flow hot path at 256x832 on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one iteration of the reference's train.py:137-152 (zero_grad, Model_flow.forward on a batch of
synthetic KITTI-shaped triplets, weighted loss, backward, Adam step) = BASELINE.json configs[1]
("kitti.yaml flow-mode training step, synthetic 256x832 frame pairs, batch=8, 1xB200 fp32").  Each triplet
holds 2 frame pairs.  N > 1: one process per GPU, gradients averaged by one NCCL all-reduce inside the
CUDA graph (`--ddp`: eager DistributedDataParallel), per-GPU batch fixed (weak scaling).

Rank 0 prints ONE JSON line (see the driver contract in the task statement): value = whole-job frame-pairs/s
with inputs resident in HBM; e2e = the same metric with pinned-host inputs copied in and the loss read back
inside the timed region; roofline = the dominant hand-written hot-path kernel's achieved algorithmic GB/s measured
with CUDA events around its launches during K extra (instrumented) steps; cpu_baseline = the CPU oracle
(a port of the reference's PyTorch path) timed on this host's cores on a bounded sample.

`--impl reference` times the reference's own CPU implementation of the path (the oracle port: the reference
is pure Python and /root/reference does not exist on the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'train frame-pairs/s @256x832'
UNIT = 'frame-pairs/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=8, help='triplets per GPU (kitti.yaml / train.py:168 default 8)')
    ap.add_argument('--global-batch', type=int, default=0,
                    help='strong scaling (BASELINE.json configs[4]): total triplets per step, split evenly over the ranks '
                         '(overrides --batch; 64 -> 64/32/16/8 per rank at N=1/2/4/8)')
    ap.add_argument('--hw', type=int, nargs=2, default=[256, 832])
    ap.add_argument('--no-gpu-baseline', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-profile', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='do not capture the step in a CUDA graph')
    ap.add_argument('--overlap-tail-ms', type=float, default=0.0,
                    help='N > 1: all-reduce the gradients that are ready this many ms before backward ends from a hook, '
                         'overlapped with the rest of backward (0 = one all-reduce after backward)')
    ap.add_argument('--ddp', action='store_true',
                    help='N > 1: eager DistributedDataParallel step instead of the graphed flat all-reduce step')
    ap.add_argument('--profile-step', action='store_true',
                    help='warm up, then run ONE step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)')
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def ncu_traffic_bytes(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel, taken from the committed ncu capture
    (profiles/ncu_traffic.json, written by profiles/make_traffic_json.py from an `ncu` run of `bench.py --profile-step`);
    None when the capture has no such launch."""
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    try:
        table = json.load(open(p))
    except Exception:
        return None
    v = table.get(kernel_key)
    return int(v['dram_bytes']) if v else None


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.tmp = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(',') for r in open(self.tmp.name).read().strip().splitlines() if r.count(',') >= 8]
        os.unlink(self.tmp.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].strip().replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[5 + i].strip().lower() == 'active' for r in rows)]
        out.update(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=float(rows[0][2]), reasons=reasons,
                   samples=len(rows), power_w_max=max(float(r[3]) for r in rows))
        return out


# ------------------------------------------------------------------------- per-kernel roofline
def algorithmic_bytes(name, args):
    """SURVEY.md 8(d) per-unit figures x the units of this call (fp32).  args = the ctypes call arguments."""
    def lv_pixels(levels, n, B):
        return sum(B * levels[i].H * levels[i].W for i in range(n))
    if name == 'uof_cost_volume_fwd':
        B, C, H, W = args[3:7]
        return (2 * C + 81) * 4 * B * H * W
    if name == 'uof_cost_volume_bwd':
        B, C, H, W = args[6:10]
        return (4 * C + 81) * 4 * B * H * W
    if name == 'uof_cost_volume_fwd_ex':
        B, C, H, W = (int(v) for v in args[4:8])
        return (2 * C + 81) * 4 * B * H * W
    if name == 'uof_cost_volume_bwd_ex':        # + the concat-slice gradient read by the epilogue, when given
        B, C, H, W = (int(v) for v in args[9:13])
        return (4 * C + 81 + (C if args[5] else 0)) * 4 * B * H * W
    if name == 'uof_warp_fwd':
        B, C, H, W = args[3:7]
        return (2 * C + 2) * 4 * B * H * W
    if name == 'uof_warp_bwd':
        B, C, H, W = args[5:9]
        need_gx = bool(args[3].value)
        return ((3 * C + 4) if need_gx else (2 * C + 4)) * 4 * B * H * W
    if name == 'uof_photo_loss_fwd':
        return 36 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_photo_loss_bwd':
        return 60 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_photo_warp_loss_fwd':       # image warps fused in: img 12 + two sources 24 + two flows 16 + two weight maps 8
        return 60 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_photo_warp_loss_bwd':       # img 12 + sources 24 + flows 16 + weights 8 + two flow gradients 16
        return 76 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_smooth_loss_fwd':
        return 20 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_smooth_loss_bwd':
        return 28 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_consis_loss_fwd':
        return 20 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_consis_loss_bwd':
        return 28 * lv_pixels(args[0], args[1], args[2])
    if name == 'uof_bias_lrelu_fwd':            # read + write the activation in place
        B, C, H, W = args[2:6]
        return 8 * B * C * H * W
    if name == 'uof_bias_lrelu_fwd2':           # read the convolution output, write one or two destinations
        B, C, H, W = (int(v) for v in args[6:10])
        return (3 if args[4] else 2) * 4 * B * C * H * W
    if name == 'uof_bias_lrelu_bwd3':           # read g1 (+ g2) and y, write gx
        B, C, H, W = (int(v) for v in args[8:12])
        return (4 if args[2] else 3) * 4 * B * C * H * W
    if name == 'uof_bias_lrelu_bwd2':           # read g1 (+ g2) and y, write gx
        B, C, H, W = (int(v) for v in args[7:11])
        return (4 if args[2] else 3) * 4 * B * C * H * W
    if name == 'uof_bias_lrelu_bwd':            # read gout and y, write gx
        B, C, H, W = args[4:8]
        return 12 * B * C * H * W
    if name in ('uof_upsample_bilinear_fwd', 'uof_upsample_bilinear_bwd'):      # read one side, write the other
        planes, h, w, H, W = (int(v) for v in args[2:7])
        return 4 * planes * (h * w + H * W)
    if name == 'uof_upsample_bilinear_fwd2':    # + the second destination
        planes, h, w, H, W = (int(v) for v in args[5:10])
        return 4 * planes * (h * w + (2 if args[2] else 1) * H * W)
    if name == 'uof_upsample_bilinear_bwd3':    # + the extra gradient sources
        planes, h, w, H, W = (int(v) for v in args[6:11])
        return 4 * planes * (h * w + (1 + bool(args[1]) + bool(args[4])) * H * W)
    if name == 'uof_img_pyramid_stacked':      # read the triplet once, write level 0 and the coarser levels
        nimg, B, C, H, W = (int(v) for v in args[9:14])
        return int(4 * nimg * B * C * H * W * (2 + 0.25 + 0.0625))
    if name == 'uof_img_pyramid':
        nimg, B, C, H, W = args[7:12]
        return int(nimg * B * C * H * W * 4 * (1 + 1 / 4 + 1 / 16))
    return 0


class KernelObserver:
    """Brackets every libuof_b200 call with CUDA events on the launching stream."""

    def __init__(self, torch):
        self.torch = torch
        self.records = []

    @staticmethod
    def key_of(name, args):
        key = name
        if name == 'uof_cost_volume_fwd_ex':        # same kernels as the plain entry points: same key
            key = 'uof_cost_volume_fwd[%s]' % 'x'.join(str(int(d)) for d in args[4:8])
        elif name == 'uof_cost_volume_bwd_ex':
            key = 'uof_cost_volume_bwd[%s]' % 'x'.join(str(int(d)) for d in args[9:13])
        elif name == 'uof_upsample_bilinear_fwd2':
            key = 'uof_upsample_bilinear_fwd[%s]' % 'x'.join(str(int(d)) for d in args[5:10])
        elif name == 'uof_upsample_bilinear_bwd3':
            key = 'uof_upsample_bilinear_bwd[%s]' % 'x'.join(str(int(d)) for d in args[6:11])
        elif name.startswith('uof_cost_volume') or name.startswith('uof_warp'):
            dims = args[3:7] if name.endswith('fwd') else (args[6:10] if 'cost' in name else args[5:9])
            key = '%s[%s]' % (name, 'x'.join(str(int(d)) for d in dims))
            if name == 'uof_warp_bwd':
                key += '+gx' if args[3].value else ''
        elif name.startswith('uof_upsample_bilinear'):
            key = '%s[%s]' % (name, 'x'.join(str(int(d)) for d in args[2:7]))
        elif name == 'uof_bias_lrelu_fwd2':
            key = 'uof_bias_lrelu_fwd[%s]%s' % ('x'.join(str(int(d)) for d in args[6:10]), '+d2' if args[4] else '')
        elif name == 'uof_bias_lrelu_bwd3':
            key = 'uof_bias_lrelu_bwd[%s]%s' % ('x'.join(str(int(d)) for d in args[8:12]), '+g2' if args[2] else '')
        elif name == 'uof_bias_lrelu_bwd2':      # same key as the one-gradient form, '+g2' when two gradients are summed
            key = 'uof_bias_lrelu_bwd[%s]%s' % ('x'.join(str(int(d)) for d in args[7:11]), '+g2' if args[2] else '')
        elif name.startswith('uof_bias_lrelu'):
            dims = args[2:6] if name.endswith('fwd') else args[4:8]
            key = '%s[%s]' % (name, 'x'.join(str(int(d)) for d in dims))
        return key

    def begin(self, name, args):
        s = self.torch.cuda.current_stream()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        key = self.key_of(name, args)
        e0.record(s)
        return (key, algorithmic_bytes(name, args), e0, e1, s)

    def end(self, token):
        key, nbytes, e0, e1, s = token
        e1.record(s)
        self.records.append((key, nbytes, e0, e1))

    def summary(self, peak_gbs):
        self.torch.cuda.synchronize()
        agg = {}
        for key, nbytes, e0, e1 in self.records:
            a = agg.setdefault(key, {'calls': 0, 'ms': 0.0, 'bytes': 0})
            a['calls'] += 1
            a['ms'] += e0.elapsed_time(e1)
            a['bytes'] += nbytes
        out = []
        for key, a in agg.items():
            us = 1e3 * a['ms'] / a['calls']
            gbs = (a['bytes'] / a['calls']) / (us * 1e-6) / 1e9 if us > 0 else 0.0
            out.append({'kernel': key, 'calls': a['calls'], 'avg_us': round(us, 2), 'total_ms': round(a['ms'], 3),
                        'alg_mb': round(a['bytes'] / a['calls'] / 1e6, 3), 'achieved_gbs': round(gbs, 1),
                        'frac': round(gbs / peak_gbs, 4)})
        out.sort(key=lambda r: -r['total_ms'])
        return out


def exchange_note(graphed):
    ex = graphed.exchange
    if ex is not None and ex.n_early:
        return ('flat gradient buffer in the CUDA graph, two NCCL all-reduces: %.1f MB (%d tensors, the decoder) launched from a '
                'gradient hook and overlapped with the rest of backward, %.1f MB after backward'
                % (ex.n_early * 4 / 1e6, ex._early_total, (ex.flat.numel() - ex.n_early) * 4 / 1e6))
    return 'one flat NCCL all-reduce of the gradients inside the CUDA graph'


# ------------------------------------------------------------------------------ reference arms
NBUF = 4          # distinct input batches rotated through every timed loop


def workload_config(B, world, H, W, scaling):
    """`config` of the JSON line: the workload only, identical for the b200 and the reference arm of one launch."""
    return {'workload': 'BASELINE.json configs[1]: kitti.yaml flow-mode training step (Model_flow fwd+bwd+Adam, train.py:137-152), '
                        'synthetic %dx%d triplets, batch %d per GPU, fp32 (TF32 off)' % (H, W, B),
            'img_hw': [H, W], 'batch_per_gpu': B, 'global_batch': B * world, 'frame_pairs_per_triplet': 2,
            'parallelism': 'dp%d (one process per GPU, batch sharded by sample, %s scaling)' % (world, scaling) if world > 1 else 'single',
            'tf32': False,
            'l2': '%d distinct input batches (%.0f MB each, %.0f MB total vs 126 MB L2) rotated; the activations a step writes '
                  'and re-reads (GBs) flush L2 between iterations' % (NBUF, B * 9 * H * W * 4 / 1e6, NBUF * B * 9 * H * W * 4 / 1e6)}


def reference_model(device):
    """(model, kind, step): the reference's Model_flow + train.py:35-39,137-152 restated around it (train.py itself does not
    import in this image).  kind 'reference' = the UNMODIFIED reference installed in baseline/_ref by
    baseline/install_ref.py (its own core.networks.get_model / core.config.generate_loss_weights_dict, stock ATen code
    path, none of this repo's kernels); kind 'port' = the oracle restatement, only when baseline/_ref is absent."""
    import torch
    from unopticalflow_b200 import train as T
    cfg = T.KITTI_CFG
    try:
        from baseline import install_ref
        have_ref = install_ref.available()
    except Exception:
        have_ref = False
    torch.manual_seed(0)
    if have_ref:
        net = install_ref.load()
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                    # the reference prints a banner from Model_flow.__init__; stdout carries ONE JSON line
        try:
            model = net.get_model(cfg.mode)(cfg)
            sys.stdout.flush()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        import core.config as ref_config                      # the reference's own loss-weight dict (config_utils.py:3-9)
        weights = ref_config.generate_loss_weights_dict(cfg)
        kind = 'reference'
    else:
        from oracle import model as omodel
        model = omodel.Model_flow(cfg)
        weights = T.generate_loss_weights_dict(cfg)
        kind = 'port'
    model = model.to(device)
    model.train()
    optimizer = torch.optim.Adam([{'params': filter(lambda p: p.requires_grad, model.parameters()), 'lr': cfg.lr}])   # train.py:39

    def step(inputs):                    # train.py:137-152
        optimizer.zero_grad()
        loss_pack = model(inputs)
        loss_list = []
        for key in list(loss_pack.keys()):
            loss_list.append((weights[key] * loss_pack[key].mean()).unsqueeze(0))
        loss = torch.cat(loss_list, 0).sum()
        loss.backward()
        optimizer.step()
        return loss.detach()
    return model, kind, step


def cpu_reference_steps(steps, warmup, H, W, batch, budget_s=240.0):
    """The reference's own CPU implementation of the training step on this host's cores, all threads.  Each step is a batch
    of `batch` triplets unless the projected run would exceed `budget_s`, in which case the sample shrinks (and says so)."""
    import warnings
    import torch
    warnings.filterwarnings('ignore')
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model, kind, step = reference_model(torch.device('cpu'))
    gen = torch.Generator().manual_seed(1234)
    sample = batch
    t0 = time.perf_counter()
    step(torch.rand(1, 3, 3 * H, W, generator=gen))              # probe (also the first warm-up): one triplet
    t1 = time.perf_counter() - t0
    while sample > 1 and 0.8 * t1 * sample * (steps + warmup) > budget_s:      # B=8 runs ~20 % faster per sample than B=1
        sample //= 2
    xs = [torch.rand(sample, 3, 3 * H, W, generator=gen) for _ in range(NBUF)]
    for i in range(max(warmup - 1, 1)):
        step(xs[i % NBUF])
    t0 = time.perf_counter()
    for i in range(steps):
        step(xs[i % NBUF])
    dt = time.perf_counter() - t0
    return dt / steps, torch.get_num_threads(), kind, sample


def describe_sample(kind, sample, batch, cores, s_per_step):
    what = ('the UNMODIFIED reference (baseline/_ref: core.networks.get_model(\'flow\') + train.py:137-152 restated)'
            if kind == 'reference' else 'the CPU oracle port of the reference PyTorch path (baseline/_ref absent)')
    return ('each step = %d triplet%s (%s batch of %d) through %s: zero_grad + fwd + bwd + Adam on %d torch threads, %.2f s/step'
            % (sample, 's' if sample > 1 else '', 'the whole' if sample == batch else 'of the', batch, what, cores, s_per_step))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    H, W = args.hw
    B = per_rank_batch(args, max(args.gpus, 1))
    s_per_step, cores, kind, sample = cpu_reference_steps(args.steps, args.warmup, H, W, B)
    value = 2.0 * sample / s_per_step
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': round(value, 4), 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(1e3 * s_per_step, 3),
        'higher_is_better': True, 'scaling': scaling_of(args), 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(B, max(args.gpus, 1), H, W, scaling_of(args)),
        'cpu_baseline': {'value': round(value, 4), 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': describe_sample(kind, sample, B, cores, s_per_step)},
        'e2e': {'value': round(value, 4), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'one CPU process on rank 0 only (the reference has no multi-process CPU path); at N > 1 compare per-GPU numbers, '
                'not this line, with the N-GPU aggregate',
    }
    print(json.dumps(line), flush=True)


def gpu_reference_baseline(torch, dev, B, H, W, host, steps=3, warmup=2):
    """The number to beat: the reference's stock PyTorch/ATen/cuDNN path -- unmodified when baseline/_ref is
    installed -- running the same training step on the SAME B200 inside this run: fp32, TF32 off, cudnn.benchmark on (the
    reference leaves it off; on is the faster setting), eager like train.py, device-resident inputs, CUDA events."""
    model, kind, step = reference_model(dev)
    xs = [h.to(dev) for h in host[:2]]
    for i in range(warmup):
        step(xs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(xs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, step, xs
    torch.cuda.empty_cache()
    return {'value': round(2 * B / (ms * 1e-3), 3), 'unit': UNIT, 'ms_per_step': round(ms, 3), 'kind': kind, 'steps': steps,
            'warmup': warmup, 'what': ('%s Model_flow fwd+bwd+Adam (train.py:137-152), eager, batch %d of %dx%d triplets on this GPU, fp32 TF32 off'
                                       % ('unmodified reference (baseline/_ref)' if kind == 'reference' else 'oracle port of the reference', B, H, W))}


def scaling_of(args):
    return 'strong' if args.global_batch else 'weak'


def per_rank_batch(args, world):
    if args.global_batch:
        assert args.global_batch % world == 0, '--global-batch must be a multiple of the number of ranks'
        return args.global_batch // world
    return args.batch


# ----------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback for the product path)'
    torch.backends.cudnn.allow_tf32 = False          # FP32 parity with the reference
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # NCCL prints its version banner (NCCL_DEBUG=VERSION/WARN/INFO) on stdout when the communicator is created, and
        # stdout carries the ONE JSON line: create the communicator with file descriptor 1 pointed at stderr.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    import unopticalflow_b200 as u
    from unopticalflow_b200 import _lib, train as T
    _lib.load()
    cfg = T.KITTI_CFG
    H, W = args.hw
    B = per_rank_batch(args, world)
    weights = T.generate_loss_weights_dict(cfg)

    torch.manual_seed(0)
    model = u.Model_flow(cfg).to(dev)
    net = model
    if world > 1 and (args.ddp or args.no_graph):
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=8, broadcast_buffers=False)
    opt = T.make_optimizer(model, cfg.lr)

    # synthetic KITTI-shaped triplets; NBUF distinct resident batches (> L2 in total) are rotated so that no
    # timed iteration re-reads inputs left in L2 by the previous one (the step's activations are GBs anyway)
    gen = torch.Generator(device='cpu').manual_seed(1234 + rank)
    host = [torch.rand(B, 3, 3 * H, W, generator=gen).pin_memory() for _ in range(NBUF)]
    resident = [h.to(dev) for h in host]
    in_bytes = host[0].numel() * 4

    def eager_step(x):
        return T.train_step(net, opt, x, weights)

    # The whole iteration is captured into a CUDA graph (unopticalflow_b200.train.GraphedTrainStep, the package's
    # public training-step API) and replayed.  N > 1: gradients live in one flat buffer and are averaged by a single
    # NCCL all-reduce captured in the same graph (train.FlatGradAllReduce); `--ddp` (or `--no-graph`) runs the
    # iteration eagerly under DistributedDataParallel instead.
    use_graph = not args.no_graph and not (world > 1 and args.ddp)
    launches_per_graph_step = 0
    if use_graph:
        n_before = _lib.launch_count()
        graphed = T.GraphedTrainStep(model, resident[0], weights, cfg.lr, warmup=3, allreduce=world > 1,
                                     overlap=args.overlap_tail_ms > 0, overlap_tail_ms=args.overlap_tail_ms)
        launches_per_graph_step = (_lib.launch_count() - n_before) // 4      # 3 eager warm-ups + 1 capture pass
        step = graphed
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for i in range(max(args.warmup, 3)):
        step(resident[i % NBUF])

    if args.profile_step:
        class CallRecorder:          # ordered list of C-ABI calls of the profiled step, to align with ncu's launch list
            def __init__(self):
                self.keys = []

            def begin(self, name, a):
                self.keys.append(KernelObserver.key_of(name, a))
                return None

            def end(self, token):
                pass
        rec = CallRecorder()
        torch.cuda.synchronize()
        _lib.call_observer = rec
        for _ in range(3):
            eager_step(resident[1])
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        eager_step(resident[0])      # eager so that every C-ABI call is seen by the recorder
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        _lib.call_observer = None
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        json.dump(rec.keys, open(os.path.join(ROOT, 'gpurun_out', 'profile_step_calls.json'), 'w'))
        return

    # ---- device-resident timing ("value") -----------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = _lib.launch_count()
    ms_step = timed(lambda i: step(resident[i % NBUF]), args.steps)
    launches = (_lib.launch_count() - n0) * world
    if use_graph:      # replayed kernels do not pass through the C ABI again: count what one captured step holds
        launches = launches_per_graph_step * args.steps * world
    clocks = sampler.stop() if sampler else {}

    # ---- end to end: pinned host inputs -> H2D every step, loss read back every step ---------------
    last = {}

    if use_graph:
        # Input pipeline of the public API: batch i+1 is staged (pinned host -> device, side stream) while iteration i
        # runs; every timed step still performs one H2D copy of a full batch and one D2H read of the loss.
        def e2e_step(i):
            loss = step()                              # consumes the staged batch i
            step.stage(host[(i + 1) % NBUF])           # H2D of batch i+1 overlaps iteration i
            last['loss'] = float(loss)                 # .item(): D2H + sync, like the reference's logging path
        step.stage(host[0])
        e2e_step(0)
    else:
        def e2e_step(i):
            x = host[i % NBUF].to(dev, non_blocking=True)
            last['loss'] = float(step(x))
        e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)

    # ---- the reference's own PyTorch path on this same GPU (rank 0 of a 1-GPU run) ------------------------
    gpu_baseline = None
    if world == 1 and not args.no_gpu_baseline:
        gpu_baseline = gpu_reference_baseline(torch, dev, B, H, W, host)

    # ---- instrumented pass: CUDA events around every hand-written kernel launch ---------------------
    peak, peak_src = measured_peak_gbs()
    kernels = []
    if not args.no_kernel_profile:
        obs = KernelObserver(torch)
        _lib.call_observer = obs
        for i in range(args.steps):
            eager_step(resident[i % NBUF])      # events cannot be recorded inside a graph replay: this pass runs eagerly
        _lib.call_observer = None
        kernels = obs.summary(peak)

    # ---- isolated per-kernel timing (CUDA-graph replay over rotating buffer sets > L2), rank 0 of a 1-GPU run ----
    isolated = []
    if not args.no_kernel_profile and world == 1:
        del resident
        torch.cuda.empty_cache()
        from unopticalflow_b200 import kernel_bench
        isolated = kernel_bench.run(peak, B=B, H=H, W=W)

    if world > 1:
        dist.barrier()
    fp_per_step = 2 * B * world
    if rank == 0:
        dominant = kernels[0] if kernels else None
        own_ms = sum(k['total_ms'] for k in kernels) / max(args.steps, 1)
        line = {
            'metric': METRIC, 'value': round(fp_per_step / (ms_step * 1e-3), 3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': round(ms_step, 3),
            'higher_is_better': True, 'scaling': scaling_of(args), 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(B, world, H, W, scaling_of(args)),
            'triplets_per_s': round(B * world / (ms_step * 1e-3), 3),
            'execution': {'cuda_graph': bool(use_graph),
                          'gradient_exchange': ((exchange_note(graphed) if use_graph else
                                                 'DistributedDataParallel, NCCL') if world > 1 else None)},
            'clocks': clocks,
            'e2e': {'value': round(fp_per_step / (ms_e2e * 1e-3), 3), 'unit': UNIT, 'ms_per_step': round(ms_e2e, 3),
                    'h2d_bytes_per_step': in_bytes * world, 'd2h_bytes_per_step': 4 * world,
                    'api': ('unopticalflow_b200.train.GraphedTrainStep(Model_flow, pinned-host batch%s)'
                            % (', allreduce=True' if world > 1 else '') if use_graph else
                            'unopticalflow_b200.train.train_step(DDP(Model_flow), Adam, pinned-host batch)')},
            'gpu_launches': int(launches),
            'own_kernels_ms_per_step': round(own_ms, 3),
        }
        def roofline_of(k, how):
            r = {'kernel': k['kernel'], 'bound': 'hbm', 'achieved': k['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
                 'frac': k['frac'], 'traffic': ncu_traffic_bytes(k['kernel']), 'peak_source': peak_src,
                 'alg_bytes_per_launch': int(k['alg_mb'] * 1e6), 'avg_us': k['avg_us'], 'calls': k['calls'],
                 'total_ms': k['total_ms'], 'how': how}
            import re
            m = re.match(r'uof_cost_volume_(fwd|bwd)\[(\d+)x(\d+)x(\d+)x(\d+)\]', k['kernel'])
            if m:      # the contraction sits at the FP32 ridge: also give it against the measured CUDA-core FMA peak
                b_, c_, h_, w_ = (int(v) for v in m.groups()[1:])
                flops = (2 if m.group(1) == 'fwd' else 4) * 81 * c_ * b_ * h_ * w_
                r['fp32_tflops'] = round(flops / (k['avg_us'] * 1e-6) / 1e12, 2)
                r['fp32_peak_tflops'] = 69.0
                r['fp32_frac'] = round(r['fp32_tflops'] / 69.0, 3)
                r['note'] = ('arithmetic intensity %.1f FLOP/B is at the FP32 ridge (69 TFLOP/s / 6.5 TB/s = 10.6): the kernel is bound '
                             'by shared-memory bandwidth and FMA issue, not HBM (DESIGN.md 4.1)' % (flops / (k['alg_mb'] * 1e6)))
            return r
        if dominant:
            # `roofline`: the dominant kernel of the HOT PATH (cost volume, warps, losses, pyramid).
            # The fused bias+LeakyReLU / up-sampling glue around the kept convolutions (8f) streams at 93-96 % of the HBM
            # peak and, summed over its ~90 launches per step, takes more time than any single hot-path kernel; it is
            # reported separately as `roofline_glue` so that the headline fraction is the one of the kernel the north star
            # is about.
            glue = ('uof_bias_lrelu', 'uof_upsample')
            hot = [k for k in kernels if not k['kernel'].startswith(glue)]
            how = ('CUDA events around each launch of this entry point during %d instrumented steps (stream backlogged, so no '
                   'host latency inside the bracket); traffic = ncu dram bytes of the same launch, profiles/ncu_traffic.json'
                   % args.steps)
            if hot:
                line['roofline'] = roofline_of(hot[0], 'dominant = largest total device time among the hand-written HOT-PATH '
                                               'kernels (SURVEY 8a); ' + how)
            gl = [k for k in kernels if k['kernel'].startswith(glue)]
            if gl:
                line['roofline_glue'] = roofline_of(gl[0], 'largest total device time among the fused glue kernels around the '
                                                    'kept convolutions (SURVEY 8f); ' + how)
            if not hot:
                line['roofline'] = line['roofline_glue']
            line['kernels'] = kernels[:24]
        if isolated:
            line['kernels_isolated'] = isolated
        if gpu_baseline:
            line['gpu_baseline'] = gpu_baseline
            line['gpu_baseline']['speedup'] = round(line['value'] / gpu_baseline['value'], 3)
        if world == 1 and not args.no_cpu_baseline:
            s_per_step, cores, kind, sample = cpu_reference_steps(3, 2, H, W, B, budget_s=30.0)
            line['cpu_baseline'] = {'value': round(2.0 * sample / s_per_step, 4), 'unit': UNIT, 'cores': cores, 'kind': kind,
                                    'sample': '3 timed steps (2 warm-up); ' + describe_sample(kind, sample, B, cores, s_per_step)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown: a CUDA graph that holds captured NCCL kernels must be released before the communicator, and a
        # communicator teardown that still blocks (seen once at N=2: the JSON line was out, the process never exited)
        # must not hang the launcher -- a daemon timer ends the process after a grace period.
        import threading
        sys.stdout.flush()
        killer = threading.Timer(30.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        torch.cuda.synchronize()
        if use_graph:
            step.graph.reset()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        killer.cancel()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
