"""GPU check + timing of the cost-volume backward paths (tcgen05 vs CUDA-core TMA kernel) at the pyramid-level shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ops as O
import unopticalflow_b200 as U

def run(shape, mode):
    for k in ('UOF_CV_NO_TC', 'UOF_CV_FORCE_TC'):
        os.environ.pop(k, None)
    if mode != 'cuda':
        os.environ['UOF_CV_FORCE_TC'] = '1'
    B, C, H, W = shape
    g = torch.Generator(device='cuda').manual_seed(sum(shape))
    f1 = torch.randn(shape, device='cuda', generator=g, requires_grad=True)
    f2 = torch.randn(shape, device='cuda', generator=g, requires_grad=True)
    ct = torch.randn(B, 81, H, W, device='cuda', generator=g)
    out = U.corr(f1, f2)
    g1, g2 = torch.autograd.grad((out * ct).sum(), (f1, f2))
    return f1, f2, ct, g1, g2

def timeit(shape, mode, reps=20):
    B, C, H, W = shape
    nset = max(2, int(300e6 // ((4 * C + 81) * 4 * B * H * W)) + 1)
    sets = []
    for i in range(nset):
        f1 = torch.randn(shape, device='cuda', requires_grad=True); f2 = torch.randn(shape, device='cuda', requires_grad=True)
        ct = torch.randn(B, 81, H, W, device='cuda')
        sets.append((f1, f2, ct, U.corr(f1, f2)))
    for k in ('UOF_CV_NO_TC', 'UOF_CV_FORCE_TC'):
        os.environ.pop(k, None)
    if mode != 'cuda':
        os.environ['UOF_CV_FORCE_TC'] = '1'
    def go(i):
        f1, f2, ct, out = sets[i % nset]
        torch.autograd.grad(out, (f1, f2), ct, retain_graph=True)
    for i in range(3): go(i)
    torch.cuda.synchronize()
    # time only the kernel: use events around the C call via the observer
    from unopticalflow_b200 import _lib
    class Obs:
        def __init__(s): s.ev = []
        def begin(s, name, a):
            if name != 'uof_cost_volume_bwd': return None
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(); return (e0, e1)
        def end(s, t):
            if t: t[1].record(); s.ev.append(t)
    o = Obs(); _lib.call_observer = o
    for i in range(reps): go(i)
    torch.cuda.synchronize(); _lib.call_observer = None
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in o.ev)
    return ts[len(ts) // 2]

shapes = [(16, 32, 64, 208), (16, 64, 32, 104), (16, 96, 16, 52), (2, 32, 64, 208), (3, 64, 40, 72), (2, 128, 16, 24), (1, 32, 17, 12)]
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
for shape in shapes:
    B, C, H, W = shape
    f1, f2, ct, a1, a2 = run(shape, 'cuda')
    _, _, _, b1, b2 = run(shape, 'tc')
    r1, r2 = torch.autograd.grad((O.cost_volume(f1, f2) * ct).sum(), (f1, f2))
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print('%-20s cuda-core vs oracle %.2e %.2e | tcgen05 vs oracle %.2e %.2e' % (shape, rel(a1, r1), rel(a2, r2), rel(b1, r1), rel(b2, r2)), flush=True)
if which != 'check':
    for shape in shapes[:3]:
        B, C, H, W = shape
        mb = (4 * C + 81) * 4 * B * H * W / 1e6
        tc, tt = timeit(shape, 'cuda'), timeit(shape, 'tc')
        print('%-20s %.1f MB  cuda-core %.1f us (%.0f GB/s)  tcgen05 %.1f us (%.0f GB/s)  speed-up %.2fx' % (shape, mb, tc, mb / tc * 1e3, tt, mb / tt * 1e3, tc / tt), flush=True)
