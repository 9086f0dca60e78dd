"""Pipeline trace of the tcgen05 cost-volume backward: clock64 stamps of CTA (0,0,0) per window row (debug tool)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['UOF_CV_FORCE_TC'] = '1'
import torch
import unopticalflow_b200 as U
from unopticalflow_b200 import _lib
lib = _lib.load()
lib.uof_cv_tc_trace_buffer.restype = ctypes.c_void_p
shape = tuple(int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (16, 32, 64, 208)
B, C, H, W = shape
f1 = torch.randn(shape, device='cuda', requires_grad=True); f2 = torch.randn(shape, device='cuda', requires_grad=True)
ct = torch.randn(B, 81, H, W, device='cuda')
out = U.corr(f1, f2)
for _ in range(2):
    torch.autograd.grad(out, (f1, f2), ct, retain_graph=True)
torch.cuda.synchronize()
ptr = lib.uof_cv_tc_trace_buffer()
torch.autograd.grad(out, (f1, f2), ct, retain_graph=True)
torch.cuda.synchronize()
import numpy as np
buf = (ctypes.c_longlong * (25 * 8))()
assert lib.uof_cv_tc_trace_read(buf) == 0
host = torch.tensor(list(buf), dtype=torch.int64)
a = host.view(25, 8).numpy()
t0 = a[a > 0].min()
names = ['ld_issue', 'ld_pub', 'cv_stage', 'cv_built', 'cv_brow', 'cv_ready', 'mma_got', 'mma_done']
print('row ' + ' '.join('%9s' % n for n in names))
for c in range(24):
    print('%3d ' % c + ' '.join('%9d' % (a[c, e] - t0 if a[c, e] else -1) for e in range(8)))
print('prologue done %d, accumulators complete %d, epilogue done %d' % tuple(a[24, :3] - t0))
