import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['UOF_CV_FORCE_TC'] = '1'
import torch
import unopticalflow_b200 as U
shape = tuple(int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (1, 32, 16, 8)
B, C, H, W = shape
f1 = torch.randn(shape, device='cuda', requires_grad=True); f2 = torch.randn(shape, device='cuda', requires_grad=True)
ct = torch.randn(B, 81, H, W, device='cuda')
out = U.corr(f1, f2)
torch.cuda.synchronize(); print('fwd ok', flush=True)
g1, g2 = torch.autograd.grad((out * ct).sum(), (f1, f2))
torch.cuda.synchronize(); print('bwd ok', float(g1.abs().max()), float(g2.abs().max()), flush=True)
