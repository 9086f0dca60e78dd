"""GPU diagnostic (not a test): where does the step-level gradient difference vs the same-GPU oracle chain come from?
Per loss term, per batch size, and against the oracle's OWN sensitivity to a different cuDNN algorithm choice."""
import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import unopticalflow_b200 as U
from oracle import model as omodel, ops as O

W = {'loss_pixel': 0.15, 'loss_ssim': 0.85, 'loss_flow_smooth': 10.0, 'loss_flow_consis': 0.01}

def grads(model, x, keys):
    model.zero_grad(set_to_none=True)
    pack = model(x)
    sum(W[k] * pack[k].mean() for k in keys).backward()
    return torch.cat([p.grad.flatten() for p in model.parameters()]).clone()

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    H, Wd = 256, 832
    torch.manual_seed(0)
    ref = omodel.Model_flow(omodel.Cfg)
    m = U.Model_flow(omodel.Cfg); m.load_state_dict(ref.state_dict()); m = m.cuda()
    ref = ref.cuda()
    x = torch.rand(B, 3, 3 * H, Wd, generator=torch.Generator().manual_seed(1234)).cuda()
    allk = list(W)
    torch.backends.cudnn.benchmark = False
    g_ref = grads(ref, x, allk)
    n = float(g_ref.norm())
    for keys in [allk] + [[k] for k in allk]:
        gr = grads(ref, x, keys); gm = grads(m, x, keys)
        print('B=%d %-28s |g_ref| %.3e  err/|g_total| %.3e  err/|g_term| %.3e' % (B, '+'.join(k[5:] for k in keys), float(gr.norm()),
              float((gm - gr).norm()) / n, float((gm - gr).norm()) / float(gr.norm())), flush=True)
    # the oracle against itself under another cuDNN algorithm choice
    torch.backends.cudnn.benchmark = True
    for _ in range(2):
        g_ref2 = grads(ref, x, allk)
    print('oracle cudnn.benchmark on vs off: %.3e' % (float((g_ref2 - g_ref).norm()) / n))
    for keys in [[k] for k in allk]:
        torch.backends.cudnn.benchmark = False
        a = grads(ref, x, keys)
        torch.backends.cudnn.benchmark = True
        b = grads(ref, x, keys)
        print('   %-20s oracle self-diff / |g_total| %.3e' % (keys[0], float((a - b).norm()) / n))
    # per-sample independence: oracle B=8 vs product B=8 on flows
    with torch.no_grad():
        _, ff_m, fb_m = m(x, output_flow=True)
        c = x[:, :, H:2 * H]; r = x[:, :, 2 * H:]
        ff_r = ref.pwc_model(ref.fpyramid(c), ref.fpyramid(r), [H, Wd])
    for s in range(3):
        print('flow level %d: max abs diff %.3e (max |flow| %.3f)' % (s, float((ff_m[s] - ff_r[s]).abs().max()), float(ff_r[s].abs().max())))

main()
