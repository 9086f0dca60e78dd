"""Not a test: times the reference's PyTorch op chain (the oracle restatement, which already avoids the reference's
CPU mesh-grid + H2D copies) on the SAME GPU as the CUDA path, for DESIGN.md's "B2" baseline.

    python tests/perf_reference_gpu.py [--batch 8] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as omodel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--channels-last', action='store_true')
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    m = omodel.Model_flow(omodel.Cfg).to(dev)
    if a.channels_last:
        m = m.to(memory_format=torch.channels_last)
    opt = omodel.make_optimizer(m)
    xs = [torch.rand(a.batch, 3, 768, 832, device=dev) for _ in range(4)]
    for i in range(3):
        omodel.train_step(m, opt, xs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        omodel.train_step(m, opt, xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({'what': 'reference PyTorch op chain (oracle restatement) on this GPU, fp32, TF32 off',
                      'batch': a.batch, 'channels_last': a.channels_last, 'ms_per_step': round(ms, 2),
                      'frame_pairs_per_s': round(2 * a.batch / (ms * 1e-3), 2)}))


if __name__ == '__main__':
    main()
