import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REL_TOL = 1e-4          # north star: loss values and gradients within 1e-4 relative in FP32
MASK_AGREE = 0.999      # north star: >= 99.9 % pixel agreement on thresholded masks


def rel_err(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def assert_close(a, b, tol=REL_TOL, what=''):
    e = rel_err(a, b)
    assert e <= tol, '%s: max rel err %.3e > %.1e' % (what, e, tol)
    return e


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: torch.from_numpy(z[k]) if z[k].dtype.kind == 'f' else z[k] for k in z.files}


def flows_like(g, B, H, W, sigma, oob=True):
    f = torch.randn(B, 2, H, W, generator=g) * sigma
    if oob and W > 3 and H > 2:
        f[:, 0, :, W - 2] += W
        f[:, 1, 1, :] -= H
    return f
