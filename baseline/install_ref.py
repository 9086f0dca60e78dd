"""Install the UNMODIFIED reference's hot-path packages into baseline/_ref/ (git-ignored, ships to the GPU box).

    python baseline/install_ref.py [--src /root/reference] [--check]

The reference has no setup.py / pyproject, so `pip install --target baseline/_ref /root/reference`
has nothing to build: this recipe is the install.  It copies, byte for byte, the two packages the training
step of BASELINE.json needs -- `core/networks` (Model_flow, PWC_tf, warp_flow, SSIM; model_flow_paper.py:14-255)
and `core/config` (config_utils.generate_loss_weights_dict) -- plus `config/*.yaml`, and writes a manifest of
sha256 digests so that "unmodified" can be checked (`--check`, also done by `load()` below).  The other
reference packages (dataset, evaluation, visualize, train.py, test.py) do not import in this image
(png, imageio, h5py, skimage are missing) and are not on the path.

baseline/_ref/ is listed in .gitignore (reference sources never enter this repository's history) and not in
.gpurunignore, so the copy travels with the working tree.  Users: `bench.py --impl reference` (CPU arm),
`bench.py`'s `gpu_baseline` leg (same code on the same B200) and tests/test_gpu_reference_dropin.py.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
MANIFEST = os.path.join(DEST, 'MANIFEST.json')
PARTS = ('core/networks', 'core/config', 'config')


def _digest(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def _files(root):
    out = []
    for part in PARTS:
        for d, _, names in os.walk(os.path.join(root, part)):
            if '__pycache__' in d:
                continue
            for n in names:
                if n.endswith(('.py', '.yaml')):
                    out.append(os.path.relpath(os.path.join(d, n), root))
    return sorted(out)


def install(src='/root/reference'):
    if not os.path.isdir(os.path.join(src, 'core', 'networks')):
        raise FileNotFoundError('no reference checkout at %s' % src)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for rel in _files(src):
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        manifest[rel] = _digest(dst)
    with open(os.path.join(DEST, 'core', '__init__.py'), 'a'):
        pass                                     # the reference's `core` is a namespace package; make it explicit
    json.dump({'source': src, 'files': manifest}, open(MANIFEST, 'w'), indent=1, sort_keys=True)
    return manifest


def available():
    return os.path.exists(MANIFEST)


def check():
    """True when every installed file still has the digest recorded at install time."""
    files = json.load(open(MANIFEST))['files']
    return all(os.path.exists(os.path.join(DEST, r)) and _digest(os.path.join(DEST, r)) == d for r, d in files.items())


def load(cpu_shim=True):
    """Import the installed reference: returns its `core.networks` module (`get_model`, `Model_flow`).

    `cpu_shim`: net_utils.py:48 does `.to(x.get_device())`, which is -1 for CPU tensors (the reference's CPU
    path crashes as shipped); the shim -- harness side, the reference files stay untouched -- makes `get_device()` return
    the device object for CPU tensors.  The reference registers its modules under top-level names (`net_utils`, `pwc_tf`,
    `model_flow_paper`, `ssim`, ...); they are reachable through sys.modules afterwards."""
    if not available():
        raise FileNotFoundError('baseline/_ref is not installed: run `python baseline/install_ref.py` where /root/reference exists')
    if not check():
        raise RuntimeError('baseline/_ref differs from its manifest: not the unmodified reference')
    import torch
    if cpu_shim and not getattr(torch.Tensor.get_device, '_uof_shim', False):
        orig = torch.Tensor.get_device

        def get_device(self):
            return orig(self) if self.is_cuda else self.device
        get_device._uof_shim = True
        torch.Tensor.get_device = get_device
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import core.networks as ref_networks
    return ref_networks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default=os.environ.get('UOF_REFERENCE', '/root/reference'))
    ap.add_argument('--check', action='store_true')
    a = ap.parse_args()
    if a.check:
        ok = available() and check()
        print('baseline/_ref: %s' % ('unmodified (%d files)' % len(json.load(open(MANIFEST))['files']) if ok else 'MISSING or MODIFIED'))
        sys.exit(0 if ok else 1)
    m = install(a.src)
    print('installed %d reference files into %s' % (len(m), DEST))


if __name__ == '__main__':
    main()
