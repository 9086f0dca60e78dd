"""Build libuof_b200.so in-tree with nvcc for sm_100a (no torch, no cmake).

    python -m unopticalflow_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/uof_b200.h); it links only against the static
CUDA runtime so that it loads next to whatever CUDA runtime PyTorch brings.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB_DIR = os.path.join(PKG, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libuof_b200.so')
SOURCES = ['runtime.cu', 'cost_volume.cu', 'cost_volume_tma.cu', 'cost_volume_small.cu', 'cost_volume_tc.cu', 'warp.cu', 'photo_loss.cu', 'photo_warp.cu', 'ssim_map.cu', 'flow_losses.cu',
           'seams.cu', 'splat.cu', 'pyramid.cu', 'act.cu', 'upsample.cu', 'io_pipeline.cu']
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    cand = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(cand):
        raise RuntimeError('nvcc not found: libuof_b200.so cannot be built (there is no CPU fallback)')
    return cand


STAMP_PATH = LIB_PATH + '.stamp'


def source_digest():
    """sha256 over every file the library is built from (csrc/*, the public header, this recipe).  Stored next to the
    .so so that staleness is decided by content, not by mtimes (a snapshot copied to another box has fresh mtimes)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h')))
    deps += [os.path.join(PKG, '..', 'include', 'uof_b200.h'), os.path.abspath(__file__)]
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP_PATH):
        return True
    return open(STAMP_PATH).read().strip() != source_digest()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a (-lineinfo for ncu source pages) and link the shared library."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = nvcc_path()
    common = [nvcc, '-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC'] + ARCH_FLAGS
    if verbose:
        common += ['-Xptxas', '-v']
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace('.cu', '.o'))
        objs.append(obj)
        procs.append((src, subprocess.Popen(common + ['-c', os.path.join(CSRC, src), '-o', obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('---- %s ----\n%s\n' % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed; see output above')
    link = [nvcc, '-shared', '-o', LIB_PATH] + objs + ARCH_FLAGS + ['-cudart', 'static']
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    with open(STAMP_PATH, 'w') as f:
        f.write(source_digest() + '\n')
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
