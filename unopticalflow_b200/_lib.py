"""ctypes binding of libuof_b200.so (the C ABI declared in include/uof_b200.h).

There is NO CPU fallback: if the library is missing or a tensor is not on a CUDA device every
operator raises.  The library is built in-tree by `python -m unopticalflow_b200.build`
(`__graft_entry__.build()` calls it).
"""
from __future__ import annotations

import ctypes
import os

from .build import LIB_PATH

c_float_p = ctypes.c_void_p   # device pointers travel as integers
MAX_LEVELS = 4


SUMS_EXTRA = 1      # include/uof_b200.h UOF_SUMS_EXTRA: floats appended to every fused-loss `sums` workspace


class PhotoLevel(ctypes.Structure):
    _fields_ = [('img', ctypes.c_void_p), ('warped_l', ctypes.c_void_p), ('warped_r', ctypes.c_void_p),
                ('weight_l', ctypes.c_void_p), ('weight_r', ctypes.c_void_p),
                ('diff_l', ctypes.c_void_p), ('diff_r', ctypes.c_void_p),
                ('gwarped_l', ctypes.c_void_p), ('gwarped_r', ctypes.c_void_p),
                ('H', ctypes.c_int), ('W', ctypes.c_int)]


class PhotoWarpLevel(ctypes.Structure):
    _fields_ = [('img', ctypes.c_void_p), ('src_l', ctypes.c_void_p), ('src_r', ctypes.c_void_p),
                ('flow_l', ctypes.c_void_p), ('flow_r', ctypes.c_void_p),
                ('warped_l', ctypes.c_void_p), ('warped_r', ctypes.c_void_p),
                ('weight_l', ctypes.c_void_p), ('weight_r', ctypes.c_void_p),
                ('diff_l', ctypes.c_void_p), ('diff_r', ctypes.c_void_p),
                ('gflow_l', ctypes.c_void_p), ('gflow_r', ctypes.c_void_p),
                ('H', ctypes.c_int), ('W', ctypes.c_int)]


class SmoothLevel(ctypes.Structure):
    _fields_ = [('flow', ctypes.c_void_p), ('img', ctypes.c_void_p), ('gflow', ctypes.c_void_p),
                ('H', ctypes.c_int), ('W', ctypes.c_int)]


class ConsisLevel(ctypes.Structure):
    _fields_ = [('flow_fwd', ctypes.c_void_p), ('flow_bwd', ctypes.c_void_p), ('weight_fwd', ctypes.c_void_p),
                ('gflow_fwd', ctypes.c_void_p), ('H', ctypes.c_int), ('W', ctypes.c_int)]


_I, _LL, _P, _F = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_float

# name -> argtypes; every function returns int except the three diagnostics
SIGNATURES = {
    'uof_cost_volume_fwd': [_P, _P, _P, _I, _I, _I, _I, _LL, _P],
    'uof_cost_volume_bwd': [_P, _LL, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    'uof_cost_volume_fwd_ex': [_P, _LL, _P, _P, _I, _I, _I, _I, _LL, _P],
    'uof_cost_volume_bwd_ex': [_P, _LL, _P, _LL, _P, _P, _LL, _P, _P, _I, _I, _I, _I, _P],
    'uof_warp_fwd': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'uof_warp_bwd': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    'uof_photo_loss_fwd': [ctypes.POINTER(PhotoLevel), _I, _I, _P, _P, _P, _P],
    'uof_photo_loss_bwd': [ctypes.POINTER(PhotoLevel), _I, _I, _P, _P, _P, _P],
    'uof_photo_warp_loss_fwd': [ctypes.POINTER(PhotoWarpLevel), _I, _I, _I, _P, _P, _P, _P],
    'uof_photo_warp_loss_bwd': [ctypes.POINTER(PhotoWarpLevel), _I, _I, _I, _P, _P, _P, _P],
    'uof_diff_weight_fwd': [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'uof_diff_weight_bwd': [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'uof_masked_mean_fwd': [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_I),
                            ctypes.POINTER(_I), _I, _I, _I, _P, _P, _P],
    'uof_masked_mean_bwd': [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                            ctypes.POINTER(_I), ctypes.POINTER(_I), _I, _I, _I, _P, _P, _P],
    'uof_weighted_mean_sum_fwd': [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_F), ctypes.POINTER(_I), _I, _P, _P],
    'uof_weighted_mean_sum_bwd': [_P, ctypes.POINTER(_F), ctypes.POINTER(_I), _I, ctypes.POINTER(ctypes.c_void_p), _P],
    'uof_ssim_fwd': [_P, _P, _P, _I, _I, _I, _P],
    'uof_ssim_bwd': [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    'uof_smooth_loss_fwd': [ctypes.POINTER(SmoothLevel), _I, _I, _I, _P, _P, _P],
    'uof_smooth_loss_bwd': [ctypes.POINTER(SmoothLevel), _I, _I, _I, _P, _P],
    'uof_smooth_loss_bwd_acc': [ctypes.POINTER(SmoothLevel), _I, _I, _I, _P, _I, _P],
    'uof_consis_loss_bwd_acc': [ctypes.POINTER(ConsisLevel), _I, _I, _P, _P, _I, _P],
    'uof_consis_loss_fwd': [ctypes.POINTER(ConsisLevel), _I, _I, _P, _P, _P],
    'uof_consis_loss_bwd': [ctypes.POINTER(ConsisLevel), _I, _I, _P, _P, _P],
    'uof_bias_lrelu_fwd': [_P, _P, _I, _I, _I, _I, _F, _P],
    'uof_bias_lrelu_bwd': [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    'uof_bias_lrelu_bwd2': [_P, _LL, _P, _LL, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    'uof_bias_lrelu_fwd2': [_P, _P, _P, _LL, _P, _LL, _I, _I, _I, _I, _F, _P],
    'uof_bias_lrelu_bwd3': [_P, _LL, _P, _LL, _P, _LL, _P, _P, _I, _I, _I, _I, _F, _P],
    'uof_upsample_bilinear_fwd': [_P, _P, _I, _I, _I, _I, _I, _F, _P],
    'uof_upsample_bilinear_bwd': [_P, _P, _I, _I, _I, _I, _I, _F, _P],
    'uof_upsample_bilinear_fwd2': [_P, _P, _P, _I, _LL, _I, _I, _I, _I, _I, _F, _P],
    'uof_upsample_bilinear_bwd3': [_P, _P, _I, _LL, _P, _P, _I, _I, _I, _I, _I, _F, _P],
    'uof_img_pyramid': [_P, _LL, _LL, _LL, _LL, ctypes.POINTER(ctypes.c_void_p), _I, _I, _I, _I, _I, _I, _P],
    'uof_img_pyramid_stacked': [_P, _LL, _LL, _LL, _LL, _P, ctypes.POINTER(_I), ctypes.POINTER(ctypes.c_void_p), _I, _I, _I, _I, _I,
                                _I, _P],
    'uof_splat_fwd': [_P, _P, _P, _I, _I, _I, _I, _P],
    'uof_splat_bwd': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    'uof_splat_targets': [_P, _P, _I, _I, _I, _P],
    'uof_clamp01': [_P, _LL, _P],
    'uof_fb_consistency_mask': [_P, _P, _P, _I, _I, _I, _F, _F, _I, _P],
    'uof_preprocess_u8': [_P, _LL, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'uof_flow_png_decode': [_P, _P, _LL, _P],
    'uof_flow_png_encode': [_P, _I, _P, _LL, _P],
    'uof_flow_eval': [_P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P],
}
DIAGNOSTICS = {'uof_abi_version': (ctypes.c_int, []), 'uof_last_error': (ctypes.c_char_p, []),
               'uof_launch_count': (ctypes.c_longlong, [])}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load(path: str | None = None):
    """Load the shared library (once) and declare every entry point of include/uof_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get('UOF_B200_LIB', LIB_PATH)
    if not os.path.exists(path):
        raise LibraryMissing(
            '%s not found. Build it with `python -m unopticalflow_b200.build` (needs nvcc, targets sm_100a). '
            'unopticalflow_b200 has no CPU or PyTorch fallback.' % path)
    lib = ctypes.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI mismatch, fail loudly
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    for name, (restype, argtypes) in DIAGNOSTICS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.uof_abi_version() != 1:
        raise RuntimeError('libuof_b200.so ABI version %d, expected 1' % lib.uof_abi_version())
    _lib = lib
    return lib


# Optional observer used by bench.py to bracket every C-ABI call with CUDA events: an object with
# begin(name, args) -> token and end(token).  None in normal operation.
call_observer = None


def call(name: str, *args):
    """Invoke an entry point; non-zero status becomes a Python exception (SURVEY 8b conventions)."""
    lib = load()
    if call_observer is not None:
        token = call_observer.begin(name, args)
        rc = getattr(lib, name)(*args)
        call_observer.end(token)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.uof_last_error().decode('utf-8', 'replace')
        if rc == 1:
            raise ValueError('%s: %s' % (name, msg))
        raise RuntimeError('%s failed (status %d): %s' % (name, rc, msg))


def launch_count() -> int:
    return int(load().uof_launch_count())
