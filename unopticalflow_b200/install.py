"""Drop the CUDA operators into a loaded checkout of the reference.

The reference has no plugin registry: its operator seams are module-global names resolved at call time plus
one instance attribute.  Because every reference package does `sys.path.append(dirname(__file__))` and bare
imports (core/networks/__init__.py:1-3, structures/__init__.py:1-6, model_flow_paper.py:1-4), the live module
objects are registered under TOP-LEVEL names (`net_utils`, `pwc_tf`, `model_flow_paper`, `ssim`).

    import core.networks                       # the reference
    from unopticalflow_b200.install import install
    install()                                  # rebinds warp_flow / SSIM / corr / loss methods
    model = core.networks.get_model('flow')(cfg).cuda()
"""
from __future__ import annotations

import sys

from . import ops

_SEAMS = (('net_utils', 'warp_flow'), ('pwc_tf', 'warp_flow'), ('model_flow_paper', 'warp_flow'),
          ('structures', 'warp_flow'), ('ssim', 'SSIM'), ('pytorch_ssim', 'SSIM'), ('model_flow_paper', 'SSIM'))


def _corr_method(self, input1, input2, d=4):
    if d != 4:
        raise ValueError('the CUDA cost volume is built for d=4')
    return ops.corr(input1, input2)


def _smooth(self, optical_flows, img_pyramid):
    return ops.flow_smooth_loss(optical_flows, img_pyramid, self.num_scales)


def _consis(self, fwd_flow_pyramid, bwd_flow_pyramid, occ_mask_list):
    return ops.flow_consis_loss(fwd_flow_pyramid, bwd_flow_pyramid, occ_mask_list, self.num_scales)


def _diff_weight(self, img_pyramid_from_l, img_pyramid, img_pyramid_from_r):
    return ops.diff_weight(img_pyramid_from_l, img_pyramid, img_pyramid_from_r, self.num_scales)


def _loss_with_mask(self, diff_list, occ_mask_list):
    return ops.loss_with_mask(diff_list, occ_mask_list, self.num_scales)


def _pyramid(self, img, num_pyramid):
    return ops.img_pyramid(img, num_pyramid)


def _warp_pyramid(self, img_pyramid, flow_pyramid):
    return [ops.warp_flow(i, f, use_mask=True) for i, f in zip(img_pyramid, flow_pyramid)]


_METHODS = (('compute_loss_flow_smooth', _smooth), ('compute_loss_flow_consis', _consis), ('generate_img_pyramid', _pyramid),
            ('warp_flow_pyramid', _warp_pyramid), ('compute_diff_weight', _diff_weight), ('compute_loss_with_mask', _loss_with_mask))
_saved = []          # (owner object, attribute name, original value) of everything install() rebound, for uninstall()


def _rebind(owner, name, value):
    _saved.append((owner, name, owner.__dict__.get(name) if isinstance(owner, type) else getattr(owner, name)))
    setattr(owner, name, value)


def install(modules=None, models=()):
    """Rebind every seam found in `modules` (default: sys.modules) and patch already-built `models`.

    Returns the list of (module, name) pairs that were rebound so callers can verify the drop-in took;
    `uninstall()` puts the reference's own callables back."""
    modules = sys.modules if modules is None else modules
    done = []
    for mod, name in _SEAMS:
        m = modules.get(mod)
        if m is not None and hasattr(m, name):
            _rebind(m, name, ops.warp_flow if name == 'warp_flow' else ops.SSIM)
            done.append((mod, name))
    pwc = modules.get('pwc_tf')
    if pwc is not None and hasattr(pwc, 'PWC_tf'):
        _rebind(pwc.PWC_tf, 'corr_naive', _corr_method)   # picked up by `self.corr = self.corr_naive` (pwc_tf.py:19)
        done.append(('pwc_tf', 'PWC_tf.corr_naive'))
    mfp = modules.get('model_flow_paper')
    if mfp is not None and hasattr(mfp, 'Model_flow'):
        for n, fn in _METHODS:
            _rebind(mfp.Model_flow, n, fn)
            done.append(('model_flow_paper', 'Model_flow.' + n))
    for model in models:                                 # instances built before install()
        inner = getattr(model, 'module', model)
        if hasattr(inner, 'pwc_model'):
            _rebind(inner.pwc_model, 'corr', ops.corr)
            done.append((type(inner).__name__, 'pwc_model.corr'))
    return done


def uninstall():
    """Undo every rebinding made by install() (most recent first).  Returns how many were restored."""
    n = len(_saved)
    while _saved:
        owner, name, orig = _saved.pop()
        if orig is None and isinstance(owner, type):
            delattr(owner, name)
        else:
            setattr(owner, name, orig)
    return n
