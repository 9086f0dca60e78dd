"""unopticalflow_b200 — B200-native (sm_100a) training hot path of UnOpticalFlow's `Model_flow`.

Public surface (mirrors the reference's `core/networks`):

    from unopticalflow_b200 import get_model, Model_flow, PWC_tf, FeaturePyramid, warp_flow, SSIM, corr
    from unopticalflow_b200 import ops            # fused losses, splat, masks
    from unopticalflow_b200.install import install   # rebind the seams of a loaded reference checkout

All compute goes through libuof_b200.so (C ABI in include/uof_b200.h); there is no CPU fallback.
"""
from . import ops  # noqa: F401
from .networks import FeaturePyramid, Model_flow, PWC_tf, get_model  # noqa: F401
from .ops import SSIM, corr, warp_flow  # noqa: F401

__all__ = ['ops', 'get_model', 'Model_flow', 'PWC_tf', 'FeaturePyramid', 'warp_flow', 'SSIM', 'corr']
