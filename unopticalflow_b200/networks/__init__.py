"""Same public surface as the reference's `core/networks` (core/networks/__init__.py:1-9)."""
from .model_flow import Model_flow
from .structures import FeaturePyramid, PWC_tf, conv, warp_flow  # noqa: F401
from ..ops import SSIM  # noqa: F401


def get_model(mode):
    if mode == 'flow':
        return Model_flow
    raise ValueError('Mode {} not found.'.format(mode))
