"""Depth Pro — B200-native drop-in for the reference's `depth_pro` package
(`/root/reference/src/depth_pro/__init__.py`): same entry points, CUDA sm_100a engine behind a C-ABI."""

from .depth_pro import (  # noqa: F401
    DEFAULT_MONODEPTH_CONFIG_DICT,
    DepthPro,
    DepthProConfig,
    create_model_and_transforms,
)
from .utils import load_rgb  # noqa: F401
