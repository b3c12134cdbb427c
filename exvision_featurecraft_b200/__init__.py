"""exvision_featurecraft_b200 -- B200-native (sm_100a) feature-extraction hot path of FeatureCraft.

Host side: Python mirrors of the reference's function signatures; arithmetic: hand-written CUDA
kernels behind the C ABI in include/featurecraft_b200.h (libfeaturecraft_b200.so).  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
