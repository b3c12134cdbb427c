"""Locate / run csrc/build.py without importing torch (used by tests and __graft_entry__.build)."""
from __future__ import annotations

import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def _module():
    spec = importlib.util.spec_from_file_location("fc_csrc_build", os.path.join(_HERE, "csrc", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ensure_built(force: bool = False) -> str:
    """Compile libfeaturecraft_b200.so for sm_100a if it is missing or stale.  On a box without nvcc
    (never the case in this image) a prebuilt library is used as is."""
    mod = _module()
    if not force and os.path.exists(mod.LIB) and not os.path.exists(mod.NVCC):
        return mod.LIB
    return mod.build(force=force)
