"""Thin torch plumbing: device buffers, raw pointers and the current stream for the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.FeatureCraftError(
            "exvision_featurecraft_b200 needs a CUDA device (sm_100a); there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "C ABI needs contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(a, dtype=None) -> torch.Tensor:
    """numpy array / tensor -> contiguous CUDA tensor (optionally cast)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        t = a.to(dev, non_blocking=True)
    else:
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def scan_scratch(n: int) -> torch.Tensor:
    elems = int(_lib.load().fc_scan_scratch_elems(int(n)))
    return torch.empty(max(elems, 1), dtype=torch.int32, device=require_cuda())
