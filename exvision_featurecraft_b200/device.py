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


def base_ptr(t):
    """Device pointer of the first element of a (possibly strided) CUDA tensor view."""
    assert t.is_cuda
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


_PINNED_SLOTS: dict = {}


def read_last_ints(tensors) -> list:
    """[int(t[-1]) for t in tensors] (int32 CUDA tensors) with ONE host synchronisation: the last elements are
    published to pinned host memory by kernels (see read_int) and a single event is waited for."""
    if not tensors:
        return []
    key = (tensors[0].device.index, torch.cuda.current_stream().cuda_stream)
    slot = _PINNED_SLOTS.get(key)
    if slot is None or slot[0].numel() < len(tensors):
        slot = (torch.zeros(max(16, len(tensors)), dtype=torch.int32).pin_memory(), torch.cuda.Event())
        _PINNED_SLOTS[key] = slot
    host, ev = slot
    for i, t in enumerate(tensors):
        src = t.reshape(-1)[-1:]
        _lib.call("fc_publish_i32", C.c_void_p(src.data_ptr()), C.c_void_p(host.data_ptr() + 4 * i), 1, stream_ptr())
    ev.record()
    ev.synchronize()
    return [int(v) for v in host[: len(tensors)].tolist()]


def read_ints_at(t: torch.Tensor, indices, *more, between=None) -> list:
    """[int(t[i]) for i in indices] (int32 CUDA tensor) with one kernel launch per 32 indices and ONE host
    synchronisation: the values are gathered into pinned host memory by a kernel (see read_int).  Further
    (tensor, indices) pairs may follow; their values are appended, still under the single synchronisation.
    between: called after the event is recorded and before the host waits for it -- work queued there (kernels that do
    not need the values) keeps the device busy across the host round trip."""
    jobs = [(t, list(indices))] + [(a, list(b)) for a, b in zip(more[0::2], more[1::2])]
    n = sum(len(ix) for _, ix in jobs)
    assert n > 0
    for a, _ in jobs:
        assert a.is_cuda and a.dtype == torch.int32 and a.is_contiguous()
    key = (t.device.index, torch.cuda.current_stream().cuda_stream)
    slot = _PINNED_SLOTS.get(key)
    if slot is None or slot[0].numel() < max(32, n):
        slot = (torch.zeros(max(32, n), dtype=torch.int32).pin_memory(), torch.cuda.Event())
        _PINNED_SLOTS[key] = slot
    host, ev = slot
    at = 0
    for a, ix in jobs:
        for lo in range(0, len(ix), 32):
            part = [int(i) for i in ix[lo:lo + 32]]
            idx = (C.c_int64 * len(part))(*part)
            _lib.call("fc_publish_gather_i32", C.c_void_p(a.data_ptr()), idx, len(part), C.c_void_p(host.data_ptr() + 4 * at),
                      stream_ptr())
            at += len(part)
    ev.record()
    if between is not None:
        between()
    ev.synchronize()
    return [int(v) for v in host[:n].tolist()]


def read_int(t: torch.Tensor, index: int = -1) -> int:
    """int(t[index]) for an int32 CUDA tensor WITHOUT a D2H memcpy: a kernel stores the value into pinned host
    memory and the host waits on an event.  A 4-byte cudaMemcpy would queue behind any bulk device-to-host
    transfer in flight on the copy engine and serialise the pipelined drivers (BatchExtractor etc.)."""
    assert t.is_cuda and t.dtype == torch.int32
    key = (t.device.index, torch.cuda.current_stream().cuda_stream)
    slot = _PINNED_SLOTS.get(key)
    if slot is None:
        slot = (torch.zeros(16, dtype=torch.int32).pin_memory(), torch.cuda.Event())
        _PINNED_SLOTS[key] = slot
    host, ev = slot
    src = t.reshape(-1)[index:] if index != -1 else t.reshape(-1)[-1:]
    _lib.call("fc_publish_i32", C.c_void_p(src.data_ptr()), C.c_void_p(host.data_ptr()), 1, stream_ptr())
    ev.record()
    ev.synchronize()
    return int(host[0])
