"""Batched, device-resident corner detectors (Harris / lambda-minus) over the C ABI.

These are the drivers the reference-shaped functions in ``backend.py`` sit on, and what the
benchmarks call: inputs and outputs are CUDA tensors, nothing synchronises except the one read of
the corner count that sizes the output list.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib
from .device import ptr, read_int, require_cuda, scan_scratch, stream_ptr, to_device


@dataclass
class CornerResult:
    """response: [B,H,W] float64; mask: [B,H,ceil(W/32)] int32 bit mask of response > threshold;
    coords: [N,2] int32 (row, col) in np.argwhere order, images concatenated; values: [N] float64;
    image_offsets: [B+1] int64 -- corners of image b are coords[image_offsets[b]:image_offsets[b+1]];
    image_max: [B] float64 (lambda-minus only)."""

    response: torch.Tensor
    mask: torch.Tensor | None = None
    coords: torch.Tensor | None = None
    values: torch.Tensor | None = None
    image_offsets: torch.Tensor | None = None
    image_max: torch.Tensor | None = None


def _as_gray_batch(gray) -> torch.Tensor:
    g = to_device(gray)
    if g.dtype != torch.uint8:
        raise TypeError("gray images must be uint8 (convert_to_gray output)")
    if g.dim() == 2:
        g = g.unsqueeze(0)
    if g.dim() != 3:
        raise ValueError("expected [H,W] or [B,H,W] gray images")
    return g.contiguous()


def rgb_to_gray(rgb) -> torch.Tensor:
    """convert_to_gray on the device: [B,H,W,3] or [H,W,3] uint8 -> uint8 gray."""
    t = to_device(rgb)
    single = t.dim() == 3
    if single:
        t = t.unsqueeze(0)
    if t.dim() != 4 or t.shape[-1] < 3 or t.dtype != torch.uint8:
        raise ValueError("expected uint8 [..., H, W, >=3]")
    t = t[..., :3].contiguous()
    b, h, w, _ = t.shape
    out = torch.empty((b, h, w), dtype=torch.uint8, device=t.device)
    _lib.call("fc_rgb_to_gray_u8", ptr(t), b, h, w, ptr(out), stream_ptr())
    return out[0] if single else out


def compact_mask(mask: torch.Tensor, width: int, response: torch.Tensor | None = None):
    """np.argwhere order compaction of a [B,H,pitch] bit mask -> (coords, values, image_offsets)."""
    b, h, _ = mask.shape
    rows = b * h
    row_offsets = torch.empty(rows + 1, dtype=torch.int32, device=mask.device)
    scratch = scan_scratch(rows)
    _lib.call("fc_mask_scan", ptr(mask), b, h, width, ptr(row_offsets), ptr(scratch), stream_ptr())
    image_offsets = row_offsets[::h].to(torch.int64)  # B+1 entries, still on the device
    total = read_int(row_offsets)  # the one host sync: sizes the output
    coords = torch.empty((total, 2), dtype=torch.int32, device=mask.device)
    values = torch.empty(total, dtype=torch.float64, device=mask.device) if response is not None else None
    if total:
        _lib.call("fc_mask_emit", ptr(mask), ptr(row_offsets), b, h, width, ptr(response), ptr(coords), ptr(values),
                  stream_ptr())
    return coords, values, image_offsets


def _new_mask(b, h, w, device):
    return torch.empty((b, h, (w + 31) // 32), dtype=torch.int32, device=device)


def harris(gray, window_size: int = 5, k: float = 0.04, threshold: float | None = 10000, compact: bool = True,
           out: torch.Tensor | None = None) -> CornerResult:
    """Harris response (reference arithmetic, bit-exact) + optional threshold / compaction."""
    if window_size % 2 == 0:
        raise ValueError("cannot reshape array: window_size must be odd (reference fails the same way)")
    g = _as_gray_batch(gray)
    b, h, w = g.shape
    if h < 2 or w < 2:
        raise ValueError(
            "Shape of array too small to calculate a numerical gradient, at least (edge_order + 1) elements are required."
        )
    resp = out if out is not None else torch.empty((b, h, w), dtype=torch.float64, device=g.device)
    mask = _new_mask(b, h, w, g.device) if threshold is not None else None
    _lib.call("fc_harris_response", ptr(g), b, h, w, int(window_size), float(k),
              float(threshold if threshold is not None else 0.0), ptr(resp), ptr(mask), stream_ptr())
    res = CornerResult(resp, mask)
    if mask is not None and compact:
        res.coords, res.values, res.image_offsets = compact_mask(mask, w, resp)
    return res


def notebook_harris(gray, window_size: int, k: float, threshold: float, compact: bool = True) -> CornerResult:
    """harris.ipynb variant: K_X gradient, corners only for offset <= y < H-offset."""
    g = _as_gray_batch(gray)
    b, h, w = g.shape
    resp = torch.empty((b, h, w), dtype=torch.float64, device=g.device)
    _lib.call("fc_notebook_harris_response", ptr(g), b, h, w, int(window_size), float(k), ptr(resp), stream_ptr())
    mask = _new_mask(b, h, w, g.device)
    _lib.call("fc_threshold_mask_abs", ptr(resp), b, h, w, float(threshold), int(window_size / 2), ptr(mask), stream_ptr())
    res = CornerResult(resp, mask)
    if compact:
        res.coords, res.values, res.image_offsets = compact_mask(mask, w, resp)
    return res


def lambda_minus(gray, threshold_percentage: float | None = 0.01, compact: bool = True,
                 out: torch.Tensor | None = None, fast: bool = False) -> CornerResult:
    """lambda-minus (min eigenvalue) map, per-image max, and map > pct * max.
    fast=True (opt-in): det / larger eigenvalue from exact 64-bit integer determinants instead of LAPACK's operation
    sequence -- a few float64 ulps from the default map, not bit-identical (fc_lambda_minus_response_fast)."""
    g = _as_gray_batch(gray)
    b, h, w = g.shape
    eig = out if out is not None else torch.empty((b, h, w), dtype=torch.float64, device=g.device)
    image_max = torch.empty(b, dtype=torch.float64, device=g.device)
    _lib.call("fc_lambda_minus_response_fast" if fast else "fc_lambda_minus_response", ptr(g), b, h, w, ptr(eig), ptr(image_max),
              stream_ptr())
    res = CornerResult(eig, image_max=image_max)
    if threshold_percentage is not None:
        res.mask = _new_mask(b, h, w, g.device)
        _lib.call("fc_threshold_mask_rel", ptr(eig), b, h, w, float(threshold_percentage), ptr(image_max), ptr(res.mask),
                  stream_ptr())
        if compact:
            res.coords, res.values, res.image_offsets = compact_mask(res.mask, w, eig)
    return res


_GRADIENTS = {"central": _lib.FC_GRAD_CENTRAL, "kx": _lib.FC_GRAD_KX, "sobel": _lib.FC_GRAD_SOBEL}
_WINDOWS = {"box": _lib.FC_WINDOW_BOX, "gaussian": _lib.FC_WINDOW_GAUSS}
_RESPONSES = {"harris": _lib.FC_RESP_HARRIS, "lambda_minus": _lib.FC_RESP_LAMBDA_MINUS}


def gaussian_window_taps(n: int, sigma: float):
    """cv2.getGaussianKernel(n, sigma) for sigma > 0: exp(-(i - (n-1)/2)^2 / (2 sigma^2)), normalised to sum 1."""
    import numpy as np

    if not sigma > 0:
        raise ValueError("the Gaussian window needs sigma > 0")
    x = np.arange(n, dtype=np.float64) - (n - 1) * 0.5
    w = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return w / w.sum()


def corner_response(gray, gradient: str = "central", window: str = "box", window_size: int = 5, sigma: float | None = None,
                    response: str = "harris", k: float = 0.04, threshold: float | None = None, nms: bool = False,
                    compact: bool = True) -> CornerResult:
    """The corner response with its OPT-IN components (BASELINE north_star; the reference uses none of them, so the
    defaults are the app's Harris): gradient "central" (np.gradient) | "kx" (5x5 K_X) | "sobel" (cv2.Sobel ksize 3),
    window "box" (zero padded ones window) | "gaussian" (cv2.GaussianBlur with `sigma`), response "harris" |
    "lambda_minus", and with a threshold optionally 3x3 non-maximum suppression."""
    import ctypes as C

    g = _as_gray_batch(gray)
    b, h, w = g.shape
    if window_size % 2 == 0:
        raise ValueError("window_size must be odd")
    taps = None
    if window == "gaussian":
        t = gaussian_window_taps(window_size, sigma)
        taps = (C.c_double * window_size)(*t.tolist())
    resp = torch.empty((b, h, w), dtype=torch.float64, device=g.device)
    mask = _new_mask(b, h, w, g.device) if threshold is not None else None
    _lib.call("fc_corner_response_ex", ptr(g), b, h, w, _GRADIENTS[gradient], _WINDOWS[window], int(window_size), taps,
              _RESPONSES[response], float(k), 1 if (nms and mask is not None) else 0,
              float(threshold if threshold is not None else 0.0), ptr(resp), ptr(mask), stream_ptr())
    res = CornerResult(resp, mask)
    if mask is not None and compact:
        res.coords, res.values, res.image_offsets = compact_mask(mask, w, resp)
    return res


def nms_threshold(response: torch.Tensor, threshold: float):
    """Opt-in 3x3 non-maximum suppression of a resident response map (the reference only thresholds): corners where
    R > threshold and R >= its neighbours -> (coords, values, image_offsets) like rethreshold()."""
    r = response if response.dim() == 3 else response.unsqueeze(0)
    b, h, w = r.shape
    mask = _new_mask(b, h, w, r.device)
    _lib.call("fc_nms3x3_mask", ptr(r), b, h, w, float(threshold), ptr(mask), stream_ptr())
    return compact_mask(mask, w, r)


def rethreshold(response: torch.Tensor, threshold: float, border: int = 0):
    """on_changing_threshold (APP/FeatureCraft_Backend.py:261-292): re-threshold a resident map with
    an ABSOLUTE value (the reference compares eigenvalues to slider/10000, not to a percentage)."""
    r = response if response.dim() == 3 else response.unsqueeze(0)
    b, h, w = r.shape
    mask = _new_mask(b, h, w, r.device)
    _lib.call("fc_threshold_mask_abs", ptr(r), b, h, w, float(threshold), int(border), ptr(mask), stream_ptr())
    return compact_mask(mask, w, r)


class BatchCornerDetector:
    """Harris + lambda-minus for a stream of equally shaped uint8 image batches that live on the HOST.

    Same three-stream structure as sift.BatchExtractor: pinned input -> device, kernels on the caller's stream,
    corner lists -> pinned output, the D2H of batch i overlapping the kernels of batch i+1.  The response maps stay
    on the device (the reference caches them for its threshold slider, APP/FeatureCraft_Backend.py:355, :437);
    what travels back is what the reference returns: Harris (row, col, R) and lambda-minus (row, col) lists."""

    def __init__(self, batch: int, height: int, width: int, max_corners: int, window_size: int = 5, k: float = 0.04,
                 threshold: float = 10000, threshold_percentage: float = 0.01, depth: int = 2):
        dev = require_cuda()
        self.batch, self.h, self.w, self.depth = batch, height, width, depth
        self.args = (window_size, k, threshold, threshold_percentage)
        self.cap = int(max_corners)
        self.copy_in = torch.cuda.Stream(device=dev)
        self.copy_out = torch.cuda.Stream(device=dev)
        self.host_in = [torch.empty((batch, height, width), dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.dev_in = [torch.empty((batch, height, width), dtype=torch.uint8, device=dev) for _ in range(depth)]
        self.resp = torch.empty((batch, height, width), dtype=torch.float64, device=dev)
        self.eig = torch.empty((batch, height, width), dtype=torch.float64, device=dev)
        self.h_coords = [torch.empty((self.cap, 2), dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.h_values = [torch.empty(self.cap, dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.h_lcoords = [torch.empty((self.cap, 2), dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.h_offs = [torch.empty((2, batch + 1), dtype=torch.int64).pin_memory() for _ in range(depth)]
        self.slots = [None] * depth
        self.staged = [None] * depth
        self.consumed = [None] * depth
        self.n_submitted = 0

    def stage(self, images, ticket: int | None = None):
        """Start the H2D copy of the batch a later submit() will process (overlaps the previous batch's kernels)."""
        ticket = self.n_submitted if ticket is None else ticket
        slot = ticket % self.depth
        if images is not None:
            src = images if isinstance(images, torch.Tensor) else torch.from_numpy(images)
            self.host_in[slot].copy_(src)
        with torch.cuda.stream(self.copy_in):
            if self.consumed[slot] is not None:
                self.copy_in.wait_event(self.consumed[slot])
            self.dev_in[slot].copy_(self.host_in[slot], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.copy_in)
        self.staged[slot] = ev_in

    def submit(self, images, prefetch_next: bool = False) -> int:
        ticket = self.n_submitted
        slot = ticket % self.depth
        if self.slots[slot] is not None:
            raise RuntimeError("collect() the batch submitted %d calls ago before reusing its buffers" % self.depth)
        if images is not None or self.staged[slot] is None:
            self.stage(images, ticket)
        ev_in, self.staged[slot] = self.staged[slot], None
        if prefetch_next and self.depth > 1:
            self.stage(None, ticket + 1)
        compute = torch.cuda.current_stream()
        compute.wait_event(ev_in)
        ws, k, thr, pct = self.args
        hr = harris(self.dev_in[slot], ws, k, thr, compact=True, out=self.resp)
        lm = lambda_minus(self.dev_in[slot], pct, compact=True, out=self.eig)
        self.consumed[slot] = torch.cuda.Event()
        self.consumed[slot].record(compute)
        n1, n2 = int(hr.coords.shape[0]), int(lm.coords.shape[0])
        if max(n1, n2) > self.cap:
            raise _lib.CapacityError("BatchCornerDetector.submit (corners of the batch)", max(n1, n2), self.cap)
        offs = torch.stack([hr.image_offsets, lm.image_offsets])
        ev_done = torch.cuda.Event()
        ev_done.record(compute)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(ev_done)
            self.h_coords[slot][:n1].copy_(hr.coords, non_blocking=True)
            self.h_values[slot][:n1].copy_(hr.values, non_blocking=True)
            self.h_lcoords[slot][:n2].copy_(lm.coords, non_blocking=True)
            self.h_offs[slot].copy_(offs, non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(self.copy_out)
        self.slots[slot] = (ev_out, n1, n2, (hr, lm, offs))
        self.n_submitted += 1
        return ticket

    def collect(self, ticket: int):
        """-> (harris_coords [n,2], harris_values [n], harris_image_offsets [B+1], lm_coords [m,2], lm_image_offsets)
        as numpy views into the pinned buffers (valid until the slot is reused)."""
        slot = ticket % self.depth
        ev_out, n1, n2, _keep = self.slots[slot]
        ev_out.synchronize()
        self.slots[slot] = None
        offs = self.h_offs[slot].numpy()
        return (self.h_coords[slot][:n1].numpy(), self.h_values[slot][:n1].numpy(), offs[0],
                self.h_lcoords[slot][:n2].numpy(), offs[1])
