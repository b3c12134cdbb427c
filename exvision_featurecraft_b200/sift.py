"""Batched, device-resident SIFT: scale space -> extrema -> orientation -> descriptors over the C ABI.

This is the driver the reference-shaped functions in ``utils/SIFT_scale_space.py`` and
``utils/keypoint_descriptor.py`` sit on, and what the benchmarks call.  All images of a batch have the
same shape; every stage is one (or a few) kernel launches over the whole batch.  The only host work is
the parameter tables the reference also builds on the host (Gaussian taps, window tables: a few KB).

Reference structure followed (APP = FeatureCraft-PyQt-App):
  generate_octaves_pyramid            APP/utils/SIFT_scale_space.py:163-210
  generate_gaussian_images_in_octave  APP/utils/SIFT_scale_space.py:109-160
  represent_keypoints                 APP/utils/keypoint_descriptor.py:42-69   (scale range quirk)
  dog_keypoints_orientations          APP/utils/keypoint_descriptor.py:116-172
  extract_sift_descriptors            APP/utils/keypoint_descriptor.py:230-289
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass, field, replace
from math import cos, sin

import numpy as np
import torch

from . import _lib
from .device import ptr, read_int, read_ints_at, read_last_ints, require_cuda, scan_scratch, stream_ptr, to_device


@dataclass(frozen=True)
class SiftParams:
    """Defaults = the app's (APP/FeatureCraft_Backend.py:58-63)."""

    n_octaves: int = 4
    s_value: int = 2
    sigma_base: float = 1.6
    contrast_th: float = 0.03
    ratio_th: float = 10
    variant: str = "app"  # "app": keep every extremum; "script": SIFT.py contrast + edge tests
    # arithmetic of the scale space:
    #   "mixed" (production): float32 bulk arithmetic with every discrete decision certified -- an extremum, an
    #           orientation bin or a descriptor bin is taken from float32 values only when their error bound cannot
    #           change it, and is otherwise re-evaluated on float64 values: the keypoint lists are the float64 path's;
    #   "f64":  everything in float64 (the reference's arithmetic).
    # (An uncertified float32 mode existed in round 1: ~1 % of its descriptors had a sample on the wrong side of a bin
    # edge and a few dozen extrema per image flipped; certified mode costs 35 % more time and has neither.)
    dtype: str = "mixed"
    # `hist >= 0.8 * hist.max()` on the float32 orientation histogram: numpy >= 2 multiplies in float32 (default; what this
    # image's numpy does and what the golden vectors were produced with), numpy 1.26 -- the version the reference's
    # requirements.txt pins -- rounds the float64 product to float32.  The cuts differ by at most one float32 ulp.
    numpy_legacy_promotion: bool = False

    def __post_init__(self):
        if self.dtype not in ("mixed", "f64"):
            raise ValueError("dtype must be 'mixed' or 'f64'")

    @property
    def torch_dtype(self):
        """element type of the octave base images (both modes start from the float64 base)"""
        return torch.float64

    @property
    def fc_dtype(self):
        return _lib.FC_F64

    @property
    def n_levels(self):
        return self.s_value + 3


# ---------------------------------------------------------------------------------------------
# host-side parameter tables (the reference builds the same tables on the host; a7 in SURVEY 8(a))
# ---------------------------------------------------------------------------------------------
def sigma_levels(sigma, s):
    """generateGaussianKernels' sigma sequence: repeated multiplication by 2**(1/s) (SS:46-56)."""
    out = [sigma]
    lvl = sigma
    k = 2 ** (1 / s)
    for _ in range(1, s + 3):
        lvl *= k
        out.append(lvl)
    return out


def kernel_radius(sigma):
    """gaussian_filter_kernel's half width: ((4*sigma)+1)//2 in float arithmetic (SS:18-20)."""
    return int(((4 * sigma) + 1) // 2)


def separable_taps(sigma):
    """g(x) = exp(-x^2/2s^2)/sqrt(2*pi*s^2): outer(g, g) is the reference's (unnormalised, truncated)
    2-D kernel exp(-r^2/2s^2)/(2*pi*s^2) to 1 ulp."""
    r = kernel_radius(sigma)
    x = np.arange(-r, r + 1, dtype=np.float64)
    return np.exp(-(x ** 2) / (2 * sigma ** 2)) / np.sqrt(2 * np.pi * (sigma ** 2))


def gaussian_window_table(sigma, ry=None, rx=None):
    """gaussian_filter_kernel(sigma) as the dense 2-D table the orientation stage multiplies by (SS:22-28).
    ry / rx restrict it to the offsets |dy| <= ry, |dx| <= rx (same element-wise expression, so the same values):
    a window clipped to a small high-octave image never reads the rest of a (2r+1)^2 table."""
    r = kernel_radius(sigma)
    ry = r if ry is None else min(r, ry)
    rx = r if rx is None else min(r, rx)
    x = np.arange(-ry, ry + 1)[:, np.newaxis]
    y = np.arange(-rx, rx + 1)
    k = np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    k /= 2 * np.pi * (sigma ** 2)
    return k


def gaussian_mask16(sigma, n=16):
    """get_gaussian_mask (KD:215-227): sum-normalised Gaussian centred at (n-1)/2."""
    if not sigma > 0:
        raise ValueError("Invalid value of Sigma")
    k = np.fromfunction(
        lambda x, y: (1 / (2 * np.pi * sigma ** 2)) * np.exp(-((x - (n - 1) / 2) ** 2 + (y - (n - 1) / 2) ** 2) / (2 * sigma ** 2)),
        (n, n),
    )
    return k / np.sum(k)


def keypoint_sigma(sigma_base, octave_idx, scale_idx, s):
    """KD:127-134 / :244-251 with the reference's operation order."""
    return 1.5 * sigma_base * (2 ** octave_idx) * ((2 ** (1 / s)) ** (scale_idx))


def trig_table():
    """Per orientation bin: cos, sin of theta*3.14159/180 (KD:182) and the fixed-point column deltas
    rint(cos*x*1024), rint(sin*x*1024), x = 0..15, of cv2.warpAffine's nearest-neighbour path; last row: the
    8 bin edges of int(rel * 8 / 360.0) (KD:263-268)."""
    t = np.zeros((37, 34), dtype=np.float64)
    xs = np.arange(16)
    for b in range(36):
        theta = ((b + 0.5) * (360.0 / 36) % 360) * (3.14159 / 180)
        c, s = cos(theta), sin(theta)
        t[b, 0], t[b, 1] = c, s
        t[b, 2:18] = np.rint(c * xs * 1024)
        t[b, 18:34] = np.rint(s * xs * 1024)
    # row 36: descriptor bin edges.  int(rel * 8 / 360.0) >= j  <=>  rel >= edge[j-1]; the smallest such double is
    # found by stepping from 45*j so that the device's comparisons reproduce the reference's two roundings exactly
    for j in range(1, 9):
        r = np.float64(45.0 * j)
        while (r * 8) / 360.0 >= j:
            r = np.nextafter(r, -np.inf)
        while (r * 8) / 360.0 < j:
            r = np.nextafter(r, np.inf)
        t[36, j - 1] = r
    return t


_TABLE_CACHE: "OrderedDict" = OrderedDict()
_TABLE_CACHE_LIMIT = 256 << 20  # bytes of device tables kept (least recently used go first)


class StageTimer:
    """Optional CUDA-event timing of the pipeline stages (bench.py installs one; None = no overhead)."""

    def __init__(self, only=None):
        self.events = []
        self.only = set(only) if only else None  # record just these stages (the others cost no events at all)

    class _Ctx:
        def __init__(self, owner, name):
            self.owner, self.name = owner, name

        def __enter__(self):
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

        def __exit__(self, *exc):
            self.b.record()
            self.owner.events.append((self.name, self.a, self.b))

    def stage(self, name):
        if self.only is not None and name not in self.only:
            return _NULL
        return StageTimer._Ctx(self, name)

    def totals(self):
        """ms per stage name, summed (call after a synchronize)."""
        out = {}
        for name, a, b in self.events:
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def reset(self):
        self.events = []


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


TIMER: StageTimer | None = None
_NULL = _NullCtx()


def _stage(name):
    return TIMER.stage(name) if TIMER is not None else _NULL


def _cached(key, builder, dtype=torch.float64):
    dev = torch.cuda.current_device()
    k = (dev,) + key + (dtype,)
    t = _TABLE_CACHE.get(k)
    if t is None:
        t = to_device(np.ascontiguousarray(builder(), dtype=np.float64)).to(dtype).contiguous()
        _TABLE_CACHE[k] = t
        total = sum(v.numel() * v.element_size() for v in _TABLE_CACHE.values())
        while total > _TABLE_CACHE_LIMIT and len(_TABLE_CACHE) > 1:
            _, old = _TABLE_CACHE.popitem(last=False)
            total -= old.numel() * old.element_size()
    else:
        _TABLE_CACHE.move_to_end(k)
    return t


# ---------------------------------------------------------------------------------------------
@dataclass
class Octave:
    gauss: torch.Tensor | None  # [B, L, h, w]; None when the certified pipeline fused the search into the blur kernel
    extrema: torch.Tensor | None  # [L-3, B, h, ceil(w/32)] int32 bit masks, index k-1
    height: int
    width: int
    batch: int = 0


@dataclass
class ScaleBlock:
    """The block-major lists of one (octave, scale): keypoints (image,row,col), oriented (image,row,col,bin)."""

    octave: int
    scale: int
    keypoints: torch.Tensor
    oriented: torch.Tensor | None = None


@dataclass
class SiftResult:
    """Outputs of detect_and_describe, all on the device and already in the reference's order: images one after the
    other, each image's keypoints in emission order (octave, scale, row-major, orientation bin)."""

    params: SiftParams
    octaves: list
    blocks: list = field(default_factory=list)
    descriptors: torch.Tensor | None = None  # [M, 128]
    records: torch.Tensor | None = None      # [M, 2] int32: packed (row, col) | (octave, scale, bin), see unpack_records
    counts: torch.Tensor | None = None       # [B] int32 descriptors per image
    undecided_extrema: int = 0               # certified mode: pixels the float64 repair kernel decided
    recheck_counts: torch.Tensor | None = None  # certified mode: per (octave, scale) block, keypoints decided in float64

    @property
    def total(self) -> int:
        return 0 if self.descriptors is None else int(self.descriptors.shape[0])

    def assemble_device(self, batch: int):
        """(records [M,2] int32, descriptors [M,128], per-image counts [B] int32) -- device tensors, no copies."""
        return self.records, self.descriptors, self.counts

    def assemble(self, batch: int):
        """Per-image (points [m,5] float64 = (i, j, octave, scale, angle), descriptors [m,128]) on the host, in the
        reference's emission order (octave, scale, row-major, bin)."""
        if self.descriptors is None or self.total == 0:
            e = np.zeros((0, 5))
            return [(e, np.zeros((0, 128), dtype=np.float32)) for _ in range(batch)]
        counts = self.counts.cpu().numpy()
        pts = unpack_records(self.records.cpu().numpy())
        d = self.descriptors.cpu().numpy()
        out, lo = [], 0
        for b in range(batch):
            out.append((pts[lo:lo + counts[b]], d[lo:lo + counts[b]]))
            lo += counts[b]
        return out


def unpack_records(rec: np.ndarray) -> np.ndarray:
    """[M,2] int32 keypoint records (uint16 row | uint16 col << 16, uint8 octave | scale << 8 | bin << 16) -> the
    reference's 5-tuples as [M,5] float64: (i, j, octave, scale, (bin + 0.5) * 10) (KD:167-170)."""
    r = np.ascontiguousarray(rec).view(np.uint32).reshape(-1, 2)
    out = np.empty((r.shape[0], 5), dtype=np.float64)
    out[:, 0] = r[:, 0] & 0xFFFF
    out[:, 1] = r[:, 0] >> 16
    out[:, 2] = r[:, 1] & 0xFF
    out[:, 3] = (r[:, 1] >> 8) & 0xFF
    out[:, 4] = ((r[:, 1] >> 16) & 0xFF) * 10.0 + 5.0
    return out


def _as_base_batch(base, params: SiftParams) -> torch.Tensor:
    t = to_device(base)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3:
        raise ValueError("expected [H,W] or [B,H,W] grayscale base images")
    return t.to(torch.float64).contiguous()


def upsample2(gray, params: SiftParams) -> torch.Tensor:
    """rescale(gray, 2, anti_aliasing=False) (KD:296) for [H,W] or [B,H,W] float64 images -> [B,2H,2W]."""
    t = to_device(np.asarray(gray, dtype=np.float64) if not isinstance(gray, torch.Tensor) else gray, torch.float64)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    b, h, w = t.shape
    out = torch.empty((b, 2 * h, 2 * w), dtype=params.torch_dtype, device=t.device)
    if params.dtype == "mixed" and w % 2 == 0:
        # certified mode wants the base in both precisions and its magnitude bound: one pass writes all three; the
        # float32 companions ride along on the tensor (gauss_octave_mixed picks them up)
        out._fc_base32 = torch.empty((b, 2 * h, 2 * w), dtype=torch.float32, device=t.device)
        out._fc_bound = torch.empty(1, dtype=torch.float32, device=t.device)
        _lib.call("fc_upsample2_dual", ptr(t), b, h, w, ptr(out), ptr(out._fc_base32), ptr(out._fc_bound), stream_ptr())
        return out
    _lib.call("fc_upsample2", ptr(t), b, h, w, params.fc_dtype, ptr(out), stream_ptr())
    return out


def _level_taps(params: SiftParams, taps=None):
    if taps is None:
        taps = [separable_taps(s) for s in sigma_levels(params.sigma_base, params.s_value)]
    lens = (C.c_int * len(taps))(*[len(t) for t in taps])
    flat = np.ascontiguousarray(np.concatenate(taps))
    return flat, lens


def _filter_code(params: SiftParams) -> int:
    if params.variant == "script" and params.n_levels - 3 >= 3:
        # SIFT.py:206 indexes a 3-deep stack with the global k: k >= 3 raises in the reference
        raise IndexError("index 3 is out of bounds for axis 2 with size 3")
    return _lib.FC_FILTER_SCRIPT if params.variant == "script" else _lib.FC_FILTER_APP


def gauss_octave(base: torch.Tensor, params: SiftParams, octave_index: int, with_extrema: bool = True, taps=None) -> Octave:
    """generate_gaussian_images_in_octave for a [B,h,w] batch (SS:109-160).  `taps` overrides the separable
    taps derived from (sigma_base, s_value) -- used when the caller hands in its own kernel list."""
    if params.dtype == "mixed":
        return gauss_octave_mixed(base, params, octave_index, with_extrema, taps)[0]
    b, h, w = base.shape
    L = params.n_levels
    flat, lens = _level_taps(params, taps)
    gauss = torch.empty((b, L, h, w), dtype=params.torch_dtype, device=base.device)
    extrema = None
    if with_extrema and L >= 4:
        extrema = torch.empty((L - 3, b, h, (w + 31) // 32), dtype=torch.int32, device=base.device)
    filt = _filter_code(params)
    with _stage("pyramid_o%d" % octave_index):
        _lib.call("fc_gauss_octave", ptr(base), params.fc_dtype, b, h, w, L, flat.ctypes.data_as(C.c_void_p),
                  C.cast(lens, C.c_void_p), 1 if octave_index > 0 else 0, filt, float(params.contrast_th),
                  float(params.ratio_th), ptr(gauss), None, stream_ptr())
    if extrema is not None:
        with _stage("extrema"):
            _lib.call("fc_gauss_octave_extrema", ptr(gauss), params.fc_dtype, b, h, w, L, filt, float(params.contrast_th),
                      float(params.ratio_th), ptr(extrema), stream_ptr())
    return Octave(gauss, extrema, h, w, b)


# --- certified ("mixed") mode -----------------------------------------------------------------------------------
#: |DoG32 - DoG64| stays below 4e-7 * max|image| (13-tap .. 27-tap float32 blurs of values in [0, 1]: measured 3.4e-7
#: on 8M pixels, rms 4e-8); a comparison of two DoG values is trusted when they differ by more than EXTREMA_EPS * max|image|
EXTREMA_EPS = 2.0 ** -19
#: entries of the undecided-pixel list per octave: this fraction of the octave's (pixel, k) pairs (app filter decides
#: ~2e-4 of them in float64; the script filter sends every extremum, ~1e-2).  More than that -> float64 fallback.
AMB_FRACTION = {"app": 1.0 / 256, "script": 1.0 / 16}


@dataclass
class MixedOctave:
    """float64 companions of a certified octave: base image, levels 1..s, and the undecided-pixel counter."""

    base: torch.Tensor        # [B, h, w] float64
    mid: torch.Tensor         # [B, s, h, w] float64: levels 1 .. s
    amb_count: torch.Tensor   # [1] int32, entries the float32 search left to float64 (may exceed amb_cap)
    amb_cap: int


def gauss_octave_mixed(base64: torch.Tensor, params: SiftParams, octave_index: int, with_extrema: bool = True, taps=None,
                       value_bound: torch.Tensor | None = None, amb_count: torch.Tensor | None = None, need_stack: bool = True):
    """One octave in certified mode -> (Octave with the float32 stack and the certified extremum masks, MixedOctave).

    base64 is the octave's float64 base.  value_bound: device float with max |image| (computed here for octave 0 while
    the float32 copy is made; deeper octaves are blurs of it and inherit the bound).  need_stack=False (the pipeline:
    nothing reads the float32 levels after the search) lets the blur kernel run the search on its own tiles and skip
    writing the stack (fc_gauss_octave_search; default radii only, otherwise the two-kernel form is used)."""
    b, h, w = base64.shape
    dev = base64.device
    L, s = params.n_levels, params.s_value
    flat, lens = _level_taps(params, taps)
    tp, lp = flat.ctypes.data_as(C.c_void_p), C.cast(lens, C.c_void_p)
    filt = _filter_code(params)
    ident = 1 if octave_index > 0 else 0
    base32 = getattr(base64, "_fc_base32", None)
    if base32 is not None and value_bound is None:
        value_bound = getattr(base64, "_fc_bound", None)
    if base32 is None or value_bound is None:
        base32 = torch.empty((b, h, w), dtype=torch.float32, device=dev)
        with _stage("pyramid_f64"):
            if value_bound is None:
                value_bound = torch.empty(1, dtype=torch.float32, device=dev)
                _lib.call("fc_convert_f64_bound", ptr(base64), C.c_int64(base64.numel()), ptr(base32), ptr(value_bound), stream_ptr())
            else:
                _lib.call("fc_convert_f64", ptr(base64), C.c_int64(base64.numel()), _lib.FC_F32, ptr(base32), stream_ptr())
    if amb_count is None:
        amb_count = torch.zeros(1, dtype=torch.int32, device=dev)
    extrema, cap, amb, gauss = None, 0, None, None
    if with_extrema and L >= 4:
        extrema = torch.empty((L - 3, b, h, (w + 31) // 32), dtype=torch.int32, device=dev)
        cap = int(b * h * w * (L - 3) * AMB_FRACTION[params.variant]) + 4096
        amb = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    fused = False
    if extrema is not None and not need_stack and L == 5 and w % 4 == 0:
        with _stage("pyramid_o%d" % octave_index):
            rc = _lib.load().fc_gauss_octave_search(ptr(base32), b, h, w, L, tp, lp, ident, filt, ptr(value_bound), float(EXTREMA_EPS),
                                                    ptr(extrema), ptr(amb), cap, ptr(amb_count), stream_ptr())
        if rc == _lib.FC_OK:
            fused = True
        elif rc != -4:  # FC_ERR_UNSUPPORTED: other radii -> the two-kernel form below
            _lib.check(rc, "fc_gauss_octave_search")
    if not fused:
        gauss = torch.empty((b, L, h, w), dtype=torch.float32, device=dev)
        with _stage("pyramid_o%d" % octave_index):
            _lib.call("fc_gauss_octave", ptr(base32), _lib.FC_F32, b, h, w, L, tp, lp, ident, filt, float(params.contrast_th),
                      float(params.ratio_th), ptr(gauss), None, stream_ptr())
    n_mid = min(s, L - 1)
    mid = torch.empty((b, n_mid, h, w), dtype=torch.float64, device=dev)
    with _stage("pyramid_f64"):
        _lib.call("fc_gauss_levels_f64", ptr(base64), b, h, w, L, tp, lp, 1, n_mid, ptr(mid), stream_ptr())
    if extrema is not None:
        if not fused:
            with _stage("extrema"):
                _lib.call("fc_gauss_octave_extrema_certified", ptr(gauss), b, h, w, L, filt, float(params.contrast_th),
                          float(params.ratio_th), ptr(value_bound), float(EXTREMA_EPS), ptr(extrema), ptr(amb), cap, ptr(amb_count),
                          stream_ptr())
        with _stage("extrema_f64"):
            _lib.call("fc_extrema_repair", ptr(base64), ident, ptr(mid), 1, n_mid, b, h, w, L, tp, lp, filt,
                      float(params.contrast_th), float(params.ratio_th), ptr(amb), ptr(amb_count), cap, ptr(extrema), stream_ptr())
    octv = Octave(gauss, extrema, h, w, b)
    octv.value_bound = value_bound
    return octv, MixedOctave(base64, mid, amb_count, cap)


def downsample(octv: Octave, params: SiftParams, mixed: MixedOctave | None = None) -> torch.Tensor:
    """next base = G[-3][::2, ::2] (SS:209); certified mode takes it from the float64 level s."""
    if mixed is not None:
        b, h, w, L = octv.batch, octv.height, octv.width, params.n_levels
        n_mid = mixed.mid.shape[1]
        if L - 3 < 1 or L - 3 > n_mid:
            raise _lib.FeatureCraftError("certified mode keeps levels 1..s; the next base is level s")
        out = torch.empty((b, (h + 1) // 2, (w + 1) // 2), dtype=torch.float64, device=mixed.mid.device)
        out._fc_base32 = torch.empty(out.shape, dtype=torch.float32, device=out.device)
        _lib.call("fc_downsample2_dual", ptr(mixed.mid), b, n_mid, L - 4, h, w, ptr(out), ptr(out._fc_base32), stream_ptr())
        return out
    b, L, h, w = octv.gauss.shape
    out = torch.empty((b, (h + 1) // 2, (w + 1) // 2), dtype=octv.gauss.dtype, device=octv.gauss.device)
    _lib.call("fc_downsample2", ptr(octv.gauss), params.fc_dtype, b, L, L - 3, h, w, ptr(out), stream_ptr())
    return out


def _build_octaves(img, params: SiftParams, with_extrema: bool = True, need_stack: bool = True):
    """-> (octaves, mixed companions or None per octave, per-octave undecided-pixel counters or None).  Certified mode
    that overflows its undecided-pixel list is detected by the callers (they read the counters with the keypoint counts)."""
    octaves, companions = [], []
    bound = None
    amb_counts = None
    if params.dtype == "mixed":
        amb_counts = torch.empty(params.n_octaves, dtype=torch.int32, device=img.device)  # zeroed by the search kernels' launcher
    for o in range(params.n_octaves):
        if params.dtype == "mixed":
            octv, mx = gauss_octave_mixed(img, params, o, with_extrema, value_bound=bound, amb_count=amb_counts[o:o + 1],
                                          need_stack=need_stack)
            bound = octv.value_bound
        else:
            octv, mx = gauss_octave(img, params, o, with_extrema), None
        octaves.append(octv)
        companions.append(mx)
        if o + 1 < params.n_octaves:
            img = downsample(octv, params, mx)
    return octaves, companions, amb_counts


def scale_space(base, params: SiftParams, with_extrema: bool = True) -> list:
    """generate_octaves_pyramid (SS:163-210) for a batch of base images."""
    img = _as_base_batch(base, params)
    octaves, companions, amb_counts = _build_octaves(img, params, with_extrema)
    if params.dtype == "mixed" and with_extrema:
        counts = amb_counts.tolist()
        if any(c > m.amb_cap for c, m in zip(counts, companions)):
            return scale_space(base, replace(params, dtype="f64"), with_extrema)
    return octaves


def keypoints_from_mask(mask: torch.Tensor, width: int) -> torch.Tensor:
    """np.argwhere order (image, row, col) int32 triples of a [B,h,pitch] bit mask."""
    b, h, _ = mask.shape
    rows = b * h
    row_offsets = torch.empty(rows + 1, dtype=torch.int32, device=mask.device)
    scratch = scan_scratch(rows)
    _lib.call("fc_mask_scan", ptr(mask), b, h, width, ptr(row_offsets), ptr(scratch), stream_ptr())
    total = read_int(row_offsets)
    kps = torch.empty((total, 3), dtype=torch.int32, device=mask.device)
    if total:
        _lib.call("fc_mask_emit_keypoints", ptr(mask), ptr(row_offsets), b, h, width, ptr(kps), stream_ptr())
    return kps


def gradient_field(octv: Octave, level: int, params: SiftParams):
    """sift_gradient of one level of a float64 (or float32) stack -> (magnitude, direction, uint8 orientation bins), three
    dense planes.  (Certified mode uses gradient_codes on its float64 levels instead.)"""
    b, L, h, w = octv.gauss.shape
    fc = _lib.FC_F32 if octv.gauss.dtype == torch.float32 else _lib.FC_F64
    mag = torch.empty((b, h, w), dtype=octv.gauss.dtype, device=octv.gauss.device)
    direction = torch.empty_like(mag)
    obin = torch.empty((b, h, w), dtype=torch.uint8, device=mag.device)
    _lib.call("fc_gradient_field", ptr(octv.gauss), fc, b, L, level, h, w, ptr(mag), ptr(direction), ptr(obin), stream_ptr())
    return mag, direction, obin


def gradient_codes(levels64: torch.Tensor, plane: int) -> torch.Tensor:
    """Certified mode's gradient field of one float64 level: [B,h,w] int32 words = magnitude | exact direction code."""
    b, n, h, w = levels64.shape
    codes = torch.empty((b, h, w), dtype=torch.int32, device=levels64.device)
    _lib.call("fc_gradient_codes", ptr(levels64), b, n, plane, h, w, ptr(codes), stream_ptr())
    return codes


def _orientation_weights(sigma, h, w, dtype):
    """(radius, table_ry, table_rx, device table): gaussian_filter_kernel(sigma) (KD:136) restricted to the offsets a
    window clipped to an h x w image can reach."""
    radius = int(round(sigma * 2))
    if 2 * kernel_radius(sigma) + 1 != 2 * radius + 1:
        # the reference multiplies a (2r+1)^2 window by the kernel: shapes must agree or numpy raises
        raise ValueError("operands could not be broadcast together with shapes (%d,%d) (%d,%d)"
                         % (2 * radius + 1, 2 * radius + 1, 2 * kernel_radius(sigma) + 1, 2 * kernel_radius(sigma) + 1))
    ry, rx = min(radius, max(h - 1, 0)), min(radius, max(w - 1, 0))
    return radius, ry, rx, _cached(("ori", float(sigma), ry, rx), lambda: gaussian_window_table(sigma, ry, rx), dtype)


def separable_window_table(sigma, ry, rx):
    """Certified mode's float32 orientation weights: gaussian_filter_kernel(sigma)[dy, dx] = wy[dy] * wx[dx] with
    wy = exp(-dy^2 / 2 sigma^2) / (2 pi sigma^2), wx = exp(-dx^2 / 2 sigma^2); table = [wy (2 ry + 1) | wx (2 rx + 1)].
    The factorisation is exact in real arithmetic; its float32 roundings are part of orientation_tolerance."""
    dy = np.arange(-ry, ry + 1, dtype=np.float64)
    dx = np.arange(-rx, rx + 1, dtype=np.float64)
    wy = np.exp(-(dy ** 2) / (2 * sigma ** 2)) / (2 * np.pi * (sigma ** 2))
    wx = np.exp(-(dx ** 2) / (2 * sigma ** 2))
    return np.concatenate([wy, wx])


def orientation_tolerance(radius: int) -> float:
    """Relative error bound of a float32 orientation bin of certified mode against the float64 sum, doubled (the bin
    and the 0.8 * max cut both carry it): magnitudes are rounded to 16 mantissa bits (2^-17), the two weight factors,
    their product and the sample product to float32 (4 x 2^-24), and a lane adds at most ceil((2r+1)/lanes) * (2r+1)
    terms before the lanes are combined."""
    lanes = 8 if radius <= 10 else (16 if radius <= 15 else 32)
    side = 2 * radius + 1
    per_lane = -(-side // lanes) * side
    return 2.0 * (2.0 ** -17 + (per_lane + lanes + 6) * 2.0 ** -24)


def orientation_bins(mag, obin, keypoints, n, sigma, bin_mask, legacy=False):
    """Launch the 36-bin histogram kernel on gradient_field's magnitude / bin planes."""
    b, h, w = mag.shape
    fc = _lib.FC_F32 if mag.dtype == torch.float32 else _lib.FC_F64
    radius, ry, rx, weights = _orientation_weights(sigma, h, w, mag.dtype)
    _lib.call("fc_orientation_bins", ptr(mag), ptr(obin), fc, b, h, w, ptr(keypoints), n, ptr(weights), radius, ry, rx,
              ptr(bin_mask), int(bool(legacy)), stream_ptr())


def orientation_bins_certified(levels64, plane, codes, keypoints, n, sigma, bin_mask, recheck_list, recheck_count, legacy=False):
    """Certified mode: float32 histograms from the code words, then the float64 decision of the keypoints whose
    0.8 * max comparison the float32 error bound cannot guarantee (recheck_list: int32 [n] scratch, recheck_count: [1])."""
    b, h, w = codes.shape
    radius, ry, rx, w64 = _orientation_weights(sigma, h, w, torch.float64)
    w32 = _cached(("ori_sep", float(sigma), ry, rx), lambda: separable_window_table(sigma, ry, rx), torch.float32)
    _lib.call("fc_orientation_bins_coded", ptr(codes), b, h, w, ptr(keypoints), n, ptr(w32), radius, ry, rx,
              float(orientation_tolerance(radius)), ptr(bin_mask), ptr(recheck_list), ptr(recheck_count), stream_ptr())
    _lib.call("fc_orientation_recheck", ptr(levels64), levels64.shape[1], plane, ptr(codes), b, h, w, ptr(keypoints), n,
              ptr(w64), radius, ry, rx, ptr(bin_mask), ptr(recheck_list), ptr(recheck_count), int(bool(legacy)), stream_ptr())


def orient(mag, obin, keypoints, sigma, params: SiftParams) -> torch.Tensor:
    """dog_keypoints_orientations for one (octave, scale): -> oriented [m,4] int32 (image,row,col,bin)."""
    n = keypoints.shape[0]
    dev = mag.device
    if n == 0:
        return torch.empty((0, 4), dtype=torch.int32, device=dev)
    bin_mask = torch.empty(n, dtype=torch.int64, device=dev)
    orientation_bins(mag, obin, keypoints, n, sigma, bin_mask, params.numpy_legacy_promotion)
    offsets = torch.empty(n + 1, dtype=torch.int32, device=dev)
    scratch = scan_scratch(n)
    _lib.call("fc_orientation_scan", ptr(bin_mask), n, ptr(offsets), ptr(scratch), stream_ptr())
    total = read_int(offsets)
    oriented = torch.empty((total, 4), dtype=torch.int32, device=dev)
    if total:
        _lib.call("fc_orientation_emit", ptr(keypoints), ptr(bin_mask), ptr(offsets), n, ptr(oriented), stream_ptr())
    return oriented


_FC_OUT = {torch.float32: _lib.FC_F32, torch.float64: _lib.FC_F64, torch.float16: _lib.FC_F16}


def describe(mag, direction, oriented, sigma, params: SiftParams, out_dtype=torch.float32, out=None, placement=None) -> torch.Tensor:
    """extract_sift_descriptors for one (octave, scale): -> [m,128].  mag = certified mode's code plane (int32) or a
    dense magnitude plane (with `direction`).
    out / placement: write into the rows of `out` that the fc_desc_placement (an _lib.DescPlacement) selects."""
    m = oriented.shape[0]
    desc = out if out is not None else torch.empty((m, 128), dtype=out_dtype, device=mag.device)
    if m == 0:
        return desc
    gmask = _cached(("gmask", float(sigma)), lambda: gaussian_mask16(sigma))
    trig = _cached(("trig",), trig_table)
    b, h, w = mag.shape
    out_fc = _FC_OUT[desc.dtype]
    pl = C.byref(placement) if placement is not None else None
    if mag.dtype == torch.int32:
        _lib.call("fc_descriptors_coded", ptr(mag), b, h, w, ptr(oriented), m, ptr(gmask), ptr(trig), ptr(desc), out_fc, pl,
                  stream_ptr())
    elif desc.dtype == torch.float16:
        raise _lib.FeatureCraftError("float16 descriptor export is implemented for the certified (mixed) pipeline")
    else:
        fc = _lib.FC_F32 if mag.dtype == torch.float32 else _lib.FC_F64
        _lib.call("fc_descriptors", ptr(mag), ptr(direction), fc, b, h, w, ptr(oriented), m, ptr(gmask), ptr(trig),
                  ptr(desc), out_fc, pl, stream_ptr())
    return desc


def refine_keypoints(octv: Octave, k: int, keypoints: torch.Tensor, contrast_th: float = 0.03, ratio_th: float = 10.0):
    """OPT-IN sub-pixel refinement (the reference keeps it commented out, SS:89-103): localize_keypoint (SS:213-261) for
    the (image, row, col) keypoints of DoG level k of a float64 octave.  -> dict of device tensors: offsets [n,3] in
    (x, y, s) order, contrast [n], ratio [n], keep [n] bool (the commented-out acceptance tests)."""
    if octv.gauss.dtype != torch.float64:
        raise _lib.FeatureCraftError("refine_keypoints works on the float64 scale space (SiftParams(dtype='f64'))")
    b, L, h, w = octv.gauss.shape
    n = int(keypoints.shape[0])
    dev = octv.gauss.device
    out = {"offsets": torch.empty((n, 3), dtype=torch.float64, device=dev), "contrast": torch.empty(n, dtype=torch.float64, device=dev),
           "ratio": torch.empty(n, dtype=torch.float64, device=dev)}
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.call("fc_refine_keypoints", ptr(octv.gauss), b, L, h, w, ptr(keypoints.contiguous()), n, int(k), float(contrast_th),
              float(ratio_th), ptr(out["offsets"]), ptr(out["contrast"]), ptr(out["ratio"]), ptr(keep), stream_ptr())
    out["keep"] = keep.bool()
    return out


def active_scales(params: SiftParams):
    """represent_keypoints builds masks for dog_idx in range(1, n_octaves-1) (KD:60: len(DoG) is the
    number of OCTAVES); get_keypoints only ever emits k in 1..s.  dog_keypoints_orientations then reads
    img_gaussians[o][dog_idx] for every mask, raising IndexError when dog_idx >= s+3."""
    ks = list(range(1, params.n_octaves - 1))
    if ks and ks[-1] >= params.n_levels:
        raise IndexError("list index out of range")
    return [k for k in ks if k <= params.s_value]


def detect_and_describe(base, params: SiftParams = SiftParams(), out_dtype=torch.float32, keep_octaves: bool = False) -> SiftResult:
    """The whole hot path after the x2 up-sample: computeKeypointsAndDescriptors (KD:297-311) for a
    [B,H,W] batch of base images.

    The (octave, scale) blocks are independent once the pyramid exists, so the work is queued in four phases with
    only TWO host synchronisations per batch (the two data-dependent output sizes): A pyramid + extrema of every
    octave, B mask scans -> all keypoint counts, C keypoint lists + gradient fields + orientation histograms + bin
    scans -> all oriented counts, D oriented lists + descriptors."""
    require_cuda()
    scales = active_scales(params)
    img = _as_base_batch(base, params)
    dev = img.device
    res = SiftResult(params, [])
    certified = params.dtype == "mixed"
    # A: scale space
    octaves, companions, amb_dev = _build_octaves(img, params, True, need_stack=keep_octaves)
    if keep_octaves:
        res.octaves = octaves
    work = [(o, k) for o in range(params.n_octaves) for k in scales]
    if not work:
        return res
    # B: row-major compaction offsets.  The row counts of every (octave, scale) mask go into ONE concatenated array,
    # one scan covers them all and one gather launch brings the cumulative keypoint counts at the block boundaries to
    # the host: keypoints of all blocks then live in one list, block after block, in the reference's order.
    # (Certified mode reads its per-octave undecided-pixel counters in the same synchronisation.)
    rows = [octaves[o].batch * octaves[o].height for o, _ in work]
    row_start = [0]
    for r in rows:
        row_start.append(row_start[-1] + r)
    n_extra = params.n_octaves if certified else 0
    with _stage("compact"):
        counts = torch.empty(row_start[-1], dtype=torch.int32, device=dev)
        for i, (o, k) in enumerate(work):
            octv = octaves[o]
            _lib.call("fc_mask_row_count", ptr(octv.extrema[k - 1]), octv.batch, octv.height, octv.width,
                      ptr(counts[row_start[i]:]), stream_ptr())
        row_offs = torch.empty(row_start[-1] + 1, dtype=torch.int32, device=dev)
        _lib.call("fc_exclusive_scan_i32", ptr(counts), row_start[-1], ptr(row_offs), ptr(scan_scratch(row_start[-1])), stream_ptr())
    # the gradient code planes need no counts: queued behind the scan, they run while the host waits for the counts
    codes_of = {}

    def queue_codes():
        with _stage("gradient_field"):
            for i, (o, k) in enumerate(work):
                codes_of[i] = gradient_codes(companions[o].mid, k - 1)

    if certified:
        got = read_ints_at(row_offs, row_start[1:], amb_dev, range(n_extra), between=queue_codes)
    else:
        got = read_ints_at(row_offs, row_start[1:])
    kp_end, amb_counts = got[:len(work)], got[len(work):]  # cumulative keypoint count after each block
    if certified and any(c > m.amb_cap for c, m in zip(amb_counts, companions)):
        # more undecided pixels than the list holds (large exactly-flat regions): the float64 kernels decide everything
        return detect_and_describe(base, replace(params, dtype="f64"), out_dtype, keep_octaves)
    res.undecided_extrema = int(sum(amb_counts))
    kp_start = [0] + kp_end[:-1]
    total_kp = kp_end[-1]
    kps_all = torch.empty((total_kp, 3), dtype=torch.int32, device=dev)
    bin_mask_all = torch.empty(max(total_kp, 1), dtype=torch.int64, device=dev)
    if certified:
        recheck_all = torch.empty(max(total_kp, 1), dtype=torch.int32, device=dev)
        recheck_counts = torch.zeros(len(work), dtype=torch.int32, device=dev)
        res.recheck_counts = recheck_counts
    # C: keypoints, gradient fields, orientation histograms of every block; then ONE scan of all per-keypoint
    # orientation counts and one gather for the cumulative oriented counts
    state = []
    for i, (o, k) in enumerate(work):
        octv = octaves[o]
        b, h = octv.batch, octv.height
        n = kp_end[i] - kp_start[i]
        kps = kps_all[kp_start[i]:kp_end[i]]
        blk = ScaleBlock(o, k, kps)
        res.blocks.append(blk)
        if n == 0:
            state.append(None)
            continue
        with _stage("compact"):
            # the row offsets are positions in the concatenated list: emit through its base pointer
            _lib.call("fc_mask_emit_keypoints", ptr(octv.extrema[k - 1]), ptr(row_offs[row_start[i]:]), b, h, octv.width,
                      ptr(kps_all), stream_ptr())
        sigma = keypoint_sigma(params.sigma_base, o, k, params.s_value)
        if certified:
            mid = companions[o].mid
            codes = codes_of[i]
            with _stage("orientation"):
                orientation_bins_certified(mid, k - 1, codes, kps, n, sigma, bin_mask_all[kp_start[i]:kp_end[i]],
                                           recheck_all[kp_start[i]:kp_end[i]], recheck_counts[i:i + 1],
                                           params.numpy_legacy_promotion)
            state.append((codes, None, sigma))
            continue
        with _stage("gradient_field"):
            mag, direction, obin = gradient_field(octv, k, params)
        with _stage("orientation"):
            orientation_bins(mag, obin, kps, n, sigma, bin_mask_all[kp_start[i]:kp_end[i]], params.numpy_legacy_promotion)
        state.append((mag, direction, sigma))
    batch = octaves[0].batch
    if total_kp == 0:
        for blk in res.blocks:
            blk.oriented = torch.empty((0, 4), dtype=torch.int32, device=dev)
        res.descriptors = torch.empty((0, 128), dtype=out_dtype, device=dev)
        res.records = torch.empty((0, 2), dtype=torch.int32, device=dev)
        res.counts = torch.zeros(batch, dtype=torch.int32, device=dev)
        return res
    with _stage("orientation"):
        ori_counts = torch.empty(total_kp, dtype=torch.int32, device=dev)
        _lib.call("fc_popcount_u64", ptr(bin_mask_all), total_kp, ptr(ori_counts), stream_ptr())
        ori_offs = torch.empty(total_kp + 1, dtype=torch.int32, device=dev)
        _lib.call("fc_exclusive_scan_i32", ptr(ori_counts), total_kp, ptr(ori_offs), ptr(scan_scratch(total_kp)), stream_ptr())
        ori_end = read_ints_at(ori_offs, kp_end)  # cumulative oriented-keypoint count after each block
    ori_start = [0] + ori_end[:-1]
    total = ori_end[-1]
    oriented_all = torch.empty((max(total, 1), 4), dtype=torch.int32, device=dev)
    # D: oriented keypoint lists (every array is concatenated in block order, so ONE launch expands them all), the
    # final row of every (block, image) run, and the descriptors -- written straight to their final image-major rows
    # together with the keypoint records, so nothing is sorted or gathered afterwards
    # (float16 export is a store format of the certified kernels; the other modes convert afterwards)
    desc_dtype = out_dtype if (certified or out_dtype != torch.float16) else torch.float32
    res.descriptors = torch.empty((total, 128), dtype=desc_dtype, device=dev)
    res.records = torch.empty((total, 2), dtype=torch.int32, device=dev)
    res.counts = torch.empty(batch, dtype=torch.int32, device=dev)
    remap = torch.empty((len(work), batch, 2), dtype=torch.int32, device=dev)
    with _stage("orientation"):
        if total > 0:
            _lib.call("fc_orientation_emit", ptr(kps_all), ptr(bin_mask_all), ptr(ori_offs), total_kp, ptr(oriented_all), stream_ptr())
        starts = (C.c_int * len(work))(*row_start[:-1])
        heights = (C.c_int * len(work))(*[octaves[o].height for o, _ in work])
        _lib.call("fc_sift_layout", ptr(row_offs), ptr(ori_offs), starts, heights, len(work), batch, ptr(remap), ptr(res.counts),
                  stream_ptr())
    for i, (blk, st) in enumerate(zip(res.blocks, state)):
        blk.oriented = oriented_all[ori_start[i]:ori_end[i]]
        if st is None or ori_end[i] == ori_start[i]:
            continue
        mag, direction, sigma = st
        place = _lib.DescPlacement(remap[i].data_ptr(), res.records.data_ptr(), ori_start[i], blk.octave, blk.scale)
        with _stage("descriptors"):
            describe(mag, direction, blk.oriented, sigma, params, out=res.descriptors, placement=place)
    if desc_dtype != out_dtype:
        res.descriptors = res.descriptors.to(out_dtype)
    return res


# ---------------------------------------------------------------------------------------------
# host <-> device pipeline for streams of image batches
# ---------------------------------------------------------------------------------------------
class BatchExtractor:
    """computeKeypointsAndDescriptors for a stream of equally shaped image batches that live on the HOST.

    Three CUDA streams: pinned input -> device (copy-in), the kernels (compute, = the caller's current stream),
    results -> pinned output (copy-out).  `submit` returns as soon as the work of a batch is queued; `collect`
    waits for its copy-out only, so the D2H of batch i overlaps the kernels of batch i+1.  A 1024x1024 image yields
    ~156 000 descriptors, so PCIe, not the kernels, bounds the end-to-end rate; the export is therefore compact:
    8-byte keypoint records (see unpack_records) and float16 descriptors by default (`descriptor_dtype`; the values
    lie in [0, 0.7], so the float16 rounding moves a unit descriptor by < 5e-4 in L2 -- float32 is available)."""

    def __init__(self, params: SiftParams, batch: int, height: int, width: int, max_descriptors: int, depth: int = 2,
                 upsample: bool = True, descriptor_dtype=torch.float16):
        dev = require_cuda()
        self.params, self.batch, self.h, self.w, self.depth, self.upsample = params, batch, height, width, depth, upsample
        if descriptor_dtype == torch.float16 and params.dtype != "mixed":
            descriptor_dtype = torch.float32
        self.desc_dtype = descriptor_dtype
        self.cap = int(max_descriptors)
        self.copy_in = torch.cuda.Stream(device=dev)
        self.copy_out = torch.cuda.Stream(device=dev)
        self.host_in = [torch.empty((batch, height, width), dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.dev_in = [torch.empty((batch, height, width), dtype=torch.float64, device=dev) for _ in range(depth)]
        self.host_desc = [torch.empty((self.cap, 128), dtype=descriptor_dtype).pin_memory() for _ in range(depth)]
        self.host_rec = [torch.empty((self.cap, 2), dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.host_cnt = [torch.empty(batch, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.slots = [None] * depth
        self.staged = [None] * depth   # H2D-done event of an input already on its way to the device
        self.consumed = [None] * depth  # event after the last kernel that read dev_in[slot]
        self.n_submitted = 0

    def stage(self, images, ticket: int | None = None):
        """Start the H2D copy of the batch that a later submit() (ticket = the next one by default) will process, so
        that it overlaps the kernels of the batch before it.  images=None: the pinned staging buffer is already filled."""
        ticket = self.n_submitted if ticket is None else ticket
        slot = ticket % self.depth
        if images is not None:
            src = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(images))
            self.host_in[slot].copy_(src)
        with torch.cuda.stream(self.copy_in):
            if self.consumed[slot] is not None:
                self.copy_in.wait_event(self.consumed[slot])
            self.dev_in[slot].copy_(self.host_in[slot], non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record(self.copy_in)
        self.staged[slot] = ev_in

    def submit(self, images, prefetch_next: bool = False) -> int:
        """images: [B,H,W] float64 numpy array / CPU tensor (copied into the pinned staging buffer) or None to
        reuse what the staging buffer of this slot already holds (or what stage() already sent).  prefetch_next:
        also start the H2D of the following ticket from ITS staging buffer before this batch's kernels are queued.
        Returns a ticket for collect().  Raises _lib.CapacityError (FC_ERR_CAPACITY, with the required count) when the
        batch yields more descriptors than max_descriptors -- before anything is copied."""
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
        base = upsample2(self.dev_in[slot], self.params) if self.upsample else _as_base_batch(self.dev_in[slot], self.params)
        self.consumed[slot] = torch.cuda.Event()
        self.consumed[slot].record(compute)
        res = detect_and_describe(base, self.params, out_dtype=self.desc_dtype)
        m = res.total
        if m > self.cap:
            raise _lib.CapacityError("BatchExtractor.submit (descriptors of the batch)", m, self.cap)
        ev_done = torch.cuda.Event()
        ev_done.record(compute)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(ev_done)
            if m:
                self.host_desc[slot][:m].copy_(res.descriptors, non_blocking=True)
                self.host_rec[slot][:m].copy_(res.records, non_blocking=True)
            self.host_cnt[slot].copy_(res.counts, non_blocking=True)
            ev_out = torch.cuda.Event()
            ev_out.record(self.copy_out)
        self.slots[slot] = (ev_out, m, res)  # keep the device tensors alive until the copy is done
        self.n_submitted += 1
        return ticket

    def collect(self, ticket: int, copy: bool = False, raw: bool = False):
        """raw=False -> list of (points [m,5] float64, descriptors [m,128]) per image, like the reference returns them;
        raw=True -> (records [M,2] int32, descriptors [M,128], counts [B] int32) exactly as they arrived (views into the
        pinned buffers, valid until the slot is reused, unless copy=True): unpack_records() turns records into 5-tuples."""
        slot = ticket % self.depth
        ev_out, m, _keep = self.slots[slot]
        ev_out.synchronize()
        self.slots[slot] = None
        cnt = self.host_cnt[slot].numpy()
        rec = self.host_rec[slot][:m].numpy()
        desc = self.host_desc[slot][:m].numpy()
        if raw:
            return (rec.copy(), desc.copy(), cnt.copy()) if copy else (rec, desc, cnt)
        pts = unpack_records(rec)
        out, lo = [], 0
        for b in range(self.batch):
            hi = lo + int(cnt[b])
            out.append((pts[lo:hi], desc[lo:hi].copy() if copy else desc[lo:hi]))
            lo = hi
        return out

    def bytes_per_batch(self, m: int):
        """(host->device, device->host) bytes of a batch with m descriptors"""
        return self.batch * self.h * self.w * 8, m * (128 * self.host_desc[0].element_size() + 8) + self.batch * 4
