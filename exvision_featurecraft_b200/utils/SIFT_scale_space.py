"""GPU-backed twins of FeatureCraft-PyQt-App/utils/SIFT_scale_space.py (same names and arguments).

numpy in, numpy / Python lists out -- like the reference.  ``DTYPE`` selects the arithmetic (see
``sift.SiftParams.dtype``): "mixed" (default) = the certified pipeline for the end-to-end drivers
(computeKeypointsAndDescriptors, apply_sift), whose keypoint lists equal the float64 path's; the stage-level
functions of this module return the reference's float64 images and therefore always run in float64.  ``VARIANT`` selects the keypoint filter: "app" (this module in the
reference keeps every extremum, :89-105) or "script" (implementation_without_ui/SIFT/SIFT.py:202-215).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, sift
from ..device import ptr, stream_ptr, to_device

DTYPE = "mixed"
VARIANT = "app"
NUMPY_LEGACY_PROMOTION = False  # orientation cut the numpy-1.26 way (the reference's pinned numpy); see sift.SiftParams


def set_dtype(dtype: str) -> None:
    global DTYPE
    assert dtype in ("mixed", "f64")
    DTYPE = dtype


def stage_dtype() -> str:
    """arithmetic of the stage-level mirrors: they hand float64 images back to the caller like the reference does"""
    return "f64"


def set_numpy_legacy_promotion(on: bool) -> None:
    global NUMPY_LEGACY_PROMOTION
    NUMPY_LEGACY_PROMOTION = bool(on)


def set_variant(variant: str) -> None:
    global VARIANT
    assert variant in ("app", "script")
    VARIANT = variant


def gaussian_filter_kernel(sigma, kernel_size=None):
    """:5-30 -- a (2r+1)^2 table of a few hundred entries, built on the host exactly like the reference."""
    return sift.gaussian_window_table(sigma)


def generateGaussianKernels(sigma, s):
    """:33-57"""
    return [gaussian_filter_kernel(x) for x in sift.sigma_levels(sigma, s)]


def _taps_from_kernels(gaussian_kernels):
    taps = []
    for k in gaussian_kernels:
        k = np.asarray(k, dtype=np.float64)
        c = k.shape[0] // 2
        if k.ndim != 2 or k.shape[0] != k.shape[1] or k.shape[0] % 2 == 0:
            raise ValueError("Gaussian kernels must be square with odd size")
        t = k[c, :] / np.sqrt(k[c, c])
        if not np.allclose(np.outer(t, t), k, rtol=1e-12, atol=1e-300):
            raise _lib.FeatureCraftError("only separable (rank-1) kernels are supported on the device path")
        taps.append(t)
    return taps


def _extrema_to_list(mask_k: torch.Tensor, width: int, k: int):
    kp = sift.keypoints_from_mask(mask_k, width).cpu().numpy()
    if kp.shape[0] == 0:
        return np.array([])
    out = np.empty((kp.shape[0], 3), dtype=np.int64)
    out[:, 0], out[:, 1], out[:, 2] = kp[:, 1], kp[:, 2], k
    return out


def get_keypoints(DOG_octave, k, contrast_th, ratio_th, DoG_full_array):
    """:60-106 (app) / SIFT.py:172-222 (script): 3x3x3 extrema of the middle of three DoG images."""
    fc_dtype, tdt = _lib.FC_F64, torch.float64
    if VARIANT == "script":
        # the reference passes an (H, W, n) array (:155); a list of 2-D images is accepted too
        if isinstance(DoG_full_array, np.ndarray) and DoG_full_array.ndim == 3:
            stack = np.moveaxis(DoG_full_array, 2, 0)
        else:
            stack = np.stack([np.asarray(d) for d in DoG_full_array])
        if k >= 3:
            raise IndexError("index %d is out of bounds for axis 2 with size 3" % k)
        sel = k
    else:
        stack = np.stack([np.asarray(d) for d in DOG_octave])
        sel = 1
    d = to_device(np.ascontiguousarray(stack), tdt).unsqueeze(0).contiguous()
    _, n, h, w = d.shape
    ext = torch.empty((n - 2, 1, h, (w + 31) // 32), dtype=torch.int32, device=d.device)
    _lib.call("fc_dog_extrema", ptr(d), fc_dtype, 1, h, w, n, 1 if VARIANT == "script" else 0, float(contrast_th),
              float(ratio_th), ptr(ext), stream_ptr())
    return _extrema_to_list(ext[sel - 1], w, k)


def localize_keypoint(D, x, y, s):
    """implementation_without_ui/SIFT/SIFT.py:225-269: the 2x2 spatial Hessian of DoG level s at (y, x) -> (H, x, y, s).

    Nine array elements and six float64 operations on the host: on the device this arithmetic is fused into the
    extremum kernels (script filter, fc_dog_extrema with FC_FILTER_SCRIPT); the function exists so that code written
    against the script's API keeps working.  The operation order is the reference's (:248-259)."""
    D = np.asarray(D).astype(np.float64)
    dxx = D[y, x + 1, s] - 2 * D[y, x, s] + D[y, x - 1, s]
    dxy = ((D[y + 1, x + 1, s] - D[y + 1, x - 1, s]) - (D[y - 1, x + 1, s] - D[y - 1, x - 1, s])) / 4
    dyy = D[y + 1, x, s] - 2 * D[y, x, s] + D[y - 1, x, s]
    return np.array([[dxx, dxy], [dxy, dyy]]), x, y, s


def generate_gaussian_images_in_octave(image, gaussian_kernels, contrast_th, ratio_th, octave_index):
    """:109-160 -> (gaussian images, DoG images, keypoints [i, j, k])."""
    taps = _taps_from_kernels(gaussian_kernels)
    p = sift.SiftParams(n_octaves=1, s_value=len(taps) - 3, contrast_th=contrast_th, ratio_th=ratio_th, variant=VARIANT,
                        dtype=stage_dtype())
    base = sift._as_base_batch(np.asarray(image, dtype=np.float64), p)
    octv = sift.gauss_octave(base, p, octave_index, True, taps=taps)
    g = octv.gauss[0].to(torch.float64).cpu().numpy()
    gaussians = [g[i] for i in range(g.shape[0])]
    dogs = [gaussians[i + 1] - gaussians[i] for i in range(len(gaussians) - 1)]
    keypoints = []
    for k in range(1, len(taps) - 2):
        keypoints.extend(_extrema_to_list(octv.extrema[k - 1], octv.width, k))
    return gaussians, dogs, keypoints


def generate_octaves_pyramid(img, num_octaves=4, s_value=2, sigma=1.6, contrast_th=0.03, ratio_th=10):
    """:163-210 -> (pyramid, DoG pyramid, keypoints per octave)."""
    kernels = generateGaussianKernels(sigma, s_value)
    pyramid, dog_pyramid, keypoints = [], [], []
    img = np.asarray(img, dtype=np.float64)
    for o in range(num_octaves):
        g, d, kp = generate_gaussian_images_in_octave(img, kernels, contrast_th, ratio_th, o)
        pyramid.append(g)
        dog_pyramid.append(d)
        keypoints.append(kp)
        img = g[-3][::2, ::2]
    return pyramid, dog_pyramid, keypoints
