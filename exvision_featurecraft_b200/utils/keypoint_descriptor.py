"""GPU-backed twins of FeatureCraft-PyQt-App/utils/keypoint_descriptor.py (same names and arguments).

numpy / Python lists in and out, like the reference; the arithmetic runs in libfeaturecraft_b200.  What stays
on the host is bookkeeping the reference also does in Python (bool masks, tuple lists) and the tiny parameter
tables.  Drawing (cv2.drawMatches, :351-364) is not part of the hot path: ``match`` returns the `good` list.
"""
from __future__ import annotations

import ctypes
from math import cos, sin

import numpy as np
import torch

from .. import _lib, matching, sift
from ..device import ptr, stream_ptr, to_device
from . import SIFT_scale_space as _ss
from .SIFT_scale_space import gaussian_filter_kernel, generate_octaves_pyramid  # noqa: F401  (KD:8 star-imports them)


def _gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage's Gaussian weights: radius int(truncate * sigma + 0.5), exp(-x^2 / 2 sigma^2) normalised to sum 1."""
    r = int(truncate * float(sigma) + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return w / w.sum(), r


def resize_device(img: torch.Tensor, newshape, anti_aliasing=True) -> torch.Tensor:
    """skimage.transform.resize(img, newshape, anti_aliasing=...) for a float64 [H,W] or [H,W,C] device tensor, as the
    scipy.ndimage calls scikit-image 0.24 makes: gaussian_filter(sigma = max(0, (in/out - 1)/2) per axis, 'mirror')
    then zoom(order=1, mode='mirror', grid_mode=True)."""
    t = img if img.dim() == 3 else img.unsqueeze(-1)
    t = t.contiguous()
    h, w, c = t.shape
    h2, w2 = int(newshape[0]), int(newshape[1])
    if anti_aliasing:
        for axis, (n_in, n_out) in enumerate(((h, h2), (w, w2))):
            sigma = max(0.0, (n_in / n_out - 1.0) / 2.0)
            if sigma > 1e-15:  # gaussian_filter skips axes with a vanishing sigma
                taps, r = _gaussian_taps(sigma)
                d_taps = to_device(taps)
                out = torch.empty_like(t)
                _lib.call("fc_blur_axis_f64", ptr(t), 1, h, w, c, axis, ptr(d_taps), r, ptr(out), stream_ptr())
                t = out
    out = torch.empty((h2, w2, c), dtype=torch.float64, device=t.device)
    _lib.call("fc_zoom_linear_f64", ptr(t), 1, h, w, c, h2, w2, ptr(out), stream_ptr())
    return out if img.dim() == 3 else out[..., 0]


def _as_float_image(img) -> torch.Tensor:
    """skimage.util.img_as_float (what resize applies first): uint8 -> value / 255, floats unchanged."""
    a = np.asarray(img) if not isinstance(img, torch.Tensor) else img
    if a.dtype in (np.uint8, torch.uint8):
        t8 = to_device(a, torch.uint8)
        out = torch.empty(t8.shape, dtype=torch.float64, device=t8.device)
        _lib.call("fc_u8_to_unit_f64", ptr(t8), ctypes.c_int64(t8.numel()), ptr(out), stream_ptr())
        return out
    return to_device(a, torch.float64)


def sift_resize(img, ratio=None):
    """:11-33 -- resize to ~1 MP (or by `ratio`) with skimage.transform.resize(anti_aliasing=True); returns
    (resized image, ratio).  The resize runs on the device as the scipy.ndimage calls scikit-image makes (the library
    itself is not in this image: pinned against scipy)."""
    shape = np.shape(img) if not isinstance(img, torch.Tensor) else tuple(img.shape)
    ratio = ratio if ratio is not None else np.sqrt((1024 * 1024) / np.prod(shape[:2]))
    newshape = list(map(lambda d: int(round(d * ratio)), shape[:2]))
    out = resize_device(_as_float_image(img), newshape, anti_aliasing=True)
    return out.cpu().numpy(), ratio


def convert_to_grayscale(image):
    """:36-39 -- float weighted sum, no cast (three multiply-adds per pixel, done where the image lives)."""
    if isinstance(image, torch.Tensor):
        return grayscale_device(image)
    image = np.asarray(image)
    if image.ndim == 3:
        return np.dot(image[..., :3], [0.2989, 0.5870, 0.1140])
    return image


def grayscale_device(image: torch.Tensor) -> torch.Tensor:
    """convert_to_grayscale for a device tensor [H,W] / [H,W,C>=3] (float64 in, float64 out)."""
    if image.dim() != 3:
        return image
    t = image.to(torch.float64).contiguous()
    h, w, c = t.shape
    out = torch.empty((h, w), dtype=torch.float64, device=t.device)
    _lib.call("fc_rgb_to_gray_f64", ptr(t), 1, h, w, c, ptr(out), stream_ptr())
    return out


def represent_keypoints(keypoints, DoG):
    """:42-69 -- bool image per (octave, dog_idx) with dog_idx in range(1, len(DoG) - 1); len(DoG) is the
    number of OCTAVES (reference quirk, kept)."""
    out = []
    for o, kps in enumerate(keypoints):
        arr = np.asarray(kps).reshape(-1, 3).astype(np.int64) if len(kps) else np.zeros((0, 3), np.int64)
        per = []
        for dog_idx in range(1, len(DoG) - 1):
            m = np.zeros(np.asarray(DoG[o][0]).shape, dtype=bool)
            sel = arr[arr[:, 2] == dog_idx]
            m[sel[:, 0], sel[:, 1]] = True
            per.append(m)
        out.append(per)
    return out


def _params(s, sigma_base=1.6, n_octaves=4):
    return sift.SiftParams(n_octaves=n_octaves, s_value=s, sigma_base=sigma_base, dtype=_ss.stage_dtype(), variant=_ss.VARIANT,
                           numpy_legacy_promotion=_ss.NUMPY_LEGACY_PROMOTION)


def _level_as_octave(img, p):
    t = sift._as_base_batch(np.asarray(img, dtype=np.float64), p)
    return sift.Octave(t.unsqueeze(1).contiguous(), None, t.shape[1], t.shape[2])


def sift_gradient(img):
    """:72-79 -> (gx, gy, magnitude, direction)."""
    p = _params(2)
    octv = _level_as_octave(img, p)
    mag, direction, _ = sift.gradient_field(octv, 0, p)
    g = octv.gauss[0, 0]
    gx = torch.empty_like(g)
    gy = torch.empty_like(g)
    gx[:, 1:-1] = g[:, :-2] - g[:, 2:]
    gx[:, 0] = g[:, 0] - g[:, min(1, g.shape[1] - 1)]
    gx[:, -1] = g[:, max(g.shape[1] - 2, 0)] - g[:, -1]
    gy[1:-1] = g[:-2] - g[2:]
    gy[0] = g[0] - g[min(1, g.shape[0] - 1)]
    gy[-1] = g[max(g.shape[0] - 2, 0)] - g[-1]
    f = lambda t: t.to(torch.float64).cpu().numpy()  # noqa: E731
    return f(gx), f(gy), f(mag[0]), f(direction[0])


def padded_slice(img, sl):
    """:82-113 -- window with zeros outside the image (host bookkeeping helper)."""
    img = np.asarray(img)
    out = np.zeros((sl[1] - sl[0], sl[3] - sl[2]), dtype=img.dtype)
    r0, r1 = max(sl[0], 0), min(sl[1], img.shape[0])
    c0, c1 = max(sl[2], 0), min(sl[3], img.shape[1])
    if r1 > r0 and c1 > c0:
        out[r0 - sl[0]:r1 - sl[0], c0 - sl[2]:c1 - sl[2]] = img[r0:r1, c0:c1]
    return out


def dog_keypoints_orientations(img_gaussians, keypoints, sigma_base, num_bins=36, s=2):
    """:116-172 -> [(i, j, octave, scale, angle)]."""
    if num_bins != 36:
        raise _lib.FeatureCraftError("the device path implements the reference's 36 orientation bins")
    p = _params(s, sigma_base)
    kps = []
    for o in range(len(img_gaussians)):
        for idx, mask in enumerate(keypoints[o]):
            scale_idx = idx + 1
            gimg = img_gaussians[o][scale_idx]  # IndexError for n_octaves > s + 4, like the reference
            ij = np.argwhere(np.asarray(mask))
            if ij.shape[0] == 0:
                continue
            sigma = sift.keypoint_sigma(sigma_base, o, scale_idx, s)
            octv = _level_as_octave(gimg, p)
            mag, _, obin = sift.gradient_field(octv, 0, p)
            rec = np.zeros((ij.shape[0], 3), dtype=np.int32)
            rec[:, 1:] = ij
            ori = sift.orient(mag, obin, to_device(rec), sigma, p).cpu().numpy()
            kps.extend((int(r), int(c), o, scale_idx, (b + 0.5) * (360.0 / num_bins) % 360) for _, r, c, b in ori.tolist())
    return kps


def rotated_subimage(image, center, theta, width, height):
    """:175-212 -- nearest-neighbour rotated window (cv2.warpAffine's fixed-point rule)."""
    t = theta * (3.14159 / 180)
    img = to_device(np.asarray(image, dtype=np.float64), torch.float64)
    out = torch.empty((height, width), dtype=torch.float64, device=img.device)
    _lib.call("fc_rotated_window", ptr(img), img.shape[0], img.shape[1], float(center[0]), float(center[1]), cos(t), sin(t),
              int(width), int(height), ptr(out), stream_ptr())
    return out.cpu().numpy()


def get_gaussian_mask(sigma, filter_size):
    """:215-227"""
    return sift.gaussian_mask16(sigma, filter_size)


def extract_sift_descriptors(img_gaussians, keypoints, base_sigma, num_bins=8, s=2):
    """:230-289 -> (points, descriptors)."""
    if num_bins != 8:
        raise _lib.FeatureCraftError("the device path implements the reference's 8 descriptor bins")
    p = _params(s, base_sigma)
    descriptors = [None] * len(keypoints)
    groups = {}
    for n, (i, j, o, sc, theta) in enumerate(keypoints):
        b = (theta - 5.0) / 10.0
        if b != int(b) or not 0 <= b < 36:
            raise _lib.FeatureCraftError("orientation %r is not one of the 36 bin centres the reference emits" % (theta,))
        groups.setdefault((o, sc), []).append((n, int(i), int(j), int(b)))
    for (o, sc), items in groups.items():
        sigma = sift.keypoint_sigma(base_sigma, o, sc, s)
        octv = _level_as_octave(img_gaussians[o][sc], p)
        mag, direction, _ = sift.gradient_field(octv, 0, p)
        rec = np.zeros((len(items), 4), dtype=np.int32)
        rec[:, 1:] = [(i, j, b) for _, i, j, b in items]
        d = sift.describe(mag, direction, to_device(rec), sigma, p, torch.float64).cpu().numpy()
        for row, (n, _, _, _) in zip(d, items):
            descriptors[n] = row
    return [tuple(k) for k in keypoints], descriptors


def extract_device(image, n_octaves, s_value, sigma_base, constract_th, r_ratio, out_dtype=torch.float32) -> "sift.SiftResult":
    """computeKeypointsAndDescriptors with everything left on the device (records + descriptors of one image)."""
    gray = convert_to_grayscale(image)
    p = sift.SiftParams(n_octaves, s_value, sigma_base, constract_th, r_ratio, _ss.VARIANT, _ss.DTYPE,
                        _ss.NUMPY_LEGACY_PROMOTION)
    base = sift.upsample2(gray if isinstance(gray, torch.Tensor) else np.asarray(gray, dtype=np.float64), p)
    return sift.detect_and_describe(base, p, out_dtype=out_dtype)


def computeKeypointsAndDescriptors(image, n_octaves, s_value, sigma_base, constract_th, r_ratio):
    """:292-311 -- gray, x2 up-sample, pyramid, orientations, descriptors; one device-resident pass."""
    res = extract_device(image, n_octaves, s_value, sigma_base, constract_th, r_ratio, out_dtype=torch.float64)
    pts, desc = res.assemble(1)[0]
    points = [(int(a), int(b), int(c), int(d), float(e)) for a, b, c, d, e in pts.tolist()]
    return points, [row for row in desc]


def kp_list_2_xy(kp_list):
    """The constructor arguments of :314-327 as plain tuples (x, y, size, angle)."""
    return [(kp[1] * (2 ** (kp[2] - 1)), kp[0] * (2 ** (kp[2] - 1)), kp[3], kp[4]) for kp in kp_list]


def kp_list_2_opencv_kp_list(kp_list):
    """:314-327 -- cv2.KeyPoint(x = j * 2^(octave-1), y = i * 2^(octave-1), size = scale, angle) per keypoint.
    (Host bookkeeping like in the reference; cv2 is only needed by this export and by the match canvas.)"""
    import cv2

    return [cv2.KeyPoint(x=float(x), y=float(y), size=float(s), angle=float(a)) for x, y, s, a in kp_list_2_xy(kp_list)]


def match_good(desc_a, desc_b, tuning_distance=0.3):
    """:333-349 -- the `good` list of match(): knnMatch(k=2) + ratio test as (queryIdx, trainIdx, distance) tuples.
    desc_a / desc_b: lists / arrays like the reference's, or device tensors (then nothing crosses PCIe but the list)."""
    pairs, dist = matching.match_indices(desc_a, desc_b, tuning_distance)
    return [(int(q), int(t), float(d)) for (q, t), d in zip(pairs.tolist(), dist.tolist())]


def match(img_a, pts_a, desc_a, img_b, pts_b, desc_b, tuning_distance=0.3):
    """:330-366 -- knnMatch(k=2) + ratio test on the device, then the reference's canvas: both images side by side with
    the good matches drawn by cv2.drawMatches (host; drawing is not part of the hot path).  Returns img_match."""
    import cv2

    good = match_good(desc_a, desc_b, tuning_distance)
    img_a, img_b = tuple(map(lambda i: np.uint8(np.asarray(i) * 255), [img_a, img_b]))
    kp_a, kp_b = kp_list_2_opencv_kp_list(pts_a), kp_list_2_opencv_kp_list(pts_b)
    dm = [cv2.DMatch(_queryIdx=q, _trainIdx=t, _distance=d) for q, t, d in good]
    img_match = np.empty((max(img_a.shape[0], img_b.shape[0]), img_a.shape[1] + img_b.shape[1], 3), dtype=np.uint8)
    cv2.drawMatches(img_a, kp_a, img_b, kp_b, dm, outImg=img_match, flags=cv2.DrawMatchesFlags_NOT_DRAW_SINGLE_POINTS)
    return img_match
