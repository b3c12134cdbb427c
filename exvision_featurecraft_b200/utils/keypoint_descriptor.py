"""GPU-backed twins of FeatureCraft-PyQt-App/utils/keypoint_descriptor.py (same names and arguments).

numpy / Python lists in and out, like the reference; the arithmetic runs in libfeaturecraft_b200.  What stays
on the host is bookkeeping the reference also does in Python (bool masks, tuple lists) and the tiny parameter
tables.  Drawing (cv2.drawMatches, :351-364) is not part of the hot path: ``match`` returns the `good` list.
"""
from __future__ import annotations

from math import cos, sin

import numpy as np
import torch

from .. import _lib, matching, sift
from ..device import ptr, stream_ptr, to_device
from . import SIFT_scale_space as _ss
from .SIFT_scale_space import gaussian_filter_kernel, generate_octaves_pyramid  # noqa: F401  (KD:8 star-imports them)


def sift_resize(img, ratio=None):
    """:11-33 needs skimage.transform.resize (anti-aliased), which is the step BEFORE the hot path and
    cannot be pinned in this image (no scikit-image): SURVEY 8(f) item 1."""
    raise NotImplementedError("sift_resize (skimage.transform.resize) is outside the B200 hot path; resize on the host")


def convert_to_grayscale(image):
    """:36-39 -- float weighted sum, no cast (three multiply-adds per pixel, done where the image lives)."""
    image = np.asarray(image)
    if image.ndim == 3:
        return np.dot(image[..., :3], [0.2989, 0.5870, 0.1140])
    return image


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
    return sift.SiftParams(n_octaves=n_octaves, s_value=s, sigma_base=sigma_base, dtype=_ss.stage_dtype(), variant=_ss.VARIANT)


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


def computeKeypointsAndDescriptors(image, n_octaves, s_value, sigma_base, constract_th, r_ratio):
    """:292-311 -- gray, x2 up-sample, pyramid, orientations, descriptors; one device-resident pass."""
    gray = np.asarray(convert_to_grayscale(image), dtype=np.float64)
    p = sift.SiftParams(n_octaves, s_value, sigma_base, constract_th, r_ratio, _ss.VARIANT, _ss.DTYPE)
    base = sift.upsample2(gray, p)
    res = sift.detect_and_describe(base, p, out_dtype=torch.float64)
    pts, desc = res.assemble(1)[0]
    points = [(int(a), int(b), int(c), int(d), float(e)) for a, b, c, d, e in pts.tolist()]
    return points, [row for row in desc]


def kp_list_2_opencv_kp_list(kp_list):
    """:314-327 without the cv2 dependency: (x, y, size, angle) tuples in cv2.KeyPoint's argument order."""
    return [(kp[1] * (2 ** (kp[2] - 1)), kp[0] * (2 ** (kp[2] - 1)), kp[3], kp[4]) for kp in kp_list]


def match(img_a, pts_a, desc_a, img_b, pts_b, desc_b, tuning_distance=0.3):
    """:330-349 -- knnMatch(k=2) + ratio test.  Returns the `good` list as (queryIdx, trainIdx, distance)
    tuples; the reference then draws them (cv2.drawMatches, :351-364), which is left to the caller."""
    pairs, dist = matching.match_indices(desc_a, desc_b, tuning_distance)
    return [(int(q), int(t), float(d)) for (q, t), d in zip(pairs.tolist(), dist.tolist())]
