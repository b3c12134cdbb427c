"""Reference-shaped corner drivers: the compute methods of BackendClass
(FeatureCraft-PyQt-App/FeatureCraft_Backend.py:306-452) and the notebook's harris_detection,
with the same names, arguments and return values -- minus the Qt calls."""
from __future__ import annotations

import numpy as np

from . import corners
from .utils.harris_utils import convert_to_gray


class FeatureCraftBackend:
    """Holds the two caches the reference keeps on `self` (harris_response_operator, eigenvalues)
    as DEVICE tensors so that the threshold slider re-runs only the tiny compaction kernels."""

    def __init__(self):
        self.harris_response_operator = None
        self.eigenvalues = None
        self._harris_dev = None
        self._eig_dev = None

    # -- FeatureCraft_Backend.py:306-380 ------------------------------------------------------------
    def apply_harris_detector_vectorized(self, img_RGB, window_size=5, k=0.04, threshold=10000):
        if not np.all(img_RGB != None):  # noqa: E711  (the reference's own guard)
            return None
        img_RGB = np.asarray(img_RGB)
        gray = convert_to_gray(img_RGB)
        res = corners.harris(gray, window_size, k, threshold)
        self._harris_dev = res.response[0]
        self.harris_response_operator = self._harris_dev.cpu().numpy()
        rc = res.coords.cpu().numpy()
        vals = res.values.cpu().numpy()
        output_img = img_RGB.copy()
        output_img[rc[:, 0], rc[:, 1]] = (0, 0, 255)  # raises for 2-D input exactly like the reference
        return list(zip(rc[:, 1], rc[:, 0], vals)), output_img

    # -- FeatureCraft_Backend.py:382-452 ------------------------------------------------------------
    def apply_lambda_minus_vectorized(self, img_RGB, window_size=5, threshold_percentage=0.01):
        """Returns (rows, cols) like np.where -- the reference only draws them (cv2.circle) and
        returns None; the circles are left to the caller."""
        img_RGB = np.asarray(img_RGB)
        gray = convert_to_gray(img_RGB)
        res = corners.lambda_minus(gray, threshold_percentage)  # window_size ignored, as in the reference
        self._eig_dev = res.response[0]
        self.eigenvalues = self._eig_dev.cpu().numpy()
        rc = res.coords.cpu().numpy()
        return rc[:, 0].astype(np.int64), rc[:, 1].astype(np.int64)

    # -- FeatureCraft_Backend.py:261-292 ------------------------------------------------------------
    def on_changing_threshold(self, threshold, img_RGB=None, operator=0):
        """operator 0: Harris R > threshold; operator 1: eigenvalues > threshold where the app passes
        slider/10000 as an ABSOLUTE value.  Returns np.argwhere-ordered (row, col)."""
        dev = self._harris_dev if operator == 0 else self._eig_dev
        if dev is None:
            return None
        coords, _, _ = corners.rethreshold(dev, threshold)
        return coords.cpu().numpy().astype(np.int64)


_default = FeatureCraftBackend()


def apply_harris_detector_vectorized(img_RGB, window_size=5, k=0.04, threshold=10000):
    return _default.apply_harris_detector_vectorized(img_RGB, window_size, k, threshold)


def apply_lambda_minus_vectorized(img_RGB, window_size=5, threshold_percentage=0.01):
    return _default.apply_lambda_minus_vectorized(img_RGB, window_size, threshold_percentage)


def harris_detection(img, window_size, k, threshold):
    """implementation_without_ui/Corner Detection/harris.ipynb:123-206.  `img` is a gray uint8 array
    or an RGB array (the notebook reads a file and converts with cv2; decoding is out of scope).
    Returns (corner_list [[x, y, r], ...], output_img) with 3x3 red squares painted like the notebook."""
    img = np.asarray(img)
    gray = convert_to_gray(img) if img.ndim == 3 else img.astype(np.uint8)
    res = corners.notebook_harris(gray, window_size, k, threshold)
    rc = res.coords.cpu().numpy()
    vals = res.values.cpu().numpy()
    out = np.repeat(gray[:, :, None], 3, axis=2)
    h, w = gray.shape
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            yy = np.clip(rc[:, 0] + dy, 0, h - 1)
            xx = np.clip(rc[:, 1] + dx, 0, w - 1)
            out[yy, xx] = (0, 0, 255)
    return [[int(x), int(y), r] for (y, x), r in zip(rc.tolist(), vals.tolist())], out


def apply_sift_matches(main_image, template, n_octaves=4, s_value=2, sigma_base=1.6, contrast_th=0.03, r_ratio=10,
                       tuning_factor=0.3, resize=True, return_keypoints=True):
    """The arithmetic of apply_sift, device resident from the first resize to the ratio test: both images are resized
    (sift_resize), described (computeKeypointsAndDescriptors) and matched without their descriptors ever leaving HBM
    -- what crosses PCIe is the two input images and the `good` list.

    Returns (good [(queryIdx, trainIdx, distance)], points_main, points_template); the point lists are the reference's
    5-tuples (None with return_keypoints=False)."""
    import torch

    from . import matching, sift
    from .utils import keypoint_descriptor as KD

    if resize:
        shape = np.shape(main_image)
        ratio = np.sqrt((1024 * 1024) / np.prod(shape[:2]))
        images = []
        for img in (main_image, template):
            t = KD._as_float_image(img)
            newshape = [int(round(d * ratio)) for d in t.shape[:2]]
            images.append(KD.resize_device(t, newshape, anti_aliasing=True))
    else:
        images = [KD._as_float_image(main_image), KD._as_float_image(template)]
    ra, rb = (KD.extract_device(img, n_octaves, s_value, sigma_base, contrast_th, r_ratio, out_dtype=torch.float32) for img in images)
    if rb.total < 2:
        raise ValueError("not enough values to unpack (expected 2, got %d)" % rb.total)  # `for m, n in matches`, KD:345
    idx, dist, good = matching.knn2(ra.descriptors, rb.descriptors, tuning_factor, "tensor")
    q = torch.nonzero(good).flatten()
    triples = torch.stack([q.to(torch.float64), idx[q, 0].to(torch.float64), dist[q, 0].to(torch.float64)], 1).cpu().numpy()
    out = [(int(a), int(b), float(c)) for a, b, c in triples.tolist()]
    if not return_keypoints:
        return out, None, None
    pa = [tuple(r) for r in sift.unpack_records(ra.records.cpu().numpy()).tolist()]
    pb = [tuple(r) for r in sift.unpack_records(rb.records.cpu().numpy()).tolist()]
    return out, pa, pb


def apply_sift(main_image, template, n_octaves, s_value, sigma_base, contrast_th, r_ratio, tuning_factor):
    """BackendClass.apply_sift (FeatureCraft_Backend.py:480-523) / SIFT.py:683-707 without the Qt calls, same arguments
    and return value: sift_resize of both images (ratio from the main image), keypoints + descriptors of both, match()
    -> the match canvas (cv2.drawMatches on the host, utils/keypoint_descriptor.py:351-364)."""
    from .utils import keypoint_descriptor as KD

    main_image, ratio = KD.sift_resize(main_image)
    template, _ = KD.sift_resize(template, ratio)
    img_kp, img_des = KD.computeKeypointsAndDescriptors(main_image, n_octaves, s_value, sigma_base, contrast_th, r_ratio)
    template_kp, template_des = KD.computeKeypointsAndDescriptors(template, n_octaves, s_value, sigma_base, contrast_th, r_ratio)
    return KD.match(main_image, img_kp, img_des, template, template_kp, template_des, tuning_factor)
