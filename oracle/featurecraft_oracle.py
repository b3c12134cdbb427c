"""CPU oracle for the FeatureCraft feature-extraction hot path.  TEST INFRASTRUCTURE ONLY.

A numpy/scipy restatement of what the reference computes, stage by stage, each function citing the
reference file:line it follows (paths relative to the reference root; ``APP`` =
``FeatureCraft-PyQt-App``).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
(``exvision_featurecraft_b200``) never does and has no CPU fallback.

Pinning: ``oracle/make_golden.py`` runs the UNMODIFIED reference in place (``oracle/ref_loader.py``)
on seeded inputs, asserts this restatement reproduces it (bit-exact for every integer / index
output, <=1e-12 for float maps) and stores the reference's outputs under ``tests/golden/``.
The reference itself ships no tests or golden vectors, so those fixtures,
produced by the reference's own code, are the pin.  Two third-party calls cannot be exercised in
this image and are restated from their published algorithm instead:

* ``cv2.warpAffine(INTER_NEAREST|WARP_INVERSE_MAP)`` (opencv-python==4.10.0.84): AB_BITS=10
  fixed-point coordinate rounding, see ``nn_rotated_window``.
* ``skimage.transform.rescale/resize`` (scikit-image==0.24.0): restated as the scipy.ndimage calls
  scikit-image makes (``sift_resize``, ``rescale2``) but PARITY UNPINNED -- no scikit-image here to run
  them against; the pinned parity boundary starts at the up-sampled base image.

Third-party arithmetic that IS available here and that the oracle calls exactly as the reference
does: ``scipy.signal.convolve2d`` (reference pins scipy==1.14.0), numpy ufuncs.  ``np.gradient`` and
``np.linalg.eigvalsh`` (LAPACK dsyevd -> dsterf -> dlae2 for 2x2) are restated explicitly because
the CUDA kernels must reproduce their rounding sequence.
"""
from __future__ import annotations

from math import cos, sin

import numpy as np
from scipy.signal import convolve2d as _scipy_convolve2d

# ======================================================================================
# Corner detectors
# ======================================================================================

#: 5x5 "Sobel-like" gradient kernel, APP/FeatureCraft_Backend.py:405-413 (same in harris.ipynb).
K_X = np.array(
    [[-1, -2, 0, 2, 1], [-2, -3, 0, 3, 2], [-3, -5, 0, 5, 3], [-2, -3, 0, 3, 2], [-1, -2, 0, 2, 1]],
    dtype=np.float64,
)


def convert_to_gray(img):
    """APP/utils/harris_utils.py:4-9 -- weighted channel sum then TRUNCATING cast to uint8."""
    img = np.asarray(img)
    if img.ndim == 3:
        return np.dot(img[..., :3], [0.2989, 0.5870, 0.1140]).astype(np.uint8)
    return img.astype(np.uint8)


def padding_matrix(matrix, width, height, pad_size):
    """APP/utils/harris_utils.py:12-30 -- float64 zero frame of pad_size around the matrix."""
    out = np.zeros((height + 2 * pad_size, width + 2 * pad_size))
    out[pad_size:pad_size + height, pad_size:pad_size + width] = matrix
    return out


def convolve2d_optimized(input_matrix, convolution_kernel, mode="same"):
    """APP/utils/harris_utils.py:33-80 -- zero-padded CORRELATION (no flip), `mode` ignored,
    kernel assumed square and odd (kernel_size = shape[0]).  The reduction is numpy's sum over the
    trailing (k, k) axes of a materialised (H, W, k, k) product, restated with a strided view."""
    input_matrix = np.asarray(input_matrix)
    kernel = np.asarray(convolution_kernel)
    h, w = input_matrix.shape
    ks = kernel.shape[0]
    pad = ks // 2
    if kernel.shape != (2 * pad + 1, 2 * pad + 1):
        # the reference's reshape(H, W, ks, ks) fails for even / non-square kernels
        raise ValueError("cannot reshape array: kernel must be square with odd size")
    padded = padding_matrix(input_matrix, w, h, pad)
    win = np.ascontiguousarray(np.lib.stride_tricks.sliding_window_view(padded, (ks, ks)))
    return np.sum(win * kernel, axis=(2, 3))


def _np_gradient_2d(f):
    """numpy.gradient(f) for a 2-D array with unit spacing, edge_order=1 (numpy function_base):
    interior (f[i+1]-f[i-1])/2, first/last one-sided differences; axis 0 first, then axis 1."""
    f = np.asarray(f, dtype=np.float64)
    if f.shape[0] < 2 or f.shape[1] < 2:
        raise ValueError(
            "Shape of array too small to calculate a numerical gradient, at least (edge_order + 1) elements are required."
        )
    g0 = np.empty_like(f)
    g0[1:-1] = (f[2:] - f[:-2]) / 2.0
    g0[0] = f[1] - f[0]
    g0[-1] = f[-1] - f[-2]
    g1 = np.empty_like(f)
    g1[:, 1:-1] = (f[:, 2:] - f[:, :-2]) / 2.0
    g1[:, 0] = f[:, 1] - f[:, 0]
    g1[:, -1] = f[:, -1] - f[:, -2]
    return g0, g1


def _box_sum_zero_pad(p, radius):
    """Sum over the (2r+1)^2 window with zeros outside the image -- what convolve2d_optimized with
    a ones window computes; exact (order-independent) on the integer-valued data of this path."""
    h, w = p.shape
    padded = np.zeros((h + 2 * radius, w + 2 * radius), dtype=np.float64)
    padded[radius:radius + h, radius:radius + w] = p
    out = np.zeros((h, w), dtype=np.float64)
    for di in range(2 * radius + 1):
        for dj in range(2 * radius + 1):
            out += padded[di:di + h, dj:dj + w]
    return out


def _correlate_zero_pad(img, kernel):
    """Zero-padded correlation with a small integer kernel, exact on integer data."""
    h, w = img.shape
    r = kernel.shape[0] // 2
    padded = np.zeros((h + 2 * r, w + 2 * r), dtype=np.float64)
    padded[r:r + h, r:r + w] = img
    out = np.zeros((h, w), dtype=np.float64)
    for di in range(kernel.shape[0]):
        for dj in range(kernel.shape[1]):
            if kernel[di, dj] != 0:
                out += kernel[di, dj] * padded[di:di + h, dj:dj + w]
    return out


def harris_response(gray, window_size=5, k=0.04):
    """APP/FeatureCraft_Backend.py:336-355 -- np.gradient, products, ones-window sums,
    R = det - k * trace**2 (exactly two roundings: k*tr^2, then the subtraction)."""
    if window_size % 2 == 0:
        raise ValueError("cannot reshape array: window_size must be odd")
    gray = np.asarray(gray)
    ix, iy = _np_gradient_2d(gray)
    r = window_size // 2
    sxx = _box_sum_zero_pad(ix * ix, r)
    sxy = _box_sum_zero_pad(iy * ix, r)
    syy = _box_sum_zero_pad(iy * iy, r)
    det = sxx * syy - sxy ** 2
    trace = sxx + syy
    return det - k * (trace ** 2)


def threshold_corners(response, threshold):
    """APP/FeatureCraft_Backend.py:358-361 -- np.argwhere(R > thr): (row, col) row-major + values."""
    idx = np.argwhere(response > threshold)
    return idx, response[response > threshold]


def apply_harris_detector_vectorized(img, window_size=5, k=0.04, threshold=10000):
    """APP/FeatureCraft_Backend.py:306-380 (arithmetic only).  Returns (corners[(x, y, r)],
    painted image, response map)."""
    img = np.asarray(img)
    gray = convert_to_gray(img)
    resp = harris_response(gray, window_size, k)
    idx, vals = threshold_corners(resp, threshold)
    out = img.copy()
    if out.ndim == 3:
        out[idx[:, 0], idx[:, 1]] = (0, 0, 255)
    return list(zip(idx[:, 1], idx[:, 0], vals)), out, resp


def dsterf_2x2_min(a, b, c):
    """Smallest eigenvalue of [[a, b], [b, c]] as numpy.linalg.eigvalsh computes it for 2x2 input:
    LAPACK dsyevd(jobz='N') -> dsterf.  dsterf first zeroes a negligible off-diagonal
    (|b| <= sqrt|a|*sqrt|c|*eps, eps = dlamch('E') = 2^-53) -- the eigenvalues are then (a, c) --
    and otherwise calls dlae2(a, b, c).  The dlae2 sequence is restated operation by operation."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)
    eps = 2.0 ** -53
    with np.errstate(all="ignore"):
        split = np.abs(b) <= (np.sqrt(np.abs(a)) * np.sqrt(np.abs(c))) * eps
        sm = a + c
        df = a - c
        adf = np.abs(df)
        tb = b + b
        ab = np.abs(tb)
        a_big = np.abs(a) > np.abs(c)
        acmx = np.where(a_big, a, c)
        acmn = np.where(a_big, c, a)
        r1 = ab / adf
        rt_gt = adf * np.sqrt(1.0 + r1 * r1)
        r2 = adf / ab
        rt_lt = ab * np.sqrt(1.0 + r2 * r2)
        rt = np.where(adf > ab, rt_gt, np.where(adf < ab, rt_lt, ab * np.sqrt(2.0)))
        # sm < 0
        rt1_neg = 0.5 * (sm - rt)
        rt2_neg = (acmx / rt1_neg) * acmn - (b / rt1_neg) * b
        # sm > 0
        rt1_pos = 0.5 * (sm + rt)
        rt2_pos = (acmx / rt1_pos) * acmn - (b / rt1_pos) * b
        lo = np.where(sm < 0, np.minimum(rt1_neg, rt2_neg),
                      np.where(sm > 0, np.minimum(rt1_pos, rt2_pos), -0.5 * rt))
        return np.where(split, np.minimum(a, c), lo)


def lambda_minus_response(gray):
    """APP/FeatureCraft_Backend.py:402-437 -- K_X / K_X.T zero-padded correlation, products,
    5x5 ones window, /25, min eigenvalue (window_size argument is ignored by the reference)."""
    gray = np.asarray(gray).astype(np.float64)
    gx = _correlate_zero_pad(gray, K_X)
    gy = _correlate_zero_pad(gray, K_X.T)
    hxx = _box_sum_zero_pad(gx * gx, 2) / 25
    hyy = _box_sum_zero_pad(gy * gy, 2) / 25
    hxy = _box_sum_zero_pad(gx * gy, 2) / 25
    return dsterf_2x2_min(hxx, hxy, hyy)


def apply_lambda_minus_vectorized(img, window_size=5, threshold_percentage=0.01):
    """APP/FeatureCraft_Backend.py:382-452 (arithmetic only).  Returns ((rows, cols), eig map)."""
    gray = convert_to_gray(img)
    eig = lambda_minus_response(gray)
    threshold = threshold_percentage * eig.max()
    return np.where(eig > threshold), eig


def _scalar_pow2(arr):
    """x**2 the way a numpy float64 SCALAR computes it: libm pow(x, 2.0).  glibc's pow is not
    correctly rounded on exact ties above 2^53, and the notebook's per-pixel scalar loop inherits
    that; array code (np.square) does not.  Only the notebook variant needs this."""
    flat = np.asarray(arr, dtype=np.float64).ravel()
    return np.array([float(v) ** 2 for v in flat.tolist()], dtype=np.float64).reshape(np.shape(arr))


def harris_detection(gray, window_size, k, threshold):
    """implementation_without_ui/Corner Detection/harris.ipynb:123-206 -- K_X gradients, window
    sums only for offset <= y < H-offset (interior), r > threshold -> [x, y, r] in scan order.
    Takes the gray uint8 image (the notebook reads a path and calls cv2.cvtColor).  Sums reach
    9.4e8 so Sxx*Syy, Sxy**2 and trace**2 exceed 2^53 and round; the two squares are SCALAR pow
    calls in the notebook (see _scalar_pow2)."""
    gray = np.asarray(gray)
    h, w = gray.shape[:2]
    dx = _correlate_zero_pad(gray.astype(np.float64), K_X)
    dy = _correlate_zero_pad(gray.astype(np.float64), K_X.T)
    off = int(window_size / 2)
    sxx = _box_sum_zero_pad(dx * dx, off)
    sxy = _box_sum_zero_pad(dy * dx, off)
    syy = _box_sum_zero_pad(dy * dy, off)
    det = (sxx * syy) - _scalar_pow2(sxy)
    trace = sxx + syy
    r = det - k * _scalar_pow2(trace)
    keep = np.zeros_like(r, dtype=bool)
    keep[off:h - off, off:w - off] = True
    idx = np.argwhere(keep & (r > threshold))
    return [[int(x), int(y), r[y, x]] for y, x in idx], r


# ======================================================================================
# SIFT: scale space
# ======================================================================================
def gaussian_filter_kernel(sigma, kernel_size=None):
    """APP/utils/SIFT_scale_space.py:5-30 -- size 2*((4*sigma+1)//2)+1, exp(-r^2/2s^2)/(2*pi*s^2),
    NOT renormalised."""
    kernel_size = (4 * sigma) + 1
    offset = kernel_size // 2
    x = np.arange(-offset, offset + 1)[:, np.newaxis]
    y = np.arange(-offset, offset + 1)
    kernel = np.exp(-(x ** 2 + y ** 2) / (2 * sigma ** 2))
    kernel /= 2 * np.pi * (sigma ** 2)
    return kernel


def generateGaussianKernels(sigma, s):
    """APP/utils/SIFT_scale_space.py:33-57 -- s+3 kernels, sigma_i by REPEATED multiplication."""
    kernels = [gaussian_filter_kernel(sigma)]
    scale_level = sigma
    k = 2 ** (1 / s)
    for _ in range(1, s + 3):
        scale_level *= k
        kernels.append(gaussian_filter_kernel(scale_level))
    return kernels


def sigma_levels(sigma, s):
    """The sigma_i sequence generateGaussianKernels walks through (same float operations)."""
    out = [sigma]
    lvl = sigma
    k = 2 ** (1 / s)
    for _ in range(1, s + 3):
        lvl *= k
        out.append(lvl)
    return out


def separable_taps(sigma):
    """1-D taps g(x)=exp(-x^2/2s^2)/sqrt(2*pi*s^2) whose outer product equals gaussian_filter_kernel
    to ~1e-18.  Used by blur(mode='separable') and, on the host side, by the product."""
    offset = int(((4 * sigma) + 1) // 2)
    x = np.arange(-offset, offset + 1, dtype=np.float64)
    return np.exp(-(x ** 2) / (2 * sigma ** 2)) / np.sqrt(2 * np.pi * (sigma ** 2))


def blur(image, kernel2d, mode="direct"):
    """scipy.signal.convolve2d(image, K, 'same', 'symm') (APP/utils/SIFT_scale_space.py:132,144).
    mode='separable' evaluates the same operator as two 1-D passes (<=2e-15 from the direct sum);
    it exists so that large-image tests finish in seconds."""
    if mode == "direct":
        return _scipy_convolve2d(image, kernel2d, "same", "symm")
    n = kernel2d.shape[0]
    c = n // 2
    taps = kernel2d[c, :] / np.sqrt(kernel2d[c, c])  # g(x) * g(0) / g(0)
    tmp = _scipy_convolve2d(image, taps[np.newaxis, :], "same", "symm")
    return _scipy_convolve2d(tmp, taps[:, np.newaxis], "same", "symm")


def extrema_mask(d0, d1, d2):
    """Vectorised APP/utils/SIFT_scale_space.py:78-88: the (3,3,3) patch is flattened with the
    scale axis fastest, so `argmax == 13` means centre > the 13 earlier entries and >= the 13 later
    ones (first maximum wins), likewise for argmin.  Loops are range(1, H-2) / range(1, W-2)."""
    h, w = d1.shape
    mask = np.zeros((h, w), dtype=bool)
    if h < 4 or w < 4:
        return mask
    stack = (d0, d1, d2)
    ctr = d1[1:h - 2, 1:w - 2]
    is_max = np.ones(ctr.shape, dtype=bool)
    is_min = np.ones(ctr.shape, dtype=bool)
    for di in range(3):
        for dj in range(3):
            for s in range(3):
                flat = di * 9 + dj * 3 + s
                if flat == 13:
                    continue
                nb = stack[s][di:di + h - 3, dj:dj + w - 3]
                if flat < 13:
                    is_max &= ctr > nb
                    is_min &= ctr < nb
                else:
                    is_max &= ctr >= nb
                    is_min &= ctr <= nb
    mask[1:h - 2, 1:w - 2] = is_max | is_min
    return mask


def script_filter_mask(mask, dogs, k, contrast_th, ratio_th):
    """implementation_without_ui/SIFT/SIFT.py:202-215 with localize_keypoint :225-269 -- contrast is
    read from the LOCAL 3-stack with the GLOBAL k (DoG_k for k=1, DoG_{k+1} for k=2; IndexError for
    k>=3), Hessian 2x2 on DoG_k, reject |contrast| < contrast_th or tr^2/det > ratio_th
    (negative / NaN ratios are kept)."""
    local = dogs[k - 1:k + 2]
    if k >= len(local):
        raise IndexError("index %d is out of bounds for axis 2 with size 3" % k)
    contrast_img = local[k]
    d = dogs[k]
    ii, jj = np.nonzero(mask)
    keep = np.zeros_like(mask)
    if ii.size == 0:
        return keep
    contrast = contrast_img[ii, jj]
    dxx = d[ii, jj + 1] - 2 * d[ii, jj] + d[ii, jj - 1]
    dyy = d[ii + 1, jj] - 2 * d[ii, jj] + d[ii - 1, jj]
    dxy = ((d[ii + 1, jj + 1] - d[ii + 1, jj - 1]) - (d[ii - 1, jj + 1] - d[ii - 1, jj - 1])) / 4
    tr = dxx + dyy
    det = dxx * dyy - dxy ** 2
    with np.errstate(all="ignore"):
        r = (tr ** 2) / det
    ok = ~(np.abs(contrast) < contrast_th) & ~(r > ratio_th)
    keep[ii[ok], jj[ok]] = True
    return keep


def localize_keypoint_full(D, x, y, s):
    """APP/utils/SIFT_scale_space.py:213-261 (the app's 3-D quadratic fit; its only call site is commented out,
    :89-103): first and second differences of the DoG stack D[y, x, s], offset = -inv(H) J.  Returns
    (offset, J, H[:2, :2], x, y, s) like the reference."""
    D = np.asarray(D).astype(np.float64)
    dx = (D[y, x + 1, s] - D[y, x - 1, s]) / 2.0
    dy = (D[y + 1, x, s] - D[y - 1, x, s]) / 2.0
    ds = (D[y, x, s + 1] - D[y, x, s - 1]) / 2.0
    dxx = D[y, x + 1, s] - 2 * D[y, x, s] + D[y, x - 1, s]
    dxy = ((D[y + 1, x + 1, s] - D[y + 1, x - 1, s]) - (D[y - 1, x + 1, s] - D[y - 1, x - 1, s])) / 4
    dxs = ((D[y, x + 1, s + 1] - D[y, x - 1, s + 1]) - (D[y, x + 1, s - 1] - D[y, x - 1, s - 1])) / 4
    dyy = D[y + 1, x, s] - 2 * D[y, x, s] + D[y - 1, x, s]
    dys = ((D[y + 1, x, s + 1] - D[y - 1, x, s + 1]) - (D[y + 1, x, s - 1] - D[y - 1, x, s - 1])) / 4
    dss = D[y, x, s + 1] - 2 * D[y, x, s] + D[y, x, s - 1]
    J = np.array([dx, dy, ds])
    H = np.array([[dxx, dxy, dxs], [dxy, dyy, dys], [dxs, dys, dss]])
    offset = -np.linalg.inv(H).dot(J)
    return offset, J, H[:2, :2], x, y, s


def refine_keypoint(D, x, y, s, contrast_th, ratio_th):
    """The acceptance tests get_keypoints keeps commented out around localize_keypoint
    (APP/utils/SIFT_scale_space.py:89-101), evaluated as written: -> (offset, contrast, ratio, keep)."""
    offset, J, H, x, y, s = localize_keypoint_full(D, x, y, s)
    contrast = np.asarray(D, dtype=np.float64)[y, x, s] + 0.5 * J.dot(offset)
    tr = H[0][0] + H[1][1]
    det = H[0][0] * H[1][1] - H[0][1] ** 2
    with np.errstate(all="ignore"):
        r = (tr ** 2) / det
    keep = not (np.max(offset) > 0.5) and not (abs(contrast) < contrast_th) and not (r > ratio_th)
    return offset, contrast, r, keep


def get_keypoints(DOG_octave, k, contrast_th, ratio_th, DoG_full_array, variant="app"):
    """APP/utils/SIFT_scale_space.py:60-106 (variant='app': every extremum is kept) or
    implementation_without_ui/SIFT/SIFT.py:172-222 (variant='script').  DoG_full_array may be the
    reference's (H, W, n) stack or a list of 2-D arrays.  Returns an (N, 3) int array [i, j, k]."""
    mask = extrema_mask(DOG_octave[0], DOG_octave[1], DOG_octave[2])
    if variant == "script":
        if isinstance(DoG_full_array, np.ndarray) and DoG_full_array.ndim == 3:
            full = [DoG_full_array[:, :, n] for n in range(DoG_full_array.shape[2])]
        else:
            full = list(DoG_full_array)
        # the reference indexes the local 3-stack (== full[k-1:k+2]) and the full stack
        mask = script_filter_mask(mask, full, k, contrast_th, ratio_th)
    ij = np.argwhere(mask)
    if ij.size == 0:
        return np.array([])
    return np.concatenate([ij, np.full((ij.shape[0], 1), k, dtype=ij.dtype)], axis=1)


def generate_gaussian_images_in_octave(image, gaussian_kernels, contrast_th, ratio_th, octave_index,
                                       variant="app", blur_mode="direct"):
    """APP/utils/SIFT_scale_space.py:109-160 -- every blur is applied to the octave BASE image;
    octaves > 0 put the unblurred base in slot 0; get_keypoints runs each time >= 3 DoGs exist."""
    gaussians = []
    if octave_index == 0:
        gaussians.append(blur(image, gaussian_kernels[0], blur_mode))
    else:
        gaussians.append(image)
    dogs = []
    keypoints = []
    for kern in gaussian_kernels[1:]:
        gaussians.append(blur(image, kern, blur_mode))
        dogs.append(gaussians[-1] - gaussians[-2])
        if len(dogs) > 2:
            keypoints.extend(get_keypoints(dogs[-3:], len(dogs) - 2, contrast_th, ratio_th, list(dogs), variant))
    return gaussians, dogs, keypoints


def generate_octaves_pyramid(img, num_octaves=4, s_value=2, sigma=1.6, contrast_th=0.03, ratio_th=10,
                             variant="app", blur_mode="direct"):
    """APP/utils/SIFT_scale_space.py:163-210 -- next base = G[-3][::2, ::2]."""
    kernels = generateGaussianKernels(sigma, s_value)
    pyramid, dog_pyramid, keypoints = [], [], []
    for o in range(num_octaves):
        g, d, kp = generate_gaussian_images_in_octave(img, kernels, contrast_th, ratio_th, o, variant, blur_mode)
        pyramid.append(g)
        dog_pyramid.append(d)
        keypoints.append(kp)
        img = g[-3][::2, ::2]
    return pyramid, dog_pyramid, keypoints


# ======================================================================================
# SIFT: orientation + descriptor
# ======================================================================================
def convert_to_grayscale(image):
    """APP/utils/keypoint_descriptor.py:36-39 (float result, no cast)."""
    image = np.asarray(image)
    if image.ndim == 3:
        return np.dot(image[..., :3], [0.2989, 0.5870, 0.1140])
    return image


def represent_keypoints(keypoints, DoG):
    """APP/utils/keypoint_descriptor.py:42-69 -- bool masks per (octave, dog_idx) with dog_idx in
    range(1, len(DoG) - 1) where len(DoG) is the NUMBER OF OCTAVES (reference quirk)."""
    out = []
    for o, kps in enumerate(keypoints):
        per_octave = []
        arr = np.asarray(kps).reshape(-1, 3).astype(np.int64) if len(kps) else np.zeros((0, 3), np.int64)
        for dog_idx in range(1, len(DoG) - 1):
            m = np.zeros(DoG[o][0].shape, dtype=bool)
            sel = arr[arr[:, 2] == dog_idx]
            m[sel[:, 0], sel[:, 1]] = True
            per_octave.append(m)
        out.append(per_octave)
    return out


def sift_gradient(img):
    """APP/utils/keypoint_descriptor.py:72-79 -- true convolution with [-1, 0, 1], 'symm' border:
    gx = I[., j-1] - I[., j+1], gy = I[i-1, .] - I[i+1, .]; direction in degrees wrapped by % 360."""
    dx = np.array([-1, 0, 1]).reshape((1, 3))
    gx = _scipy_convolve2d(img, dx, boundary="symm", mode="same")
    gy = _scipy_convolve2d(img, dx.T, boundary="symm", mode="same")
    magnitude = np.sqrt(gx * gx + gy * gy)
    direction = np.rad2deg(np.arctan2(gy, gx)) % 360
    return gx, gy, magnitude, direction


def padded_slice(img, sl):
    """APP/utils/keypoint_descriptor.py:82-113 -- window [r0, r1) x [c0, c1) with zeros outside."""
    out = np.zeros((sl[1] - sl[0], sl[3] - sl[2]), dtype=img.dtype)
    r0, r1 = max(sl[0], 0), min(sl[1], img.shape[0])
    c0, c1 = max(sl[2], 0), min(sl[3], img.shape[1])
    if r1 > r0 and c1 > c0:
        out[r0 - sl[0]:r1 - sl[0], c0 - sl[2]:c1 - sl[2]] = img[r0:r1, c0:c1]
    return out


def keypoint_sigma(sigma_base, octave_idx, scale_idx, s):
    """The window sigma both stages use (APP/utils/keypoint_descriptor.py:127-134, :244-251),
    evaluated with the reference's operation order."""
    return 1.5 * sigma_base * (2 ** octave_idx) * ((2 ** (1 / s)) ** (scale_idx))


def dog_keypoints_orientations(img_gaussians, keypoints, sigma_base, num_bins=36, s=2, numpy_legacy_promotion=False):
    """APP/utils/keypoint_descriptor.py:116-172.  Bin index by np.round (half to even) so bin
    `num_bins` exists and is dropped; float32 histogram; EVERY bin >= 0.8*max emits a keypoint.
    numpy_legacy_promotion: evaluate `hist >= 0.8 * hist.max()` the way numpy 1.26 (the reference's pinned version) does --
    float64 scalar product, rounded to float32 for the comparison with the float32 array -- instead of this image's
    numpy >= 2 rule (float32 product); an emulation, the reference itself can only be run under the numpy installed here."""
    kps = []
    for o in range(len(img_gaussians)):
        for idx, mask in enumerate(keypoints[o]):
            scale_idx = idx + 1
            gimg = img_gaussians[o][scale_idx]  # IndexError here mirrors the reference for n_oct > s+4
            sigma = keypoint_sigma(sigma_base, o, scale_idx, s)
            kernel = gaussian_filter_kernel(sigma)
            radius = int(round(sigma * 2))
            _, _, magnitude, direction = sift_gradient(gimg)
            didx = np.round(direction * num_bins / 360).astype(int)
            for i, j in map(tuple, np.argwhere(mask).tolist()):
                win = [i - radius, i + radius + 1, j - radius, j + radius + 1]
                weight = padded_slice(magnitude, win) * kernel
                dwin = padded_slice(didx, win)
                hist = np.zeros(num_bins, dtype=np.float32)
                for b in range(num_bins):
                    hist[b] = np.sum(weight[dwin == b])
                cut = np.float32(0.8 * float(hist.max())) if numpy_legacy_promotion else 0.8 * hist.max()
                for b in np.argwhere(hist >= cut).tolist():
                    kps.append((i, j, o, scale_idx, (b[0] + 0.5) * (360.0 / num_bins) % 360))
    return kps


def nn_rotated_window(image, center, theta, width=16, height=16):
    """rotated_subimage (APP/utils/keypoint_descriptor.py:175-212) with cv2.warpAffine restated:
    OpenCV (modules/imgproc/src/imgwarp.cpp, warpAffine, INTER_NEAREST) rounds source coordinates in
    AB_BITS=10 fixed point: X = (rint((M01*y+M02)*1024) + 512 + rint(M00*x*1024)) >> 10 (likewise
    Y), border constant 0.  theta is converted with the reference's 3.14159/180."""
    theta = theta * (3.14159 / 180)
    v_x = (cos(theta), sin(theta))
    v_y = (-sin(theta), cos(theta))
    s_x = center[0] - v_x[0] * ((width - 1) / 2) - v_y[0] * ((height - 1) / 2)
    s_y = center[1] - v_x[1] * ((width - 1) / 2) - v_y[1] * ((height - 1) / 2)
    xs = np.arange(width)
    ys = np.arange(height)
    adelta = np.rint(v_x[0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(v_x[1] * xs * 1024).astype(np.int64)
    x0 = np.rint((v_y[0] * ys + s_x) * 1024).astype(np.int64) + 512
    y0 = np.rint((v_y[1] * ys + s_y) * 1024).astype(np.int64) + 512
    X = (x0[:, None] + adelta[None, :]) >> 10
    Y = (y0[:, None] + bdelta[None, :]) >> 10
    ok = (X >= 0) & (X < image.shape[1]) & (Y >= 0) & (Y < image.shape[0])
    out = np.zeros((height, width), dtype=image.dtype)
    out[ok] = image[Y[ok], X[ok]]
    return out


def get_gaussian_mask(sigma, filter_size):
    """APP/utils/keypoint_descriptor.py:215-227 -- sum-normalised Gaussian centred at (n-1)/2."""
    if sigma > 0:
        kernel = np.fromfunction(
            lambda x, y: (1 / (2 * np.pi * sigma ** 2))
            * np.exp(-((x - (filter_size - 1) / 2) ** 2 + (y - (filter_size - 1) / 2) ** 2) / (2 * sigma ** 2)),
            (filter_size, filter_size),
        )
        return kernel / np.sum(kernel)
    raise ValueError("Invalid value of Sigma")


def extract_sift_descriptors(img_gaussians, keypoints, base_sigma, num_bins=8, s=2):
    """APP/utils/keypoint_descriptor.py:230-289 -- 16x16 NN-rotated magnitude x Gaussian mask,
    bin = int(((dir - theta) % 360) * 8 / 360) (bin 8 possible -> dropped), 4x4 cells x 8 bins with
    float32 cell histograms, normalise, clip to [2^-10 (float16 eps), 0.2], normalise again."""
    descriptors, points = [], []
    cur = None
    kernel = magnitude = direction = None
    for i, j, o, sc, theta in keypoints:
        if cur != (o, sc):
            cur = (o, sc)
            sigma = keypoint_sigma(base_sigma, o, sc, s)
            kernel = get_gaussian_mask(sigma=sigma, filter_size=16)
            _, _, magnitude, direction = sift_gradient(img_gaussians[o][sc])
        wmag = nn_rotated_window(magnitude, (j, i), theta) * kernel
        wdir = nn_rotated_window(direction, (j, i), theta)
        wbin = (((wdir - theta) % 360) * num_bins / 360.0).astype(int)
        feats = []
        for ci in range(4):
            for cj in range(4):
                sw = wmag[ci * 4:(ci + 1) * 4, cj * 4:(cj + 1) * 4]
                sb = wbin[ci * 4:(ci + 1) * 4, cj * 4:(cj + 1) * 4]
                hist = np.zeros(num_bins, dtype=np.float32)
                for b in range(num_bins):
                    hist[b] = np.sum(sw[sb == b])
                feats.extend(hist.tolist())
        feats = np.array(feats)
        with np.errstate(all="ignore"):
            feats /= np.linalg.norm(feats)
            np.clip(feats, np.finfo(np.float16).eps, 0.2, out=feats)
            feats /= np.linalg.norm(feats)
        descriptors.append(feats)
        points.append((i, j, o, sc, theta))
    return points, descriptors


def compute_keypoints_and_descriptors_from_base(base_image, n_octaves, s_value, sigma_base, contrast_th, r_ratio,
                                                variant="app", blur_mode="direct", numpy_legacy_promotion=False):
    """computeKeypointsAndDescriptors (APP/utils/keypoint_descriptor.py:292-311) from the point
    AFTER rescale(x2): pyramid -> masks -> orientations -> descriptors."""
    pyramid, dog, kps = generate_octaves_pyramid(base_image, n_octaves, s_value, sigma_base, contrast_th, r_ratio,
                                                 variant, blur_mode)
    masks = represent_keypoints(kps, dog)
    oriented = dog_keypoints_orientations(pyramid, masks, sigma_base, 36, s_value, numpy_legacy_promotion)
    return extract_sift_descriptors(pyramid, oriented, sigma_base, 8, s_value)


def sift_resize(img, ratio=None):
    """APP/utils/keypoint_descriptor.py:11-33 with skimage.transform.resize(img, newshape, anti_aliasing=True)
    (scikit-image==0.24.0, absent here) restated as the scipy.ndimage calls it makes: img_as_float (uint8 -> /255),
    gaussian_filter(sigma = max(0, (in/out - 1)/2) per spatial axis, mode='mirror'), zoom(order=1, mode='mirror',
    grid_mode=True).  UNPINNED against skimage itself (see the module docstring)."""
    from scipy import ndimage

    img = np.asarray(img)
    ratio = ratio if ratio is not None else np.sqrt((1024 * 1024) / np.prod(img.shape[:2]))
    newshape = list(map(lambda d: int(round(d * ratio)), img.shape[:2]))
    image = img.astype(np.float64) / 255.0 if img.dtype == np.uint8 else img.astype(np.float64)
    factors = [image.shape[a] / newshape[a] for a in range(2)]
    sig = [max(0.0, (f - 1.0) / 2.0) for f in factors] + [0.0] * (image.ndim - 2)
    if any(s > 0 for s in sig):
        image = ndimage.gaussian_filter(image, sig, mode="mirror")
    zoom = [newshape[a] / image.shape[a] for a in range(2)] + [1.0] * (image.ndim - 2)
    return ndimage.zoom(image, zoom, order=1, mode="mirror", grid_mode=True), ratio


def rescale2(gray):
    """skimage.transform.rescale(gray, 2, anti_aliasing=False) (APP/utils/keypoint_descriptor.py:296) as the
    scipy.ndimage call it makes; UNPINNED against skimage itself."""
    from scipy import ndimage

    return ndimage.zoom(np.asarray(gray, dtype=np.float64), 2, order=1, mode="mirror", grid_mode=True)


def apply_sift_good(main_image, template, n_octaves, s_value, sigma_base, contrast_th, r_ratio, tuning_factor,
                    blur_mode="direct"):
    """apply_sift (APP/FeatureCraft_Backend.py:480-523 / SIFT.py:683-707) up to the `good` list that match() draws:
    sift_resize both (ratio from the main image), gray, x2, keypoints + descriptors, knn + ratio test.
    Returns (good (M, 2) int array, points_main, points_template)."""
    main_image, ratio = sift_resize(main_image)
    template, _ = sift_resize(template, ratio)
    out = []
    for img in (main_image, template):
        base = rescale2(convert_to_grayscale(img))
        out.append(compute_keypoints_and_descriptors_from_base(base, n_octaves, s_value, sigma_base, contrast_th, r_ratio,
                                                               "app", blur_mode))
    (pa, da), (pb, db) = out
    da, db = np.array(da, dtype=np.float32).reshape(-1, 128), np.array(db, dtype=np.float32).reshape(-1, 128)
    ok = ~np.isnan(da).any(1)  # a NaN query row makes BFMatcher undefined; defined here as "no match" 
    pairs, _ = match_indices(da[ok], db, tuning_factor)
    pairs[:, 0] = np.nonzero(ok)[0][pairs[:, 0]]
    return pairs, pa, pb


def kp_list_2_xy(kp_list):
    """APP/utils/keypoint_descriptor.py:314-327 without cv2.KeyPoint: (x, y, size, angle)."""
    return [(kp[1] * (2 ** (kp[2] - 1)), kp[0] * (2 ** (kp[2] - 1)), kp[3], kp[4]) for kp in kp_list]


# ======================================================================================
# SIFT: matching
# ======================================================================================
def knn2_l2(desc_a, desc_b, chunk=2048):
    """cv2.BFMatcher().knnMatch(a, b, k=2) (APP/utils/keypoint_descriptor.py:338-342): float32
    descriptors, L2 distance, two nearest train rows per query row, ties -> lower train index.
    Returns (idx (Ka, 2) int64, dist (Ka, 2) float32)."""
    a = np.asarray(desc_a, dtype=np.float32).reshape(-1, 128)
    b = np.asarray(desc_b, dtype=np.float32).reshape(-1, 128)
    if b.shape[0] < 2:
        raise ValueError("not enough values to unpack (expected 2, got %d)" % b.shape[0])
    idx = np.empty((a.shape[0], 2), dtype=np.int64)
    dist = np.empty((a.shape[0], 2), dtype=np.float32)
    for lo in range(0, a.shape[0], chunk):
        blk = a[lo:lo + chunk]
        d2 = np.zeros((blk.shape[0], b.shape[0]), dtype=np.float32)
        for c in range(128):  # float32 accumulation of squared differences, like cv::normL2Sqr_
            diff = blk[:, c:c + 1] - b[None, :, c].reshape(1, -1)
            d2 += diff * diff
        d = np.sqrt(d2)
        order = np.argsort(d, axis=1, kind="stable")[:, :2]
        idx[lo:lo + chunk] = order
        dist[lo:lo + chunk] = np.take_along_axis(d, order, axis=1)
    return idx, dist


def match_indices(desc_a, desc_b, tuning_distance=0.3):
    """The `good` list of match() (APP/utils/keypoint_descriptor.py:330-349) as an (M, 2) array of
    (queryIdx, trainIdx) plus the float32 distances; strict `<`, ratio evaluated in double."""
    idx, dist = knn2_l2(desc_a, desc_b)
    good = dist[:, 0].astype(np.float64) < tuning_distance * dist[:, 1].astype(np.float64)
    q = np.nonzero(good)[0]
    return np.stack([q, idx[q, 0]], axis=1), dist[q, 0]
