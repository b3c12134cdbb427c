"""Loads the UNMODIFIED reference from /root/reference, in place, for pinning the oracle.

TEST INFRASTRUCTURE ONLY.  This module is imported by ``oracle/make_golden.py`` /
``oracle/make_golden_large.py`` (authoring container) and by the one case of ``tests/test_cabi_cpu.py`` that uses
the script's own ``localize_keypoint`` when /root/reference is mounted (it is absent on the GPU box).  Nothing here
is shipped or timed.

No reference source is copied into this repository: function definitions are read from the
reference files at run time (``importlib`` for the modules that import cleanly, ``ast`` extraction
for the ones that drag in PyQt5 / matplotlib / a demo script) and executed with three patches that
SURVEY.md section 8(c) documents:

* ``skimage.transform`` is absent in this image -> a stub built on ``scipy.ndimage.zoom``.
  The parity boundary therefore starts AFTER ``rescale`` (both sides get the same base image).
* ``cv2.warpAffine`` on CV_64F + INTER_NEAREST returns garbage in cv2 4.13.0 -> ``rotated_subimage``
  is replaced by the 10-bit fixed-point nearest-neighbour gather OpenCV 4.10 (the pinned version,
  requirements.txt:2) performs, validated against cv2's working float32 path.
* Qt-bound methods (``BackendClass.apply_harris_detector_vectorized`` / ``apply_lambda_minus_vectorized``,
  FeatureCraft_Backend.py:306-452) are bound to a dummy ``self``.
"""
from __future__ import annotations

import ast
import importlib
import os
import sys
import types
from math import cos, sin

import numpy as np

REF_ROOT = os.environ.get("FEATURECRAFT_REFERENCE", "/root/reference")
APP_DIR = os.path.join(REF_ROOT, "FeatureCraft-PyQt-App")
SCRIPT = os.path.join(REF_ROOT, "implementation_without_ui", "SIFT", "SIFT.py")
BACKEND = os.path.join(APP_DIR, "FeatureCraft_Backend.py")
HARRIS_NB = os.path.join(REF_ROOT, "implementation_without_ui", "Corner Detection", "harris.ipynb")


def available() -> bool:
    return os.path.isfile(BACKEND)


# --------------------------------------------------------------------------------------
# patches
# --------------------------------------------------------------------------------------
def _install_skimage_stub() -> None:
    if "skimage.transform" in sys.modules:
        return
    try:  # a real skimage wins if it ever appears
        import skimage.transform  # noqa: F401
        return
    except Exception:
        pass
    from scipy import ndimage

    def rescale(image, scale, anti_aliasing=False, **_kw):
        # skimage.transform.rescale(order=1, mode='reflect') on a 2-D image == bilinear zoom with
        # pixel-centre alignment; UNVERIFIED restatement (no skimage here), so no parity claim
        # is made across this call.
        return ndimage.zoom(np.asarray(image, dtype=np.float64), scale, order=1, mode="mirror", grid_mode=True)

    def resize(image, output_shape, anti_aliasing=True, **_kw):
        image = np.asarray(image, dtype=np.float64)
        factors = [image.shape[a] / output_shape[a] for a in range(2)]
        if anti_aliasing:
            sig = [max(0.0, (f - 1.0) / 2.0) for f in factors] + [0.0] * (image.ndim - 2)
            image = ndimage.gaussian_filter(image, sig, mode="mirror")
        zoom = [output_shape[a] / image.shape[a] for a in range(2)] + [1.0] * (image.ndim - 2)
        return ndimage.zoom(image, zoom, order=1, mode="mirror", grid_mode=True)

    pkg = types.ModuleType("skimage")
    sub = types.ModuleType("skimage.transform")
    sub.rescale, sub.resize = rescale, resize
    pkg.transform = sub
    sys.modules["skimage"] = pkg
    sys.modules["skimage.transform"] = sub


def fixed_point_rotated_subimage(image, center, theta, width, height):
    """cv2.warpAffine(INTER_NEAREST | WARP_INVERSE_MAP, BORDER_CONSTANT) as OpenCV 4.10 computes it.

    Same argument list as the reference's rotated_subimage (keypoint_descriptor.py:175-212); the
    matrix is built with the same expressions, the gather follows imgwarp.cpp's AB_BITS=10 scheme.
    """
    theta = theta * (3.14159 / 180)
    v_x = (cos(theta), sin(theta))
    v_y = (-sin(theta), cos(theta))
    s_x = center[0] - v_x[0] * ((width - 1) / 2) - v_y[0] * ((height - 1) / 2)
    s_y = center[1] - v_x[1] * ((width - 1) / 2) - v_y[1] * ((height - 1) / 2)
    M = np.array([[v_x[0], v_y[0], s_x], [v_x[1], v_y[1], s_y]])
    xs = np.arange(width)
    ys = np.arange(height)
    adelta = np.rint(M[0, 0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(M[1, 0] * xs * 1024).astype(np.int64)
    X0 = np.rint((M[0, 1] * ys + M[0, 2]) * 1024).astype(np.int64) + 512
    Y0 = np.rint((M[1, 1] * ys + M[1, 2]) * 1024).astype(np.int64) + 512
    X = (X0[:, None] + adelta[None, :]) >> 10
    Y = (Y0[:, None] + bdelta[None, :]) >> 10
    ok = (X >= 0) & (X < image.shape[1]) & (Y >= 0) & (Y < image.shape[0])
    out = np.zeros((height, width), dtype=image.dtype)
    out[ok] = image[Y[ok], X[ok]]
    return out


# --------------------------------------------------------------------------------------
# loaders
# --------------------------------------------------------------------------------------
_cache: dict = {}


def app_modules():
    """(harris_utils, SIFT_scale_space, keypoint_descriptor) imported from the reference app."""
    if "app" in _cache:
        return _cache["app"]
    _install_skimage_stub()
    if APP_DIR not in sys.path:
        sys.path.insert(0, APP_DIR)
    hu = importlib.import_module("utils.harris_utils")
    ss = importlib.import_module("utils.SIFT_scale_space")
    kd = importlib.import_module("utils.keypoint_descriptor")
    kd.rotated_subimage = fixed_point_rotated_subimage
    _cache["app"] = (hu, ss, kd)
    return _cache["app"]


def _function_defs(path: str, names=None) -> dict:
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and (names is None or node.name in names):
            out[node.name] = ast.get_source_segment(src, node)
    return out


def script_namespace() -> dict:
    """Function defs of implementation_without_ui/SIFT/SIFT.py executed without its demo code."""
    if "script" in _cache:
        return _cache["script"]
    _install_skimage_stub()
    import cv2
    from scipy.signal import convolve2d
    from skimage.transform import rescale, resize

    ns = {"np": np, "cv2": cv2, "convolve2d": convolve2d, "rescale": rescale, "resize": resize,
          "cos": cos, "sin": sin}
    for name, code in _function_defs(SCRIPT).items():
        if name.startswith("visualize"):
            continue
        exec(compile(code, SCRIPT, "exec"), ns)
    ns["rotated_subimage"] = fixed_point_rotated_subimage
    _cache["script"] = ns
    return ns


class _DummyUI:
    def __getattr__(self, _name):
        return None


class _DummySelf:
    """Stands in for BackendClass: the two detector methods only touch these attributes."""

    def __init__(self):
        self.ui = _DummyUI()
        self.harris_response_operator = None
        self.eigenvalues = None

    def display_image(self, *a, **k):
        return None


def backend_methods():
    """(self, apply_harris_detector_vectorized, apply_lambda_minus_vectorized) from the Qt class."""
    if "backend" in _cache:
        return _cache["backend"]
    import cv2

    hu, _, _ = app_modules()
    src = open(BACKEND, encoding="utf-8").read()
    tree = ast.parse(src)
    ns = {"np": np, "cv2": cv2, "convert_to_gray": hu.convert_to_gray,
          "convolve2d_optimized": hu.convolve2d_optimized}
    wanted = ("apply_harris_detector_vectorized", "apply_lambda_minus_vectorized")
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "BackendClass":
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in wanted:
                    exec(compile(ast.get_source_segment(src, item), BACKEND, "exec"), ns)
    me = _DummySelf()
    _cache["backend"] = (me, ns[wanted[0]], ns[wanted[1]])
    return _cache["backend"]


def notebook_harris():
    """harris_detection from harris.ipynb with cv2.imread/cvtColor replaced by an array argument."""
    if "nb" in _cache:
        return _cache["nb"]
    import json
    import cv2

    hu, _, _ = app_modules()
    nb = json.load(open(HARRIS_NB, encoding="utf-8"))
    code = None
    for cell in nb["cells"]:
        s = "".join(cell["source"])
        if cell["cell_type"] == "code" and "def harris_detection" in s:
            code = s
    assert code is not None

    class _CV:
        COLOR_BGR2GRAY = 0
        COLOR_GRAY2RGB = 1

        @staticmethod
        def imread(arr):
            return arr  # the "path" is already a gray uint8 array

        @staticmethod
        def cvtColor(a, code_):
            if code_ == _CV.COLOR_BGR2GRAY:
                return a
            return cv2.cvtColor(a, cv2.COLOR_GRAY2RGB)

        rectangle = staticmethod(cv2.rectangle)

    ns = {"np": np, "cv2": _CV, "convolve2d_optimized": hu.convolve2d_optimized, "print": lambda *a, **k: None}
    exec(compile(code, HARRIS_NB, "exec"), ns)
    _cache["nb"] = ns["harris_detection"]
    return _cache["nb"]
