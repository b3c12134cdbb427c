"""ctypes binding of libfeaturecraft_b200.so (the C ABI declared in include/featurecraft_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
Build it with ``python exvision_featurecraft_b200/csrc/build.py`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfeaturecraft_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "featurecraft_b200.h")

FC_OK = 0
FC_F32, FC_F64, FC_F16 = 0, 1, 2
FC_FILTER_APP, FC_FILTER_SCRIPT = 0, 1
FC_GRAD_CENTRAL, FC_GRAD_KX, FC_GRAD_SOBEL = 0, 1, 2
FC_WINDOW_BOX, FC_WINDOW_GAUSS = 0, 1
FC_RESP_HARRIS, FC_RESP_LAMBDA_MINUS = 0, 1


class FeatureCraftError(RuntimeError):
    pass


class CapacityError(FeatureCraftError):
    """FC_ERR_CAPACITY: a caller-provided output buffer is too small; `required` is the count that would fit."""

    status = -2

    def __init__(self, what: str, required: int, capacity: int):
        super().__init__("%s: output capacity too small (%d needed, %d available)" % (what, required, capacity))
        self.required, self.capacity = int(required), int(capacity)


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_d = C.c_double

#: name -> (restype, argtypes); must list every function the header declares
SIGNATURES = {
    "fc_version": (_i, []),
    "fc_status_string": (C.c_char_p, [_i]),
    "fc_last_cuda_error": (_i, []),
    "fc_launch_count": (C.c_ulonglong, []),
    "fc_path_count": (C.c_ulonglong, [_i]),
    "fc_rgb_to_gray_u8": (_i, [_p, _i, _i, _i, _p, _p]),
    "fc_correlate_zero_pad_f64": (_i, [_p, _i, _i, _i, _p, _i, _p, _p]),
    "fc_harris_response": (_i, [_p, _i, _i, _i, _i, _d, _d, _p, _p, _p]),
    "fc_notebook_harris_response": (_i, [_p, _i, _i, _i, _i, _d, _p, _p]),
    "fc_lambda_minus_response": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "fc_lambda_minus_response_fast": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "fc_corner_response_ex": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _i, _d, _i, _d, _p, _p, _p]),
    "fc_nms3x3_mask": (_i, [_p, _i, _i, _i, _d, _p, _p]),
    "fc_selftest_div25": (_i, [_p, _p]),
    "fc_selftest_divsqrt": (_i, [_p, _p]),
    "fc_threshold_mask_abs": (_i, [_p, _i, _i, _i, _d, _i, _p, _p]),
    "fc_threshold_mask_rel": (_i, [_p, _i, _i, _i, _d, _p, _p, _p]),
    "fc_scan_scratch_elems": (_i64, [_i64]),
    "fc_mask_scan": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "fc_publish_i32": (_i, [_p, _p, _i, _p]),
    "fc_mask_row_count": (_i, [_p, _i, _i, _i, _p, _p]),
    "fc_popcount_u64": (_i, [_p, _i64, _p, _p]),
    "fc_exclusive_scan_i32": (_i, [_p, _i64, _p, _p, _p]),
    "fc_publish_gather_i32": (_i, [_p, _p, _i, _p, _p]),
    "fc_mask_emit": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "fc_mask_emit_keypoints": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "fc_gauss_octave": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _d, _d, _p, _p, _p]),
    "fc_gauss_octave_extrema": (_i, [_p, _i, _i, _i, _i, _i, _i, _d, _d, _p, _p]),
    "fc_dog_extrema": (_i, [_p, _i, _i, _i, _i, _i, _i, _d, _d, _p, _p]),
    "fc_upsample2": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "fc_refine_keypoints": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _d, _d, _p, _p, _p, _p, _p]),
    "fc_rotated_window": (_i, [_p, _i, _i, _d, _d, _d, _d, _i, _i, _p, _p]),
    "fc_downsample2": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "fc_convert_f64": (_i, [_p, _i64, _i, _p, _p]),
    "fc_gradient_field": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "fc_orientation_bins": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _p, _i, _i, _i, _p, _i, _p]),
    "fc_sift_layout": (_i, [_p, _p, _p, _p, _i, _i, _p, _p, _p]),
    "fc_orientation_bins_coded": (_i, [_p, _i, _i, _i, _p, _i, _p, _i, _i, _i, _d, _p, _p, _p, _p]),
    "fc_orientation_recheck": (_i, [_p, _i, _i, _p, _i, _i, _i, _p, _i, _p, _i, _i, _i, _p, _p, _p, _i, _p]),
    "fc_gradient_codes": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "fc_descriptors_coded": (_i, [_p, _i, _i, _i, _p, _i, _p, _p, _p, _i, _p, _p]),
    "fc_convert_f64_bound": (_i, [_p, _i64, _p, _p, _p]),
    "fc_upsample2_dual": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "fc_downsample2_dual": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "fc_gauss_levels_f64": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p]),
    "fc_gauss_octave_extrema_certified": (_i, [_p, _i, _i, _i, _i, _i, _d, _d, _p, _d, _p, _p, _i, _p, _p]),
    "fc_gauss_octave_search": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _d, _p, _p, _i, _p, _p]),
    "fc_extrema_repair": (_i, [_p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _p, _i, _d, _d, _p, _p, _i, _p, _p]),
    "fc_orientation_scan": (_i, [_p, _i, _p, _p, _p]),
    "fc_orientation_emit": (_i, [_p, _p, _p, _i, _p, _p]),
    "fc_descriptors": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _p, _p, _p, _i, _p, _p]),
    "fc_bgr_to_rgb_u8": (_i, [_p, _i, _i, _i, _p, _p]),
    "fc_rgb_to_gray_f64": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "fc_u8_to_unit_f64": (_i, [_p, _i64, _p, _p]),
    "fc_blur_axis_f64": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _p, _p]),
    "fc_zoom_linear_f64": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "fc_match_workspace_bytes": (_i64, [_i, _i]),
    "fc_match_workspace_stats_offset": (_i64, [_i, _i]),
    "fc_match_knn2": (_i, [_p, _i, _p, _i, _d, _p, _p, _p, _p, _p]),
    "fc_match_knn2_exact": (_i, [_p, _i, _p, _i, _d, _p, _p, _p, _p]),
}



class DescPlacement(C.Structure):
    """fc_desc_placement of include/featurecraft_b200.h"""

    _fields_ = [("remap", C.c_void_p), ("points", C.c_void_p), ("list_base", C.c_int), ("octave", C.c_int), ("scale", C.c_int)]


_lib = None


def load() -> C.CDLL:
    """Load the shared library once; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FeatureCraftError(
            "libfeaturecraft_b200.so is not built (%s). Run `python exvision_featurecraft_b200/csrc/build.py`; "
            "this package has no CPU or PyTorch fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != FC_OK:
        lib = load()
        msg = lib.fc_status_string(status).decode()
        extra = ""
        if status == -3:
            extra = " (cudaError %d)" % lib.fc_last_cuda_error()
        raise FeatureCraftError("%s failed: %s%s" % (what, msg, extra))


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args), name)


def launch_count() -> int:
    return int(load().fc_launch_count())


PATHS = ("packed_blur", "fused_search", "mid_f64", "generic_octave", "repair_patch")  # fc_path enumerators


def path_counts() -> dict:
    """How often each scale-space kernel family has been launched in this process (fc_path_count)."""
    lib = load()
    return {name: int(lib.fc_path_count(i)) for i, name in enumerate(PATHS)}
