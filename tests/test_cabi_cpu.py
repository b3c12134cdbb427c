"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports every symbol the
header declares (no compute calls -- there is no GPU here)."""
import ctypes
import re

import pytest

from exvision_featurecraft_b200 import _lib


def header_functions():
    src = open(_lib.HEADER_PATH).read()
    return sorted(set(re.findall(r"FC_API\s+[\w\s\*]+?\b(fc_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    names = header_functions()
    assert len(names) >= 20
    assert sorted(_lib.SIGNATURES) == names


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), name


def test_version_and_status_strings(lib):
    assert lib.fc_version() >= 100
    assert lib.fc_status_string(0) == b"ok"
    assert lib.fc_status_string(-1) == b"bad argument"
    assert lib.fc_launch_count() >= 0


def test_bad_arguments_are_rejected_without_a_gpu(lib):
    null = ctypes.c_void_p(0)
    assert lib.fc_harris_response(null, 1, 8, 8, 5, 0.04, 1.0, null, null, null) == -1
    assert lib.fc_harris_response(ctypes.c_void_p(16), 1, 8, 8, 4, 0.04, 1.0, ctypes.c_void_p(16), null, null) == -1
    assert lib.fc_lambda_minus_response(null, 1, 8, 8, null, null, null) == -1
    assert lib.fc_mask_scan(null, 1, 8, 8, null, null, null) == -1


def test_scan_scratch_is_monotone(lib):
    prev = 0
    for n in (0, 1, 2048, 2049, 10 ** 6, 5 * 10 ** 6):
        v = lib.fc_scan_scratch_elems(n)
        assert v >= prev and v >= n
        prev = v


def test_no_cpu_fallback_in_the_product():
    """The product package must never import the oracle."""
    import os

    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", src, flags=re.M), f


def test_compute_call_without_gpu_raises():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from exvision_featurecraft_b200 import corners

    with pytest.raises(_lib.FeatureCraftError):
        corners.harris(np.zeros((8, 8), np.uint8))


def test_localize_keypoint_mirror_follows_the_script():
    """SIFT.py:225-269 restated literally here (float64, the script's operation order) against the mirror; when the
    reference tree is mounted (authoring container) the script's own function is used as well."""
    import os

    import numpy as np

    from exvision_featurecraft_b200.utils.SIFT_scale_space import localize_keypoint

    rng = np.random.default_rng(2)
    D = rng.standard_normal((9, 11, 4)).astype(np.float32)
    funcs = []
    from oracle import ref_loader

    if ref_loader.available():
        funcs.append(ref_loader.script_namespace()["localize_keypoint"])
    for y, x, s in [(1, 1, 0), (4, 7, 2), (7, 9, 3)]:
        H, xo, yo, so = localize_keypoint(D, x, y, s)
        Dd = D.astype(np.float64)
        dxx = Dd[y, x + 1, s] - 2 * Dd[y, x, s] + Dd[y, x - 1, s]
        dyy = Dd[y + 1, x, s] - 2 * Dd[y, x, s] + Dd[y - 1, x, s]
        dxy = ((Dd[y + 1, x + 1, s] - Dd[y + 1, x - 1, s]) - (Dd[y - 1, x + 1, s] - Dd[y - 1, x - 1, s])) / 4
        assert (xo, yo, so) == (x, y, s)
        assert np.array_equal(H, np.array([[dxx, dxy], [dxy, dyy]]))
        for f in funcs:
            Hr, xr, yr, sr = f(D, x, y, s)
            assert np.array_equal(H, Hr) and (xr, yr, sr) == (x, y, s)
