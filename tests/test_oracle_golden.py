"""The oracle against the golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py).
Runs on CPU; this is what pins the oracle."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import featurecraft_oracle as O

CORNER_CASES = ["synthetic_72x88", "synthetic_rgb_65x67", "flat_edges_40x48", "checkerboard_crop", "box_in_scene_crop"]
SIFT_CASES = ["blobs_96x112", "blobs_s3_80x80", "blobs_oct5_sigma2_128x96", "box_crop_96x120"]


@pytest.mark.parametrize("case", CORNER_CASES)
def test_corner_maps_bit_exact(case):
    g = load_golden("corners")
    rgb = g[case + "/rgb"]
    gray = O.convert_to_gray(rgb)
    assert np.array_equal(gray, g[case + "/gray"])
    assert np.array_equal(O.harris_response(gray, 5, 0.04), g[case + "/harris_k0.04"])
    assert np.array_equal(O.harris_response(gray, 5, 0.06), g[case + "/harris_k0.06"])
    assert np.array_equal(O.harris_response(gray, 7, 0.04), g[case + "/harris_w7"])
    assert np.array_equal(O.lambda_minus_response(gray), g[case + "/lambda_minus"])
    corners, r = O.harris_detection(gray, 5, 0.04, 1e8)
    assert np.array_equal(r, g[case + "/notebook_r"])
    assert np.array_equal(np.array([c[:2] for c in corners], dtype=np.int64).reshape(-1, 2), g[case + "/notebook_corners"])


def test_eigvalsh_2x2_restatement():
    g = load_golden("eig2x2")
    assert np.array_equal(O.dsterf_2x2_min(g["a"], g["b"], g["c"]), g["min_eig"])


def test_convolve2d_optimized():
    g = load_golden("convolve2d_optimized")
    for ks in (3, 5, 7):
        assert np.array_equal(O.convolve2d_optimized(g["img"], g["ker%d" % ks]), g["out%d" % ks])
    with pytest.raises(ValueError):
        O.convolve2d_optimized(g["img"], np.ones((4, 4)))


def test_harris_rejects_what_the_reference_rejects():
    with pytest.raises(ValueError):
        O.harris_response(np.zeros((1, 8), np.uint8))
    with pytest.raises(ValueError):
        O.harris_response(np.zeros((8, 8), np.uint8), window_size=4)


@pytest.mark.parametrize("case", SIFT_CASES)
def test_sift_app_variant(case):
    g = load_golden("sift")
    base = g[case + "/base"]
    n_oct, s, sigma, cth, rth = g[case + "/params"]
    n_oct, s = int(n_oct), int(s)
    pyr, dog, kps = O.generate_octaves_pyramid(base, n_oct, s, sigma, cth, rth)
    for o in range(n_oct):
        assert np.array_equal(np.stack(pyr[o]), g["%s/G_oct%d" % (case, o)])
        assert np.array_equal(np.asarray(kps[o], dtype=np.int64).reshape(-1, 3), g["%s/kps_oct%d" % (case, o)])
    masks = O.represent_keypoints(kps, dog)
    ori = O.dog_keypoints_orientations(pyr, masks, sigma, 36, s)
    pts, desc = O.extract_sift_descriptors(pyr, ori, sigma, 8, s)
    assert np.array_equal(np.array(pts, dtype=np.float64).reshape(-1, 5), g[case + "/points"])
    assert np.array_equal(np.array(desc).reshape(-1, 128), g[case + "/descriptors"], equal_nan=True)


@pytest.mark.parametrize("case", SIFT_CASES)
def test_sift_script_variant(case):
    g = load_golden("sift")
    base = g[case + "/base"]
    n_oct, s, sigma, cth, rth = g[case + "/params"]
    if case + "/script_raises_indexerror" in g.files:
        with pytest.raises(IndexError):
            O.generate_octaves_pyramid(base, int(n_oct), int(s), sigma, cth, rth, variant="script")
        return
    _, _, kps = O.generate_octaves_pyramid(base, int(n_oct), int(s), sigma, cth, rth, variant="script")
    for o in range(int(n_oct)):
        assert np.array_equal(np.asarray(kps[o], dtype=np.int64).reshape(-1, 3), g["%s/script_kps_oct%d" % (case, o)])


def test_separable_blur_matches_direct():
    g = load_golden("sift")
    base = g["blobs_96x112/base"]
    for kern in O.generateGaussianKernels(1.6, 2):
        d = O.blur(base, kern, "direct")
        s = O.blur(base, kern, "separable")
        assert np.abs(d - s).max() < 5e-15
    for sigma, kern in zip(O.sigma_levels(1.6, 2), O.generateGaussianKernels(1.6, 2)):
        taps = O.separable_taps(sigma)
        assert taps.shape[0] == kern.shape[0]
        assert np.abs(np.outer(taps, taps) - kern).max() < 4 * np.finfo(np.float64).eps * kern.max()


def test_sift_pieces():
    g = load_golden("sift_pieces")
    _, _, mag, direction = O.sift_gradient(g["img"])
    assert np.array_equal(mag, g["mag"]) and np.array_equal(direction, g["direction"])
    cases = ((10, 12, 5.0), (0, 0, 95.0), (51, 39, 355.0), (25.0, 20.0, 185.0))
    for (j, i, th), w in zip(cases, g["windows"]):
        assert np.array_equal(O.nn_rotated_window(g["img"], (j, i), th), w)


def test_matching_against_bfmatcher_golden():
    g = load_golden("matching")
    idx, dist = O.knn2_l2(g["desc_a"], g["desc_b"])
    assert np.array_equal(idx, g["knn_idx"])
    assert np.abs(dist - g["knn_dist"]).max() < 5e-6
    for t in (0.3, 0.5, 0.8):
        good, _ = O.match_indices(g["desc_a"], g["desc_b"], t)
        assert np.array_equal(good, g["good_%g" % t])
    with pytest.raises(ValueError):
        O.knn2_l2(g["desc_a"], g["desc_b"][:1])
