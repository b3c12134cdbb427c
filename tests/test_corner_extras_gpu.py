"""The opt-in components north_star names and the reference does not use (SURVEY.md trap 1): Sobel gradients, Gaussian
window, 3x3 non-maximum suppression, sub-pixel refinement.  They default OFF everywhere; with the reference's own
components selected the opt-in kernel must equal the reference kernels bit for bit, and the extras are held to their
library definitions (cv2.Sobel / cv2.GaussianBlur / cv2.dilate, and the reference's dead localize_keypoint code)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import featurecraft_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C(lib):
    from exvision_featurecraft_b200 import corners

    return corners


def _gray(shape, seed):
    from exvision_featurecraft_b200 import synthetic

    return synthetic.corner_image(shape[0], shape[1], seed)


@pytest.mark.parametrize("shape", [(72, 88), (65, 67), (130, 33)])
def test_reference_components_equal_the_reference_kernels(C, shape):
    gray = _gray(shape, 3)
    for ws in (3, 5, 7):
        a = C.corner_response(gray, "central", "box", ws, response="harris", k=0.04).response[0].cpu().numpy()
        assert np.array_equal(a, O.harris_response(gray, ws, 0.04))
    lam = C.corner_response(gray, "kx", "box", 5, response="lambda_minus").response[0].cpu().numpy()
    assert np.array_equal(lam, O.lambda_minus_response(gray))
    _, r_nb = O.harris_detection(gray, 5, 0.04, 1e8)
    nb = C.corner_response(gray, "kx", "box", 5, response="harris", k=0.04).response[0].cpu().numpy()
    assert np.abs(nb - r_nb).max() <= 1e-12 * np.abs(r_nb).max()  # the notebook squares with libm pow


@pytest.mark.parametrize("shape,ws,sigma", [((72, 88), 5, 1.0), ((65, 67), 7, 1.5), ((40, 130), 3, 0.8)])
def test_sobel_gaussian_window_against_cv2(C, shape, ws, sigma):
    import cv2

    gray = _gray(shape, 5)
    gx = cv2.Sobel(gray, cv2.CV_64F, 1, 0, ksize=3)
    gy = cv2.Sobel(gray, cv2.CV_64F, 0, 1, ksize=3)
    sxx = cv2.GaussianBlur(gx * gx, (ws, ws), sigma)
    sxy = cv2.GaussianBlur(gx * gy, (ws, ws), sigma)
    syy = cv2.GaussianBlur(gy * gy, (ws, ws), sigma)
    ref_h = (sxx * syy - sxy * sxy) - 0.04 * (sxx + syy) ** 2
    got_h = C.corner_response(gray, "sobel", "gaussian", ws, sigma, "harris", 0.04).response[0].cpu().numpy()
    assert np.abs(got_h - ref_h).max() <= 1e-11 * np.abs(ref_h).max()
    ref_l = 0.5 * (sxx + syy) - np.sqrt((0.5 * (sxx - syy)) ** 2 + sxy ** 2)
    got_l = C.corner_response(gray, "sobel", "gaussian", ws, sigma, "lambda_minus").response[0].cpu().numpy()
    assert np.abs(got_l - ref_l).max() <= 1e-9 * np.abs(sxx + syy).max()
    # Sobel + box window: zero-padded ones window over the Sobel products
    box = C.corner_response(gray, "sobel", "box", ws, response="harris", k=0.04).response[0].cpu().numpy()
    bxx, bxy, byy = (O._box_sum_zero_pad(p, ws // 2) for p in (gx * gx, gx * gy, gy * gy))
    assert np.array_equal(box, (bxx * byy - bxy ** 2) - 0.04 * ((bxx + byy) ** 2))


def test_nms_is_a_3x3_dilation_test(C):
    import cv2

    gray = np.stack([_gray((96, 120), 1), _gray((96, 120), 2)])
    res = C.harris(gray, 5, 0.04, 10000)
    for fn in ("resident", "fused"):
        if fn == "resident":
            coords, values, offs = C.nms_threshold(res.response, 10000)
        else:
            r = C.corner_response(gray, "central", "box", 5, None, "harris", 0.04, threshold=10000, nms=True)
            coords, values, offs = r.coords, r.values, r.image_offsets
        offs = offs.cpu().numpy()
        for b in range(2):
            R = res.response[b].cpu().numpy()
            padded = cv2.copyMakeBorder(R, 1, 1, 1, 1, cv2.BORDER_CONSTANT, value=-np.inf)
            dil = cv2.dilate(padded, np.ones((3, 3), np.uint8))[1:-1, 1:-1]
            ref = np.argwhere((R > 10000) & (R >= dil))
            got = coords[offs[b]:offs[b + 1]].cpu().numpy()
            assert np.array_equal(got, ref) and len(ref) >= 3
            assert len(ref) < (R > 10000).sum()  # it suppresses something: the reference's plain threshold keeps more


def test_subpixel_refinement_equals_localize_keypoint(lib):
    """fc_refine_keypoints against the app's dead localize_keypoint code + its commented-out acceptance tests."""
    from exvision_featurecraft_b200 import sift

    g = load_golden("sift")
    case = "blobs_96x112"
    p = sift.SiftParams(dtype="f64")
    octs = sift.scale_space(g[case + "/base"], p)
    checked = kept = 0
    for o, octv in enumerate(octs[:2]):
        G = octv.gauss[0].cpu().numpy()
        D = np.stack([G[l + 1] - G[l] for l in range(G.shape[0] - 1)], axis=2)
        for k in (1, 2):
            kps = sift.keypoints_from_mask(octv.extrema[k - 1], octv.width)
            out = sift.refine_keypoints(octv, k, kps, 0.03, 10.0)
            off, con, rat, keep = (out[n].cpu().numpy() for n in ("offsets", "contrast", "ratio", "keep"))
            for i, (_, y, x) in enumerate(kps.cpu().numpy().tolist()):
                ro, rc, rr, rk = O.refine_keypoint(D, x, y, k, 0.03, 10.0)
                scale = max(1.0, np.abs(ro).max())
                assert np.abs(off[i] - ro).max() <= 1e-8 * scale
                assert abs(con[i] - rc) <= 1e-10 + 1e-8 * abs(rc)
                assert (np.isnan(rat[i]) and np.isnan(rr)) or abs(rat[i] - rr) <= 1e-8 * abs(rr)
                decisive = abs(np.max(ro) - 0.5) > 1e-6 and abs(abs(rc) - 0.03) > 1e-9 and not abs(rr - 10.0) < 1e-6
                if decisive:
                    assert bool(keep[i]) == rk
                checked += 1
                kept += int(keep[i])
    assert checked > 200 and 0 < kept < checked
