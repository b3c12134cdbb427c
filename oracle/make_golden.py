"""Generate tests/golden/*.npz by running the UNMODIFIED reference in place.  TEST INFRASTRUCTURE.

Run in the authoring container only (needs /root/reference):

    python oracle/make_golden.py

For every fixture the script (1) executes the reference's own functions (oracle/ref_loader.py
explains the three unavoidable patches), (2) asserts that oracle/featurecraft_oracle.py reproduces
the reference output (bit-exact for integer / index data and for the float maps, where both sides
perform the same IEEE operations), and (3) stores inputs + reference outputs.  The GPU box has no
/root/reference; tests there check the oracle and the CUDA path against these files.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import featurecraft_oracle as O  # noqa: E402
from oracle import ref_loader as RL  # noqa: E402
from exvision_featurecraft_b200 import synthetic as S  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(RL.APP_DIR, "assets", "test_images")


def _save(name, **arrays):
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KiB" % (name + ".npz", os.path.getsize(path) / 1024))


def corners():
    import cv2

    me, harris, lam = RL.backend_methods()
    nb_harris = RL.notebook_harris()
    cases = {
        "synthetic_72x88": np.stack([S.corner_image(72, 88, seed=11)] * 3, -1),
        "synthetic_rgb_65x67": np.random.default_rng(12).integers(0, 256, (65, 67, 3)).astype(np.uint8),
        "flat_edges_40x48": None,
        "checkerboard_crop": cv2.imread(os.path.join(ASSETS, "checkerboard.jpg"))[40:136, 60:188, ::-1],
        "box_in_scene_crop": cv2.imread(os.path.join(ASSETS, "box_in_scene.png"))[100:196, 200:312, ::-1],
    }
    flat = np.zeros((40, 48, 3), np.uint8)
    flat[10:25, 12:40] = 200
    flat[30:33, :] = 77
    flat[:, 5] = 255
    cases["flat_edges_40x48"] = flat
    out = {}
    for name, rgb in cases.items():
        rgb = np.ascontiguousarray(rgb)
        gray = RL.app_modules()[0].convert_to_gray(rgb)
        assert np.array_equal(gray, O.convert_to_gray(rgb))
        for k, thr in ((0.04, 10000), (0.06, 2.5e5)):
            corner_list, _ = harris(me, rgb.copy(), 5, k, thr)
            resp = me.harris_response_operator
            assert np.array_equal(resp, O.harris_response(gray, 5, k)), name
            idx, vals = O.threshold_corners(resp, thr)
            assert [c[:2] for c in corner_list] == [(x, y) for y, x in idx.tolist()]
            out["%s/harris_k%g" % (name, k)] = resp
        corner7, _ = harris(me, rgb.copy(), 7, 0.04, 10000)
        assert np.array_equal(me.harris_response_operator, O.harris_response(gray, 7, 0.04))
        out[name + "/harris_w7"] = me.harris_response_operator
        lam(me, rgb.copy())
        eig = me.eigenvalues
        assert np.array_equal(eig, O.lambda_minus_response(gray)), name
        out[name + "/lambda_minus"] = eig
        cl, _ = nb_harris(gray, 5, 0.04, 1e8)
        clo, r_nb = O.harris_detection(gray, 5, 0.04, 1e8)
        assert cl == clo, name
        out[name + "/notebook_r"] = r_nb
        out[name + "/notebook_corners"] = np.array([[c[0], c[1]] for c in cl], dtype=np.int64).reshape(-1, 2)
        out[name + "/rgb"] = rgb
        out[name + "/gray"] = gray
    _save("corners", **out)

    # LAPACK 2x2 path: eigvalsh itself on adversarial inputs
    rng = np.random.default_rng(0)
    n = 20000
    a = rng.uniform(0, 1e6, n)
    c = rng.uniform(0, 1e6, n)
    b = rng.normal(0, 1, n) * 10.0 ** rng.uniform(-20, 6, n)
    a[:200] = 0
    c[100:300] = 0
    b[::7] = 0
    a[300:400] = c[300:400]
    ev = np.linalg.eigvalsh(np.stack([a, b, b, c], -1).reshape(-1, 2, 2)).min(-1)
    assert np.array_equal(ev, O.dsterf_2x2_min(a, b, c))
    _save("eig2x2", a=a, b=b, c=c, min_eig=ev)


def convolve():
    hu = RL.app_modules()[0]
    rng = np.random.default_rng(3)
    img = rng.standard_normal((23, 31))
    out = {"img": img}
    for ks in (3, 5, 7):
        ker = rng.standard_normal((ks, ks))
        ref = hu.convolve2d_optimized(img, ker)
        assert np.array_equal(ref, O.convolve2d_optimized(img, ker))
        out["ker%d" % ks] = ker
        out["out%d" % ks] = ref
    assert np.array_equal(hu.padding_matrix(img, 31, 23, 2), O.padding_matrix(img, 31, 23, 2))
    _save("convolve2d_optimized", **out)


def _kps_arrays(kps):
    return [np.asarray(k, dtype=np.int64).reshape(-1, 3) for k in kps]


def sift():
    hu, ss, kd = RL.app_modules()
    ns = RL.script_namespace()
    # Gaussian kernels
    for sigma, s in ((1.6, 2), (2.0, 3), (3.0, 5)):
        for a, b in zip(ss.generateGaussianKernels(sigma, s), O.generateGaussianKernels(sigma, s)):
            assert np.array_equal(a, b)
    out = {}
    cases = {
        "blobs_96x112": (S.sift_image(96, 112, seed=5), 4, 2, 1.6),
        "blobs_s3_80x80": (S.sift_image(80, 80, seed=6), 4, 3, 1.6),
        "blobs_oct5_sigma2_128x96": (S.sift_image(128, 96, seed=8), 5, 2, 2.0),
    }
    import cv2

    box = cv2.imread(os.path.join(ASSETS, "box_in_scene.png"), cv2.IMREAD_GRAYSCALE)[120:216, 180:300]
    cases["box_crop_96x120"] = (box.astype(np.float64) / 255.0, 4, 2, 1.6)
    for name, (base, n_oct, s, sigma) in cases.items():
        pyr, dog, kps = ss.generate_octaves_pyramid(base, n_oct, s, sigma, 0.03, 10)
        pyo, dogo, kpso = O.generate_octaves_pyramid(base, n_oct, s, sigma, 0.03, 10)
        for A, B in zip(pyr, pyo):
            for a, b in zip(A, B):
                assert np.array_equal(a, b)
        for a, b in zip(_kps_arrays(kps), _kps_arrays(kpso)):
            assert np.array_equal(a, b), name
        masks = kd.represent_keypoints(kps, dog)
        for A, B in zip(masks, O.represent_keypoints(kpso, dogo)):
            for a, b in zip(A, B):
                assert np.array_equal(a, b)
        ori = kd.dog_keypoints_orientations(pyr, masks, sigma, 36, s)
        assert ori == O.dog_keypoints_orientations(pyo, masks, sigma, 36, s), name
        pts, desc = kd.extract_sift_descriptors(pyr, ori, sigma, 8, s)
        pts_o, desc_o = O.extract_sift_descriptors(pyo, ori, sigma, 8, s)
        assert pts == pts_o
        desc = np.array(desc, dtype=np.float64).reshape(-1, 128)
        assert np.array_equal(desc, np.array(desc_o).reshape(-1, 128), equal_nan=True), name
        out[name + "/base"] = base
        out[name + "/params"] = np.array([n_oct, s, sigma, 0.03, 10.0])
        for o in range(n_oct):
            out["%s/kps_oct%d" % (name, o)] = _kps_arrays(kps)[o]
            out["%s/G_oct%d" % (name, o)] = np.stack(pyr[o]).astype(np.float64)
        out[name + "/points"] = np.array(pts, dtype=np.float64).reshape(-1, 5)
        out[name + "/descriptors"] = desc
        # script variant (contrast + edge filter on); s>=3 raises IndexError in the reference
        try:
            _, _, kps_s = ns["generate_octaves_pyramid"](base, n_oct, s, sigma, 0.03, 10)
            _, _, kps_so = O.generate_octaves_pyramid(base, n_oct, s, sigma, 0.03, 10, variant="script")
            for o, (a, b) in enumerate(zip(_kps_arrays(kps_s), _kps_arrays(kps_so))):
                assert np.array_equal(a, b), name
                out["%s/script_kps_oct%d" % (name, o)] = a
        except IndexError:
            try:
                O.generate_octaves_pyramid(base, n_oct, s, sigma, 0.03, 10, variant="script")
                raise AssertionError("oracle should raise IndexError like the reference")
            except IndexError:
                out[name + "/script_raises_indexerror"] = np.array([1])
    _save("sift", **out)

    # sift_gradient / rotated window pieces
    img = S.sift_image(40, 52, seed=9)
    ref = kd.sift_gradient(img)
    for a, b in zip(ref, O.sift_gradient(img)):
        assert np.array_equal(a, b)
    wins = []
    for (j, i, th) in ((10, 12, 5.0), (0, 0, 95.0), (51, 39, 355.0), (25.0, 20.0, 185.0)):
        w = RL.fixed_point_rotated_subimage(img, (j, i), th, 16, 16)
        assert np.array_equal(w, O.nn_rotated_window(img, (j, i), th))
        wins.append(w)
    assert np.array_equal(kd.get_gaussian_mask(3.39, 16), O.get_gaussian_mask(3.39, 16))
    # kp_list_2_opencv_kp_list (:314-327): the cv2.KeyPoint fields against the oracle's tuple twin
    kl = [(3, 7, 0, 1, 45.0), (10, 2, 3, 2, 355.0), (0, 0, 1, 1, 5.0)]
    for kp, (x, y, size, angle) in zip(kd.kp_list_2_opencv_kp_list(kl), O.kp_list_2_xy(kl)):
        assert (kp.pt[0], kp.pt[1], kp.size, kp.angle) == (np.float32(x), np.float32(y), np.float32(size), np.float32(angle))
    # localize_keypoint (:213-261, dead code in the app): the oracle twin against the reference's own function
    rngl = np.random.default_rng(21)
    Dl = rngl.standard_normal((9, 11, 4))
    for (yy, xx, sl) in ((1, 1, 1), (4, 7, 2), (7, 9, 1)):
        a = ss.localize_keypoint(Dl, xx, yy, sl)
        b = O.localize_keypoint_full(Dl, xx, yy, sl)
        assert all(np.array_equal(u, v) for u, v in zip(a[:3], b[:3])) and a[3:] == b[3:]
    # sift_resize / rescale: the reference functions run through the scipy stub (ref_loader) -- this pins the ORACLE's
    # restatement to that stub, not to scikit-image (absent): documented as unpinned
    rgb = np.random.default_rng(4).random((50, 70, 3))
    a, ra = kd.sift_resize(rgb)
    b, rb = O.sift_resize(rgb)
    assert ra == rb and np.array_equal(a, b)
    _save("sift_pieces", img=img, mag=ref[2], direction=ref[3], windows=np.stack(wins))


def validate_fixed_point_warp():
    """The fixed-point restatement against cv2's (working) float32 warpAffine path."""
    import cv2
    from math import cos, sin

    rng = np.random.default_rng(1)
    img = rng.random((60, 70)).astype(np.float32)
    bad = 0
    total = 0
    for _ in range(300):
        j, i = int(rng.integers(0, 70)), int(rng.integers(0, 60))
        th = (int(rng.integers(0, 36)) + 0.5) * 10.0
        t = th * (3.14159 / 180)
        vx = (cos(t), sin(t))
        vy = (-sin(t), cos(t))
        sx = j - vx[0] * 7.5 - vy[0] * 7.5
        sy = i - vx[1] * 7.5 - vy[1] * 7.5
        M = np.array([[vx[0], vy[0], sx], [vx[1], vy[1], sy]])
        ref = cv2.warpAffine(img, M, (16, 16), flags=cv2.INTER_NEAREST + cv2.WARP_INVERSE_MAP,
                             borderMode=cv2.BORDER_CONSTANT)
        mine = O.nn_rotated_window(img, (j, i), th)
        bad += int((ref != mine).sum())
        total += ref.size
    print("fixed-point NN warp vs cv2 float32 path: %d / %d samples differ" % (bad, total))
    assert bad == 0


def matching():
    import cv2

    base = S.sift_image(96, 112, seed=5)
    tmpl = np.ascontiguousarray(np.rot90(base[20:76, 30:86]))
    _, da = O.compute_keypoints_and_descriptors_from_base(base, 4, 2, 1.6, 0.03, 10)
    _, db = O.compute_keypoints_and_descriptors_from_base(tmpl, 4, 2, 1.6, 0.03, 10)
    a = np.array(da, dtype=np.float32)
    b = np.array(db, dtype=np.float32)
    m = cv2.BFMatcher().knnMatch(a, b, k=2)
    knn_idx = np.array([[x.trainIdx, y.trainIdx] for x, y in m], dtype=np.int64)
    knn_dist = np.array([[x.distance, y.distance] for x, y in m], dtype=np.float32)
    idx, dist = O.knn2_l2(a, b)
    assert np.array_equal(idx, knn_idx)
    assert np.abs(dist - knn_dist).max() < 5e-6
    out = {"desc_a": a, "desc_b": b, "knn_idx": knn_idx, "knn_dist": knn_dist}
    for t in (0.3, 0.5, 0.8):
        good = np.array([(x.queryIdx, x.trainIdx) for x, y in m if x.distance < t * y.distance], dtype=np.int64).reshape(-1, 2)
        go, _ = O.match_indices(a, b, t)
        assert np.array_equal(good, go)
        out["good_%g" % t] = good
    _save("matching", **out)


if __name__ == "__main__":
    assert RL.available(), "reference not mounted at " + RL.REF_ROOT
    os.makedirs(GOLDEN, exist_ok=True)
    corners()
    convolve()
    validate_fixed_point_warp()
    sift()
    matching()
    print("all reference-vs-oracle assertions passed")
