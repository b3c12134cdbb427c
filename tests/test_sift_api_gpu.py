"""The reference-shaped SIFT functions (utils/SIFT_scale_space.py, utils/keypoint_descriptor.py mirrors)
against the oracle and the golden vectors, called the way the reference's own drivers call them."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import featurecraft_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def SS(lib):
    from exvision_featurecraft_b200.utils import SIFT_scale_space

    SIFT_scale_space.set_dtype("f64")
    SIFT_scale_space.set_variant("app")
    return SIFT_scale_space


@pytest.fixture(scope="module")
def KD(SS):
    from exvision_featurecraft_b200.utils import keypoint_descriptor

    return keypoint_descriptor


def test_kernel_tables(SS):
    for sigma, s in ((1.6, 2), (2.0, 3), (3.0, 5)):
        for a, b in zip(SS.generateGaussianKernels(sigma, s), O.generateGaussianKernels(sigma, s)):
            assert np.array_equal(a, b)


def test_generate_octaves_pyramid_mirror(SS):
    g = load_golden("sift")
    case = "blobs_96x112"
    pyr, dog, kps = SS.generate_octaves_pyramid(g[case + "/base"], 4, 2, 1.6, 0.03, 10)
    for o in range(4):
        assert np.abs(np.stack(pyr[o]) - g["%s/G_oct%d" % (case, o)]).max() < 2e-14
        assert len(dog[o]) == 4
        assert np.array_equal(np.asarray(kps[o], dtype=np.int64).reshape(-1, 3), g["%s/kps_oct%d" % (case, o)])


def test_get_keypoints_direct_both_variants(SS):
    g = load_golden("sift")
    base = g["box_crop_96x120/base"]
    _, dogs, _ = O.generate_octaves_pyramid(base, 1, 2, 1.6, 0.03, 10)
    d = dogs[0]
    full = np.concatenate([x[:, :, None] for x in d], axis=2)
    for k in (1, 2):
        got = SS.get_keypoints(d[k - 1:k + 2], k, 0.03, 10, full[:, :, :k + 2])
        ref = O.get_keypoints(d[k - 1:k + 2], k, 0.03, 10, d[:k + 2])
        assert np.array_equal(np.asarray(got).reshape(-1, 3), np.asarray(ref).reshape(-1, 3))
    SS.set_variant("script")
    try:
        for k in (1, 2):
            got = SS.get_keypoints(d[k - 1:k + 2], k, 0.03, 10, full)
            ref = O.get_keypoints(d[k - 1:k + 2], k, 0.03, 10, d, variant="script")
            assert np.array_equal(np.asarray(got).reshape(-1, 3), np.asarray(ref).reshape(-1, 3))
        with pytest.raises(IndexError):
            SS.get_keypoints(d[1:4], 3, 0.03, 10, full)
    finally:
        SS.set_variant("app")


def test_orientation_and_descriptor_mirrors(SS, KD):
    g = load_golden("sift")
    case = "box_crop_96x120"
    base = g[case + "/base"]
    pyr, dog, kps = O.generate_octaves_pyramid(base, 4, 2, 1.6, 0.03, 10)
    masks = KD.represent_keypoints(kps, dog)
    for A, B in zip(masks, O.represent_keypoints(kps, dog)):
        for a, b in zip(A, B):
            assert np.array_equal(a, b)
    ori = KD.dog_keypoints_orientations(pyr, masks, 1.6, 36, 2)
    ref_ori = O.dog_keypoints_orientations(pyr, masks, 1.6, 36, 2)
    assert ori == ref_ori
    pts, desc = KD.extract_sift_descriptors(pyr, ori, 1.6, 8, 2)
    assert np.array_equal(np.array(pts, dtype=np.float64).reshape(-1, 5), g[case + "/points"])
    ref_d = g[case + "/descriptors"]
    desc = np.array(desc).reshape(-1, 128)
    ok = ~np.isnan(ref_d).any(1)
    assert np.array_equal(np.isnan(desc).any(1), ~ok)
    assert np.abs(desc[ok] - ref_d[ok]).max() < 1e-9


def test_small_pieces(KD):
    g = load_golden("sift_pieces")
    gx, gy, mag, direction = KD.sift_gradient(g["img"])
    rgx, rgy, rmag, rdir = O.sift_gradient(g["img"])
    assert np.array_equal(gx, rgx) and np.array_equal(gy, rgy)
    assert np.abs(mag - g["mag"]).max() < 1e-15
    dd = np.abs(direction - g["direction"])
    assert np.minimum(dd, 360 - dd).max() < 1e-12
    cases = ((10, 12, 5.0), (0, 0, 95.0), (51, 39, 355.0), (25.0, 20.0, 185.0))
    for (j, i, th), w in zip(cases, g["windows"]):
        assert np.array_equal(KD.rotated_subimage(g["img"], (j, i), th, 16, 16), w)
    assert np.array_equal(KD.rotated_subimage(g["img"], (7, 9), 33.0, 5, 9), O.nn_rotated_window(g["img"], (7, 9), 33.0, 5, 9))
    assert np.array_equal(KD.get_gaussian_mask(3.39, 16), O.get_gaussian_mask(3.39, 16))
    sl = [-3, 9, 45, 60]
    assert np.array_equal(KD.padded_slice(g["img"], sl), O.padded_slice(g["img"], sl))
    kl = [(3, 4, 0, 1, 15.0), (10, 2, 3, 2, 355.0)]
    assert KD.kp_list_2_xy(kl) == O.kp_list_2_xy(kl)
    assert [(k.pt[0], k.pt[1], k.size, k.angle) for k in KD.kp_list_2_opencv_kp_list(kl)] == \
        [tuple(np.float32(v) for v in t) for t in O.kp_list_2_xy(kl)]
    rgb = np.random.default_rng(0).random((9, 11, 3))
    assert np.array_equal(KD.convert_to_grayscale(rgb), O.convert_to_grayscale(rgb))


def test_upsample_and_whole_driver(KD):
    from scipy import ndimage

    from exvision_featurecraft_b200 import sift, synthetic

    img = synthetic.sift_image(41, 57, seed=4)
    up = sift.upsample2(img, sift.SiftParams(dtype="f64"))[0].cpu().numpy()
    ref_up = ndimage.zoom(img, 2, order=1, mode="mirror", grid_mode=True)
    assert up.shape == ref_up.shape and np.abs(up - ref_up).max() < 4e-16
    pts, desc = KD.computeKeypointsAndDescriptors(img, 4, 2, 1.6, 0.03, 10)
    rp, rd = O.compute_keypoints_and_descriptors_from_base(up, 4, 2, 1.6, 0.03, 10)
    assert pts == [tuple(x) for x in rp]
    rd = np.array(rd).reshape(-1, 128)
    ok = ~np.isnan(rd).any(1)
    assert np.abs(np.array(desc).reshape(-1, 128)[ok] - rd[ok]).max() < 1e-9


@pytest.mark.parametrize("engine", ["exact", "tensor"])
def test_matching_golden(KD, engine):
    from exvision_featurecraft_b200 import matching

    g = load_golden("matching")
    idx, dist, good = matching.knn2(g["desc_a"], g["desc_b"], 0.5, engine)
    if engine == "exact":
        assert np.array_equal(idx.cpu().numpy(), g["knn_idx"])
        assert np.abs(dist.cpu().numpy() - g["knn_dist"]).max() < 5e-6
    for t in (0.3, 0.5, 0.8):
        pairs, d = matching.match_indices(g["desc_a"], g["desc_b"], t, engine)
        assert np.array_equal(pairs, g["good_%g" % t])
    with pytest.raises(ValueError):
        matching.knn2(g["desc_a"], g["desc_b"][:1], 0.3, engine)
    m = KD.match_good(g["desc_a"], g["desc_b"], 0.8)
    assert [x[:2] for x in m] == [tuple(r) for r in g["good_0.8"].tolist()]


@pytest.mark.parametrize("engine", ["exact", "tensor"])
def test_matching_random_vs_oracle(engine):
    from exvision_featurecraft_b200 import matching

    rng = np.random.default_rng(5)
    for nq, nt in ((1, 2), (70, 130), (300, 257), (513, 1000)):
        b = rng.random((nt, 128)).astype(np.float32)
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        a = b[rng.integers(0, nt, nq)] + rng.normal(0, 0.02, (nq, 128)).astype(np.float32)
        a[::7] = rng.random((len(a[::7]), 128)).astype(np.float32)
        if nt > 10:
            b[5] = b[3]  # exact duplicate train rows: ties go to the lower index
        ref_idx, ref_dist = O.knn2_l2(a, b)
        idx, dist, good = matching.knn2(a, b, 0.7, engine)
        idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
        ref_good = ref_dist[:, 0].astype(np.float64) < 0.7 * ref_dist[:, 1].astype(np.float64)
        assert np.array_equal(good.cpu().numpy(), ref_good)
        assert np.array_equal(idx[ref_good, 0], ref_idx[ref_good, 0])
        if engine == "exact":
            assert np.array_equal(idx, ref_idx)
            assert np.array_equal(dist, ref_dist)


def test_batch_matcher_pipeline_equals_single_calls():
    """matching.BatchMatcher (host -> device -> host pipeline over a stream of pairs) returns, pair by pair, what
    match_indices returns for that pair alone -- with pairs in flight overlapping each other."""
    from exvision_featurecraft_b200 import matching

    rng = np.random.default_rng(9)
    nq, nt = 700, 900
    pairs_in = []
    for s in range(5):
        b = rng.random((nt, 128)).astype(np.float32)
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        a = b[rng.integers(0, nt, nq)] + rng.normal(0, 0.002 * (s + 1), (nq, 128)).astype(np.float32)
        a[::5] = rng.random((len(a[::5]), 128)).astype(np.float32)
        pairs_in.append((a, b))
    bm = matching.BatchMatcher(nq, nt, 0.6, depth=2)
    tickets, results = [], []
    for a, b in pairs_in:  # keep two pairs in flight
        tickets.append(bm.submit(a, b))
        if len(tickets) >= 2:
            results.append(bm.collect(tickets[len(results)]))
    while len(results) < len(tickets):
        results.append(bm.collect(tickets[len(results)]))
    for (a, b), (pairs, dist) in zip(pairs_in, results):
        ref_pairs, ref_dist = matching.match_indices(a, b, 0.6, "exact")
        assert np.array_equal(pairs, ref_pairs)
        assert np.array_equal(dist, ref_dist)
    assert sum(len(p) for p, _ in results) > 500  # the comparison above is not between empty lists


@pytest.mark.parametrize("n_dup", [3, 400])
def test_exact_recheck_with_split_train_range(n_dup):
    """Rows the tensor path cannot prove go to the exact kernel, which cuts the train range into pieces when only a few
    query blocks are flagged (one CTA walking 20 000 train rows alone is a millisecond-long latency chain).  Near-duplicate
    train rows 1e-5 apart are below what the fp16 candidates can separate: those queries must be re-checked, and the
    merged pieces must give the exact engine's answer (the fp16-overflow case of the next test sends every row through the
    same path: many query blocks, fewer pieces)."""
    import torch

    from exvision_featurecraft_b200 import matching

    rng = np.random.default_rng(21)
    nq, nt = 3000, 20000
    b = np.abs(rng.normal(0, 1, (nt, 128))).astype(np.float32)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    dup = rng.choice(nt // 2, n_dup, replace=False)
    b[nt // 2 + np.arange(n_dup)] = b[dup] + rng.normal(0, 1e-5, (n_dup, 128)).astype(np.float32)  # twins far apart in the range
    a = np.abs(rng.normal(0, 1, (nq, 128))).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    a[:n_dup] = b[dup] + rng.normal(0, 0.01, (n_dup, 128)).astype(np.float32)  # queries whose two best neighbours are twins
    for ratio in (0.3, 0.999):
        ei, ed, eg = matching.knn2(a, b, ratio, "exact")
        ti, td, tg = matching.knn2(a, b, ratio, "tensor")
        st = matching.last_stats()
        assert st["overflow"] == 0
        assert torch.equal(eg, tg)
        g = eg.nonzero().flatten()
        assert torch.equal(ei[g, 0], ti[g, 0]) and torch.equal(ed[g, 0], td[g, 0])
        if ratio > 0.9:
            assert st["rechecked_rows"] >= 1, st  # some rows near the ratio cut cannot be proven from the candidates
        # a second call on the same workspace sizes: the arrival counters were left at zero
        ti2, td2, tg2 = matching.knn2(a, b, ratio, "tensor")
        assert torch.equal(tg2, tg) and torch.equal(ti2[g, 0], ti[g, 0])


def test_tensor_matcher_rarely_needs_the_exact_recheck():
    """The tcgen05 candidates + bounds must decide almost every row on their own (otherwise the test above
    would pass on the exact fallback alone), and agree with the exact engine at a size that spans many tiles,
    several splits and a ragged tail."""
    import torch

    from exvision_featurecraft_b200 import matching

    rng = np.random.default_rng(11)
    nq, nt = 5000, 7777
    b = np.abs(rng.normal(0, 1, (nt, 128))).astype(np.float32)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    a = np.abs(rng.normal(0, 1, (nq, 128))).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    pick = rng.integers(0, nt, nq // 2)
    a[: nq // 2] = b[pick] + rng.normal(0, 0.01, (nq // 2, 128)).astype(np.float32)
    for ratio in (0.3, 0.8):
        ei, ed, eg = matching.knn2(a, b, ratio, "exact")
        ti, td, tg = matching.knn2(a, b, ratio, "tensor")
        st = matching.last_stats()
        assert st["overflow"] == 0
        assert st["rechecked_rows"] <= 0.05 * nq, st
        assert torch.equal(eg, tg)
        g = eg.nonzero().flatten()
        assert g.numel() > nq // 4
        assert torch.equal(ei[g, 0], ti[g, 0]) and torch.equal(ed[g, 0], td[g, 0])
        # the nearest neighbour itself is almost always identical even for rows that fail the ratio test
        assert (ei[:, 0] == ti[:, 0]).float().mean() > 0.99
    # fp16 overflow -> the call falls back to the exact kernel for every row
    big = a * 1e5
    ei, ed, eg = matching.knn2(big, b, 0.8, "exact")
    ti, td, tg = matching.knn2(big, b, 0.8, "tensor")
    assert matching.last_stats()["overflow"] == 1
    assert torch.equal(ei, ti) and torch.equal(eg, tg)
