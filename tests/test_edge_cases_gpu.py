"""Edge cases the reference's code paths imply: empty results, degenerate images, parameter quirks, error behaviour."""
import numpy as np
import pytest
import torch

from oracle import featurecraft_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M(lib):
    from exvision_featurecraft_b200 import backend, corners, matching, sift
    from exvision_featurecraft_b200.utils import SIFT_scale_space, keypoint_descriptor

    SIFT_scale_space.set_dtype("f64")
    SIFT_scale_space.set_variant("app")

    class NS:
        pass

    ns = NS()
    ns.backend, ns.corners, ns.matching, ns.sift, ns.SS, ns.KD = backend, corners, matching, sift, SIFT_scale_space, keypoint_descriptor
    return ns


def test_constant_image_has_no_corners_and_no_keypoints(M):
    flat = np.full((64, 96), 77, np.uint8)
    res = M.corners.harris(flat, 5, 0.04, 10000)
    assert res.coords.shape == (0, 2) and float(res.response.abs().max()) == 0.0
    lam = M.corners.lambda_minus(flat, 0.01)
    # K_X is applied with ZERO padding: a constant image still has gradients along its frame (and only there)
    ref = O.lambda_minus_response(flat)
    assert np.array_equal(lam.response[0].cpu().numpy(), ref) and ref[8:-8, 8:-8].max() == 0.0
    rr, cc = np.where(ref > 0.01 * ref.max())
    assert np.array_equal(lam.coords.cpu().numpy(), np.stack([rr, cc], 1))
    zero = np.zeros((32, 64), np.uint8)
    lz = M.corners.lambda_minus(zero, 0.01)
    # eig == 0 everywhere, threshold = 0.01 * 0 = 0, strict '>' -> nothing
    assert lz.coords.shape == (0, 2) and float(lz.image_max[0]) == 0.0
    r = M.sift.detect_and_describe(np.full((40, 56), 0.5), M.sift.SiftParams(dtype="f64"), out_dtype=torch.float64)
    pts, desc = r.assemble(1)[0]
    assert pts.shape == (0, 5) and desc.shape == (0, 128)


def test_images_smaller_than_the_kernels(M):
    """13x9 image: every Gaussian kernel is wider than the image ('symm' reflection wraps more than once)."""
    rng = np.random.default_rng(0)
    base = rng.random((13, 9))
    pyr, dog, kps = O.generate_octaves_pyramid(base, 2, 2, 1.6, 0.03, 10)
    octs = M.sift.scale_space(base, M.sift.SiftParams(n_octaves=2, dtype="f64"))
    for o in range(2):
        assert np.abs(octs[o].gauss[0].cpu().numpy() - np.stack(pyr[o])).max() < 2e-14
        ref = np.asarray(kps[o], dtype=np.int64).reshape(-1, 3)
        for k in (1, 2):
            got = M.sift.keypoints_from_mask(octs[o].extrema[k - 1], octs[o].width).cpu().numpy()[:, 1:]
            assert np.array_equal(got, ref[ref[:, 2] == k][:, :2])
    tiny = rng.integers(0, 255, (3, 4)).astype(np.uint8)
    assert np.array_equal(M.corners.harris(tiny, 5, 0.04, None).response[0].cpu().numpy(), O.harris_response(tiny, 5, 0.04))
    assert np.array_equal(M.corners.lambda_minus(tiny, None).response[0].cpu().numpy(), O.lambda_minus_response(tiny))


def test_keypoints_on_the_image_border_use_zero_padded_windows(M):
    """Orientation windows hang over the edge (zero padding, KD:82-113) and descriptor windows sample outside
    (constant border 0, KD:206-212): a small image makes every keypoint a border keypoint."""
    from exvision_featurecraft_b200 import synthetic

    base = synthetic.sift_image(24, 30, seed=5)
    rp, rd = O.compute_keypoints_and_descriptors_from_base(base, 4, 2, 1.6, 0.03, 10)
    pts, desc = M.sift.detect_and_describe(base, M.sift.SiftParams(dtype="f64"), out_dtype=torch.float64).assemble(1)[0]
    assert np.array_equal(pts, np.array(rp, dtype=np.float64).reshape(-1, 5))
    rd = np.array(rd).reshape(-1, 128)
    nan = np.isnan(rd).any(1)
    assert np.array_equal(np.isnan(desc).any(1), nan)
    assert np.abs(desc[~nan] - rd[~nan]).max() < 1e-9


@pytest.mark.parametrize("n_oct,s,sigma", [(2, 2, 1.6), (3, 2, 2.5), (4, 4, 1.6), (6, 2, 1.6)])
def test_parameter_ranges_of_the_ui(M, n_oct, s, sigma):
    """UI ranges (FeatureCraft_UI.py:197-229): octaves 4-8, s 2-5, sigma 1.6-3; also fewer octaves.  The quirk of
    represent_keypoints (scale masks = range(1, n_octaves-1)) decides which scales reach the descriptor stage."""
    from exvision_featurecraft_b200 import synthetic

    base = synthetic.sift_image(96, 80, seed=n_oct * 10 + s)
    rp, rd = O.compute_keypoints_and_descriptors_from_base(base, n_oct, s, sigma, 0.03, 10)
    p = M.sift.SiftParams(n_oct, s, sigma, 0.03, 10, "app", "f64")
    pts, desc = M.sift.detect_and_describe(base, p, out_dtype=torch.float64).assemble(1)[0]
    assert np.array_equal(pts, np.array(rp, dtype=np.float64).reshape(-1, 5).reshape(pts.shape))
    if len(rd):
        rd = np.array(rd).reshape(-1, 128)
        ok = ~np.isnan(rd).any(1)
        assert np.abs(desc[ok] - rd[ok]).max() < 1e-9


def test_too_many_octaves_raises_like_the_reference(M):
    """n_octaves >= s + 5 makes dog_keypoints_orientations index img_gaussians[o][s + 3] -> IndexError."""
    base = np.random.default_rng(1).random((64, 64))
    with pytest.raises(IndexError):
        O.compute_keypoints_and_descriptors_from_base(base, 7, 2, 1.6, 0.03, 10)
    with pytest.raises(IndexError):
        M.sift.detect_and_describe(base, M.sift.SiftParams(n_octaves=7, dtype="f64"))


def test_matching_degenerate_inputs(M):
    rng = np.random.default_rng(2)
    a = rng.random((5, 128)).astype(np.float32)
    b = rng.random((9, 128)).astype(np.float32)
    for engine in ("exact", "tensor"):
        idx, dist, good = M.matching.knn2(a[:0], b, 0.5, engine)
        assert idx.shape == (0, 2) and good.shape == (0,)
        with pytest.raises(ValueError):
            M.matching.knn2(a, b[:1], 0.5, engine)
        # a NaN query row (the reference's NaN descriptors of all-zero windows) matches nothing
        an = a.copy()
        an[2] = np.nan
        idx, dist, good = M.matching.knn2(an, b, 0.9, engine)
        assert not bool(good[2]) and int(idx[2, 0]) == -1
        ref_idx, _ = O.knn2_l2(a, b)
        keep = [0, 1, 3, 4]
        assert np.array_equal(idx.cpu().numpy()[keep, 0], ref_idx[keep, 0])
        # NaN train rows are never selected
        bn = b.copy()
        bn[4] = np.nan
        idx, dist, good = M.matching.knn2(a, bn, 0.9, engine)
        assert not (idx.cpu().numpy() == 4).any()
    with pytest.raises(ValueError):
        M.matching.knn2(rng.random((4, 64)).astype(np.float32), b)


def test_c_abi_rejects_bad_arguments_on_the_gpu(M, lib):
    import ctypes as C

    g = torch.zeros((1, 8, 8), dtype=torch.uint8, device="cuda")
    r = torch.zeros((1, 8, 8), dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    z = C.c_void_p(0)
    assert lib.fc_harris_response(p(g), 1, 8, 8, 33, 0.04, 0.0, p(r), z, z) == -1   # window too large
    assert lib.fc_harris_response(p(g), 0, 8, 8, 5, 0.04, 0.0, p(r), z, z) == -1    # empty batch
    assert lib.fc_harris_response(p(g), 1, 1, 8, 5, 0.04, 0.0, p(r), z, z) == -1    # np.gradient needs 2 rows
    assert lib.fc_gauss_octave(p(r), 7, 1, 8, 8, 5, z, z, 0, 0, 0.0, 0.0, p(r), z, z) == -1  # bad dtype / null taps
    assert lib.fc_match_knn2_exact(z, 1, z, 1, 0.5, z, z, z, z) == -1
    assert lib.fc_status_string(-2) == b"output capacity too small"
    taps = (C.c_double * 3)(0.2, 0.5, 0.3)  # asymmetric taps are not a Gaussian: unsupported, not silently wrong
    lens = (C.c_int * 2)(3, 3)
    f = torch.zeros((1, 2, 8, 8), dtype=torch.float64, device="cuda")
    taps2 = (C.c_double * 6)(0.2, 0.5, 0.3, 0.2, 0.5, 0.3)
    assert lib.fc_gauss_octave(p(r), 1, 1, 8, 8, 2, C.cast(taps2, C.c_void_p), C.cast(lens, C.c_void_p), 0, 0, 0.0, 0.0, p(f), z, z) == -4


def test_rethreshold_and_batch_drivers(M):
    from exvision_featurecraft_b200 import synthetic

    imgs = np.stack([synthetic.corner_image(64, 128, seed=s) for s in range(4)])
    det = M.corners.BatchCornerDetector(4, 64, 128, max_corners=4 * 64 * 128)
    t = det.submit(imgs)
    hc, hv, ho, lc, lo = det.collect(t)
    for b in range(4):
        ref = O.harris_response(imgs[b], 5, 0.04)
        idx, vals = O.threshold_corners(ref, 10000)
        assert np.array_equal(hc[ho[b]:ho[b + 1]], idx) and np.array_equal(hv[ho[b]:ho[b + 1]], vals)
        lam = O.lambda_minus_response(imgs[b])
        rr, cc = np.where(lam > 0.01 * lam.max())
        assert np.array_equal(lc[lo[b]:lo[b + 1]], np.stack([rr, cc], 1))
    sift_imgs = np.stack([synthetic.sift_image(48, 64, seed=s) for s in range(3)])
    p = M.sift.SiftParams(dtype="f64")
    ex = M.sift.BatchExtractor(p, 3, 48, 64, max_descriptors=200000, depth=2)
    t0 = ex.submit(sift_imgs, prefetch_next=False)
    t1 = ex.submit(sift_imgs[::-1].copy())
    r0, r1 = ex.collect(t0, copy=True), ex.collect(t1, copy=True)
    for b in range(3):
        assert np.array_equal(r0[b][0], r1[2 - b][0]) and np.array_equal(r0[b][1], r1[2 - b][1], equal_nan=True)
        base = M.sift.upsample2(sift_imgs[b], p)[0].cpu().numpy()
        rp, _ = O.compute_keypoints_and_descriptors_from_base(base, 4, 2, 1.6, 0.03, 10, "app", "separable")
        assert r0[b][0].shape[0] == len(rp)


def test_batched_compaction_pieces(M, lib):
    """fc_mask_row_count / fc_popcount_u64 / fc_exclusive_scan_i32 / fc_publish_gather_i32: the pieces the batched SIFT
    driver concatenates; compared with numpy, large enough for the multi-level scan."""
    import ctypes as C

    from exvision_featurecraft_b200 import _lib
    from exvision_featurecraft_b200.device import ptr, read_ints_at, scan_scratch, stream_ptr

    rng = np.random.default_rng(11)
    B, H, W = 3, 37, 200
    pitch = (W + 31) // 32
    words = rng.integers(0, 2 ** 32, size=(B, H, pitch), dtype=np.uint64).astype(np.uint32)
    mask = torch.from_numpy(words.view(np.int32)).cuda()
    counts = torch.empty(B * H, dtype=torch.int32, device="cuda")
    _lib.call("fc_mask_row_count", ptr(mask), B, H, W, ptr(counts), stream_ptr())
    ref = np.unpackbits(words.view(np.uint8), axis=-1).reshape(B * H, -1).sum(1)
    assert np.array_equal(counts.cpu().numpy(), ref)

    n = 300_001  # > 2048 * 2048 / 16: three scan levels are not needed, two are
    m64 = rng.integers(0, 2 ** 36, size=n, dtype=np.int64)
    masks = torch.from_numpy(m64).cuda()
    pc = torch.empty(n, dtype=torch.int32, device="cuda")
    _lib.call("fc_popcount_u64", ptr(masks), n, ptr(pc), stream_ptr())
    ref_pc = np.array([bin(int(v)).count("1") for v in m64[:5000]])
    assert np.array_equal(pc.cpu().numpy()[:5000], ref_pc)
    offs = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    _lib.call("fc_exclusive_scan_i32", ptr(pc), n, ptr(offs), ptr(scan_scratch(n)), stream_ptr())
    ref_offs = np.concatenate([[0], np.cumsum(pc.cpu().numpy(), dtype=np.int64)])
    assert np.array_equal(offs.cpu().numpy().astype(np.int64), ref_offs)
    idx = [0, 1, 2047, 2048, 2049, 123_456, n] + [int(v) for v in rng.integers(0, n + 1, 40)]  # > 32: two gather launches
    assert read_ints_at(offs, idx) == [int(ref_offs[i]) for i in idx]
    # bad arguments
    assert lib.fc_publish_gather_i32(ptr(offs), (C.c_int64 * 1)(0), 33, ptr(offs), stream_ptr()) != 0
    assert lib.fc_exclusive_scan_i32(None, 10, ptr(offs), None, stream_ptr()) != 0
    assert lib.fc_popcount_u64(ptr(masks), 0, ptr(pc), stream_ptr()) != 0
    assert lib.fc_mask_row_count(None, B, H, W, ptr(counts), stream_ptr()) != 0


def test_capacity_errors_carry_the_required_count(M, lib):
    """FC_ERR_CAPACITY: the batch drivers know the output size after their one synchronisation and refuse BEFORE
    anything is copied into the caller's (too small) buffers."""
    from exvision_featurecraft_b200 import _lib, synthetic

    imgs = np.stack([synthetic.sift_image(48, 64, seed=s) for s in range(2)])
    ex = M.sift.BatchExtractor(M.sift.SiftParams(), 2, 48, 64, max_descriptors=10, depth=2)
    with pytest.raises(_lib.CapacityError) as ei:
        ex.submit(imgs)
    assert ei.value.status == -2 and ei.value.required > 10 and ei.value.capacity == 10
    big = M.sift.BatchExtractor(M.sift.SiftParams(), 2, 48, 64, max_descriptors=ei.value.required, depth=2)
    out = big.collect(big.submit(imgs))
    assert sum(len(p) for p, _ in out) == ei.value.required
    gray = np.stack([synthetic.corner_image(64, 128, seed=s) for s in range(2)])
    det = M.corners.BatchCornerDetector(2, 64, 128, max_corners=3)
    with pytest.raises(_lib.CapacityError) as ec:
        det.submit(gray)
    assert ec.value.required > 3


def test_notebook_harris_accepts_even_windows(M):
    """harris.ipynb:170-176: offset = int(window_size / 2), window = 2 * offset + 1 -- any size works in the notebook;
    only the app's convolve2d_optimized path fails on even sizes."""
    from exvision_featurecraft_b200 import synthetic

    gray = synthetic.corner_image(60, 72, seed=8)
    for ws in (4, 6):
        got = M.corners.notebook_harris(gray, ws, 0.04, 1e8)
        corners_ref, r_ref = O.harris_detection(gray, ws, 0.04, 1e8)
        assert np.abs(got.response[0].cpu().numpy() - r_ref).max() <= 1e-12 * np.abs(r_ref).max()
        assert got.coords.cpu().numpy().tolist() == [[y, x] for x, y, _ in corners_ref]
    with pytest.raises(ValueError):
        M.corners.harris(gray, 4, 0.04, 10000)
