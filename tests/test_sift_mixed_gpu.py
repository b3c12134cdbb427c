"""Certified ("mixed") SIFT mode: float32 bulk arithmetic whose every discrete decision is either guaranteed by its
error bound or re-evaluated in float64.  The bar is therefore the float64 bar -- IDENTICAL keypoint lists (extrema,
orientations, NaN rows) against the golden vectors the unmodified reference produced, no percentile allowances --
and descriptors within float32 rounding (<= 2e-5 L2; north_star asks 1e-3).  Each certified piece is also tested on
its own, with the error margins turned up so that the float64 repair kernels decide (nearly) everything."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import featurecraft_oracle as O

pytestmark = pytest.mark.gpu

SIFT_CASES = ["blobs_96x112", "blobs_s3_80x80", "blobs_oct5_sigma2_128x96", "box_crop_96x120"]


@pytest.fixture(scope="module")
def S(lib):
    from exvision_featurecraft_b200 import sift

    return sift


def _params(S, g, case, **kw):
    n_oct, s, sigma, cth, rth = g[case + "/params"]
    return S.SiftParams(int(n_oct), int(s), float(sigma), float(cth), float(rth), **kw)


def _kps_of(S, octv, k):
    return S.keypoints_from_mask(octv.extrema[k - 1], octv.width).cpu().numpy()[:, 1:]


@pytest.mark.parametrize("case", SIFT_CASES)
def test_full_pipeline_mixed_equals_reference(S, case):
    g = load_golden("sift")
    p = _params(S, g, case, dtype="mixed")
    pts, desc = S.detect_and_describe(g[case + "/base"], p, out_dtype=torch.float32).assemble(1)[0]
    ref_p, ref_d = g[case + "/points"], g[case + "/descriptors"]
    assert pts.shape == ref_p.shape
    assert np.array_equal(pts, ref_p)
    nan_ref = np.isnan(ref_d).any(1)
    assert np.array_equal(np.isnan(desc).any(1), nan_ref)
    assert np.linalg.norm(desc[~nan_ref] - ref_d[~nan_ref], axis=1).max() < 2e-5


@pytest.mark.parametrize("case", SIFT_CASES)
@pytest.mark.parametrize("variant", ["app", "script"])
def test_certified_extrema_equal_reference(S, case, variant):
    g = load_golden("sift")
    p = _params(S, g, case, dtype="mixed", variant=variant)
    key = "%s/kps_oct%d" if variant == "app" else "%s/script_kps_oct%d"
    if variant == "script" and case + "/script_raises_indexerror" in g.files:
        with pytest.raises(IndexError):
            S.scale_space(g[case + "/base"], p)
        return
    octs = S.scale_space(g[case + "/base"], p)
    for o, octv in enumerate(octs):
        assert octv.gauss.dtype == torch.float32  # the certified path ran (no float64 fallback)
        ref_k = g[key % (case, o)]
        for k in range(1, p.s_value + 1):
            assert np.array_equal(_kps_of(S, octv, k), ref_k[ref_k[:, 2] == k][:, :2]), (o, k)


@pytest.mark.parametrize("eps", [2.0 ** -19, 1e-4, 3e-3])
def test_repair_kernel_decides_like_float64(S, eps, monkeypatch):
    """With a large margin most candidates go through fc_extrema_repair (float64 levels read back or blurred on
    demand): the masks must still be the float64 path's, for both filters and on odd image sizes."""
    from exvision_featurecraft_b200 import synthetic

    monkeypatch.setattr(S, "EXTREMA_EPS", eps)
    monkeypatch.setitem(S.AMB_FRACTION, "app", 1.0)
    monkeypatch.setitem(S.AMB_FRACTION, "script", 1.0)
    for shape, variant in (((96, 112), "app"), ((75, 131), "app"), ((64, 88), "script")):
        base = synthetic.sift_image(*shape, seed=31 + shape[0])
        ref = S.scale_space(base, S.SiftParams(dtype="f64", variant=variant))
        got = S.scale_space(base, S.SiftParams(dtype="mixed", variant=variant))
        for o, (a, b) in enumerate(zip(got, ref)):
            assert a.gauss.dtype == torch.float32
            assert torch.equal(a.extrema, b.extrema), (shape, variant, o)


def test_certified_extrema_on_ties(S):
    """Integer-valued images: the DoG is full of exact ties and flat runs.  Every tie is undecidable in float32 by
    construction and must come back from the float64 repair with the reference's first-wins rule."""
    rng = np.random.default_rng(7)
    base = rng.integers(0, 4, size=(72, 96)).astype(np.float64) / 4.0
    base[20:50, 30:70] = 0.5  # a flat block: all-equal neighbourhoods are no extrema
    ref = S.scale_space(base, S.SiftParams(dtype="f64"))
    got = S.scale_space(base, S.SiftParams(dtype="mixed"))
    pyr, dog, kps = O.generate_octaves_pyramid(base, 4, 2, 1.6, 0.03, 10, blur_mode="separable")
    for o, (a, b) in enumerate(zip(got, ref)):
        assert torch.equal(a.extrema, b.extrema), o


def test_float64_levels_match_the_oracle(S):
    from exvision_featurecraft_b200 import _lib, synthetic
    from exvision_featurecraft_b200.device import ptr, stream_ptr
    import ctypes as C

    for shape, (sigma, s) in (((150, 260), (1.6, 2)), ((77, 131), (1.6, 2)), ((64, 96), (2.0, 3))):
        img = np.stack([synthetic.sift_image(*shape, seed=3), synthetic.sift_image(*shape, seed=4)])
        p = S.SiftParams(sigma_base=sigma, s_value=s)
        flat, lens = S._level_taps(p)
        t = torch.from_numpy(img).cuda()
        out = torch.empty((2, s) + shape, dtype=torch.float64, device="cuda")
        _lib.call("fc_gauss_levels_f64", ptr(t), 2, shape[0], shape[1], p.n_levels, flat.ctypes.data_as(C.c_void_p),
                  C.cast(lens, C.c_void_p), 1, s, ptr(out), stream_ptr())
        kern = O.generateGaussianKernels(sigma, s)
        for b in range(2):
            for l in range(s):
                ref = O.blur(img[b], kern[1 + l], "separable")
                assert np.abs(out[b, l].cpu().numpy() - ref).max() < 2e-15, (shape, b, l)


def test_gradient_codes_are_the_float64_codes(S):
    """Every pixel's code must be the one the reference's float64 direction gives: orientation bin by np.round
    (half to even), 5-degree sector by comparison.  Directions that sit on a multiple of 5 degrees to within an ulp
    (integer-valued test images: 0, 45, 90 .. degrees) depend on the last bit of atan2, which libdevice and glibc do
    not share: there the code may be the neighbouring one, everywhere else it must be exact."""
    g = load_golden("sift_pieces")
    rng = np.random.default_rng(11)
    imgs = [g["img"], rng.integers(0, 5, size=(40, 52)).astype(np.float64), rng.random((37, 50))]
    for n, img in enumerate(imgs):
        t = torch.from_numpy(np.ascontiguousarray(img)).cuda().reshape(1, 1, *img.shape)
        codes = S.gradient_codes(t, 0)[0].cpu().numpy().view(np.uint32)
        _, _, mag, direction = O.sift_gradient(img)
        ref_bin = np.round(direction * 36 / 360).astype(int)
        ref_sector = np.floor(direction / 5).astype(int)
        on_edge = np.abs(direction / 5 - np.round(direction / 5)) < 1e-9
        if n != 1:
            assert not on_edge.any()
        got_bin = (codes & 63).astype(int)
        sector = 2 * got_bin - 1 + ((codes >> 6) & 3).astype(int)
        assert np.array_equal(got_bin[~on_edge], ref_bin[~on_edge])
        assert np.array_equal(sector[~on_edge], ref_sector[~on_edge])
        assert np.abs(got_bin - ref_bin)[on_edge].max(initial=0) <= 1
        assert np.isin(np.abs(sector - ref_sector)[on_edge], (0, 1, 71)).all()
        m = ((codes & ~np.uint32(255)) >> 1).view(np.float32).astype(np.float64)
        assert np.abs(m - mag).max() <= 2.0 ** -16 * mag.max()
        assert np.array_equal(m == 0, mag == 0)


@pytest.mark.parametrize("tol_scale", [1.0, 1e4])
def test_certified_orientations_equal_float64(S, tol_scale, monkeypatch):
    """tol_scale = 1e4 marks every keypoint for the float64 recheck kernel; 1.0 is production.  Both must reproduce
    the float64 pipeline's oriented keypoint list (and so the reference's, see the golden test above)."""
    from exvision_featurecraft_b200 import synthetic

    real = S.orientation_tolerance
    monkeypatch.setattr(S, "orientation_tolerance", lambda r: real(r) * tol_scale)
    base = synthetic.sift_image(120, 136, seed=17)
    a = S.detect_and_describe(base, S.SiftParams(dtype="mixed"), out_dtype=torch.float32).assemble(1)[0]
    b = S.detect_and_describe(base, S.SiftParams(dtype="f64"), out_dtype=torch.float64).assemble(1)[0]
    assert np.array_equal(a[0], b[0])
    ok = ~np.isnan(b[1]).any(1)
    assert np.array_equal(np.isnan(a[1]).any(1), ~ok)
    assert np.linalg.norm(a[1][ok] - b[1][ok], axis=1).max() < 2e-5


def test_batched_final_order_equals_single_images(S):
    """The descriptor kernel writes straight to image-major rows (fc_sift_layout): a batch must equal its images."""
    from exvision_featurecraft_b200 import synthetic

    imgs = np.stack([synthetic.sift_image(72, 100, seed=s) for s in (1, 2, 3, 4, 5)])
    p = S.SiftParams(dtype="mixed")
    res = S.detect_and_describe(imgs, p, out_dtype=torch.float32)
    batched = res.assemble(5)
    assert int(res.counts.sum()) == res.total
    for b in range(5):
        pts, desc = S.detect_and_describe(imgs[b], p, out_dtype=torch.float32).assemble(1)[0]
        assert np.array_equal(batched[b][0], pts)
        assert np.array_equal(batched[b][1], desc, equal_nan=True)
    rp, _ = O.compute_keypoints_and_descriptors_from_base(imgs[3], 4, 2, 1.6, 0.03, 10)
    assert np.array_equal(batched[3][0], np.array(rp, dtype=np.float64).reshape(-1, 5))


def test_float16_export_is_within_tolerance(S):
    from exvision_featurecraft_b200 import synthetic

    base = synthetic.sift_image(96, 128, seed=3)
    p = S.SiftParams(dtype="mixed")
    a = S.detect_and_describe(base, p, out_dtype=torch.float16).assemble(1)[0]
    b = S.detect_and_describe(base, p, out_dtype=torch.float32).assemble(1)[0]
    assert np.array_equal(a[0], b[0])
    ok = ~np.isnan(b[1]).any(1)
    assert np.linalg.norm(a[1][ok].astype(np.float64) - b[1][ok], axis=1).max() < 5e-4  # north_star: 1e-3 L2


def test_undecided_list_overflow_falls_back_to_float64(S, monkeypatch):
    """A list that is too small must not lose keypoints: the driver re-runs the batch with the float64 kernels."""
    from exvision_featurecraft_b200 import synthetic

    monkeypatch.setattr(S, "EXTREMA_EPS", 1e-2)
    monkeypatch.setitem(S.AMB_FRACTION, "app", 0.0)
    monkeypatch.setattr(S, "gauss_octave_mixed", _tiny_cap(S.gauss_octave_mixed))
    base = synthetic.sift_image(200, 264, seed=5)
    a = S.detect_and_describe(base, S.SiftParams(dtype="mixed"), out_dtype=torch.float64).assemble(1)[0]
    b = S.detect_and_describe(base, S.SiftParams(dtype="f64"), out_dtype=torch.float64).assemble(1)[0]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1], equal_nan=True)


def _tiny_cap(fn):
    def wrapped(*a, **k):
        octv, mx = fn(*a, **k)
        mx.amb_cap = 8  # pretend the list was this small: the counters say "overflow"
        return octv, mx
    return wrapped


@pytest.mark.parametrize("shape", [(37, 53), (64, 33), (19, 90), (101, 77)])
def test_odd_and_tiny_sizes_equal_float64(S, shape):
    """Widths that are not multiples of 4 take the tiled certified extremum kernel, the scalar code kernel and the
    generic float64 level kernel: the lists must still be the float64 path's."""
    from exvision_featurecraft_b200 import synthetic

    base = synthetic.sift_image(*shape, seed=shape[0] + shape[1])
    a = S.detect_and_describe(base, S.SiftParams(dtype="mixed"), out_dtype=torch.float32).assemble(1)[0]
    b = S.detect_and_describe(base, S.SiftParams(dtype="f64"), out_dtype=torch.float64).assemble(1)[0]
    assert np.array_equal(a[0], b[0])
    ok = ~np.isnan(b[1]).any(1)
    assert np.array_equal(np.isnan(a[1]).any(1), ~ok)
    if ok.any():
        assert np.linalg.norm(a[1][ok] - b[1][ok], axis=1).max() < 2e-5


@pytest.mark.parametrize("params", [dict(n_octaves=5, s_value=3, sigma_base=2.5), dict(n_octaves=4, s_value=2, sigma_base=3.0),
                                    dict(n_octaves=6, s_value=4, sigma_base=1.6), dict(n_octaves=4, s_value=5, sigma_base=3.0),
                                    dict(n_octaves=4, s_value=5, sigma_base=1.6), dict(n_octaves=4, s_value=3, sigma_base=1.9),
                                    dict(n_octaves=8, s_value=4, sigma_base=2.2),
                                    dict(n_octaves=4, s_value=2, sigma_base=1.6, variant="script"),
                                    dict(n_octaves=4, s_value=2, sigma_base=2.0, variant="script", contrast_th=0.01, ratio_th=6.0)])
def test_ui_parameter_ranges_equal_float64(S, params):
    """Points of the app's parameter ranges (FeatureCraft_UI.py:197-273: octaves 4-8, s 2-5, sigma 1.6-3, r 6-12) and the
    script filter.  Certified mode must equal float64 mode everywhere, and it must get there on the fast kernels: the
    float64 levels 1 .. s (radii 5 .. 12 over the whole range) through the two-level kernel in its radius classes, never
    through the generic level-by-level one, and the float64 repair from shared-memory patches (radii up to 24 = the
    whole range)."""
    from exvision_featurecraft_b200 import _lib, synthetic

    base = synthetic.sift_image(160, 192, seed=41)
    before = _lib.path_counts()
    a = S.detect_and_describe(base, S.SiftParams(dtype="mixed", **params), out_dtype=torch.float32)
    after = _lib.path_counts()
    took = {k: after[k] - before[k] for k in after}
    p = S.SiftParams(dtype="mixed", **params)
    radii = [S.kernel_radius(x) for x in S.sigma_levels(p.sigma_base, p.s_value)]
    assert took["mid_f64"] == p.n_octaves * ((p.s_value + 1) // 2), took
    # the only generic launches are the float32 stacks of non-default radii (one per octave)
    assert took["generic_octave"] == (0 if radii == [3, 5, 6, 9, 13] else p.n_octaves), took
    if radii == [3, 5, 6, 9, 13]:
        assert took["fused_search"] == p.n_octaves, took
    assert max(radii) <= 24 and took["repair_patch"] == p.n_octaves, took
    b = S.detect_and_describe(base, S.SiftParams(dtype="f64", **params), out_dtype=torch.float64)
    pa, da = a.assemble(1)[0]
    pb, db = b.assemble(1)[0]
    assert pb.shape[0] > 50
    assert np.array_equal(pa, pb)
    ok = ~np.isnan(db).any(1)
    assert np.array_equal(np.isnan(da).any(1), ~ok)
    assert np.linalg.norm(da[ok] - db[ok], axis=1).max() < 2e-5


@pytest.mark.parametrize("seed", range(6))
def test_random_images_certified_equals_float64(S, seed):
    """Different image statistics (blob fields, pure noise, smooth ramps with saturated plateaus, quantised values):
    whatever the data, every decision of certified mode must be the float64 decision."""
    rng = np.random.default_rng(100 + seed)
    h, w = int(rng.integers(48, 140)), int(rng.integers(12, 36)) * 4
    kind = seed % 3
    if kind == 0:
        base = rng.random((h, w))
    elif kind == 1:
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.clip(0.5 + 0.8 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + rng.normal(0, 0.01, (h, w)), 0.0, 1.0)  # clipped plateaus
    else:
        base = np.round(rng.random((h, w)) * 8) / 8  # nine grey levels: ties everywhere
    a = S.detect_and_describe(base, S.SiftParams(dtype="mixed"), out_dtype=torch.float32).assemble(1)[0]
    b = S.detect_and_describe(base, S.SiftParams(dtype="f64"), out_dtype=torch.float64).assemble(1)[0]
    assert np.array_equal(a[0], b[0])
    ok = ~np.isnan(b[1]).any(1)
    assert np.array_equal(np.isnan(a[1]).any(1), ~ok)
    if ok.any():
        assert np.linalg.norm(a[1][ok] - b[1][ok], axis=1).max() < 2e-5


def test_fused_search_equals_two_kernel_form(S):
    """fc_gauss_octave_search (blur + certified search in one kernel, no float32 stack written; what the pipeline runs)
    against the two-kernel form (keep_octaves=True keeps the stack): same masks, same undecided decisions, same result;
    sizes that exercise partial tiles (92 x 30 centre stride), several images, both filters."""
    from exvision_featurecraft_b200 import synthetic

    for shape, variant in (((2 * 60, 2 * 100), "app"), ((188, 376), "app"), ((96, 112), "script"), ((300, 96), "app")):
        imgs = np.stack([synthetic.sift_image(*shape, seed=s) for s in (3, 4)])
        p = S.SiftParams(dtype="mixed", variant=variant)
        a = S.detect_and_describe(imgs, p, out_dtype=torch.float32)
        b = S.detect_and_describe(imgs, p, out_dtype=torch.float32, keep_octaves=True)
        assert b.octaves[0].gauss is not None
        assert torch.equal(a.records, b.records) and torch.equal(a.counts, b.counts)
        assert torch.equal(a.descriptors.nan_to_num(), b.descriptors.nan_to_num())
        for blk_a, blk_b in zip(a.blocks, b.blocks):
            assert torch.equal(blk_a.keypoints, blk_b.keypoints)
