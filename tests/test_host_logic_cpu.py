"""Host-side logic and the integer identities the kernels rely on, checked without a GPU."""
import numpy as np
import pytest

from oracle import featurecraft_oracle as O


def _sift():
    from exvision_featurecraft_b200 import sift

    return sift


def test_separable_orientation_weights_factor_the_reference_table():
    """Certified mode multiplies wy[dy] * wx[dx] in float32; in float64 the factorisation must reproduce
    gaussian_filter_kernel (APP/utils/SIFT_scale_space.py:22-28) to rounding, for every radius of the UI ranges."""
    S = _sift()
    for sigma in (2.4, 3.39, 4.8, 7.2, 13.6, 38.4):
        r = S.kernel_radius(sigma)
        table = S.gaussian_window_table(sigma)
        sep = S.separable_window_table(sigma, r, r)
        wy, wx = sep[: 2 * r + 1], sep[2 * r + 1:]
        assert table.shape == (2 * r + 1, 2 * r + 1)
        assert np.abs(np.outer(wy, wx) / table - 1).max() < 1e-15
        assert np.array_equal(table, O.gaussian_filter_kernel(sigma))
        # the float32 product carries three roundings: inside what orientation_tolerance budgets per sample (4 x 2^-24)
        f32 = np.outer(wy.astype(np.float32), wx.astype(np.float32)).astype(np.float64)
        assert np.abs(f32 / table - 1).max() < 3.1 * 2.0 ** -24


def test_orientation_tolerance_covers_its_error_terms():
    S = _sift()
    for radius in (5, 7, 10, 14, 15, 20, 40, 77):
        tol = S.orientation_tolerance(radius)
        assert tol > 2 * 2.0 ** -17          # the magnitude rounding alone, doubled
        assert tol < 1e-3                    # and far below anything that would send every keypoint to the re-check
    assert S.orientation_tolerance(77) > S.orientation_tolerance(7)


def test_radii_of_the_ui_parameter_ranges_fit_the_kernel_classes():
    """FeatureCraft_UI.py:197-229: sigma 1.6 .. 3.0 in steps of 0.1, s 2 .. 5.  The float64 class kernel covers radii up to
    14 (levels 1 .. s stay below 12), the repair patches and the generic kernels radii up to 24, 8 levels at most."""
    S = _sift()
    seen = set()
    for s in range(2, 6):
        for i in range(15):
            sigma = 1.6 + 0.1 * i
            radii = [S.kernel_radius(x) for x in S.sigma_levels(sigma, s)]
            seen.add(tuple(radii))
            assert len(radii) == s + 3 <= 8
            assert max(radii) <= 24 and min(radii) >= 3
            assert max(radii[1:s + 1]) <= 12
            assert radii == sorted(radii)
    assert (3, 5, 6, 9, 13) in seen  # the default point, the one the packed kernel is specialised for


def test_descriptor_bin_identity():
    """descriptor_coded_kernel: sector difference sc in [0, 144] (offset 72 folded in), bin = (sc mod 72) / 9 computed as
    ((sc * 57) >> 9) & 7 -- valid because 72 = 8 * 9 and 57 / 512 approximates 1 / 9 closely enough on that range."""
    for sc in range(0, 200):
        assert (((sc * 57) >> 9) & 7) == (sc % 72) // 9
    # and the reference's own expression for directions exactly on a 5-degree grid: int(rel * 8 / 360.0)
    for o in range(37):
        for sub in range(3):
            for b in range(36):
                sector = (2 * o - 1 + sub) - (2 * b + 1)           # 5-degree sector of direction - theta
                sc = 2 * o + sub + 70 - 2 * b                      # what the kernel forms
                assert sc == sector + 72 and 0 <= sc <= 144
                assert (((sc * 57) >> 9) & 7) == (sector % 72) // 9


def _prmt(x, y, sel):
    bs = [(int(x) >> (8 * i)) & 255 for i in range(4)] + [(int(y) >> (8 * i)) & 255 for i in range(4)]
    out = 0
    for i in range(4):
        n = (sel >> (4 * i)) & 15
        v = bs[n & 7]
        if n & 8:
            v = 255 if v & 128 else 0
        out |= v << (8 * i)
    return out


def test_lambda_minus_packed_lane_arithmetic_reproduces_the_kx_gradients():
    """lambda5_packed_kernel keeps two pixels per 32-bit register (hi * 65536 + lo, lo signed), builds both 5-tap kernels
    from box filters and gets gy from per-row horizontal kernels: the same steps in Python integers with 32-bit
    wrap-around must give the K_X / K_X^T correlations of the oracle (APP/FeatureCraft_Backend.py:405-413) exactly."""
    rng = np.random.default_rng(0)
    H, W, c0 = 24, 16, 8
    img = rng.integers(0, 256, (H, W)).astype(np.int64)
    img[5:9, 3:12] = 255   # saturated block: the largest gradients the lanes must hold
    img[12:15, :] = 0
    gx_ref = O._correlate_zero_pad(img.astype(np.float64), O.K_X)
    gy_ref = O._correlate_zero_pad(img.astype(np.float64), O.K_X.T)
    M = 0xFFFFFFFF

    def unpack(w):
        e = []
        for q in range(3):
            e += [_prmt(w[q], 0, 0x4140), _prmt(w[q], 0, 0x4342)]
        return e

    def lo(p):
        v = _prmt(p, 0, 0x9910)
        return v - (1 << 32) if v & 0x80000000 else v

    def hi(p):
        v = (p + 0x8000) & M
        return (v - (1 << 32) if v & 0x80000000 else v) >> 16

    def words(r):
        if r < 0 or r >= H:
            return [0, 0, 0]
        return [sum((int(img[r, c]) if 0 <= c < W else 0) << (8 * b) for b, c in enumerate(range(c0 - 4 + 4 * q, c0 + 4 * q)))
                for q in range(3)]

    vA, vB, sP, sC, hor, checked = [0] * 6, [0] * 6, [0] * 6, [0] * 6, {}, 0
    for q in range(-4, H + 4):
        vN = unpack(words(q))
        od = [_prmt(vN[j], vN[j + 1], 0x5432) for j in range(5)]
        so = [(vN[j] + od[j] + vN[j + 1]) & M for j in range(5)]
        se = [None] + [(od[j - 1] + vN[j] + od[j]) & M for j in range(1, 5)]
        An = [(so[p] + se[p + 1] + so[p + 1]) & M for p in range(4)]
        Bn = [(2 * An[p] - se[p + 1]) & M for p in range(4)]
        hor[q] = (An, Bn)
        sN = [(vA[j] + vB[j] + vN[j]) & M for j in range(6)]
        A = [(sP[j] + sC[j] + sN[j]) & M for j in range(6)]
        B = [(2 * A[j] - sC[j]) & M for j in range(6)]
        hr = q - 2
        if 0 <= hr < H and (q - 4) in hor:
            bo = [_prmt(B[j], B[j + 1], 0x5432) for j in range(5)]
            ao, bu, bd = hor[q - 4][0], hor[q - 3][1], hor[q - 1][1]
            for p in range(4):
                GX = ((A[p + 2] - A[p]) + (bo[p + 1] - bo[p])) & M
                GY = ((An[p] - ao[p]) + (bd[p] - bu[p])) & M
                for half, f in enumerate((lo, hi)):
                    c = c0 - 2 + 2 * p + half
                    if 0 <= c < W:
                        assert f(GX) == int(gx_ref[hr, c]) and f(GY) == int(gy_ref[hr, c])
                        checked += 1
        vA, vB, sP, sC = vB, vN, sC, sN
    assert checked == H * 8
    assert np.abs(gx_ref).max() <= 6120 and np.abs(gy_ref).max() <= 6120  # 255 x the positive weights: fits a signed 16-bit lane


def test_lambda_minus_integer_sums_fit_their_types():
    """|g| <= 6120 -> products <= 3.75e7, 5x5 sums <= 9.4e8 < 2^31 (int32 cross sums), determinant < 2^60 and the radicand
    (Sxx - Syy)^2 + 4 Sxy^2 < 2^63 (the fast mode's exact 64-bit integers)."""
    g = 6120
    s = 25 * g * g
    assert s < 2 ** 31
    assert s * s < 2 ** 60
    assert s * s + 4 * s * s < 2 ** 63


@pytest.mark.parametrize("legacy", [False, True])
def test_orientation_cut_rules_of_the_oracle(legacy):
    hist = np.array([1.2345678, 0.98765434, 0.5], dtype=np.float32)
    cut = np.float32(0.8 * float(hist.max())) if legacy else 0.8 * hist.max()
    assert np.asarray(cut).dtype == np.float32
    ref = np.float32(0.8 * float(hist.max())) if legacy else np.float32(0.8) * hist.max()
    assert cut == ref
