"""The steps in front of the hot path: channel handling, sift_resize, and apply_sift as a drop-in.

sift_resize is skimage.transform.resize in the reference; scikit-image is not in this image, so the device kernels
are held to the scipy.ndimage calls scikit-image 0.24 makes (oracle.sift_resize) -- "parity unpinned" against the
library itself, stated in DESIGN.md.  Everything after the resize is pinned: the keypoint lists and the `good` list
of apply_sift must be the oracle chain's."""
import numpy as np
import pytest
import torch

from oracle import featurecraft_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(lib):
    from exvision_featurecraft_b200 import backend, synthetic
    from exvision_featurecraft_b200.utils import SIFT_scale_space, helper_functions, keypoint_descriptor

    SIFT_scale_space.set_dtype("mixed")
    SIFT_scale_space.set_variant("app")
    return backend, synthetic, helper_functions, keypoint_descriptor


def test_bgr_to_rgb_and_grayscale(mods):
    backend, synthetic, HF, KD = mods
    rng = np.random.default_rng(1)
    bgr = rng.integers(0, 256, (37, 53, 3)).astype(np.uint8)
    assert np.array_equal(HF.convert_BGR_to_RGB(bgr), bgr[..., ::-1])
    dev = HF.convert_BGR_to_RGB(torch.from_numpy(bgr).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), bgr[..., ::-1])
    assert np.array_equal(HF.convert_to_gray(bgr[..., ::-1].copy()), O.convert_to_gray(bgr[..., ::-1]))
    rgb = rng.random((29, 31, 3))
    got = KD.grayscale_device(torch.from_numpy(rgb).cuda()).cpu().numpy()
    assert np.abs(got - O.convert_to_grayscale(rgb)).max() < 3e-16  # BLAS dot vs explicit sum: last-bit order
    rgba = rng.random((8, 9, 4))
    assert np.abs(KD.grayscale_device(torch.from_numpy(rgba).cuda()).cpu().numpy() - O.convert_to_grayscale(rgba)).max() < 3e-16
    gray = rng.random((5, 6))
    assert KD.convert_to_grayscale(gray) is gray


@pytest.mark.parametrize("shape,ratio", [((300, 420, 3), None), ((1500, 1100), None), ((64, 48, 3), 2.5), ((97, 131), 0.37),
                                          ((40, 40), 1.0)])
def test_sift_resize_matches_the_scipy_restatement(mods, shape, ratio):
    backend, synthetic, HF, KD = mods
    rng = np.random.default_rng(shape[0])
    img = rng.random(shape)
    got, r = KD.sift_resize(img, ratio)
    ref, rr = O.sift_resize(img, ratio)
    assert r == rr and got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-15
    if len(shape) == 3:  # uint8 input goes through img_as_float
        u8 = (img * 255).astype(np.uint8)
        got8, _ = KD.sift_resize(u8, ratio)
        ref8, _ = O.sift_resize(u8, ratio)
        assert np.abs(got8 - ref8).max() < 2e-15 and got8.max() <= 1.0


def test_apply_sift_chain_equals_oracle_chain(mods):
    """computeKeypointsAndDescriptors x 2 + match on the device against the oracle's chain from the same working images
    (resize=False: the 1 MP working size of sift_resize costs the CPU oracle minutes per image; the resize itself is
    pinned by the test above and the full-size chain by tests/test_fullsize_gpu.py's reference goldens)."""
    backend, synthetic, HF, KD = mods
    main = synthetic.sift_image(72, 96, seed=12)
    tmpl = np.ascontiguousarray(main[10:58, 30:78])
    good, pa, pb = backend.apply_sift_matches(main, tmpl, 4, 2, 1.6, 0.03, 10, 0.6, resize=False)
    outs = [O.compute_keypoints_and_descriptors_from_base(O.rescale2(im), 4, 2, 1.6, 0.03, 10) for im in (main, tmpl)]
    (rpa, rda), (rpb, rdb) = outs
    assert pa == [tuple(map(float, p)) for p in rpa] and pb == [tuple(map(float, p)) for p in rpb]
    rda, rdb = np.array(rda, dtype=np.float32), np.array(rdb, dtype=np.float32)
    ok = ~np.isnan(rda).any(1)
    ref_pairs, ref_dist = O.match_indices(rda[ok], rdb, 0.6)
    ref_pairs[:, 0] = np.nonzero(ok)[0][ref_pairs[:, 0]]
    a, b = {g[:2] for g in good}, set(map(tuple, ref_pairs.tolist()))
    assert len(b) > 50 and len(a & b) >= 0.99 * len(a | b), (len(a), len(b), len(a & b))


def test_apply_sift_is_a_drop_in(mods):
    """Reference signature and return value (APP/FeatureCraft_Backend.py:480-523 / SIFT.py:683-707): two images and six
    parameters in, the cv2.drawMatches canvas of the resized pair out."""
    backend, synthetic, HF, KD = mods
    main = np.repeat(synthetic.sift_image(120, 160, seed=5)[:, :, None], 3, axis=2)
    tmpl = np.ascontiguousarray(main[20:80, 40:120])
    canvas = backend.apply_sift(main, tmpl, 4, 2, 1.6, 0.03, 10, 0.5)
    ratio = np.sqrt(1024 * 1024 / (120 * 160))
    hm, wm = int(round(120 * ratio)), int(round(160 * ratio))
    ht, wt = int(round(60 * ratio)), int(round(80 * ratio))
    assert canvas.dtype == np.uint8 and canvas.shape == (max(hm, ht), wm + wt, 3)
    # the device-resident variant of the same call returns the matches themselves
    good, pa, pb = backend.apply_sift_matches(main, tmpl, 4, 2, 1.6, 0.03, 10, 0.5, resize=True)
    assert len(good) > 500 and len(pa) > len(pb) > 1000
    kps = KD.kp_list_2_opencv_kp_list(pa[:3])
    assert [(k.pt[0], k.pt[1], k.size, k.angle) for k in kps] == [tuple(np.float32(v) for v in t) for t in O.kp_list_2_xy(pa[:3])]
