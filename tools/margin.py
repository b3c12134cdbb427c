"""largest float32-vs-float64 DoG error on the BASELINE-size image relative to the certification margin (development helper)"""
import sys
sys.path.insert(0, ".")
import torch
from exvision_featurecraft_b200 import sift, synthetic
for params in (dict(), dict(sigma_base=3.0, s_value=2), dict(sigma_base=2.5, s_value=3)):
    for seed in (2, 5):
        img = synthetic.sift_image(1024, 1024, seed=seed, max_blobs=1024)
        pm, pf = sift.SiftParams(dtype="mixed", **params), sift.SiftParams(dtype="f64", **params)
        base = sift.upsample2(img, pf)
        bound = float(base.abs().max())
        o32, o64 = sift.scale_space(sift.upsample2(img, pm), pm), sift.scale_space(base, pf)
        worst = 0.0
        for a, b in zip(o32, o64):
            worst = max(worst, float(((a.gauss[:, 1:] - a.gauss[:, :-1]).double() - (b.gauss[:, 1:] - b.gauss[:, :-1])).abs().max()))
        print(params, "seed", seed, "max DoG error %.3g = %.3f of the margin (%.3g)" % (worst, worst / (sift.EXTREMA_EPS * bound), sift.EXTREMA_EPS * bound))
