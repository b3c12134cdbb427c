"""worst descriptors of the config2_main golden comparison (diagnostic)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.make_golden_large import inputs, upsample2_host, projection_matrix
from exvision_featurecraft_b200 import sift
g = np.load("tests/golden/sift_large.npz")
base = upsample2_host(inputs()["config2_main"])
for dt in ("mixed", "f64"):
    r = sift.detect_and_describe(base, sift.SiftParams(dtype=dt), out_dtype=torch.float64 if dt == "f64" else torch.float32)
    d = r.descriptors.cpu().numpy().astype(np.float64)
    proj = np.nan_to_num(d) @ projection_matrix()
    err = np.abs(proj - g["config2_main/desc_proj"]).max(1)
    worst = np.argsort(err)[-5:]
    pts = sift.unpack_records(r.records.cpu().numpy())
    print(dt, "worst projection errors", err[worst], "rows", worst)
    print(pts[worst])
    print("descriptor min/max of worst", d[worst].min(1), d[worst].max(1))
    if dt == "mixed":
        keep = d
    else:
        l2 = np.linalg.norm(keep - d, axis=1)
        w = np.argsort(l2)[-5:]
        print("mixed vs f64 L2 worst", l2[w], w, pts[w])
        print("count above 2e-5:", int((l2 > 2e-5).sum()), "of", len(l2))
