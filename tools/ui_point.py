"""time detect_and_describe at one point of the UI parameter ranges (development helper; needs a GPU)
usage: python tools/ui_point.py <sigma> <s> [n_octaves] [batch] [dtype]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from exvision_featurecraft_b200 import sift, synthetic

sigma, s = float(sys.argv[1]), int(sys.argv[2])
n_oct = int(sys.argv[3]) if len(sys.argv) > 3 else 4
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dtype = sys.argv[5] if len(sys.argv) > 5 else "mixed"
p = sift.SiftParams(n_octaves=n_oct, s_value=s, sigma_base=sigma, dtype=dtype)
imgs = np.stack([synthetic.sift_image(1024, 1024, seed=i % 4, max_blobs=1024) for i in range(batch)])
dev = torch.from_numpy(imgs).cuda()

def run():
    return sift.detect_and_describe(sift.upsample2(dev, p), p, out_dtype=torch.float32)

for _ in range(3):
    res = run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    res = run()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
t = sift.StageTimer()
sift.TIMER = t
for _ in range(3):
    run()
sift.TIMER = None
torch.cuda.synchronize()
print("sigma %.2f s %d octaves %d batch %d %s: %.3f ms/step = %.1f MP/s, %d descriptors/image" % (
    sigma, s, n_oct, batch, dtype, ms, batch * 1.048576 / ms * 1e3, res.total // batch))
print("  radii", [sift.kernel_radius(x) for x in sift.sigma_levels(sigma, s)])
print("  stages", {k: round(v / 3, 3) for k, v in sorted(t.totals().items())})
