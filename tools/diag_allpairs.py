import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from exvision_featurecraft_b200 import matching
from bench_sift import ring_block, match_sets
n = 65536
a = torch.from_numpy(ring_block(n, 0)).cuda(); b = torch.from_numpy(ring_block(n, 1)).cuda()
ms_a, ms_b = match_sets(n, n, 0)
ma, mb = torch.from_numpy(ms_a).cuda(), torch.from_numpy(ms_b).cuda()
def t(q, tr, label):
    for _ in range(2): matching.knn2(q, tr, 0.3, "tensor")
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): out = matching.knn2(q, tr, 0.3, "tensor")
    e1.record(); torch.cuda.synchronize()
    print(label, "%.3f ms" % (e0.elapsed_time(e1) / 5), matching.last_stats(), int(out[2].sum()))
t(ma, mb, "match_sets")
t(a, b, "ring a-b")
t(a, a, "ring a-a")
