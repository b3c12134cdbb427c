"""per-phase timing of one all-pairs ring step with the bench's descriptor blocks (development helper)"""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from exvision_featurecraft_b200 import sharding, matching
from bench_sift import ring_block
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = 65536
a = torch.from_numpy(ring_block(n, rank)).cuda()
other = torch.from_numpy(ring_block(n, 1 - rank)).cuda()
def sync_time(fn, label, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): out = fn()
    torch.cuda.synchronize()
    print("rank", rank, label, "%.3f ms" % ((time.perf_counter() - t0) * 1e3 / reps), matching.last_stats(), flush=True)
    return out
sync_time(lambda: matching.knn2(a, a, 0.3, "tensor"), "knn2(a, a)")
sync_time(lambda: matching.knn2(a, other, 0.3, "tensor"), "knn2(a, local copy of the other rank's block)")
recv = sharding.ring_shift(a, n)
if rank == 0: print("received == other:", bool(torch.equal(recv, other)), recv.shape, recv.dtype, recv.is_contiguous(), recv.data_ptr() % 256, flush=True)
sync_time(lambda: matching.knn2(a, recv, 0.3, "tensor"), "knn2(a, received block)")
rc = recv.clone()
sync_time(lambda: matching.knn2(a, rc, 0.3, "tensor"), "knn2(a, clone of received block)")
sync_time(lambda: sharding.all_pairs_match(a, lambda q, t: matching.knn2(q, t, 0.3, "tensor"), max_rows=n), "all_pairs_match")
sync_time(lambda: sharding.ring_shift(a, n), "ring_shift")
dist.destroy_process_group()
