"""2-rank timing of the descriptor ring step (development helper): torchrun --nproc-per-node 2 tools/diag_ring.py"""
import os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from exvision_featurecraft_b200 import sharding, matching
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
n = 65536
a = torch.rand((n, 128), device="cuda")
a /= a.norm(dim=1, keepdim=True)
def timeit(fn, label, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    if rank == 0: print(label, "%.3f ms" % ((time.perf_counter() - t0) * 1e3 / reps), flush=True)
timeit(lambda: sharding.ring_shift(a, n), "ring_shift blocking")
timeit(lambda: matching.knn2(a, a, 0.3, "tensor"), "knn2 alone")
timeit(lambda: sharding.all_pairs_match(a, lambda q, t: matching.knn2(q, t, 0.3, "tensor"), max_rows=n), "all_pairs_match")
def raw():
    send = a; recv = torch.empty_like(a)
    ops = [dist.P2POp(dist.isend, send, (rank - 1) % 2), dist.P2POp(dist.irecv, recv, (rank + 1) % 2)]
    for r in dist.batch_isend_irecv(ops): r.wait()
timeit(raw, "raw batch_isend_irecv 34 MB")
import argparse
from bench_sift import AllPairsWorkload, ring_block
w = AllPairsWorkload(argparse.Namespace(batch=None))
w.setup(rank, torch)
timeit(lambda: w.step_device(), "AllPairsWorkload.step_device")
timeit(lambda: w.step_device(record=True), "AllPairsWorkload.step_device(record)")
b = torch.from_numpy(ring_block(n, rank)).cuda()
timeit(lambda: sharding.all_pairs_match(b, lambda q, t: matching.knn2(q, t, 0.3, "tensor"), max_rows=n), "all_pairs_match ring_block data")
timeit(lambda: matching.knn2(b, b, 0.3, "tensor"), "knn2 ring_block self")
dist.destroy_process_group()
