"""Two-rank NCCL run of the ring all-pairs matching (skipped unless the box has >= 2 GPUs; the same logic runs on
gloo in test_sharding_cpu.py).  Launch: torchrun --nproc-per-node 2 -m pytest tests/test_multigpu.py -m gpu, or
plain pytest (spawns the two ranks itself)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from bench_sift import match_sets
        from exvision_featurecraft_b200 import matching, sharding

        a, _ = match_sets(3000 + 500 * rank, 16, seed=40 + rank)
        local = torch.from_numpy(a).cuda()
        res = sharding.all_pairs_match(local, lambda q, t: matching.knn2(q, t, 0.8, "tensor"), max_rows=4096)
        out.put((rank, [(tr, idx.cpu().numpy(), good.cpu().numpy()) for tr, (idx, d, good) in res]))
    finally:
        dist.destroy_process_group()


def test_two_rank_ring_matching_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from bench_sift import match_sets
    from exvision_featurecraft_b200 import matching

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    blocks = [match_sets(3000 + 500 * r, 16, seed=40 + r)[0] for r in range(2)]
    for r in range(2):
        assert sorted(tr for tr, _, _ in got[r]) == [0, 1]
        for tr, idx, good in got[r]:
            ei, ed, eg = matching.knn2(blocks[r], blocks[tr], 0.8, "exact")
            assert np.array_equal(good, eg.cpu().numpy())
            g = np.nonzero(good)[0]
            assert np.array_equal(idx[g, 0], ei.cpu().numpy()[g, 0])
