"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes exercise the shard partition, the count
gather and the ring exchange of the all-pairs matching step (the oracle is the matcher here)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exvision_featurecraft_b200 import sharding


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 4, 8):
            b = sharding.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_ring_schedule_is_a_latin_square():
    for w in (1, 2, 4, 8):
        sched = sharding.ring_schedule(w)
        pairs = [p for step in sched for p in step]
        assert sorted(pairs) == [(a, b) for a in range(w) for b in range(w)]
        for step in sched:
            assert sorted(t for _, t in step) == list(range(w))  # every block is visited by exactly one rank per step


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import featurecraft_oracle as O

        n_images = 5
        lo, hi = sharding.my_shard(n_images)
        # "extraction": image i yields 10 + i descriptors, deterministic content
        descs = []
        for i in range(lo, hi):
            rng = np.random.default_rng(100 + i)
            d = np.abs(rng.standard_normal((10 + i, 128))).astype(np.float32)
            descs.append(d / np.linalg.norm(d, axis=1, keepdims=True))
        counts = sharding.gather_counts([d.shape[0] for d in descs])
        assert counts == [10 + i for i in range(n_images)], counts
        local = torch.from_numpy(np.concatenate(descs))

        def match_fn(q, t):
            idx, dist_ = O.knn2_l2(q.numpy(), t.numpy())
            return idx[:, 0].copy(), dist_[:, 0].copy(), t.shape[0]

        res = sharding.all_pairs_match(local, match_fn, max_rows=64)
        out.put((rank, lo, hi, [(tr, r[0], r[1], r[2]) for tr, r in res]))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_extract_counts_and_ring_matching():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from oracle import featurecraft_oracle as O

    bounds = sharding.shard_bounds(5, world)
    blocks = []
    for lo, hi in bounds:
        ds = []
        for i in range(lo, hi):
            rng = np.random.default_rng(100 + i)
            d = np.abs(rng.standard_normal((10 + i, 128))).astype(np.float32)
            ds.append(d / np.linalg.norm(d, axis=1, keepdims=True))
        blocks.append(np.concatenate(ds))
    seen = set()
    for rank, lo, hi, results in got:
        assert (lo, hi) == bounds[rank]
        for train_rank, idx0, dist0, n_train in results:
            seen.add((rank, train_rank))
            assert n_train == blocks[train_rank].shape[0]
            ref_idx, ref_dist = O.knn2_l2(blocks[rank], blocks[train_rank])
            assert np.array_equal(idx0, ref_idx[:, 0]) and np.array_equal(dist0, ref_dist[:, 0])
    assert seen == {(a, b) for a in range(world) for b in range(world)}
