"""SIFT and matching workloads for bench.py (kept in a separate module so bench.py stays readable)."""
from __future__ import annotations

import numpy as np

from bench import _cpu_sift_task


class SiftWorkload:
    """BASELINE config[1]: SIFT keypoints + descriptors on 1024x1024 synthetic images (4 octaves, 5 scales),
    app filter variant, batched and sharded by image."""

    name = "sift"
    metric = "sift_megapixels_per_sec"
    unit = "MP/s"
    dtype = "mixed"
    H = W = 1024
    cpu_sample_items = 16
    cpu_sample_note = "(256x256 crops of the same generator, oracle with separable blurs: the reference's direct 2-D convolution is ~5x slower and a 1024x1024 image takes ~5 min per core)"

    def __init__(self, args):
        self.batch = args.batch or 16
        self.mode = getattr(args, "sift_dtype", None) or "mixed"
        self.dtype = {"mixed": "f32 arithmetic + f64-certified decisions", "f32": "f32", "f64": "f64"}[self.mode]

    def describe(self, world):
        return {"workload": "config[1]: SIFT keypoints + descriptors, %d synthetic 1024x1024 float64 images per step per GPU "
                            "(x2 up-sample -> 2048^2 base, 4 octaves, 5 Gaussian levels, app filter variant, %s scale space)"
                            % (self.batch, self.mode),
                "images_per_step_per_gpu": self.batch, "image": [self.H, self.W], "sharding": "images, no collective",
                "descriptors_per_image": getattr(self, "desc_per_image", None),
                "l2": "per-step working set (%.1f GB) exceeds the 126 MB L2; no flush needed" % (self.batch * 0.35)}

    def units_per_step(self):
        return self.batch * self.H * self.W / 1e6

    def setup(self, rank, torch):
        from exvision_featurecraft_b200 import sift, synthetic

        self.torch, self.sift = torch, sift
        self.params = sift.SiftParams(dtype=self.mode)
        imgs = np.stack([synthetic.sift_image(self.H, self.W, seed=rank * 4 + i, max_blobs=1024) for i in range(min(4, self.batch))])
        host = np.concatenate([imgs] * ((self.batch + len(imgs) - 1) // len(imgs)))[: self.batch]
        self.host_in = torch.from_numpy(np.ascontiguousarray(host)).pin_memory()
        self.dev_in = self.host_in.cuda(non_blocking=True)
        # inside the timed region only the dominant kernel is bracketed by CUDA events (one pair per step); the full
        # per-stage breakdown comes from extra untimed steps afterwards (two events per stage and block are ~0.5 ms of
        # host time per step, which would be charged to `value`)
        self.timer = sift.StageTimer(only={"pyramid_o0"})
        self.recorded_steps = 0

    def _run(self, dev_in):
        base = self.sift.upsample2(dev_in, self.params)
        return self.sift.detect_and_describe(base, self.params, out_dtype=self.torch.float32)  # descriptors stay on the device

    def step_device(self, record=False):
        if record:
            self.sift.TIMER = self.timer
            self.recorded_steps += 1
        try:
            res = self._run(self.dev_in)
        finally:
            self.sift.TIMER = None
        self.desc_per_image = res.total / self.batch
        return res

    def e2e_begin(self):
        """Pipelined end-to-end driver: pinned input -> H2D -> kernels -> D2H into pinned output, the D2H of batch
        i overlapping the kernels of batch i+1 (sift.BatchExtractor, the public batch API)."""
        cap = int(self.desc_per_image * self.batch * 1.25) + 4096
        self.extractor = self.sift.BatchExtractor(self.params, self.batch, self.H, self.W, cap, depth=2)
        for s in range(2):
            self.extractor.host_in[s].copy_(self.host_in)
        self.pending = None
        self.io_bytes = (0, 0)

    def step_e2e(self):
        t = self.extractor.submit(None, prefetch_next=True)  # the pinned staging buffers already hold the images
        if self.pending is not None:
            rec, desc, cnt = self.extractor.collect(self.pending, raw=True)
            self.io_bytes = self.extractor.bytes_per_batch(rec.shape[0])
        self.pending = t
        return self.io_bytes

    def e2e_end(self):
        rec, desc, cnt = self.extractor.collect(self.pending, raw=True)
        self.pending = None
        self.io_bytes = self.extractor.bytes_per_batch(rec.shape[0])
        return self.io_bytes

    def dominant_kernel(self):
        tot = self.timer.totals()
        n = max(self.recorded_steps, 1)
        ms = tot.get("pyramid_o0", 0.0) / n
        alg = 24.0 * self.batch * (2 * self.H) * (2 * self.W)  # 4 B base read + 5 x 4 B levels written per octave pixel
        if self.mode == "f64":
            alg *= 2
        return "gauss_octave_packed_kernel (octave 0)" if self.mode != "f64" else "gauss_octave_kernel (octave 0)", alg, ms, "hbm"

    def traffic(self, tr):
        return tr["gauss_octave_packed_kernel"]["bytes_per_octave_pixel"] * self.batch * (2 * self.H) * (2 * self.W) * (2 if self.mode == "f64" else 1)

    def extra(self):
        full = self.sift.StageTimer()
        n = 3
        self.sift.TIMER = full
        try:
            for _ in range(n):
                self._run(self.dev_in)
        finally:
            self.sift.TIMER = None
        self.torch.cuda.synchronize()
        return {k: round(v / n, 4) for k, v in sorted(full.totals().items())}

    cpu_task = staticmethod(_cpu_sift_task)

    def cpu_items(self, n):
        # the oracle needs ~5 min per 1024^2 image per core: the CPU arm uses 256^2 crops of the same generator
        return [(256, 256, s) for s in range(n)]

    def cpu_units(self, n):
        return n * 256 * 256 / 1e6


def _cpu_match_task(args):
    nq, nt, seed = args
    import time

    from oracle import featurecraft_oracle as O

    a, b = match_sets(nq, nt, seed)
    t0 = time.perf_counter()
    O.match_indices(a, b, 0.3)
    return time.perf_counter() - t0, nq


def match_sets(nq, nt, seed):
    """Unit-norm non-negative descriptor sets; half of the queries are noisy copies of train rows."""
    rng = np.random.default_rng(seed)
    b = np.abs(rng.standard_normal((nt, 128), dtype=np.float32))
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    a = np.abs(rng.standard_normal((nq, 128), dtype=np.float32))
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    h = nq // 2
    a[:h] = b[rng.integers(0, nt, h)] + rng.normal(0, 0.01, (h, 128)).astype(np.float32)
    return a, b


class MatchWorkload:
    """Brute-force 2-NN + ratio test between two descriptor sets of the size one 1024x1024 image yields
    (BASELINE configs[2]/[4] matching leg): value = descriptor pairs compared per second."""

    name = "match"
    metric = "match_descriptor_pairs_per_sec"
    unit = "Gpairs/s"
    dtype = "fp16 tcgen05 candidates, fp32 exact re-rank"
    cpu_sample_items = 16
    cpu_sample_note = "(2048 x 8192 descriptor sets per task: float32 brute force of the oracle)"

    def __init__(self, args):
        self.nq = self.nt = (args.batch or 128) * 1024

    def describe(self, world):
        return {"workload": "2-NN ratio-test matching of %d query x %d train 128-d float32 descriptors per step per GPU "
                            "(synthetic unit vectors, half of the queries have a true match)" % (self.nq, self.nt),
                "sharding": "independent pairs of descriptor sets, no collective",
                "l2": "operands per step (%.0f MB fp32 + %.0f MB fp16) vs 126 MB L2: train tiles are re-read from L2 by design"
                      % ((self.nq + self.nt) * 512 / 1e6, (self.nq + self.nt) * 288 / 1e6)}

    def units_per_step(self):
        return self.nq * self.nt / 1e9

    def setup(self, rank, torch):
        from exvision_featurecraft_b200 import matching

        self.torch, self.matching = torch, matching
        a, b = match_sets(self.nq, self.nt, rank)
        self.host_a = torch.from_numpy(a).pin_memory()
        self.host_b = torch.from_numpy(b).pin_memory()
        self.a = self.host_a.cuda()
        self.b = self.host_b.cuda()
        self.ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(64)]
        self.ev_i = 0

    def step_device(self, record=False):
        if record and self.ev_i < len(self.ev):
            e0, e1 = self.ev[self.ev_i]
            self.ev_i += 1
            e0.record()
            out = self.matching.knn2(self.a, self.b, 0.3, "tensor")
            e1.record()
            return out
        return self.matching.knn2(self.a, self.b, 0.3, "tensor")

    def e2e_begin(self):
        """Pipelined end-to-end driver (matching.BatchMatcher, the public batch API): the H2D of pair i+1 runs while
        pair i is matched, results land in pinned host memory."""
        self.matcher = self.matching.BatchMatcher(self.nq, self.nt, 0.3, depth=2)
        for s in range(2):
            self.matcher.host_q[s].copy_(self.host_a)
            self.matcher.host_t[s].copy_(self.host_b)
        self.pending = None
        self.n_good = None

    def step_e2e(self):
        t = self.matcher.submit(None, None)  # the pinned staging buffers already hold the descriptors
        if self.pending is not None:
            pairs, _ = self.matcher.collect(self.pending)
            self.n_good = len(pairs)
        self.pending = t
        return self.matcher.bytes_per_pair()

    def e2e_end(self):
        pairs, _ = self.matcher.collect(self.pending)
        self.pending = None
        self.n_good = len(pairs)
        return self.matcher.bytes_per_pair()

    def dominant_kernel(self):
        ms = [a.elapsed_time(b) for a, b in self.ev[: self.ev_i]]
        # whole knn2 call (prep + GEMM + re-rank + re-check); algorithmic FLOPs = 2 * nq * nt * 128
        return "match_tc_kernel (+prep, re-rank, re-check)", 2.0 * self.nq * self.nt * 128, float(np.mean(ms)) if ms else None, "tensor"

    def traffic(self, tr):
        t = tr["match_tc_kernel"]
        return t["bytes_per_launch"] if (self.nq, self.nt) == (t["n_query"], t["n_train"]) else None

    def extra(self):
        return self.matching.last_stats()

    cpu_task = staticmethod(_cpu_match_task)

    def cpu_items(self, n):
        return [(2048, 8192, s) for s in range(n)]

    def cpu_units(self, n):
        return n * 2048 * 8192 / 1e9


class AllPairsWorkload(MatchWorkload):
    """BASELINE configs[4] matching leg, scaled: every rank keeps its block of descriptors resident and matches it
    against the block of every rank; the blocks travel around a ring of NCCL send/recv (sharding.all_pairs_match).
    The one workload with a real exchange step; value = descriptor pairs compared per second over all ranks."""

    name = "allpairs"
    metric = "allpairs_match_descriptor_pairs_per_sec"

    def __init__(self, args):
        self.nq = self.nt = (args.batch or 64) * 1024
        self.world = 1

    def describe(self, world):
        return {"workload": "all-pairs 2-NN ratio-test matching: %d descriptors per GPU against the blocks of all %d GPUs per step "
                            "(ring of NCCL send/recv, %d-row blocks)" % (self.nq, world, self.nq),
                "sharding": "descriptor blocks per rank; exchange = (world-1)-step ring, next block in flight while the current one is matched",
                "l2": "blocks (%.0f MB fp32 each) exceed L2 together with the fp16 operands" % (self.nq * 512 / 1e6)}

    def units_per_step(self):
        return self.nq * self.nt * self.world / 1e9

    def setup(self, rank, torch):
        super().setup(rank, torch)
        from exvision_featurecraft_b200 import sharding

        self.sharding = sharding
        self.world = sharding.world_info()[1]

    def _run(self):
        return self.sharding.all_pairs_match(self.a, lambda q, t: self.matching.knn2(q, t, 0.3, "tensor"), max_rows=self.nq)

    def step_device(self, record=False):
        if record and self.ev_i < len(self.ev):
            e0, e1 = self.ev[self.ev_i]
            self.ev_i += 1
            e0.record()
            out = self._run()
            e1.record()
            return out
        return self._run()

    def step_e2e(self):
        a = self.host_a.cuda(non_blocking=True)
        keep, self.a = self.a, a
        res = self._run()
        self.a = keep
        d2h = 0
        for _, (idx, dist, good) in res:
            q = self.torch.nonzero(good).flatten()
            out = (q.cpu(), idx[q, 0].cpu(), dist[q, 0].cpu())
            d2h += sum(t.numel() * t.element_size() for t in out)
        return self.nq * 512, d2h

    def dominant_kernel(self):
        ms = [a.elapsed_time(b) for a, b in self.ev[: self.ev_i]]
        return ("match_tc_kernel (+prep, re-rank, ring exchange)", 2.0 * self.nq * self.nt * self.world * 128,
                float(np.mean(ms)) if ms else None, "tensor")
