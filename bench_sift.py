"""SIFT, matching and pipeline workloads for bench.py (kept in a separate module so bench.py stays readable)."""
from __future__ import annotations

import numpy as np

from bench import _cpu_sift_task


class SiftWorkload:
    """BASELINE config[1]: SIFT keypoints + descriptors on 1024x1024 synthetic images (4 octaves, 5 scales),
    app filter variant, batched and sharded by image."""

    name = "sift"
    metric = "sift_megapixels_per_sec"
    unit = "MP/s"
    H = W = 1024
    CROP = 192  # CPU arms: crops of the same images (a whole 1024x1024 image costs the reference ~5 min per core)

    def __init__(self, args):
        self.batch = args.batch or 16
        self.mode = getattr(args, "sift_dtype", None) or "mixed"
        self.overrides = dict(getattr(args, "sift_params", None) or {})  # non-default point of the UI parameter ranges
        self.dtype = {"mixed": "f32 arithmetic, f64-certified decisions", "f64": "f64"}[self.mode]
        self.desc_per_image = None

    def describe(self, world):
        if self.overrides:
            what = "SIFT keypoints + descriptors at another point of the app's parameter ranges %s, %d synthetic 1024x1024 float64 images per step per GPU" % (
                self.overrides, self.batch)
        else:
            what = ("config[1]: SIFT keypoints + descriptors, %d synthetic 1024x1024 float64 images per step per GPU "
                    "(x2 up-sample -> 2048^2 base, 4 octaves, 5 Gaussian levels, sigma 1.6, app filter variant)" % self.batch)
        return {"workload": what,
                "images_per_step_per_gpu": self.batch, "image": [self.H, self.W], "sharding": "images, no collective",
                "l2": "per-step working set (%.1f GB) exceeds the 126 MB L2; no flush needed" % (self.batch * 0.5)}

    def units_per_step(self):
        return self.batch * self.H * self.W / 1e6

    def setup(self, rank, torch):
        from exvision_featurecraft_b200 import sift, synthetic

        self.torch, self.sift = torch, sift
        self.params = sift.SiftParams(dtype=self.mode, **self.overrides)
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
        self.last = res
        return res

    def e2e_begin(self):
        """Pipelined end-to-end driver: pinned input -> H2D -> kernels -> D2H into pinned output, the D2H of batch
        i overlapping the kernels of batch i+1 (sift.BatchExtractor, the public batch API; compact export: 8-byte
        keypoint records + float16 descriptors in certified mode)."""
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

    def kernels(self):
        tot = self.timer.totals()
        n = max(self.recorded_steps, 1)
        ms = tot.get("pyramid_o0", 0.0) / n
        # SURVEY 8(d), pyramid + DoG + extrema row: 4 B base read + 5 x 4 B levels per octave pixel
        alg = 24.0 * self.batch * (2 * self.H) * (2 * self.W)
        if self.mode == "f64":
            return [("gauss_mid_f64_kernel x 3 (octave 0, float64 blur only: levels in pairs through the radius-class kernel)", 2 * alg, ms, "hbm")]
        if self.overrides:
            nl = self.params.n_levels
            return [("gauss_octave_kernel<float> (octave 0, %d levels, blur only: non-default radii)" % nl,
                     (4.0 + 4.0 * nl) * self.batch * (2 * self.H) * (2 * self.W), ms, "hbm")]
        out = [("gauss_octave_packed_kernel<SEARCH> (octave 0: five blurs + DoG + certified 3x3x3 search fused; the float32 "
                "levels stay in shared memory)", alg, ms, "hbm")]
        blur = self._blur_only_ms()
        if blur:
            out.append(("gauss_octave_packed_kernel (octave 0, blur only, levels written: the stand-alone fc_gauss_octave)", alg, blur, "hbm"))
        return out

    def _blur_only_ms(self):
        """the same octave through the stand-alone blur kernel (TMA-staged tiles, writes the five levels), CUDA events"""
        import ctypes as C

        from exvision_featurecraft_b200 import _lib
        from exvision_featurecraft_b200.device import ptr, stream_ptr

        torch, sift = self.torch, self.sift
        b, h, w = self.batch, 2 * self.H, 2 * self.W
        base = torch.rand((b, h, w), dtype=torch.float32, device="cuda")
        out = torch.empty((b, 5, h, w), dtype=torch.float32, device="cuda")
        flat, lens = sift._level_taps(self.params)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(4):
            if i == 1:
                ev[0].record()
            _lib.call("fc_gauss_octave", ptr(base), _lib.FC_F32, b, h, w, 5, flat.ctypes.data_as(C.c_void_p), C.cast(lens, C.c_void_p),
                      0, 0, 0.03, 10.0, ptr(out), None, stream_ptr())
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / 3

    def traffic(self, tr):
        if self.overrides:
            return None
        key = "gauss_octave_search_kernel" if self.mode == "mixed" else "gauss_octave_kernel_f64"
        return tr[key]["bytes_per_octave_pixel"] * self.batch * (2 * self.H) * (2 * self.W)

    def stats(self):
        out = {"descriptors_per_image": self.desc_per_image, "scale_space": self.mode}
        if self.mode == "mixed" and getattr(self, "last", None) is not None:
            kp = sum(int(blk.keypoints.shape[0]) for blk in self.last.blocks)
            out["keypoints_per_image"] = kp / self.batch
            out["float64_decided_extrema_per_image"] = self.last.undecided_extrema / self.batch
            if self.last.recheck_counts is not None:
                out["float64_decided_orientations_per_image"] = int(self.last.recheck_counts.sum().item()) / self.batch
        return out

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
    cpu_sample_items = None  # one crop per core

    def cpu_items(self, n, step=0):
        # crop k = tile k % 25 of image seed (k // 25) % 4: non-overlapping 192x192 crops of the images the GPU arm uses
        out = []
        for i in range(n):
            k = step * n + i
            tile = k % 25
            out.append((self.H, self.W, (k // 25) % 4, (tile // 5) * self.CROP, (tile % 5) * self.CROP, self.CROP, self.CROP))
        return out

    def cpu_units(self, n):
        return n * self.CROP * self.CROP / 1e6

    def cpu_sample(self, n):
        return ("%d crops of %dx%d pixels of the config[1] images per step, one per process, through the oracle with the "
                "reference's direct 2-D convolutions (a whole 1024x1024 image takes the reference ~5 min on one core)"
                % (n, self.CROP, self.CROP))


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

    def _call(self):
        return self.matching.knn2(self.a, self.b, 0.3, "tensor")

    def step_device(self, record=False):
        if record and self.ev_i < len(self.ev):
            e0, e1 = self.ev[self.ev_i]
            self.ev_i += 1
            e0.record()
            out = self._call()
            e1.record()
            return out
        return self._call()

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

    def kernels(self):
        ms = [a.elapsed_time(b) for a, b in self.ev[: self.ev_i]]
        # whole knn2 call (prep + GEMM + re-rank + re-check); algorithmic FLOPs = 2 * nq * nt * 128
        return [("match_tc2_kernel (+prep, re-rank, re-check: the whole fc_match_knn2 call)", 2.0 * self.nq * self.nt * 128,
                 float(np.mean(ms)) if ms else None, "tensor")]

    def traffic(self, tr):
        t = tr["match_tc_kernel"]
        return t["bytes_per_launch"] if (self.nq, self.nt) == (t["n_query"], t["n_train"]) else None

    def extra(self):
        return self.matching.last_stats()

    cpu_task = staticmethod(_cpu_match_task)

    def cpu_items(self, n, step=0):
        return [(2048, 8192, step * n + s) for s in range(n)]

    def cpu_units(self, n):
        return n * 2048 * 8192 / 1e9

    def cpu_sample(self, n):
        return "%d tasks of 2048 x 8192 descriptors: float32 brute force + ratio test of the oracle (cv2.BFMatcher's arithmetic)" % n


def ring_block(n, rank):
    """Descriptor block of one rank for the all-pairs workload: noisy copies of rows of ONE base set every rank
    generates identically, so that true matches exist between any two ranks' blocks (unrelated random blocks would
    make every ratio test fail and the ring check vacuous)."""
    base = np.abs(np.random.default_rng(1234).standard_normal((n, 128), dtype=np.float32))
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    rng = np.random.default_rng(500 + rank)
    rows = rng.permutation(n)
    a = base[rows] + rng.normal(0, 0.01, (n, 128)).astype(np.float32)
    a[::3] = np.abs(rng.standard_normal((len(a[::3]), 128), dtype=np.float32)) / 11.3  # a third without a partner
    return np.ascontiguousarray(a)


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
        from exvision_featurecraft_b200 import matching, sharding

        self.torch, self.matching, self.sharding = torch, matching, sharding
        self.world = sharding.world_info()[1]
        self.host_a = torch.from_numpy(ring_block(self.nq, rank)).pin_memory()
        self.a = self.host_a.cuda()
        self.ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(64)]
        self.ev_i = 0
        # before anything is timed: the ring's answer for the LAST block to arrive (it crossed world-1 hops) must be the
        # exact engine's answer for that pair of blocks
        res = self._run()
        train_rank, (idx, dist, good) = res[-1]
        ei, ed, eg = self.matching.knn2(self.a, torch.from_numpy(ring_block(self.nq, train_rank)).cuda(), 0.3, "exact")
        g = eg.nonzero().flatten()
        assert torch.equal(good, eg) and torch.equal(idx[g, 0], ei[g, 0]), "ring all-pairs result differs from the exact engine"
        self.verified = {"train_rank": int(train_rank), "good": int(g.numel()), "equal_to_exact_engine": True}

    def stats(self):
        return {"ring_check": self.verified}

    def _run(self):
        return self.sharding.all_pairs_match(self.a, lambda q, t: self.matching.knn2(q, t, 0.3, "tensor"), max_rows=self.nq)

    def _call(self):
        return self._run()

    def e2e_begin(self):
        self.pending = None

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

    def e2e_end(self):
        return self.nq * 512, 0

    def kernels(self):
        ms = [a.elapsed_time(b) for a, b in self.ev[: self.ev_i]]
        return [("match_tc2_kernel (+prep, re-rank, ring exchange)", 2.0 * self.nq * self.nt * self.world * 128,
                 float(np.mean(ms)) if ms else None, "tensor")]


class Pipeline4kWorkload:
    """BASELINE configs[4] as a pipeline, one step of it per GPU: SIFT keypoints + descriptors of one 3840x2160 image
    (certified mode, descriptors stay in HBM) and 2-NN ratio-test matching of that image against the images of ALL ranks
    -- the descriptor blocks visit around the NCCL ring (sharding.all_pairs_match).  The images are overlapping crops of
    one synthetic scene, so true matches exist between every pair.  value = input megapixels per second through the
    whole pipeline; the matching grows with the number of ranks (every image meets every image), which `stats` breaks out."""

    name = "pipeline4k"
    metric = "sift4k_extract_and_allpairs_match_megapixels_per_sec"
    unit = "MP/s"
    dtype = "f32 arithmetic, f64-certified decisions; fp16 tcgen05 matching"
    H, W = 2160, 3840

    def __init__(self, args):
        self.world = 1
        self.n_desc = None
        self.n_good = None

    def describe(self, world):
        return {"workload": "config[4] pipeline: per step and GPU one synthetic 3840x2160 image -> SIFT (x2 up-sample, 4 octaves) "
                            "-> its descriptors matched against the images of all %d GPUs (ratio 0.3), ring of NCCL send/recv" % world,
                "image": [self.H, self.W], "sharding": "one image per rank and step; descriptor blocks travel the ring",
                "l2": "one image's scale space is 2.7 GB: far beyond the 126 MB L2"}

    def units_per_step(self):
        return self.H * self.W / 1e6

    def setup(self, rank, torch):
        from exvision_featurecraft_b200 import matching, sharding, sift, synthetic

        self.torch, self.sift, self.matching, self.sharding = torch, sift, matching, sharding
        self.world = sharding.world_info()[1]
        scene = synthetic.sift_image(self.H + 64 * 8, self.W + 64 * 8, seed=77, max_blobs=3000)
        off = 64 * (rank % 8)
        img = np.ascontiguousarray(scene[off:off + self.H, off:off + self.W])
        self.host_in = torch.from_numpy(img[None]).pin_memory()
        self.dev_in = self.host_in.cuda()
        self.params = sift.SiftParams(dtype="mixed")
        self.ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(16)]
        self.ev_i = 0

    def _run(self, dev_in, ev=None):
        if ev:
            ev[0].record()
        res = self.sift.detect_and_describe(self.sift.upsample2(dev_in, self.params), self.params, out_dtype=self.torch.float32)
        if ev:
            ev[1].record()
        cap = int(res.total * 1.2) + 4096  # ring buffers: the other ranks' images yield about as many descriptors
        if self.world > 1:
            t = self.torch.tensor([cap], dtype=self.torch.int64, device="cuda")
            self.torch.distributed.all_reduce(t, op=self.torch.distributed.ReduceOp.MAX)
            cap = int(t.item())
        out = self.sharding.all_pairs_match(res.descriptors, lambda q, t: self.matching.knn2(q, t, 0.3, "tensor"), max_rows=cap)
        if ev:
            ev[2].record()
        self.n_desc = res.total
        return res, out

    def step_device(self, record=False):
        ev = None
        if record and self.ev_i < len(self.ev):
            ev = self.ev[self.ev_i]
            self.ev_i += 1
        return self._run(self.dev_in, ev)

    def step_e2e(self):
        dev = self.host_in.cuda(non_blocking=True)
        res, out = self._run(dev)
        d2h = 0
        self.n_good = []
        for _, (idx, dist, good) in out:
            q = self.torch.nonzero(good).flatten()
            host = (q.cpu(), idx[q, 0].cpu(), dist[q, 0].cpu())
            self.n_good.append(int(q.numel()))
            d2h += sum(t.numel() * t.element_size() for t in host)
        return self.host_in.numel() * 8, d2h

    def kernels(self):
        n = self.ev_i
        if not n:
            return []
        ms = float(np.mean([e[1].elapsed_time(e[2]) for e in self.ev[:n]]))
        return [("match_tc2_kernel (+prep, re-rank, ring exchange): %d image pairs per GPU and step" % self.world,
                 2.0 * float(self.n_desc) ** 2 * self.world * 128, ms, "tensor")]

    def traffic(self, tr):
        return None

    def stats(self):
        n = max(self.ev_i, 1)
        return {"descriptors_per_image": self.n_desc, "good_matches_per_pair": self.n_good,
                "extract_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in self.ev[:n]])) if self.ev_i else None,
                "match_ms": float(np.mean([e[1].elapsed_time(e[2]) for e in self.ev[:n]])) if self.ev_i else None,
                "image_pairs_per_step": self.world * self.world}
