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
    dtype = "f32"
    H = W = 1024
    cpu_sample_items = 16
    cpu_sample_note = "(256x256 crops of the same generator: a 1024x1024 image takes ~5 min per core)"

    def __init__(self, args):
        self.batch = args.batch or 8
        self.mode = getattr(args, "sift_dtype", None) or "f32"
        self.dtype = self.mode

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
        self.timer = sift.StageTimer()
        self.recorded_steps = 0

    def _run(self, dev_in):
        base = self.sift.upsample2(dev_in, self.params)
        return self.sift.detect_and_describe(base, self.params, out_dtype=self.torch.float32)

    def step_device(self, record=False):
        if record:
            self.sift.TIMER = self.timer
            self.recorded_steps += 1
        try:
            res = self._run(self.dev_in)
        finally:
            self.sift.TIMER = None
        self.desc_per_image = sum(b.descriptors.shape[0] for b in res.blocks if b.descriptors is not None) / self.batch
        return res

    def step_e2e(self):
        d = self.host_in.cuda(non_blocking=True)
        res = self._run(d)
        out = res.assemble(self.batch)
        d2h = sum(p.nbytes + q.nbytes for p, q in out)
        return self.host_in.numel() * 8, d2h

    def dominant_kernel(self):
        tot = self.timer.totals()
        n = max(self.recorded_steps, 1)
        ms = tot.get("pyramid_o0", 0.0) / n
        alg = 24.0 * self.batch * (2 * self.H) * (2 * self.W)  # 4 B base read + 5 x 4 B levels written per octave pixel
        if self.mode == "f64":
            alg *= 2
        return "gauss_octave_kernel (octave 0)", alg, ms, "hbm"

    def extra(self):
        n = max(self.recorded_steps, 1)
        return {k: round(v / n, 4) for k, v in sorted(self.timer.totals().items())}

    cpu_task = staticmethod(_cpu_sift_task)

    def cpu_items(self, n):
        # the oracle needs ~5 min per 1024^2 image per core: the CPU arm uses 256^2 crops of the same generator
        return [(256, 256, s) for s in range(n)]

    def cpu_units(self, n):
        return n * 256 * 256 / 1e6


class MatchWorkload:
    pass
