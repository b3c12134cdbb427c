#!/usr/bin/env python
"""Benchmark of the B200 feature-extraction hot path (contract: see DESIGN.md, "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One rank per GPU (torchrun for N > 1).  A "step" is one pass of the hot path over one batch of synthetic inputs that
is already resident in HBM (`value`) or starts in pinned host memory and ends with the results back on the host
(`e2e`).  Without --workload the line is BASELINE.json's headline -- SIFT keypoints + descriptors on 1024x1024 images
(configs[1]) -- and carries the rest of the metric under `secondary`: Harris + lambda-minus (configs[3]), descriptor
matching, the 4K extract-and-match pipeline with its ring exchange (configs[4]) and the float64 SIFT mode, each with
its own value / e2e / roofline / cpu_baseline.  Images are independent, so ranks never exchange data except in the
ring of the all-pairs matching: weak scaling, value = units of all ranks / max-over-ranks device time.

`--impl reference` times the CPU oracle (oracle/featurecraft_oracle.py, the restatement of the reference's NumPy
path with its direct 2-D convolutions -- the reference itself is Python and cannot travel to the GPU box) on all host
cores, on the same `config`, each step a bounded sample of the workload's step (stated in cpu_baseline.sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


# ----------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi in the background during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load" = samples in the upper half of the observed power range
        if sm:
            cut = (max(power) + min(power)) / 2 if len(power) > 2 else -1
            loaded = [s for s, p in zip(sm, power) if p >= cut] or sm
            return {"sm_mhz": float(np.median(loaded)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": max(power)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons), "samples": 0}


# ----------------------------------------------------------------------------------------------
# CPU arms (oracle) -- run in worker processes, one item per task
# ----------------------------------------------------------------------------------------------
def _cpu_harris_task(args):
    h, w, seed = args
    from exvision_featurecraft_b200 import synthetic
    from oracle import featurecraft_oracle as O

    gray = synthetic.corner_image(h, w, seed)
    t0 = time.perf_counter()
    r = O.harris_response(gray, 5, 0.04)
    idx, vals = O.threshold_corners(r, 10000)
    lam = O.lambda_minus_response(gray)
    rows, cols = np.where(lam > 0.01 * lam.max())
    return time.perf_counter() - t0, int(idx.shape[0]) + int(rows.shape[0])


_SIFT_IMAGE_CACHE: dict = {}


def _cpu_sift_task(args):
    """One crop of a config[1] image through the oracle with the reference's own direct 2-D convolutions."""
    h, w, seed, y0, x0, ch, cw = args
    from exvision_featurecraft_b200 import synthetic
    from oracle import featurecraft_oracle as O

    key = (h, w, seed)
    if key not in _SIFT_IMAGE_CACHE:
        _SIFT_IMAGE_CACHE.clear()
        _SIFT_IMAGE_CACHE[key] = synthetic.sift_image(h, w, seed=seed, max_blobs=1024)
    img = np.ascontiguousarray(_SIFT_IMAGE_CACHE[key][y0:y0 + ch, x0:x0 + cw])
    t0 = time.perf_counter()
    base = O.rescale2(img)
    pts, desc = O.compute_keypoints_and_descriptors_from_base(base, 4, 2, 1.6, 0.03, 10, "app", "direct")
    return time.perf_counter() - t0, len(desc)


def cpu_pool_run(task, items, cores):
    """Run `task` over `items` on `cores` processes; returns (wall seconds, results)."""
    import multiprocessing as mp

    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    if cores <= 1:
        res = [task(it) for it in items]
    else:
        with ctx.Pool(cores) as pool:
            res = pool.map(task, items, chunksize=1)
    return time.perf_counter() - t0, res


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
class HarrisWorkload:
    """BASELINE config 4: batched Harris + lambda-minus over 2048x2048 uint8 images, sharded by image."""

    name = "harris"
    metric = "harris_lambda_minus_megapixels_per_sec"
    unit = "MP/s"
    dtype = "u8 -> exact f32/int sums -> f64 response"
    H = W = 2048
    cpu_sample_items = 8

    def __init__(self, args):
        self.batch = args.batch or 64

    def describe(self, world):
        return {"workload": "config[3]: batched Harris (k=0.04, 5x5, thr 1e4) + lambda-minus (1%% of max) over %d synthetic "
                            "2048x2048 u8 images per step per GPU (1024-image job = 16 steps/GPU at 1 GPU)" % self.batch,
                "images_per_step_per_gpu": self.batch, "image": [self.H, self.W], "sharding": "images, no collective",
                "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed" % (self.batch * self.H * self.W * 17 / 1e9)}

    def units_per_step(self):
        return self.batch * self.H * self.W / 1e6  # megapixels

    def setup(self, rank, torch):
        from exvision_featurecraft_b200 import corners, synthetic

        self.torch, self.corners = torch, corners
        # 8 distinct seeded images tiled to the batch (generation cost), distinct per rank
        base = np.stack([synthetic.corner_image(self.H, self.W, seed=rank * 8 + i) for i in range(8)])
        host = np.concatenate([base] * ((self.batch + 7) // 8))[: self.batch]
        self.host_in = torch.from_numpy(host).pin_memory()
        self.dev_in = self.host_in.cuda(non_blocking=True)
        self.resp = torch.empty((self.batch, self.H, self.W), dtype=torch.float64, device="cuda")
        self.eig = torch.empty((self.batch, self.H, self.W), dtype=torch.float64, device="cuda")
        self.ev = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(64)]
        self.ev_i = 0

    def step_device(self, record=False):
        c = self.corners
        ev = self.ev[self.ev_i] if record and self.ev_i < len(self.ev) else None
        if ev:
            ev[0].record()
        res = c.harris(self.dev_in, 5, 0.04, 10000, compact=False, out=self.resp)
        if ev:
            ev[1].record()
        lam = c.lambda_minus(self.dev_in, None, compact=False, out=self.eig)
        if ev:
            ev[2].record()
            self.ev_i += 1
        coords, values, offs = c.compact_mask(res.mask, self.W, res.response)
        lam.mask = c._new_mask(self.batch, self.H, self.W, self.eig.device)
        from exvision_featurecraft_b200 import _lib
        from exvision_featurecraft_b200.device import ptr, stream_ptr

        _lib.call("fc_threshold_mask_rel", ptr(self.eig), self.batch, self.H, self.W, 0.01, ptr(lam.image_max), ptr(lam.mask),
                  stream_ptr())
        lcoords, _, _ = c.compact_mask(lam.mask, self.W, self.eig)
        return coords, values, lcoords

    def e2e_begin(self):
        """Pipelined driver (corners.BatchCornerDetector): pinned input -> H2D -> kernels -> corner lists into pinned
        output, the D2H of step i overlapping the kernels of step i+1."""
        self.det = self.corners.BatchCornerDetector(self.batch, self.H, self.W, max_corners=int(0.06 * self.batch * self.H * self.W))
        for s in range(2):
            self.det.host_in[s].copy_(self.host_in)
        self.pending = None
        self.io_bytes = (0, 0)

    def _account(self, out):
        return self.host_in.numel(), sum(a.nbytes for a in out)

    def step_e2e(self):
        t = self.det.submit(None, prefetch_next=True)
        if self.pending is not None:
            self.io_bytes = self._account(self.det.collect(self.pending))
        self.pending = t
        return self.io_bytes

    def e2e_end(self):
        self.io_bytes = self._account(self.det.collect(self.pending))
        self.pending = None
        return self.io_bytes

    def kernels(self):
        """[(kernel, algorithmic units per launch, ms per launch, bound)]: 1 B in + 8 B out per pixel"""
        n = self.ev_i
        alg = 9.0 * self.batch * self.H * self.W
        if not n:
            return []
        h = float(np.mean([e[0].elapsed_time(e[1]) for e in self.ev[:n]]))
        lm = float(np.mean([e[1].elapsed_time(e[2]) for e in self.ev[:n]]))
        out = [("harris5_f32_kernel", alg, h, "hbm"), ("lambda5_packed_kernel (+ per-image max)", alg, lm, "hbm")]
        fast = self._lambda_fast_ms()
        if fast:
            out.append(("lambda5_packed_kernel<fast> (opt-in response mode, NOT part of `value`: exact integer determinant / "
                        "larger eigenvalue, ~1e-15 relative to the default map)", alg, fast, "hbm"))
        return out

    def _lambda_fast_ms(self):
        """the opt-in fast lambda-minus mode on the same batch, timed separately after the measured steps (CUDA events)"""
        torch, c = self.torch, self.corners
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for i in range(4):
            if i == 1:
                ev[0].record()
            c.lambda_minus(self.dev_in, None, compact=False, out=self.eig, fast=True)
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / 3

    def traffic(self, tr):
        """dram bytes per launch of the dominant kernel, from the committed ncu capture scaled to this launch"""
        return tr["harris5_f32_kernel"]["bytes_per_pixel"] * self.batch * self.H * self.W

    cpu_task = staticmethod(_cpu_harris_task)

    def cpu_items(self, n, step=0):
        return [(self.H, self.W, step * n + s) for s in range(n)]

    def cpu_units(self, n):
        return n * self.H * self.W / 1e6

    def cpu_sample(self, n):
        return "%d full 2048x2048 images per step, Harris + lambda-minus through the oracle (NumPy restatement of the reference)" % n


def workloads():
    from bench_sift import AllPairsWorkload, MatchWorkload, Pipeline4kWorkload, SiftWorkload

    return {"sift": SiftWorkload, "harris": HarrisWorkload, "match": MatchWorkload, "allpairs": AllPairsWorkload,
            "pipeline4k": Pipeline4kWorkload}


# ----------------------------------------------------------------------------------------------
# measurement of one workload
# ----------------------------------------------------------------------------------------------
class Env:
    def __init__(self, torch, dist, rank, local, world):
        self.torch, self.dist, self.rank, self.local, self.world = torch, dist, rank, local, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_entry(kname, alg, kms, bound, pk, traffic=None, sustained=False):
    if bound == "hbm":
        peak, punit, src = pk.get("hbm_gbs", 6650.0), "GB/s", ("measured" if pk else "fallback")
        ach = alg / (kms * 1e-3) / 1e9 if kms else None
    else:
        key = "bf16_tflops_sustained" if sustained else "bf16_tflops"
        peak, punit = pk.get(key, 1400.0 if sustained else 1590.0), "TFLOP/s"
        src = ("measured %s" % ("sustained" if sustained else "burst")) if pk else "fallback"
        ach = alg / (kms * 1e-3) / 1e12 if kms else None
    return {"kernel": kname, "bound": bound, "achieved": ach, "peak": peak, "unit": punit, "frac": (ach / peak) if ach else None,
            "traffic": traffic, "peak_source": src, "kernel_ms": kms, "algorithmic_per_launch": alg}


def measure(wl, env, steps, warmup, sample_clocks):
    """Warm-up, K device-resident steps (CUDA events, max over ranks), K end-to-end steps x 3 (median), roofline."""
    torch = env.torch
    from exvision_featurecraft_b200 import _lib

    wl.setup(env.rank, torch)
    torch.cuda.synchronize()
    for _ in range(warmup):
        wl.step_device()
    env.barrier()
    sampler = ClockSampler(env.local) if (sample_clocks and env.rank == 0) else None
    if sampler:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        wl.step_device(record=True)
    e1.record()
    env.barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    # end to end: pinned host -> device -> hot path -> results on the host (workloads with a pipelined driver overlap the
    # D2H of step i with the kernels of step i+1; the timed region still contains every step's H2D, kernels and D2H,
    # closed by e2e_end() + synchronize).  K steps, host clock around synchronised ends, 3 repetitions, median kept.
    begin, end = getattr(wl, "e2e_begin", None), getattr(wl, "e2e_end", None)
    if begin:
        begin()
    wl.step_e2e()
    if end:
        end()
    env.barrier()
    reps = []
    h2d = d2h = 0
    for _ in range(3):
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            h2d, d2h = wl.step_e2e()
        if end:
            h2d, d2h = end()
        env.barrier()
        reps.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = sorted(reps)[1]
    clocks = sampler.stop() if sampler else None
    ms, e2e_ms = env.max_over_ranks([ms, e2e_ms])
    units = wl.units_per_step() * steps * env.world
    pk = peaks()
    traffic = None
    try:
        traffic = wl.traffic(json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))))
    except Exception:
        traffic = None
    kernels = wl.kernels()
    roofs = [roofline_entry(k, alg, kms, bound, pk, traffic if i == 0 else None) for i, (k, alg, kms, bound) in enumerate(kernels)]
    out = {"metric": wl.metric, "value": units / (ms * 1e-3), "unit": wl.unit, "n_gpus": env.world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype,
           "data": "synthetic", "config": wl.describe(env.world),
           "e2e": {"value": units / (e2e_ms * 1e-3), "unit": wl.unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_ms / steps},
           "gpu_launches": int(launches), "roofline": roofs[0] if roofs else None, "clocks": clocks}
    if len(roofs) > 1:
        out["roofline_more"] = roofs[1:]
    if kernels and kernels[0][3] == "tensor":
        out["roofline_vs_sustained"] = roofline_entry(*kernels[0], pk, None, sustained=True)
    stats = getattr(wl, "stats", None)
    if stats:
        out["stats"] = stats()
    extra = getattr(wl, "extra", None)
    if extra:
        out["stages"] = extra()
    return out


def cpu_baseline(wl):
    try:  # the CPU arm uses every host core again (the GPU arm was pinned next to its GPU)
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except OSError:
        pass
    cores = os.cpu_count() or 1
    n_items = getattr(wl, "cpu_sample_items", None) or cores
    used = min(cores, n_items)
    wall, res = cpu_pool_run(wl.cpu_task, wl.cpu_items(n_items), used)
    return {"value": wl.cpu_units(n_items) / wall, "unit": wl.unit, "cores": used, "kind": "port",
            "sample": "%s; %.1f s wall" % (wl.cpu_sample(n_items), wall)}


def run_reference(args, wl):
    """The reference arm: the oracle on all host cores, same config / metric / unit, W warm-up + K timed steps, each step
    a bounded sample of the workload's step."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    n_items = cores * max(1, args.ref_items_per_core)
    times = []
    for s in range(args.warmup + args.steps):
        wall, _ = cpu_pool_run(wl.cpu_task, wl.cpu_items(n_items, step=s), cores)
        if s >= args.warmup:
            times.append(wall)
    t = float(np.mean(times))
    v = wl.cpu_units(n_items) / t
    line = {"impl": "reference", "metric": wl.metric, "value": v, "unit": wl.unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": wl.describe(max(1, args.gpus)),
            "cpu_baseline": {"value": v, "unit": wl.unit, "cores": cores, "kind": "port", "sample": wl.cpu_sample(n_items)},
            "e2e": {"value": v, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


SECONDARY = (("harris", 5), ("match", 5), ("pipeline4k", 2), ("sift_f64", 3), ("sift_ui", 3))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="one of sift | harris | match | allpairs | pipeline4k (default: sift + secondary)")
    ap.add_argument("--batch", type=int, default=None, help="images per step per GPU")
    ap.add_argument("--ref-items-per-core", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--sift-dtype", default=None, choices=["mixed", "f64"])
    args = ap.parse_args()
    table = workloads()
    wl = table[args.workload or "sift"](args)
    if args.impl == "reference":
        args.steps = args.steps or 1
        args.warmup = 0 if args.warmup is None else args.warmup
        run_reference(args, wl)
        return
    args.steps = args.steps or 10
    args.warmup = 3 if args.warmup is None else max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    from exvision_featurecraft_b200 import sharding

    numa_bound = sharding.bind_to_gpu_numa(local)  # before any pinned staging buffer is allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from exvision_featurecraft_b200 import _lib

    _lib.load()
    env = Env(torch, dist, rank, local, world)
    line = measure(wl, env, args.steps, args.warmup, sample_clocks=True)
    pcie = None
    if rank == 0:
        # context for the end-to-end number: what one pinned 256 MB copy achieves on this box, each direction
        hp = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        dp = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        pcie = {}
        for name, (dst, src) in (("h2d_gbs", (dp, hp)), ("d2h_gbs", (hp, dp))):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            tp = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            pcie[name] = round(3 * hp.numel() / (time.perf_counter() - tp) / 1e9, 1)
        del hp, dp
    line["pcie"] = pcie
    line["numa_bound"] = bool(numa_bound)
    if rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl)
    del wl
    torch.cuda.empty_cache()
    if args.workload is None and not args.no_secondary:
        # the rest of BASELINE.json's metric, each as its own small measurement (same rules: >= 3 warm-up steps, CUDA
        # events, max over ranks); CPU arms on rank 0 only
        line["secondary"] = {}
        todo = list(SECONDARY) + ([("allpairs", 3)] if world > 1 else [])
        for name, k in todo:
            sub_args = argparse.Namespace(**vars(args))
            sub_args.batch = None
            sub_args.sift_dtype = "f64" if name == "sift_f64" else None
            cls = table["sift" if name in ("sift_f64", "sift_ui") else name]
            if name in ("sift_f64", "sift_ui"):
                sub_args.batch = 8
            # another point of the app's parameter ranges (FeatureCraft_UI.py:197-229): 6 levels, radii 5 .. 16
            sub_args.sift_params = {"sigma_base": 2.5, "s_value": 3} if name == "sift_ui" else None
            w2 = cls(sub_args)
            try:
                sub = measure(w2, env, k, 3, sample_clocks=False)
                for drop in ("higher_is_better", "scaling", "vs_baseline", "data", "n_gpus", "clocks"):
                    sub.pop(drop, None)
                if rank == 0 and not args.no_cpu_baseline and name not in ("sift_f64", "sift_ui", "pipeline4k", "allpairs"):
                    sub["cpu_baseline"] = cpu_baseline(w2)
                line["secondary"][name] = sub
            except Exception as exc:  # a secondary workload must not take the headline down with it
                line["secondary"][name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            del w2
            torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
