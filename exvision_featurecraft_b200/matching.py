"""Brute-force 2-NN descriptor matching with the ratio test (APP/utils/keypoint_descriptor.py:330-349).

``knn2`` is cv2.BFMatcher().knnMatch(query, train, k=2) + ``m.distance < t * n.distance``:
the tcgen05 path (``engine="tensor"``: fp16 GEMM candidate search, exact float32 re-rank, exact re-check
of ambiguous rows) and its CUDA-core twin (``engine="exact"``) return the same arrays.  "Exact" is meant against the
oracle's definition (sequential float32 sum of squares); OpenCV sums with SIMD partial accumulators, so its distances
agree to the last float32 ulp only and a ratio test decided inside that ulp can differ.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_ptr, to_device


def _as_desc(d) -> torch.Tensor:
    if not isinstance(d, torch.Tensor):
        d = np.asarray(d, dtype=np.float32)  # match() casts both sides to float32 (KD:333-334)
        if d.ndim == 2 and d.shape[1] != 128:
            raise ValueError("descriptors must be [n, 128]")
        d = d.reshape(-1, 128) if d.size else np.zeros((0, 128), np.float32)
    t = to_device(d, torch.float32)
    if t.dim() != 2 or t.shape[1] != 128:
        raise ValueError("descriptors must be [n, 128]")
    return t.contiguous()


#: diagnostics of the last tensor-engine call: {"rechecked_rows": int tensor (device), "overflow": ...}; read lazily
LAST_STATS = None


def last_stats() -> dict:
    """Rows the last tensor-engine call re-checked with the exact kernel, and the fp16 overflow flag (syncs)."""
    if LAST_STATS is None:
        return {}
    v = LAST_STATS.cpu().numpy()
    return {"overflow": int(v[1]), "rechecked_rows": int(v[2])}


def knn2(query, train, ratio: float = 0.3, engine: str = "tensor"):
    """-> (knn_idx [nq,2] int32, knn_dist [nq,2] float32, good [nq] bool), all CUDA tensors.

    engine="tensor": good[], and for rows with good=1 knn_idx[:,0] / knn_dist[:,0], are exactly the brute-force
    float32 results (proved per row by the candidate bounds, otherwise the row is re-done exactly); the second
    neighbour of rows that clearly fail the ratio test is the best candidate found.  engine="exact": everything."""
    dev = require_cuda()
    q, t = _as_desc(query), _as_desc(train)
    nq, nt = q.shape[0], t.shape[0]
    if nt < 2:
        # `for m, n in matches` cannot unpack a 1-element row (KD:345)
        raise ValueError("not enough values to unpack (expected 2, got %d)" % nt)
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    dist = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    good = torch.empty(nq, dtype=torch.uint8, device=dev)
    if nq == 0:
        return idx, dist, good.bool()
    if engine == "exact":
        _lib.call("fc_match_knn2_exact", ptr(q), nq, ptr(t), nt, float(ratio), ptr(idx), ptr(dist), ptr(good), stream_ptr())
    elif engine == "tensor":
        nbytes = int(_lib.load().fc_match_workspace_bytes(nq, nt))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        _lib.call("fc_match_knn2", ptr(q), nq, ptr(t), nt, float(ratio), ptr(idx), ptr(dist), ptr(good), ptr(ws), stream_ptr())
        global LAST_STATS
        off = int(_lib.load().fc_match_workspace_stats_offset(nq, nt))
        LAST_STATS = ws[off:off + 16].view(torch.int32)
    else:
        raise ValueError("engine must be 'tensor' or 'exact'")
    return idx, dist, good.bool()


def match_indices(desc_a, desc_b, tuning_distance: float = 0.3, engine: str = "tensor"):
    """The `good` list of match(): (M,2) int64 (queryIdx, trainIdx) and the float32 distances, query order."""
    idx, dist, good = knn2(desc_a, desc_b, tuning_distance, engine)
    q = torch.nonzero(good).flatten()
    pairs = torch.stack([q, idx[q, 0].to(torch.int64)], 1)
    return pairs.cpu().numpy(), dist[q, 0].cpu().numpy()


class BatchMatcher:
    """match() for a stream of equally sized (query, train) descriptor-set pairs that live on the HOST.

    Three CUDA streams like the extraction drivers: pinned descriptors -> device (copy-in), the matching kernels
    (compute = the caller's current stream), results -> pinned host memory (copy-out).  `submit` returns as soon as
    the work of a pair is queued; `collect` waits for that pair's copy-out only, so the 2 x nq x 512 bytes of the
    next pair cross PCIe while the current pair is on the tensor cores."""

    def __init__(self, n_query: int, n_train: int, ratio: float = 0.3, depth: int = 2, engine: str = "tensor"):
        dev = require_cuda()
        self.nq, self.nt, self.ratio, self.depth, self.engine = int(n_query), int(n_train), float(ratio), depth, engine
        self.copy_in = torch.cuda.Stream(device=dev)
        self.copy_out = torch.cuda.Stream(device=dev)
        self.host_q = [torch.empty((self.nq, 128), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.host_t = [torch.empty((self.nt, 128), dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.dev_q = [torch.empty((self.nq, 128), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.dev_t = [torch.empty((self.nt, 128), dtype=torch.float32, device=dev) for _ in range(depth)]
        self.host_good = [torch.empty(self.nq, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self.host_idx = [torch.empty(self.nq, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self.host_dist = [torch.empty(self.nq, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.consumed = [None] * depth  # event after the kernels that read dev_q/dev_t[slot]
        self.done = [None] * depth      # event after the copy-out of a slot
        self.keep = [None] * depth      # device results kept alive until their copy-out has run
        self.slots = [None] * depth     # ticket that occupies a slot until collect() releases it
        self.n_submitted = 0

    def submit(self, query=None, train=None) -> int:
        """Queue one pair.  query / train: [n,128] float32 arrays or tensors on the host; None = the pinned staging
        buffers host_q / host_t of this ticket's slot were filled by the caller."""
        ticket = self.n_submitted
        slot = ticket % self.depth
        if self.slots[slot] is not None:
            # the staging buffers and result buffers of this slot still belong to an uncollected pair: overwriting
            # them would corrupt its pending H2D copy and make collect(old ticket) return this pair's matches
            raise RuntimeError("collect() ticket %d (submitted %d calls ago) before reusing its buffers" % (self.slots[slot], self.depth))
        self.slots[slot] = ticket
        self.n_submitted += 1
        if query is not None:
            self.host_q[slot].copy_(query if isinstance(query, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(query, np.float32)))
        if train is not None:
            self.host_t[slot].copy_(train if isinstance(train, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(train, np.float32)))
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_in):
            if self.consumed[slot] is not None:
                self.copy_in.wait_event(self.consumed[slot])
            self.dev_q[slot].copy_(self.host_q[slot], non_blocking=True)
            self.dev_t[slot].copy_(self.host_t[slot], non_blocking=True)
            staged = torch.cuda.Event()
            staged.record()
        compute.wait_event(staged)
        idx, dist, good = knn2(self.dev_q[slot], self.dev_t[slot], self.ratio, self.engine)
        best_idx, best_dist, good_u8 = idx[:, 0].contiguous(), dist[:, 0].contiguous(), good.to(torch.uint8)
        ready = torch.cuda.Event()
        ready.record()
        self.consumed[slot] = ready
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(ready)
            self.host_good[slot].copy_(good_u8, non_blocking=True)
            self.host_idx[slot].copy_(best_idx, non_blocking=True)
            self.host_dist[slot].copy_(best_dist, non_blocking=True)
            out = torch.cuda.Event()
            out.record()
        self.done[slot] = out
        self.keep[slot] = (best_idx, best_dist, good_u8)
        return ticket

    def collect(self, ticket: int):
        """-> (pairs [M,2] int64 (queryIdx, trainIdx), distances [M] float32): the `good` list of match(), query order."""
        slot = ticket % self.depth
        if self.slots[slot] != ticket:
            raise RuntimeError("ticket %d is not pending (already collected, or never submitted)" % ticket)
        self.done[slot].synchronize()
        q = np.flatnonzero(self.host_good[slot].numpy())
        pairs = np.stack([q, self.host_idx[slot].numpy()[q].astype(np.int64)], 1)
        dist = self.host_dist[slot].numpy()[q].copy()
        self.slots[slot] = None
        self.keep[slot] = None
        return pairs, dist

    def bytes_per_pair(self):
        return (self.nq + self.nt) * 512, self.nq * 9
