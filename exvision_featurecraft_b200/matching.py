"""Brute-force 2-NN descriptor matching with the ratio test (APP/utils/keypoint_descriptor.py:330-349).

``knn2`` is cv2.BFMatcher().knnMatch(query, train, k=2) + ``m.distance < t * n.distance``:
the tcgen05 path (``engine="tensor"``: fp16 GEMM candidate search, exact float32 re-rank, exact re-check
of ambiguous rows) and its CUDA-core twin (``engine="exact"``) return the same arrays.
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
