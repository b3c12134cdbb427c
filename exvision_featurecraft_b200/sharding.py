"""Image-level sharding across the GPUs of one box: one process per GPU, torch.distributed
for the plumbing.

Independent units are images (Harris / lambda-minus, SIFT extraction) and ordered image pairs (matching).
Extraction needs NO data-path collective: rank r owns a contiguous block of images and only the per-image
result counts are gathered.  All-pairs matching has ONE real exchange step: every rank keeps its block of
descriptor sets resident and the other blocks visit in a (world-1)-step ring of send/recv -- NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin this process to the CPUs next to `cuda:device_index` (NVML's ideal affinity) before it allocates pinned
    staging memory: pages are placed on the node of the allocating thread, and on a two-socket box a staging buffer
    on the far socket sends every H2D / D2H byte across the socket interconnect.  With 8 ranks that link, not PCIe,
    bounded the end-to-end throughput.  Returns False (and changes nothing) when NVML cannot tell."""
    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


def shard_bounds(n_items: int, world: int):
    """Contiguous, balanced blocks: the first n % world ranks get one extra item.  -> [(lo, hi)] * world."""
    base, extra = divmod(n_items, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def my_shard(n_items: int, rank: int | None = None, world: int | None = None):
    r, w = world_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    return shard_bounds(n_items, world)[rank]


def gather_counts(local_counts, device=None):
    """Per-image result counts of every rank, in global image order (ranks own contiguous blocks).
    Works before init_process_group (single process) and with any backend."""
    rank, world = world_info()
    t = torch.as_tensor(list(local_counts), dtype=torch.int64)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return t.tolist()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n = torch.tensor([t.numel()], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros(m, dtype=torch.int64, device=device)
    pad[: t.numel()] = t.to(device)
    bufs = [torch.zeros(m, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = []
    for s, b in zip(sizes, bufs):
        out.extend(b[: int(s.item())].tolist())
    return out


def ring_schedule(world: int):
    """All-pairs over blocks: at step s rank r matches its own (query) block against block (r + s) % world.
    Every ordered (query block, train block) pair appears exactly once."""
    return [[(r, (r + s) % world) for r in range(world)] for s in range(world)]


def ring_shift_start(block: torch.Tensor, max_rows: int):
    """Post one ring step: send `block` ([n, d], n <= max_rows) to rank-1 and receive the block of rank+1 (after s
    steps a rank holds the block of rank (r + s) % world).  Returns a handle for ring_shift_finish; the transfers run
    on the communication backend's own stream, so kernels queued in between overlap them."""
    rank, world = world_info()
    if world == 1:
        return ("local", block)
    d = block.shape[1]
    send = torch.zeros((max_rows + 1, d), dtype=block.dtype, device=block.device)
    send[0, 0] = block.shape[0]  # row count travels in the header row
    send[1:1 + block.shape[0]] = block
    recv = torch.empty_like(send)
    dst, src = (rank - 1) % world, (rank + 1) % world
    ops = [dist.P2POp(dist.isend, send, dst), dist.P2POp(dist.irecv, recv, src)]
    return ("posted", dist.batch_isend_irecv(ops), send, recv)


def ring_shift_finish(handle):
    if handle[0] == "local":
        return handle[1]
    _, reqs, _send, recv = handle
    for req in reqs:
        req.wait()
    n = int(recv[0, 0].item())
    return recv[1:1 + n].contiguous()


def ring_shift(block: torch.Tensor, max_rows: int):
    """One blocking ring step (see ring_shift_start)."""
    return ring_shift_finish(ring_shift_start(block, max_rows))


def all_pairs_match(local_desc: torch.Tensor, match_fn, max_rows: int):
    """Match the local descriptor block (queries) against the block of every rank (train), ring order.  The transfer
    of the NEXT block is posted before the current block is matched, so NVLink traffic hides under the GEMM.
    match_fn(query, train) -> any result; returns [(train_rank, result)] with world entries."""
    rank, world = world_info()
    results = []
    visiting = local_desc
    for s in range(world):
        handle = ring_shift_start(visiting, max_rows) if s + 1 < world else None
        results.append(((rank + s) % world, match_fn(local_desc, visiting)))
        if handle is not None:
            visiting = ring_shift_finish(handle)
    return results
