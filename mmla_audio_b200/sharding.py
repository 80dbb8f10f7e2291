"""Clip sharding across the GPUs of one box (one process per GPU) and the single exchange step
of the path: gather per-clip labels, sum tallies.  No collective touches the
data path — clips are independent — so scaling is weak and communication is a few KB..MB.

Works with any initialised ``torch.distributed`` backend: NCCL on GPUs (the product), gloo on
CPU ranks (the unit tests of this host logic).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Rank r of R owns [floor(r*N/R), floor((r+1)*N/R))."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def gather_labels(labels_local, n_total: int, rank: int, world: int):
    """all_gather of int32 labels sharded with shard_range → full [n_total] tensor on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return labels_local
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.full((mx,), -2, dtype=torch.int32, device=labels_local.device)
    pad[: labels_local.numel()] = labels_local.to(torch.int32)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


def allreduce_counts(counts, world: int):
    """all_reduce(SUM) of the int64 tally vector."""
    import torch.distributed as dist
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


_XCHG_CACHE = {}


def exchange_labels_and_counts(labels_local, counts_local, n_total: int, rank: int, world: int):
    """The path's one exchange step as ONE collective: every rank contributes its int32 labels (padded to the largest
    shard) followed by its int64 tally vector (as int32 pairs); after the all_gather every rank holds all labels and
    sums the tallies locally.  Same results as ``gather_labels`` + ``allreduce_counts`` with one collective launch and
    no per-step allocations instead of two collectives and a dozen small tensor ops (at 8 GPUs the step's exchange
    cost ~0.1 ms of a 1.2 ms step)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return labels_local, counts_local
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    mx, nc = max(sizes), counts_local.numel()
    width = mx + 2 * nc
    key = (labels_local.device, world, width)
    bufs = _XCHG_CACHE.get(key)
    if bufs is None:
        bufs = (torch.empty((width,), dtype=torch.int32, device=labels_local.device),
                torch.empty((world * width,), dtype=torch.int32, device=labels_local.device))
        _XCHG_CACHE[key] = bufs
    send, recv = bufs
    n_local = labels_local.numel()
    send[:n_local].copy_(labels_local)
    if n_local < mx:
        send[n_local:mx].fill_(-2)
    send[mx:].copy_(counts_local.to(torch.int64).contiguous().view(torch.int32))
    dist.all_gather_into_tensor(recv, send)
    table = recv.view(world, width)
    if min(sizes) == mx:
        labels = table[:, :mx].reshape(-1)
    else:
        labels = torch.cat([table[r, :sz] for r, sz in enumerate(sizes)])
    counts = table[:, mx:].contiguous().view(torch.int64).view(world, nc).sum(0)
    return labels, counts


def bind_host_thread_to_gpu(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as closest to ``device_index`` (same NUMA
    node / PCIe root), so pinned staging buffers are first-touched next to the GPU that will read them.
    With one process per GPU the eight uploads otherwise contend for one socket's memory controllers.
    Returns False (and changes nothing) when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False
