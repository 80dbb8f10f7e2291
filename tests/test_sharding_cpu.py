"""world_size-2 gloo test of the multi-GPU host logic (clip sharding, label gather, tally
all-reduce) on CPU ranks — the N>1 path without GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmla_audio_b200.sharding import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 4096, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, n_classes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mmla_audio_b200.sharding import allreduce_counts, gather_labels, shard_range as sr
    full = torch.from_numpy(np.random.default_rng(0).integers(-1, n_classes, n_total).astype(np.int32))
    lo, hi = sr(n_total, rank, world)
    local = full[lo:hi].clone()
    gathered = gather_labels(local, n_total, rank, world)
    counts = torch.bincount(torch.where(local < 0, torch.tensor(n_classes), local.long()), minlength=n_classes + 1)
    counts = allreduce_counts(counts, world)
    ref = torch.bincount(torch.where(full < 0, torch.tensor(n_classes), full.long()), minlength=n_classes + 1)
    # the same exchange as ONE collective (what bench.py and the pipelines use), twice to exercise the cached buffers
    from mmla_audio_b200.sharding import exchange_labels_and_counts
    ok_x = True
    for _ in range(2):
        local_counts = torch.bincount(torch.where(local < 0, torch.tensor(n_classes), local.long()), minlength=n_classes + 1)
        g2, c2 = exchange_labels_and_counts(local, local_counts, n_total, rank, world)
        ok_x = ok_x and bool(torch.equal(g2, full)) and bool(torch.equal(c2, ref)) and c2.dtype == torch.int64
    q.put((rank, bool(torch.equal(gathered, full)) and ok_x, bool(torch.equal(counts, ref))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1001, 4096])        # ragged and equal shards
def test_gather_and_allreduce_world2(n_total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, 10, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] for r in res)
