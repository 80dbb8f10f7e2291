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


# ---------------------------------------------------------------------------------------------
# long single recording: 256-frame chunk ranges + read-only halo (second half)
# ---------------------------------------------------------------------------------------------
def _slice_features(rec, rank, world):
    """What a rank computes, with the CPU oracle standing in for the device MFCC: rows of its own chunks."""
    from oracle import psf
    from mmla_audio_b200.sharding import session_slice
    sl = session_slice(len(rec), rank, world)
    n_own = sl["chunk_hi"] - sl["chunk_lo"]
    if n_own <= 0:
        return np.zeros((0, 256, 39)), sl
    feat = psf.mfcc39(rec[sl["sample_lo"]:sl["sample_hi"]])
    rows = feat[sl["skip_rows"]:sl["skip_rows"] + n_own * 256]
    if sl["chunk_hi"] == sl["n_chunks_total"]:
        rows = rows[: sl["n_frames_total"] - sl["chunk_lo"] * 256]
    out = np.zeros((n_own * 256, 39))
    out[: len(rows)] = rows
    return out.reshape(n_own, 256, 39), sl


@pytest.mark.parametrize("n_samples", [40960 * 7 + 1234, 40960 * 3, 50000, 300])
def test_session_slices_reproduce_whole_file_features(n_samples):
    """Concatenating every rank's chunks gives the whole-file chunks: the 5-frame / 4-frame halo covers the pre-emphasis
    sample and the delta-delta context across chunk borders.  (Oracle arithmetic, float64: equal up to BLAS blocking of
    the filterbank product, 1e-10 relative — a missing halo would show up at 1e-2; the device path is checked for exact
    equality in tests/test_pipeline_gpu.py.)"""
    from oracle import psf, synth
    rec = synth.synth_clips(900, -(-n_samples // 40960), 40960).reshape(-1)[:n_samples]
    whole = psf.chunked_features(rec)
    for world in (1, 2, 3, 8):
        parts, covered = [], []
        for r in range(world):
            feats, sl = _slice_features(rec, r, world)
            parts.append(feats)
            covered.append((sl["chunk_lo"], sl["chunk_hi"]))
            assert 0 <= sl["sample_lo"] <= sl["sample_hi"] <= n_samples
            if r > 0 and sl["chunk_hi"] > sl["chunk_lo"]:
                assert sl["chunk_lo"] * 256 * 160 - sl["sample_lo"] in (0, 5 * 160)      # 800-sample halo before
        assert covered[0][0] == 0 and covered[-1][1] == whole.shape[0]
        np.testing.assert_allclose(np.concatenate(parts), whole, rtol=1e-10, atol=1e-10)


def _session_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import nets as onets, synth
    from mmla_audio_b200 import weights as W
    from mmla_audio_b200.sharding import exchange_labels_and_counts
    rec = synth.synth_clips(950, 5, 40960).reshape(-1)[: 5 * 40960 - 999]
    spec = W.speaker_spec(4, "sigmoid")
    w = W.synthetic_weights(spec, 21)
    feats, sl = _slice_features(rec, rank, world)
    lab = torch.from_numpy(onets.speaker_forward(feats.astype(np.float32), w, spec).argmax(1).astype(np.int32)) \
        if len(feats) else torch.empty((0,), dtype=torch.int32)
    cnt = torch.bincount(lab.long(), minlength=5)
    all_lab, all_cnt = exchange_labels_and_counts(lab, cnt, sl["n_chunks_total"], rank, world)
    q.put((rank, all_lab.tolist(), all_cnt.tolist()))
    dist.destroy_process_group()


def test_long_session_world2_labels_equal_single_rank():
    """world_size 2 over gloo: each rank labels its chunk range of ONE recording; the gathered labels and tallies equal
    the unsharded session's."""
    from oracle import nets as onets, psf, synth
    from mmla_audio_b200 import weights as W
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_session_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rec = synth.synth_clips(950, 5, 40960).reshape(-1)[: 5 * 40960 - 999]
    spec = W.speaker_spec(4, "sigmoid")
    w = W.synthetic_weights(spec, 21)
    want = onets.speaker_forward(psf.chunked_features(rec).astype(np.float32), w, spec).argmax(1)
    for _rank, labels, counts in res:
        assert labels == want.tolist()
        assert counts[:4] == np.bincount(want, minlength=4).tolist()
