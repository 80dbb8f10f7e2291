"""GPU parity for the end-to-end pipelines: BASELINE configs 1, 4 and 5 (scaled so the CPU
oracle finishes in seconds) — labels and tallies must equal the oracle's."""
from datetime import datetime

import numpy as np
import pytest

from oracle import librosa_mel as lm, nets as onets, psf, synth, tally as otally

pytestmark = pytest.mark.gpu


def _clear(prob, margin=1e-3):
    s = np.sort(prob, axis=1)
    return (s[:, -1] - s[:, -2]) > margin


def test_config1_overlap_single_clip(cuda):
    """One 2.5 s clip: features (first 24000 samples) + classifier forward, label vs oracle."""
    from mmla_audio_b200 import models, weights as W
    from mmla_audio_b200.pipeline import OverlapPipeline
    w = W.synthetic_weights(W.OVERLAP, 1234)
    pipe = OverlapPipeline(models.Model(W.OVERLAP, w, precision="fp32"))
    sig = synth.synth_clips(77, 1, 40000)
    labels, prob = pipe.run_device(cuda.from_numpy(sig).cuda())
    ref = onets.overlap_forward(lm.classifier_input(sig[0])[None], w, W.OVERLAP)
    assert np.abs(prob.cpu().numpy() - ref).max() <= 1e-3        # image may differ by 1 LSB on <1% pixels
    if _clear(ref)[0]:
        assert labels.cpu().numpy()[0] == ref.argmax(1)[0]
    short = synth.synth_clips(78, 2, 3999)                        # < 4000 samples => 'silent'
    l2, _ = pipe.run_device(cuda.from_numpy(short).cuda())
    assert l2.cpu().tolist() == [-1, -1]


def test_config4_long_session_overlap(cuda, tmp_path):
    """A 60 s recording cut into 1.5 s windows (segmentation index math), labels + tallies."""
    from mmla_audio_b200 import models, tally, weights as W
    from mmla_audio_b200.pipeline import OverlapPipeline, segmentation_windows
    w = W.synthetic_weights(W.OVERLAP, 1234)
    pipe = OverlapPipeline(models.Model(W.OVERLAP, w, precision="fp32"))
    rec = synth.synth_clips(500, 40, 24000).reshape(-1)[: 40 * 24000 - 5000]   # ragged tail is dropped
    n = segmentation_windows(len(rec), 24000, 24000)
    assert n == otally.num_windows(len(rec), 24000, 24000) == 39
    t0 = datetime(2021, 6, 1, 12, 0, 0, 654321)
    log = tmp_path / "experiment" / "logs" / "session.txt"
    labels, (counts, secs, total) = pipe.run_session(rec, t0=t0, log_path=str(log))
    assert labels.numel() == n
    x = np.stack([lm.classifier_input(rec[i * 24000:(i + 1) * 24000]) for i in range(n)])
    ref = onets.overlap_forward(x, w, W.OVERLAP)
    got = labels.cpu().numpy()
    clear = _clear(ref, 5e-3)
    assert (got[clear] == ref.argmax(1)[clear]).all() and (got == ref.argmax(1)).mean() >= 0.9
    names = [tally.OVERLAP_DEGREE_DICT[str(int(l))] for l in got]
    lines = otally.log_rows(names, t0, 1.5, "overlapped degree", add_before_first=False)
    rc, rs, rt = otally.tally_from_log(lines, list(tally.OVERLAP_DEGREE_DICT.values()))
    assert (counts, secs, total) == (rc, rs, rt)                  # integer tallies: bit exact
    # the log file on disk is the one the reference would have written, and the file-based
    # visualization() counts it to the same tallies (silent is pre-seeded there, :11)
    assert log.read_text().splitlines() == lines
    from mmla_audio_b200 import overlap_degree_distribution as odd
    res = odd.visualization(str(log.parent))["session.txt"]
    assert dict(zip(res["labels"], res["counts"])) == {**rc, "silent": 0}
    assert res["total_seconds"] == rt


def test_config4_long_session_speaker(cuda, tmp_path):
    """Whole-file MFCC-39 -> 256-frame chunks -> one predict -> rows every 2.56 s."""
    from mmla_audio_b200 import models, weights as W
    from mmla_audio_b200.pipeline import SpeakerPipeline
    spec = W.speaker_spec(10, "sigmoid")
    w = W.synthetic_weights(spec, 4321)
    pipe = SpeakerPipeline(models.Model(spec, w, precision="fp32"))
    rec = synth.synth_clips(700, 30, 40960).reshape(-1)           # 76.8 s -> 30 chunks
    names = {i: f"spk{i}" for i in range(10)}
    t0 = datetime(2022, 3, 3, 8, 30, 0, 111111)
    log = tmp_path / "experiment" / "logs" / "session.txt"
    labels, (counts, secs, total) = pipe.run_session(rec, names, t0=t0, silent_index=(2, 7), log_path=str(log))
    chunks = psf.chunked_features(rec).astype(np.float32)
    ref = onets.speaker_forward(chunks, w, spec)
    assert labels.numel() == ref.shape[0] == 30
    got = labels.cpu().numpy()
    assert got[2] == -1 and got[7] == -1
    keep = np.ones(30, bool)
    keep[[2, 7]] = False
    clear = _clear(ref) & keep
    assert (got[clear] == ref.argmax(1)[clear]).all()
    lab_names = [names.get(int(l), "silent") for l in got]
    lines = otally.log_rows(lab_names, t0, 2.56, "speaker", add_before_first=True)
    rc, rs, rt = otally.tally_from_log(lines)
    assert (counts, secs, total) == (rc, rs, rt)
    assert log.read_text().splitlines() == lines
    from mmla_audio_b200 import speaker_time_distribution as std
    res = std.visualization(str(log.parent))["session.txt"]
    assert dict(zip(res["labels"], res["counts"])) == rc and dict(zip(res["labels"], res["seconds"])) == rs


def test_config5_enrollment_features(cuda):
    """make_feature_experiment: per-speaker corpora -> [M,256,39] chunks, one-hot labels, id dict."""
    from mmla_audio_b200 import speaker_identification as si
    corpora = [(f"spk{i}", synth.synth_clips(900 + 10 * i, 4, 40000).reshape(-1)) for i in range(3)]   # 10 s each
    x, y, ids = si.make_feature_experiment(corpora)
    ref = np.concatenate([psf.chunked_features(sig) for _, sig in corpora])
    assert x.shape == ref.shape == (12, 256, 39) and y.shape == (12, 3)
    tol = 1e-4 * np.abs(ref).max()
    assert np.all(np.abs(x - ref) <= 1e-4 * np.abs(ref) + tol)
    assert ids == {"0": "spk0", "1": "spk1", "2": "spk2"}
    assert (y.argmax(1) == np.repeat(np.arange(3), 4)).all()


def test_config2_speaker_batch_labels_bit_exact_on_clear_margins(cuda):
    """256 clips of 1.5 s, 10 speakers: labels vs oracle (fp32 path) and tf32 agreement."""
    from mmla_audio_b200 import models, tally, weights as W
    from mmla_audio_b200.pipeline import SpeakerPipeline
    spec = W.speaker_spec(10, "sigmoid")
    w = W.synthetic_weights(spec, 4321)
    pcm = synth.synth_clips(0, 256, 24000)
    x = np.concatenate([psf.input_feature_gen(pcm[i]) for i in range(256)]).astype(np.float32)
    ref = onets.speaker_forward(x, w, spec)
    dev = cuda.from_numpy(pcm).cuda()
    for prec, min_agree in (("fp32", 0.99), ("tf32", 0.95)):
        pipe = SpeakerPipeline(models.Model(spec, w, precision=prec))
        labels, prob = pipe.run_device(dev)
        got = labels.cpu().numpy()
        assert (got == ref.argmax(1)).mean() >= min_agree, prec
        if prec == "fp32":
            clear = _clear(ref)
            assert (got[clear] == ref.argmax(1)[clear]).all()
        counts = tally.device_counts(labels, 10).cpu().numpy()
        np.testing.assert_array_equal(counts[:10], np.bincount(got, minlength=10))


def test_host_pipeline_sync_and_pipelined(cuda):
    """run_host / submit_host (pinned host PCM in, labels + tallies out; uploads overlap compute, several
    batches in flight) give exactly what the device-resident path gives, batch after batch."""
    import torch
    from mmla_audio_b200 import models, tally, weights as W
    from mmla_audio_b200 import synth as dsynth
    from mmla_audio_b200.pipeline import SpeakerPipeline
    spec = W.speaker_spec(10, "sigmoid")
    pipe = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision="tf32"))
    batches = [dsynth.synth_clips(1000 + 300 * i, 300, 24000) for i in range(4)]
    want = []
    for b in batches:
        labels, _ = pipe.run_device(b)
        want.append((labels.cpu().numpy(), tally.device_counts(labels, 10).cpu().numpy()))
    hosts = [b.cpu().pin_memory() for b in batches]
    lab, cnt = pipe.run_host(hosts[0], 10, n_chunks=3)
    assert np.array_equal(lab, want[0][0]) and np.array_equal(cnt, want[0][1]) and cnt.sum() == 300
    pending = [pipe.submit_host(h, 10, n_chunks=2, depth=3) for h in hosts[:3]]      # three batches in flight
    pending.append(pipe.submit_host(hosts[3], 10, n_chunks=2, depth=3))              # reuses the first slot: waits for it
    for i in (1, 2, 3):
        lab, cnt = pending[i].result()
        assert np.array_equal(lab, want[i][0]) and np.array_equal(cnt, want[i][1])


def test_config4_full_size_8h_overlap_session_properties(cuda):
    """BASELINE configs[3] at full size: an 8 h recording (460.8 M samples) -> 19 200 windows of 1.5 s through the
    tensor-core overlap pipeline.  Too large for the oracle, so the checks are size-independent properties:
    the window count of the reference's index math, tallies that sum to the window count, a second pass that gives
    identical labels (determinism), the same labels when the session is processed in two halves with another chunk
    size (batch independence across micro-batch boundaries), and agreement with the oracle-checked small-session path
    on the first windows."""
    from mmla_audio_b200 import models, synth as dsynth, weights as W
    from mmla_audio_b200.pipeline import OverlapPipeline, segmentation_windows
    torch = cuda
    pipe = OverlapPipeline(models.Model(W.OVERLAP, W.synthetic_weights(W.OVERLAP, 1234), precision="tf32"))
    n_win, win = 19200, 24000
    rec = dsynth.synth_clips(7000, n_win, win).reshape(-1)          # int16 CUDA, 8 h at 16 kHz
    assert rec.numel() == 8 * 3600 * 16000
    assert segmentation_windows(rec.numel(), win, win) == otally.num_windows(rec.numel(), win, win) == n_win
    t0 = datetime(2021, 6, 1, 9, 0, 0, 123456)
    labels, (counts, secs, total) = pipe.run_session(rec, t0=t0)
    assert labels.numel() == n_win and sum(counts.values()) == n_win
    assert abs(total - (n_win - 1) * 1.5) <= 1                      # t_last - t_first in whole seconds
    labels2, (counts2, _, _) = pipe.run_session(rec, t0=t0)
    assert torch.equal(labels, labels2) and counts == counts2
    half = (n_win // 2) * win
    la, _ = pipe.run_session(rec[:half], t0=t0, chunk=1000)
    lb, _ = pipe.run_session(rec[half:], t0=t0, chunk=777)
    assert torch.equal(labels, torch.cat([la, lb]))
    small, _ = pipe.run_session(rec[: 40 * win], t0=t0)
    assert torch.equal(labels[:40], small)


def test_long_session_halo_sharding_equals_single_gpu(cuda):
    """SURVEY §8e, long single recording: the chunks each of R ranks computes from its sample range + halo are
    bit-identical to the same chunks of the single-GPU whole-file pass (emulated ranks on one device: the sharding is
    pure index math, nothing is exchanged on the data path)."""
    from mmla_audio_b200 import models, weights as W
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.pipeline import SpeakerPipeline
    torch = cuda
    spec = W.speaker_spec(10, "sigmoid")
    pipe = SpeakerPipeline(models.Model(spec, W.synthetic_weights(spec, 4321), precision="tf32"))
    for n_samples in (40960 * 9 + 777, 40960 * 4, 46000):
        rec = synth.synth_clips(1200, -(-n_samples // 40960), 40960).reshape(-1)[:n_samples]
        rec_dev = torch.from_numpy(rec).cuda()
        whole = si.whole_file_chunks(rec_dev)
        for world in (2, 3, 8):
            parts = [pipe.session_chunks_sharded(rec_dev, r, world)[0] for r in range(world)]
            got = torch.cat(parts)
            assert got.shape == whole.shape
            assert torch.equal(got, whole), (n_samples, world, (got - whole).abs().max().item())
