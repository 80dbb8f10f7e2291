"""GPU tests of the offline session drivers: `post_anlysing()` (overlap) and `post_analysing()`
(speaker) over a directory tree of WAV files laid out as the reference expects, against the same steps composed from
the oracle (segmentation index math -> features -> classifier -> arg-max -> log rows -> visualization tallies)."""
import os
from datetime import datetime

import numpy as np
import pytest

from oracle import librosa_mel as lm, nets as onets, psf, synth, tally as otally, webrtc_vad as ovad

pytestmark = pytest.mark.gpu


def _clear(p, m):
    s = np.sort(p, axis=1)
    return (s[:, -1] - s[:, -2]) > m


def test_overlap_post_anlysing_over_wav_directories(cuda, tmp_path):
    from mmla_audio_b200 import models, overlap_degree_distribution as odd, weights as W
    from mmla_audio_b200 import overlap_detection_post_processing as pp
    from mmla_audio_b200.audio_io import read_wav_int16, write_wav_int16
    from mmla_audio_b200.offline_common import apply_dbfs_gain
    root = tmp_path
    whole = root / "experiment" / "recordings" / "post-time" / "whole"
    whole.mkdir(parents=True)
    rec = (synth.synth_clips(3000, 9, 24000).reshape(-1)[: 9 * 24000 - 7000] // 4).astype(np.int16)
    write_wav_int16(str(whole / "zoom_meeting.wav"), rec)
    write_wav_int16(str(whole / "notes.wav"), rec)                       # neither 'zoom' nor 'audio': skipped like the reference
    w = models.save_synthetic_model(str(root / "timit" / "models" / "timit2.0"), W.OVERLAP, seed=1234)
    t0 = datetime(2021, 6, 1, 9, 30, 0, 250000)
    logs = pp.post_anlysing(str(root), precision="fp32", t0=t0, write_png=True)
    assert list(logs) == [str(root / "experiment" / "logs" / "zoom_meeting.txt")]
    # the standardised file: gain to 0 dBFS with audioop.mul semantics
    _, std_sig = read_wav_int16(str(root / "experiment/recordings/post-time/standardized/zoom_meeting.wav"))
    np.testing.assert_array_equal(std_sig, apply_dbfs_gain(rec, 0))
    n = otally.num_windows(len(std_sig), 24000, 24000)
    assert n == 8 and len(os.listdir(root / "experiment/recordings/post-time/segments/zoom_meeting")) == n
    assert len(os.listdir(root / "experiment/recordings/post-time/features/zoom_meeting")) == n
    x = np.stack([lm.classifier_input(std_sig[i * 24000:(i + 1) * 24000]) for i in range(n)])
    ref = onets.overlap_forward(x, w, W.OVERLAP)
    rows = logs[str(root / "experiment" / "logs" / "zoom_meeting.txt")]
    got = [r.split("\t")[1] for r in rows[1:]]
    want = [pp.overlap_degree_dict[str(int(k))] for k in ref.argmax(1)]
    clear = _clear(ref, 5e-3)
    assert all(g == wv for g, wv, c in zip(got, want, clear) if c)
    assert rows == otally.log_rows(got, t0, 1.5, "overlapped degree", add_before_first=False)
    odd.Root_Dir = str(root)
    res = odd.visualization()["zoom_meeting.txt"]
    counts, secs, total = otally.tally_from_log(rows, ["non-overlapped", "overlapped", "silent"])
    assert dict(zip(res["labels"], res["counts"])) == counts and dict(zip(res["labels"], res["seconds"])) == secs


def test_speaker_post_analysing_over_wav_directories(cuda, tmp_path):
    from mmla_audio_b200 import models, speaker_time_distribution as std, weights as W
    from mmla_audio_b200 import speaker_identification_post_processing as pp
    from mmla_audio_b200.audio_io import write_wav_int16
    root = tmp_path
    for sub in ("experiment/corpus", "experiment/recordings/post-time/standardized", "experiment/recordings/post-time/segments"):
        (root / sub).mkdir(parents=True)
    for who in ("ann", "bob", "cy"):
        write_wav_int16(str(root / "experiment/corpus" / (who + ".wav")), np.zeros(16, np.int16))
    names = sorted(os.listdir(root / "experiment/corpus"))            # the reference uses os.listdir order; fix it for the test
    rec = synth.synth_clips(4000, 6, 40960)
    rec[2] = (np.random.default_rng(1).standard_normal(40960) * 12).astype(np.int16)     # a silent segment
    rec[4, 2500:] = 0                                                                      # too little voice
    rec = rec.reshape(-1)[: 6 * 40960 - 3000]
    write_wav_int16(str(root / "experiment/recordings/post-time/standardized/meeting.wav"), rec)
    written = pp.segmentation(str(root / "experiment/recordings/post-time/standardized"),
                              str(root / "experiment/recordings/post-time/segments"), 2.56, 2.56)
    assert len(written) == 5
    spec = W.speaker_spec(3, "sigmoid")
    w = W.synthetic_weights(spec, 77)
    model = models.Model(spec, w, precision="fp32")
    t0 = datetime(2021, 6, 1, 9, 30, 0, 250000)
    import unittest.mock as mock
    with mock.patch("os.listdir", side_effect=lambda p, _l=os.listdir: sorted(_l(p))):
        logs = pp.post_analysing(str(root), model=model, t0=t0)
    rows = logs[str(root / "experiment/logs/meeting.txt")]
    # oracle: VAD chained over the segments in order -> silent_index; whole-file chunks -> classifier -> rows
    vad = ovad.Vad(3)
    silent = [i for i in range(5) if len(ovad.remove_silence(rec[i * 40960:(i + 1) * 40960], vad)[0]) < 4000]
    assert silent == [2, 4]
    chunks = psf.chunked_features(rec).astype(np.float32)
    assert chunks.shape[0] == 6
    prob = onets.speaker_forward(chunks, w, spec)
    got = [r.split("\t")[1] for r in rows[1:]]
    assert len(got) == 6
    clear = _clear(prob, 1e-3)
    for i in range(6):
        if i in silent:
            assert got[i] == "silent"
        elif clear[i]:
            assert got[i] == names[int(prob[i].argmax())][:-4]
    assert rows == otally.log_rows(got, t0, 2.56, "speaker", add_before_first=True)
    std.Root_Dir = str(root)
    res = std.visualization()["meeting.txt"]
    counts, secs, _ = otally.tally_from_log(rows)
    assert dict(zip(res["labels"], res["counts"])) == counts and dict(zip(res["labels"], res["seconds"])) == secs
