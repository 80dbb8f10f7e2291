"""CPU tests against golden vectors produced by RUNNING THE REFERENCE'S OWN CODE
(tests/golden/make_reference_vectors.py: the reference modules imported unmodified from /root/reference with their
un-installable third-party imports stubbed; fixtures committed as tests/golden/reference_vectors.{npz,json}).

What these vectors pin, and what they cannot:
  * fully pinned (no third-party arithmetic on the path): `delta`, `binarizer`, `frame_generator`, `vad_collector`,
    `segmentation`, both `visualization()` tallies, `normalize_matrix`;
  * composition pinned (the library call inside was the oracle's restatement): `input_feature_gen`,
    `make_feature_experiment`, `generate_mels` / `generate_zcr` / `generate_zcr_image`;
  * still unpinned: the arithmetic of python_speech_features / librosa / TensorFlow / webrtcvad themselves.
This file checks the ORACLE and the host-side product logic (no GPU); test_reference_golden_gpu.py checks the CUDA path.
"""
import json
import os
import zlib
from datetime import datetime

import numpy as np
import pytest

from oracle import librosa_mel as olm, psf, synth, tally as otally, webrtc_vad as ovad

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLDEN, "reference_vectors.npz"))


@pytest.fixture(scope="module")
def meta():
    return json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))


def test_oracle_delta_equals_reference_delta(vec):
    for name in ("delta_T37", "delta_T3", "delta_T1"):
        d1 = psf.delta(vec[name + "_in"], 2)
        np.testing.assert_allclose(d1, vec[name + "_out"], rtol=0, atol=1e-12)
        np.testing.assert_allclose(psf.delta(d1, 2), vec[name + "_out2"], rtol=0, atol=1e-12)


def test_binarizer_equals_reference(vec, meta):
    from mmla_audio_b200 import speaker_identification as si
    np.testing.assert_array_equal(si.binarizer(meta["binarizer_in"], dim=4), vec["binarizer_out"])


def test_oracle_input_feature_gen_equals_reference_composition(vec, meta):
    for k in ("ifg_1p5s", "ifg_2p56s", "ifg_2p9s_truncated", "ifg_4000"):
        sig = synth.synth_clips(meta[k]["synth_clip"], 1, meta[k]["samples"])[0]
        np.testing.assert_array_equal(psf.input_feature_gen(sig), vec[k])
    assert psf.input_feature_gen(synth.synth_clips(15, 1, 3999)[0]) == meta["ifg_3999_samples"] == "silent"


def test_oracle_chunked_features_equal_reference_make_feature_experiment(vec, meta):
    chunks, labels = [], []
    for f in meta["mfe_files"]:
        c = psf.chunked_features(synth.synth_clips(f["synth_clip"], 1, f["samples"])[0])
        chunks.append(c)
        labels += [f["label"]] * len(c)
    np.testing.assert_array_equal(np.concatenate(chunks), vec["mfe_x"])
    # the label dictionary is {str(argmax(one-hot row)): label} in order of first appearance
    from mmla_audio_b200 import speaker_identification as si
    yy = si.binarizer(labels, dim=len(set(labels)))
    np.testing.assert_array_equal(yy, vec["mfe_y"])
    assert {str(int(np.argmax(yy[i]))): labels[i] for i in range(len(labels))} == meta["mfe_speaker_id"]


def test_oracle_overlap_features_equal_reference_composition(vec, meta):
    assert meta["ofg_attributes"] == {"get_attributes": [400, 160, 16000], "time_dim": 150, "mel_dim": 128}
    np.testing.assert_array_equal(olm.normalize_matrix(vec["normalize_in"]), vec["normalize_out"])
    for k in ("ofg_2p56s", "ofg_1s_padded"):
        sig = synth.synth_clips(meta[k]["synth_clip"], 1, meta[k]["samples"])[0]
        s_db, s_db_norm = olm.generate_mels(sig)
        np.testing.assert_array_equal(s_db, vec[k + "_s_db"])
        np.testing.assert_array_equal(s_db_norm, vec[k + "_s_db_norm"])          # python double loop == vectorised float32 ops
        np.testing.assert_array_equal(olm.generate_zcr(sig), vec[k + "_zcr"])
        np.testing.assert_array_equal(olm.generate_zcr_image(sig), vec[k + "_image_f64"])
        assert meta[k]["imsave_origin"] == "lower"                                # rows are flipped on save


def test_oracle_frame_count_and_collector_equal_reference(meta):
    fg = meta["frame_generator"]
    for n in ("40960", "24000", "481", "480"):
        assert ovad.num_frames(int(n)) == fg[n]
    assert fg["frame_bytes"] == 960 and abs(fg["timestamp_3"] - 0.09) < 1e-12
    for case in meta["vad_collector"]:
        flags = np.asarray(case["flags"], np.uint8)
        keep = ovad.vad_collector_mask(flags)
        kept = sorted(i for seg in case["segments"] for i in seg)
        assert kept == list(np.nonzero(keep)[0]), case["flags"]
        # the joined output is in frame order (segments are consecutive runs, yielded in order)
        assert [i for seg in case["segments"] for i in seg] == kept


def test_segmentation_index_math_equals_reference_files(meta):
    from mmla_audio_b200.pipeline import segmentation_windows
    inp = meta["segmentation_input"]
    rec = synth.synth_clips(inp["synth_first_clip"], inp["clips"], inp["clip_len"]).reshape(-1)
    rec = rec[: inp["clips"] * inp["clip_len"] - inp["drop_tail"]]
    for case in meta["segmentation"]:
        win, step = int(16000 * case["win"]), int(16000 * case["step"])
        n = segmentation_windows(len(rec), win, step)
        assert n == otally.num_windows(len(rec), win, step) == len(case["segments"])
        for seg in case["segments"]:
            w = rec[seg["j"] * step: seg["j"] * step + win]
            assert len(w) == seg["samples"] and (zlib.crc32(w.tobytes()) & 0xFFFFFFFF) == seg["crc32"]
            assert seg["file_suffix"] == "session_%d_16000_split.wav" % seg["j"]


def _check_viz(lines, golden, initial):
    counts, secs, total = otally.tally_from_log(lines, initial)
    assert [[k, v] for k, v in secs.items()] == golden["pie"]
    return counts


def test_oracle_tallies_equal_reference_visualization(meta):
    for fname, lines in meta["viz_overlap_logs"].items():
        _check_viz(lines, meta["viz_overlap"][fname], ["non-overlapped", "overlapped", "silent"])
    for fname, lines in meta["viz_speaker_logs"].items():
        counts = _check_viz(lines, meta["viz_speaker"][fname], [])
        series = {name: y for name, y in meta["viz_speaker"][fname]["bar_series"]}
        assert {k: sum(v for v in y if v) for k, y in series.items()} == counts


def test_product_log_rows_equal_the_rows_the_reference_parsed(meta):
    """The golden logs were parsed by the reference's own visualization(); the product writes the same rows."""
    from mmla_audio_b200.tally import log_rows
    t0 = datetime(2021, 6, 1, 12, 0, 0, 654321)
    lines = meta["viz_overlap_logs"]["a.txt"]
    labels = [ln.split("\t")[1] for ln in lines[1:]]
    assert log_rows(labels, t0, 1.5, "overlapped degree", add_before_first=False) == lines
    lines = meta["viz_speaker_logs"]["s.txt"]
    labels = [ln.split("\t")[1] for ln in lines[1:]]
    assert log_rows(labels, t0, 2.56, "speaker", add_before_first=True) == lines


def test_oracle_vad_basic_behaviour():
    """No reference vector exists for the detector itself (webrtcvad is not installable): sanity properties only."""
    vad = ovad.Vad(3)
    assert not ovad.remove_silence(np.zeros(40960, np.int16), vad)[1].any()
    vad.reset()
    speechy = synth.synth_clips(0, 1, 40960)[0]
    flags = ovad.remove_silence(speechy, vad)[1]
    assert flags.sum() > 40
    vad.reset()
    assert np.array_equal(ovad.remove_silence(speechy, vad)[1], flags)              # reset => deterministic
