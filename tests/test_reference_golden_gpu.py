"""GPU parity against golden vectors produced by RUNNING THE REFERENCE'S OWN CODE
(tests/golden/make_reference_vectors.py; see tests/test_reference_golden_cpu.py for what they pin).  The product is
called through the reference's own signatures (WAV paths in, numpy out), i.e. through the C-ABI.

Tolerances: MFCC-derived values 1e-4 relative + 1e-4 * max|ref| (fp32 kernel vs the float64 reference composition);
dB maps 0.02 dB, normalised maps 3e-4 (fp32 kernel vs float32 librosa restatement); ZCR, one-hot labels, dictionaries,
window counts, tallies: exact."""
import json
import os
from datetime import datetime

import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLDEN, "reference_vectors.npz"))


@pytest.fixture(scope="module")
def meta():
    return json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))


def _close_mfcc(got, ref):
    tol = 1e-4 * np.abs(ref) + 1e-4 * np.abs(ref).max()
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= tol), float(np.abs(got - ref).max())


def test_delta_signature(cuda, vec):
    from mmla_audio_b200 import speaker_identification as si
    for name in ("delta_T37", "delta_T3", "delta_T1"):
        d1 = si.delta(vec[name + "_in"], 2)
        assert d1.dtype == np.float64
        np.testing.assert_allclose(d1, vec[name + "_out"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(si.delta(d1, 2), vec[name + "_out2"], rtol=1e-5, atol=1e-5)


def test_input_feature_gen_from_wav_paths(cuda, vec, meta, tmp_path):
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.audio_io import write_wav_int16
    for k in ("ifg_1p5s", "ifg_2p56s", "ifg_2p9s_truncated", "ifg_4000"):
        p = str(tmp_path / (k + ".wav"))
        write_wav_int16(p, synth.synth_clips(meta[k]["synth_clip"], 1, meta[k]["samples"])[0])
        out = si.input_feature_gen(p)
        assert out.dtype == np.float64 and out.shape == (1, 256, 39)
        _close_mfcc(out, vec[k])
        assert not out[0, 256 - 1:].any() or meta[k]["samples"] > 41000     # zero rows below the real frames
    p = str(tmp_path / "short.wav")
    write_wav_int16(p, synth.synth_clips(15, 1, 3999)[0])
    assert si.input_feature_gen(p) == "silent" == meta["ifg_3999_samples"]


def test_make_feature_experiment_from_wav_paths(cuda, vec, meta, tmp_path):
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.audio_io import write_wav_int16
    files = []
    for f in meta["mfe_files"]:
        p = str(tmp_path / (f["label"] + ".wav"))
        write_wav_int16(p, synth.synth_clips(f["synth_clip"], 1, f["samples"])[0])
        files.append(p)
    x, y, spk = si.make_feature_experiment(files)
    _close_mfcc(x, vec["mfe_x"])
    np.testing.assert_array_equal(y, vec["mfe_y"])
    assert spk == meta["mfe_speaker_id"]


def test_overlap_features_generator_from_wav_paths(cuda, vec, meta, tmp_path):
    from mmla_audio_b200.audio_io import write_wav_int16
    from mmla_audio_b200.overlap_features_generator import OverlapFeaturesGenerator
    ofg = OverlapFeaturesGenerator(25, 10)
    assert list(ofg.get_attributes()) == meta["ofg_attributes"]["get_attributes"]
    assert (ofg.time_dim, ofg.mel_dim) == (meta["ofg_attributes"]["time_dim"], meta["ofg_attributes"]["mel_dim"])
    np.testing.assert_allclose(ofg.normalize_matrix(vec["normalize_in"]), vec["normalize_out"], rtol=0, atol=1e-6)
    for k in ("ofg_2p56s", "ofg_1s_padded"):
        p = str(tmp_path / (k + ".wav"))
        write_wav_int16(p, synth.synth_clips(meta[k]["synth_clip"], 1, meta[k]["samples"])[0])
        s_db, s_db_norm = ofg.generate_mels(p)
        assert s_db.dtype == np.float32 and s_db.shape == (128, 151)
        assert np.abs(s_db - vec[k + "_s_db"]).max() <= 0.02
        assert np.abs(s_db_norm - vec[k + "_s_db_norm"]).max() <= 3e-4
        np.testing.assert_array_equal(ofg.generate_zcr(p), vec[k + "_zcr"])
        img = ofg.generate_zcr_image(p, str(tmp_path) + "/png/")
        assert img.dtype == np.float64 and img.shape == (128, 151, 3)
        np.testing.assert_array_equal(img[:, :, 0], vec[k + "_image_f64"][:, :, 0])
        assert np.abs(img - vec[k + "_image_f64"]).max() <= 3e-4
        # the PNG: rows flipped (origin='lower'), trunc(v*255); within 1 LSB of the reference image on <= 1 % of pixels
        assert ofg.generate_zcr_image(p, str(tmp_path) + "/png/", k + ".png") is None
        from PIL import Image
        png = np.asarray(Image.open(str(tmp_path) + "/png/" + k + ".png"))
        assert png.shape == (128, 151, 4) and (png[:, :, 3] == 255).all()
        want = (vec[k + "_image_f64"][::-1] * 255).astype(np.uint8)
        d = np.abs(png[:, :, :3].astype(int) - want.astype(int))
        assert d.max() <= 1 and (d > 0).mean() <= 0.01


def test_segmentation_drop_in_writes_the_reference_files(cuda, meta, tmp_path):
    import wave
    import zlib
    from mmla_audio_b200.audio_io import write_wav_int16
    from mmla_audio_b200.overlap_detection_post_processing import segmentation
    from mmla_audio_b200.speaker_identification_post_processing import segmentation as seg_si
    assert seg_si is segmentation
    inp = meta["segmentation_input"]
    rec = synth.synth_clips(inp["synth_first_clip"], inp["clips"], inp["clip_len"]).reshape(-1)
    rec = rec[: inp["clips"] * inp["clip_len"] - inp["drop_tail"]]
    for case in meta["segmentation"]:
        src, dst = tmp_path / ("src_" + case["tag"]), tmp_path / ("dst_" + case["tag"])
        src.mkdir()
        dst.mkdir()
        write_wav_int16(str(src / "session.wav"), rec)
        written = segmentation(str(src), str(dst), case["win"], case["step"])
        assert len(written) == len(case["segments"])
        for seg in case["segments"]:
            path = dst / "session" / seg["file_suffix"]
            with wave.open(str(path), "rb") as wf:
                data = wf.readframes(wf.getnframes())
            assert len(data) // 2 == seg["samples"] and (zlib.crc32(data) & 0xFFFFFFFF) == seg["crc32"]


def test_file_based_visualization_equals_reference(cuda, meta, tmp_path):
    """`odd.visualization()` / `std.visualization()` over experiment/logs/*: the label -> seconds series the reference
    hands to its pie chart, and the bar series, from the reference's own run on the same log files."""
    from mmla_audio_b200 import overlap_degree_distribution as odd, speaker_time_distribution as std
    for mod, logs_key, gold_key, sub in ((odd, "viz_overlap_logs", "viz_overlap", "o"), (std, "viz_speaker_logs", "viz_speaker", "s")):
        root = tmp_path / sub
        (root / "experiment" / "logs").mkdir(parents=True)
        for fname, lines in meta[logs_key].items():
            (root / "experiment" / "logs" / fname).write_text("\n".join(lines) + "\n")
        mod.Root_Dir = str(root)
        res = mod.visualization()
        for fname, gold in meta[gold_key].items():
            r = res[fname]
            assert [[l, s] for l, s in zip(r["labels"], r["seconds"])] == gold["pie"]
            assert r["x_bar"] == gold["bar_xaxis"]
            for name, y in gold["bar_series"]:
                assert r["bars"][name] == y, name
