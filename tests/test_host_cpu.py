"""CPU tests of the host layer: C-ABI exports, TF tensor-bundle reader/writer, weight spec vs
the reference's own variables.index (decoded fixture), parameter defaults, loud failure without
a GPU.  No compute calls are made."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from mmla_audio_b200 import _lib, params, tf_bundle, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mmla_b200.h")).read()
    declared = set(re.findall(r"\b(mmla_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mmla_abi_version() == 1


def test_host_only_entry_points():
    lib = _lib.load()
    p = _lib.MfccParams(samplerate=16000, frame_len=400, frame_step=160, nfft=512)
    for n, t in ((40000, 249), (24000, 149), (40960, 255), (400, 1), (401, 2), (0, 1)):
        assert lib.mmla_psf_num_frames(n, ctypes.byref(p)) == t == params.MfccConfig().num_frames(n)
    data = b"123456789"
    assert lib.mmla_crc32c_host(data, len(data)) == 0xE3069283 == tf_bundle.crc32c(data)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    from mmla_audio_b200 import speaker_identification as si
    with pytest.raises(_lib.MmlaError):
        si.mfcc_batch(np.zeros((1, 8000), np.int16))
    with pytest.raises(_lib.MmlaError):
        si.input_feature_gen(np.zeros(8000, np.int16))
    lib = _lib.load()
    rc = lib.mmla_tally(None, 0, 0, None, None)
    assert rc != 0 and lib.mmla_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mmla_audio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_bundle_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"layer_with_weights-{i}/kernel/.ATTRIBUTES/VARIABLE_VALUE": rng.normal(size=(3, 4, i + 1)).astype(np.float32)
               for i in range(40)}
    tensors["variables/116/.ATTRIBUTES/VARIABLE_VALUE"] = rng.normal(size=(128, 1024)).astype(np.float32)
    tensors["scalar/.ATTRIBUTES/VARIABLE_VALUE"] = np.float32(3.5).reshape(())
    prefix = str(tmp_path / "variables" / "variables")
    tf_bundle.write_bundle(prefix, tensors)
    header, entries = tf_bundle.read_index(prefix + ".index")
    assert header["num_shards"] == 1 and len(entries) == len(tensors)
    assert [e.key for e in entries] == sorted(tensors, key=lambda s: s.encode())
    back = tf_bundle.read_bundle(prefix, verify_crc=True)
    for k, v in tensors.items():
        np.testing.assert_array_equal(back[k], v)
    os.remove(prefix + ".data-00000-of-00001")
    with pytest.raises(FileNotFoundError):                      # stripped shard: no silent fallback
        tf_bundle.read_bundle(prefix)


@pytest.mark.parametrize("fixture,spec", [("index_overlap_timit2.json", W.OVERLAP),
                                          ("index_overlap_timit1.json", W.OVERLAP),
                                          ("index_speaker_timit.json", W.SPEAKER_BASE)])
def test_weight_spec_matches_reference_index(fixture, spec):
    """Every tensor the spec names exists in the reference's variables.index with that shape, and
    every float weight tensor of the model graph in the index is named by the spec."""
    doc = json.load(open(os.path.join(GOLDEN, fixture)))
    idx = {e["key"]: tuple(e["shape"]) for e in doc["entries"]}
    spec = W.resolve_lstm_keys(spec, idx)          # LSTM tensor names differ between checkpoints
    shapes = W.weight_shapes(spec)
    for k, s in shapes.items():
        assert idx.get(k) == tuple(s), (k, idx.get(k), s)
    model_keys = {k for k in idx if k.startswith("layer_with_weights-") or k.startswith("variables/")
                  or k.startswith("trainable_variables/")}
    assert model_keys == set(shapes)
    # byte size of the inference weights = what load_model will read from the .data shard
    assert sum(int(np.prod(s)) for s in shapes.values()) in (1548706, 1491382)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference mount absent (GPU box)")
def test_reader_on_real_reference_index():
    for fixture in ("index_overlap_timit2.json", "index_speaker_timit.json"):
        doc = json.load(open(os.path.join(GOLDEN, fixture)))
        _, entries = tf_bundle.read_index(os.path.join("/root/reference", doc["source"]))
        live = {e.key: list(e.shape) for e in entries}
        for e in doc["entries"]:
            assert live[e["key"]] == e["shape"]
    with pytest.raises(FileNotFoundError):
        tf_bundle.read_bundle("/root/reference/OverlapDetection/timit/models/timit2.0/variables/variables")


def test_pack_weights_layout():
    from mmla_audio_b200.models import pack_weights
    for spec in (W.OVERLAP, W.speaker_spec(10, "sigmoid")):
        w = W.synthetic_weights(spec, 5)
        blob = pack_weights(spec, w)
        assert blob.dtype == np.float32 and blob.size == sum(v.size for v in w.values())
        np.testing.assert_array_equal(blob[:w[W.lw(0, "kernel")].size], w[W.lw(0, "kernel")].reshape(-1))
        bad = dict(w)
        bad[W.lw(0, "bias")] = np.zeros(3, np.float32)
        with pytest.raises(ValueError):
            pack_weights(spec, bad)


def test_param_defaults_equal_reference_constants():
    c = params.MfccConfig()
    assert (c.samplerate, c.frame_len, c.frame_step, c.nfft, c.nfilt, c.numcep) == (16000, 400, 160, 512, 26, 13)
    assert (c.preemph, c.ceplifter, c.appendEnergy, c.window) == (0.97, 22, True, "rect")
    assert params.OVERLAP_CLIP_SAMPLES == 150 * 160 and params.SILENT_MIN_SAMPLES == 4000
    assert params.round_half_up(2.5) == 3 and params.round_half_up(400.0) == 400


def test_wav_io_roundtrip(tmp_path):
    from mmla_audio_b200.audio_io import read_wav_int16, write_wav_int16
    sig = np.random.default_rng(0).integers(-32768, 32767, 5000).astype(np.int16)
    path = str(tmp_path / "x.wav")
    write_wav_int16(path, sig)
    rate, back = read_wav_int16(path)
    assert rate == 16000
    np.testing.assert_array_equal(back, sig)


def test_segmentation_index_math():
    from mmla_audio_b200.pipeline import segmentation_windows, window_view
    assert segmentation_windows(460800000, 24000, 24000) == 19200
    assert segmentation_windows(460800000, 40960, 40960) == 11250
    x = torch.arange(100, dtype=torch.int16)
    v = window_view(x, 30, 20)                                   # overlapping windows, zero copy
    assert v.shape == (4, 30) and v[3, 0] == 60 and v.data_ptr() == x.data_ptr()


def test_seconds_from_counts_is_the_reference_expression():
    from mmla_audio_b200.tally import seconds_from_counts
    norm, secs = seconds_from_counts([1, 2, 4], 1000.0)
    assert norm == [round(1 / 7, 4), round(2 / 7, 4), round(4 / 7, 4)]
    assert secs == [int(n * 1000.0) for n in norm]


def test_log_rows_literal_reference_strings():
    """a16: the TSV rows, transcribed by hand from the reference's write statements —
    overlap_detection_post_processing.py:213-224 (`'segment' + '\\t' + 'overlapped degree' + '\\t' +
    'timestamp'`; row 0 carries `time` as is, every later row first does `time = time +
    timedelta(seconds=1.5)`) and speaker_identification_post_processing.py:278-312 (`time = time +
    timedelta(seconds=2.56)` BEFORE every row, the first included; header 'speaker').  `str(datetime)`
    drops the fraction when microsecond == 0, which is what the `[:-7]` parsing trips over."""
    from datetime import datetime
    from mmla_audio_b200.tally import log_rows
    t0 = datetime(2021, 6, 1, 12, 0, 0, 654321)
    assert log_rows(["non-overlapped", "overlapped", "silent"], t0, 1.5, "overlapped degree", add_before_first=False) == [
        "segment\toverlapped degree\ttimestamp",
        "0\tnon-overlapped\t2021-06-01 12:00:00.654321",
        "1\toverlapped\t2021-06-01 12:00:02.154321",
        "2\tsilent\t2021-06-01 12:00:03.654321",
    ]
    assert log_rows(["alice", "silent", "bob"], t0, 2.56, "speaker", add_before_first=True) == [
        "segment\tspeaker\ttimestamp",
        "0\talice\t2021-06-01 12:00:03.214321",
        "1\tsilent\t2021-06-01 12:00:05.774321",
        "2\tbob\t2021-06-01 12:00:08.334321",
    ]
    whole = datetime(2021, 12, 31, 23, 59, 59)                   # microsecond == 0: no ".000000" in str()
    assert log_rows(["overlapped", "overlapped"], whole, 1.5, "overlapped degree", add_before_first=False) == [
        "segment\toverlapped degree\ttimestamp",
        "0\toverlapped\t2021-12-31 23:59:59",
        "1\toverlapped\t2022-01-01 00:00:00.500000",
    ]


def test_normalize_names_accepts_reference_str_keys():
    """make_feature_experiment / speaker_id_dict are keyed by str(idx) (speaker_identification.py:360-369)."""
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.tally import normalize_names
    yy = si.binarizer(["bob", "bob", "alice"], dim=2)
    speaker_id = {str(int(np.argmax(yy[i]))): n for i, n in enumerate(["bob", "bob", "alice"])}
    assert speaker_id == {"0": "bob", "1": "alice"}
    assert normalize_names(speaker_id) == {0: "bob", 1: "alice"}
    assert normalize_names({0: "x"}) == {0: "x"} and normalize_names(None) == {}


def test_synthetic_bundle_carries_real_crcs(tmp_path):
    """save_synthetic_model writes masked crc32c values TensorFlow's BundleReader would accept."""
    w = {"a/.ATTRIBUTES/VARIABLE_VALUE": np.arange(70000, dtype=np.float32),
         "b/.ATTRIBUTES/VARIABLE_VALUE": np.ones((3, 5), np.float32)}
    prefix = str(tmp_path / "variables" / "variables")
    tf_bundle.write_bundle(prefix, w)
    _, entries = tf_bundle.read_index(prefix + ".index")
    assert all(e.crc32c != 0 for e in entries)
    back = tf_bundle.read_bundle(prefix, verify_crc=True)
    for k in w:
        np.testing.assert_array_equal(back[k], w[k])
    # the slice-by-8 python loop and the library routine agree
    blob = w["a/.ATTRIBUTES/VARIABLE_VALUE"].tobytes()
    native = tf_bundle.crc32c(blob)
    saved, tf_bundle._crc_native = tf_bundle._crc_native, False
    try:
        assert tf_bundle.crc32c(blob) == native
    finally:
        tf_bundle._crc_native = saved
