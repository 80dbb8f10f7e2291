#!/usr/bin/env python
"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE (lizaibeim/mmla-audio, /root/reference).

The reference is pure Python, but every script imports third-party packages that are neither vendored nor installable
here (tensorflow, librosa, python_speech_features, webrtcvad, noisereduce, soundfile, pydub, pyaudio, pyecharts, ...).
Those imports are satisfied with stub modules so that the reference's modules IMPORT UNMODIFIED from where they lie, and
then the reference's own functions are called:

  * with NO third-party arithmetic on their path (pure numpy / stdlib) — pinned completely by this script:
      speaker_identification.delta, .binarizer                                   (:141-151, :122-138)
      record_on_pc.frame_generator, .vad_collector   (is_speech replaced by a table of flags)   (:229-295)
      overlap_detection_post_processing.segmentation (writes real WAV segments with `wave`)      (:23-85)
      overlap_degree_distribution.visualization, speaker_time_distribution.visualization
          (pyecharts stubbed by a recorder: the label / seconds series handed to Pie.add, the bar series)
      OverlapFeaturesGenerator.normalize_matrix                                   (:103-117)
  * with the third-party call replaced by THIS REPO'S ORACLE restatement of it (so the vector pins the reference's
    composition AROUND the library call — deltas, concatenation, zero padding / truncation to 256 rows, chunking, label
    dictionary, image channel order, 1 - norm, origin='lower' — not the library arithmetic itself, which stays unpinned):
      speaker_identification.input_feature_gen, .make_feature_experiment   (mfcc := oracle.psf.mfcc)
      OverlapFeaturesGenerator.generate_mels / generate_zcr / generate_zcr_image
          (librosa.load / melspectrogram / power_to_db / zero_crossing_rate := oracle.librosa_mel; plt.imsave recorded)

Run HERE (the GPU box has no /root/reference):   python tests/golden/make_reference_vectors.py
Writes tests/golden/reference_vectors.npz and reference_vectors.json; tests/test_reference_golden_cpu.py checks the
oracle and the host-side product code against them, tests/test_reference_golden_gpu.py the CUDA path.
"""
from __future__ import annotations

import importlib
import json
import os
import sys
import tempfile
import types
import wave
import zlib
from datetime import datetime
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MMLA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import librosa_mel as olm, psf as opsf, synth as osynth, tally as otally  # noqa: E402


# ---------------------------------------------------------------------------------------------
# stubs for the un-installable third-party imports
# ---------------------------------------------------------------------------------------------
class _Recorder:
    """pyecharts chart stand-in: chainable, records add / add_xaxis / add_yaxis payloads."""
    log = []

    def __init__(self, *a, **k):
        self.kind = type(self).__name__
        self.calls = []
        _Recorder.log.append(self)

    def __getattr__(self, name):
        def call(*a, **k):
            self.calls.append((name, a, k))
            return self
        return call


class Pie(_Recorder):
    pass


class Bar(_Recorder):
    pass


class Page(_Recorder):
    pass


def install_stubs():
    def stub(name, **attrs):
        m = mock.MagicMock(name=name)
        m.__name__ = name
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    for name in ["tensorflow", "tensorflow.keras", "tensorflow.keras.backend", "tensorflow.keras.callbacks",
                 "tensorflow.keras.layers", "tensorflow.keras.metrics", "tensorflow.keras.models",
                 "tensorflow.keras.optimizers", "tensorflow.keras.regularizers", "webrtcvad", "pydub", "noisereduce",
                 "soundfile", "pyaudio", "requests", "skimage", "skimage.metrics",
                 "skimage.metrics._structural_similarity", "matplotlib", "matplotlib.pyplot", "genericpath_stub"]:
        if name not in sys.modules:
            stub(name)
    # keras Callback must be a real class (the reference subclasses it at import time)
    sys.modules["tensorflow.keras.callbacks"].Callback = type("Callback", (), {})
    charts = types.ModuleType("pyecharts.charts")
    charts.Pie, charts.Bar, charts.Page = Pie, Bar, Page
    pe = types.ModuleType("pyecharts")
    pe.options = mock.MagicMock(name="pyecharts.options")
    pe.charts = charts
    sys.modules["pyecharts"], sys.modules["pyecharts.charts"], sys.modules["pyecharts.options"] = pe, charts, pe.options
    # python_speech_features.mfcc := the oracle's restatement
    psf_mod = types.ModuleType("python_speech_features")
    psf_mod.mfcc = opsf.mfcc
    sys.modules["python_speech_features"] = psf_mod
    # librosa := the oracle's restatement of the four calls the reference makes
    import scipy.io.wavfile as wavfile

    def load(path, sr=None):
        rate, sig = wavfile.read(path)
        return olm.load_pcm(sig), rate

    def melspectrogram(y, sr, hop_length, n_fft, n_mels):
        assert (sr, hop_length, n_fft) == (16000, 160, 400)
        return olm.melspectrogram(np.asarray(y, np.float32), n_mels)

    def power_to_db(s, ref):
        assert ref is np.max
        return olm.power_to_db_refmax(s)

    def zero_crossing_rate(y, frame_length, hop_length):
        assert (frame_length, hop_length) == (400, 160)
        yp = np.pad(np.asarray(y), frame_length // 2, mode="edge")
        n_frames = 1 + (len(yp) - frame_length) // hop_length
        idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
        fr = yp[idx].copy()
        fr[np.abs(fr) <= 1e-10] = 0
        sign = np.signbit(fr)
        cross = np.zeros(fr.shape, dtype=bool)
        cross[1:] = sign[1:] != sign[:-1]
        return np.mean(cross, axis=0, keepdims=True)

    lib = types.ModuleType("librosa")
    lib.load, lib.power_to_db = load, power_to_db
    lib.feature = types.ModuleType("librosa.feature")
    lib.feature.melspectrogram, lib.feature.zero_crossing_rate = melspectrogram, zero_crossing_rate
    sys.modules["librosa"], sys.modules["librosa.feature"] = lib, lib.feature
    try:
        import cv2  # noqa: F401
    except Exception:
        stub("cv2")
    try:
        import sklearn.model_selection  # noqa: F401
    except Exception:
        stub("sklearn")
        stub("sklearn.model_selection")


def import_reference(subdir: str, module: str):
    path = os.path.join(REF, subdir, "scripts")
    sys.path.insert(0, path)
    try:
        sys.modules.pop(module, None)
        return importlib.import_module(module)
    finally:
        sys.path.remove(path)


def write_wav(path, sig, rate=16000):
    with wave.open(path, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(rate)
        wf.writeframes(np.ascontiguousarray(sig, np.int16).tobytes())


def crc(a) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: this script runs where the reference is mounted")
    install_stubs()
    vec, meta = {}, {"reference": "lizaibeim/mmla-audio", "made_by": "tests/golden/make_reference_vectors.py",
                     "inputs": "oracle.synth.synth_clips (seed 0x6D6D6C61) and numpy default_rng seeds named per entry"}
    tmp = tempfile.mkdtemp(prefix="mmla_golden_")

    # ---- SpeakerIdentification/scripts/speaker_identification.py -------------------------------------------
    si = import_reference("SpeakerIdentification", "speaker_identification")
    rng = np.random.default_rng(20211018)
    for name, T in (("delta_T37", 37), ("delta_T3", 3), ("delta_T1", 1)):
        feat = rng.standard_normal((T, 13)) * 10.0
        vec[name + "_in"] = feat
        vec[name + "_out"] = si.delta(feat, 2)
        vec[name + "_out2"] = si.delta(si.delta(feat, 2), 2)
    names = ["bob", "ann", "bob", "cy", "ann", "ann", "dee"]
    vec["binarizer_out"] = si.binarizer(list(names), dim=4)
    meta["binarizer_in"] = names

    clips = {"ifg_1p5s": osynth.synth_clips(11, 1, 24000)[0], "ifg_2p56s": osynth.synth_clips(12, 1, 40960)[0],
             "ifg_2p9s_truncated": osynth.synth_clips(13, 1, 46400)[0], "ifg_4000": osynth.synth_clips(14, 1, 4000)[0]}
    for k, sig in clips.items():
        p = os.path.join(tmp, k + ".wav")
        write_wav(p, sig)
        out = si.input_feature_gen(p)
        assert out.shape == (1, 256, 39)
        vec[k] = out
        meta[k] = {"synth_clip": {"ifg_1p5s": 11, "ifg_2p56s": 12, "ifg_2p9s_truncated": 13, "ifg_4000": 14}[k],
                   "samples": int(len(sig))}
    p = os.path.join(tmp, "short.wav")
    write_wav(p, osynth.synth_clips(15, 1, 3999)[0])
    meta["ifg_3999_samples"] = si.input_feature_gen(p)
    assert meta["ifg_3999_samples"] == "silent"

    files = []
    for j, (who, n) in enumerate((("speakerB", 48000), ("speakerA", 41200), ("speakerC", 90000))):
        p = os.path.join(tmp, who + ".wav")
        write_wav(p, osynth.synth_clips(100 + j, 1, n)[0])
        files.append(p)
    x, y, spk = si.make_feature_experiment(files)
    vec["mfe_x"], vec["mfe_y"] = x, y
    meta["mfe_speaker_id"] = spk
    meta["mfe_files"] = [{"label": os.path.basename(f)[:-4], "synth_clip": 100 + j, "samples": n}
                         for j, (f, n) in enumerate(zip(files, (48000, 41200, 90000)))]

    # ---- OverlapDetection/scripts/overlap_features_generator.py --------------------------------------------
    ofg_mod = import_reference("OverlapDetection", "overlap_features_generator")
    ofg = ofg_mod.OverlapFeaturesGenerator(25, 10)
    meta["ofg_attributes"] = {"get_attributes": list(ofg.get_attributes()), "time_dim": ofg.time_dim, "mel_dim": ofg.mel_dim}
    m = (np.random.default_rng(5).standard_normal((7, 9)) * 30).astype(np.float32)
    vec["normalize_in"], vec["normalize_out"] = m, ofg.normalize_matrix(m)
    saved = {}

    def imsave(path, arr, origin=None, cmap=None):
        saved["path"], saved["arr"], saved["origin"] = path, np.array(arr), origin
    ofg_mod.plt.imsave = imsave
    for k, clip, n in (("ofg_2p56s", 21, 40960), ("ofg_1s_padded", 22, 16000)):
        p = os.path.join(tmp, k + ".wav")
        write_wav(p, osynth.synth_clips(clip, 1, n)[0])
        s_db, s_db_norm = ofg.generate_mels(p)
        vec[k + "_s_db"], vec[k + "_s_db_norm"] = s_db, s_db_norm
        vec[k + "_zcr"] = ofg.generate_zcr(p)
        img = ofg.generate_zcr_image(p, tmp + "/png/")
        vec[k + "_image_f64"] = img
        assert ofg.generate_zcr_image(p, tmp + "/png/", "x.png") is None
        assert saved["origin"] == "lower" and np.array_equal(saved["arr"], img) and saved["path"] == tmp + "/png/x.png"
        meta[k] = {"synth_clip": clip, "samples": n, "imsave_origin": saved["origin"]}

    # ---- OverlapDetection/scripts/record_on_pc.py: frame_generator + vad_collector --------------------------
    rec = import_reference("OverlapDetection", "record_on_pc")
    audio = osynth.synth_clips(31, 1, 40960)[0].tobytes()
    frames = list(rec.frame_generator(30, audio, 16000))
    meta["frame_generator"] = {"40960": len(frames),
                               "24000": len(list(rec.frame_generator(30, bytes(48000), 16000))),
                               "481": len(list(rec.frame_generator(30, bytes(962), 16000))),
                               "480": len(list(rec.frame_generator(30, bytes(960), 16000))),
                               "frame_bytes": len(frames[0].bytes), "timestamp_3": frames[3].timestamp,
                               "duration": frames[0].duration}

    class FlagVad:
        def __init__(self, flags):
            self.flags, self.i = list(flags), 0

        def is_speech(self, buf, sample_rate):
            v = bool(self.flags[self.i])
            self.i += 1
            return v
    rng = np.random.default_rng(77)
    cases = []
    for n, p_on in ((85, 0.95), (85, 0.8), (85, 0.5), (85, 0.1), (49, 0.97), (9, 1.0), (10, 1.0), (11, 1.0), (200, 0.9), (0, 1.0)):
        flags = (rng.random(n) < p_on).astype(np.uint8)
        cases.append(flags)
    cases.append(np.r_[np.zeros(20, np.uint8), np.ones(30, np.uint8), np.zeros(35, np.uint8)])      # one burst
    cases.append(np.r_[np.ones(12, np.uint8), np.zeros(10, np.uint8), np.ones(15, np.uint8), np.zeros(9, np.uint8)])
    coll = []
    for ci, flags in enumerate(cases):
        fr = [rec.Frame(bytes([i % 256, i // 256]) * 480, i * 0.03, 0.03) for i in range(len(flags))]
        segs = list(rec.vad_collector(16000, 30, 300, FlagVad(flags), fr))
        kept = []
        for s in segs:                                           # recover frame indices from the 2-byte tags
            a = np.frombuffer(s, dtype=np.uint8).reshape(-1, 960)
            kept.append([int(r[0]) + 256 * int(r[1]) for r in a])
        coll.append({"flags": flags.tolist(), "segments": kept})
    meta["vad_collector"] = coll

    # ---- OverlapDetection/scripts/overlap_detection_post_processing.py: segmentation -------------------------
    pp = import_reference("OverlapDetection", "overlap_detection_post_processing")

    # Two compatibility shims, neither touching the reference's source: (1) it calls ndarray.tostring(), an alias of
    # tobytes() that NumPy 2.3 removed -> the module's `np.frombuffer` hands out an ndarray subclass that still has it;
    # (2) it joins paths with a literal backslash (`src_dir + "\\" + f`, written on Windows) -> on POSIX that names a
    # sibling file "<src_dir>\<f>", which is created next to the directory it lists.
    class _Compat(np.ndarray):
        def tostring(self):
            return self.tobytes()

    class _NpProxy:
        def __getattr__(self, k):
            return getattr(np, k)

        def frombuffer(self, *a, **k):
            return np.frombuffer(*a, **k).view(_Compat)
    pp.np = _NpProxy()
    seg_meta = []
    rec_sig = osynth.synth_clips(500, 7, 24000).reshape(-1)[: 7 * 24000 - 5000]
    for tag, win, step in (("w1.5_s1.5", 1.5, 1.5), ("w1.5_s0.5", 1.5, 0.5), ("w2.56_s2.56", 2.56, 2.56)):
        src, dst = os.path.join(tmp, "seg_src_" + tag), os.path.join(tmp, "seg_dst_" + tag)
        os.makedirs(src)
        os.makedirs(dst)
        write_wav(os.path.join(src, "session.wav"), rec_sig)            # what os.listdir(src_dir) finds
        write_wav(src + "\\" + "session.wav", rec_sig)                  # what `src_dir + "\\" + f` opens on POSIX
        pp.segmentation(src, dst, win, step)
        outs = []
        for root_, _d, fs in os.walk(dst):
            for f in fs:
                with wave.open(os.path.join(root_, f), "rb") as wf:
                    data = np.frombuffer(wf.readframes(wf.getnframes()), dtype=np.int16)
                j = int(f.split("_")[-3])                                # '<name>_<j>_<framerate>_split.wav'
                outs.append({"j": j, "file_suffix": f[f.index("session"):], "samples": int(len(data)), "crc32": crc(data)})
        outs.sort(key=lambda d: d["j"])
        seg_meta.append({"tag": tag, "win": win, "step": step, "n_samples": int(len(rec_sig)), "segments": outs})
    meta["segmentation"] = seg_meta
    meta["segmentation_input"] = {"synth_first_clip": 500, "clips": 7, "clip_len": 24000, "drop_tail": 5000}

    # ---- visualization() of both *_distribution.py -------------------------------------------------------------
    def run_visualization(subdir, module, logs):
        mod = import_reference(subdir, module)
        root = tempfile.mkdtemp(prefix="mmla_viz_", dir=tmp)
        os.makedirs(os.path.join(root, "experiment", "logs"))
        for fname, lines in logs.items():
            with open(os.path.join(root, "experiment", "logs", fname), "w") as f:
                f.write("\n".join(lines) + "\n")
        mod.Root_Dir = root
        _Recorder.log = []
        mod.visualization()
        out = {}
        order = sorted(os.listdir(os.path.join(root, "experiment", "logs")))
        listed = os.listdir(os.path.join(root, "experiment", "logs"))
        pies = [r for r in _Recorder.log if r.kind == "Pie"]
        bars = [r for r in _Recorder.log if r.kind == "Bar"]
        assert len(pies) == len(listed) == len(bars)
        for fname, pie, bar in zip(listed, pies, bars):
            add = [c for c in pie.calls if c[0] == "add"][0]
            series = [[c[2]["series_name"], c[2]["y_axis"]] for c in bar.calls if c[0] == "add_yaxis"]
            xaxis = [c for c in bar.calls if c[0] == "add_xaxis"][0][1][0]
            out[fname] = {"pie": add[1][1], "bar_series": series, "bar_xaxis": xaxis}
        assert sorted(out) == order
        return out

    t0 = datetime(2021, 6, 1, 12, 0, 0, 654321)
    lab_o = (["non-overlapped"] * 5 + ["overlapped"] * 3 + ["silent"] * 2 + ["overlapped"] * 7 + ["non-overlapped"]) * 3
    log_o = otally.log_rows(lab_o, t0, 1.5, "overlapped degree", add_before_first=False)
    lab_o2 = ["overlapped"] * 4
    log_o2 = otally.log_rows(lab_o2, datetime(2022, 1, 2, 3, 4, 5, 999999), 1.5, "overlapped degree", add_before_first=False)
    meta["viz_overlap_logs"] = {"a.txt": log_o, "b.txt": log_o2}
    meta["viz_overlap"] = run_visualization("OverlapDetection", "overlap_degree_distribution", meta["viz_overlap_logs"])
    lab_s = (["cy"] * 2 + ["ann"] * 6 + ["silent"] + ["bob"] * 4 + ["ann"] * 3) * 5
    log_s = otally.log_rows(lab_s, t0, 2.56, "speaker", add_before_first=True)
    meta["viz_speaker_logs"] = {"s.txt": log_s}
    meta["viz_speaker"] = run_visualization("SpeakerIdentification", "speaker_time_distribution", meta["viz_speaker_logs"])

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **vec)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(meta, f, indent=1, default=str)
    print("wrote", len(vec), "arrays and", len(meta), "json entries")


if __name__ == "__main__":
    main()
