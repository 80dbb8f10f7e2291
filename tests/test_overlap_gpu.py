"""GPU parity: fused overlap-feature kernel (C-ABI mmla_overlap_features) vs the librosa/
matplotlib oracle.

Stated tolerances (SURVEY.md §7 H3 — uint8 quantisation is a cliff, so bit-exact images are not
attainable across implementations):
  * zero-crossing rate: exact (integer count / 400)
  * s_db: |diff| <= 0.02 dB ; s_db_norm: |diff| <= 3e-4
  * image: every pixel within 1 LSB, at most 1 % of pixels differing at all
"""
import numpy as np
import pytest

from oracle import librosa_mel as lm, synth

pytestmark = pytest.mark.gpu

MAX_DIFF_FRACTION = 0.01


def _check_clip(sig, feats, i, name=""):
    s_db, norm = lm.generate_mels(sig)
    zcr = lm.generate_zcr(sig)
    img = lm.imsave_rgb_uint8(lm.generate_zcr_image(sig))
    got_db = feats["s_db"][i].cpu().numpy()
    got_norm = feats["s_db_norm"][i].cpu().numpy()
    got_zcr = feats["zcr"][i].cpu().numpy()
    got_img = feats["image"][i].cpu().numpy()
    np.testing.assert_array_equal(np.rint(got_zcr.astype(np.float64) * 400), np.rint(zcr[0] * 400), err_msg=name)
    assert np.abs(got_db - s_db).max() <= 0.02, (name, np.abs(got_db - s_db).max())
    assert np.abs(got_norm - norm).max() <= 3e-4, (name, np.abs(got_norm - norm).max())
    d = np.abs(got_img.astype(np.int32) - img.astype(np.int32))
    assert d.max() <= 1, (name, d.max())
    frac = (d > 0).mean()
    assert frac <= MAX_DIFF_FRACTION, (name, frac)
    return frac


def test_overlap_features_match_oracle(cuda):
    from mmla_audio_b200.overlap_features_generator import OverlapFeaturesGenerator
    ofg = OverlapFeaturesGenerator(wl=25, hl=10)
    assert ofg.get_attributes() == (400, 160, 16000)
    pcm = synth.synth_clips(0, 8, 40000)            # 2.5 s clips: only the first 24000 samples count
    feats = ofg.features_batch(pcm, want=("s_db", "s_db_norm", "zcr", "image"))
    assert tuple(feats["image"].shape) == (8, 128, 151, 3)
    fr = [_check_clip(pcm[i], feats, i, f"clip{i}") for i in range(8)]
    print("differing-pixel fractions:", fr)


def test_overlap_short_and_edge_clips(cuda):
    """shorter than 24000 (zero-padded), exactly 24000, full-scale square wave, DC + step."""
    from mmla_audio_b200.overlap_features_generator import OverlapFeaturesGenerator
    ofg = OverlapFeaturesGenerator(wl=25, hl=10)
    rng = np.random.default_rng(3)
    cases = {
        "short": synth.synth_clips(11, 1, 9000)[0],
        "exact": synth.synth_clips(12, 1, 24000)[0],
        "square": np.where((np.arange(24000) // 37) % 2 == 0, 32767, -32768).astype(np.int16),
        "noise": rng.integers(-20000, 20000, 30000).astype(np.int16),
        "step": np.concatenate([np.full(12000, 500), rng.integers(-9000, 9000, 12000)]).astype(np.int16),
    }
    for name, sig in cases.items():
        feats = ofg.features_batch(sig, want=("s_db", "s_db_norm", "zcr", "image"))
        _check_clip(sig, feats, 0, name)


def test_overlap_reference_signatures(cuda, tmp_path):
    from mmla_audio_b200.overlap_features_generator import OverlapFeaturesGenerator
    from mmla_audio_b200.audio_io import write_wav_int16
    from PIL import Image
    ofg = OverlapFeaturesGenerator(wl=25, hl=10)
    sig = synth.synth_clips(5, 1, 40960)[0]
    path = str(tmp_path / "c.wav")
    write_wav_int16(path, sig)
    s_db, norm = ofg.generate_mels(path)
    assert s_db.shape == norm.shape == (128, 151) and s_db.dtype == np.float32
    zcr = ofg.generate_zcr(path)
    assert zcr.shape == (1, 151) and zcr.dtype == np.float64
    np.testing.assert_array_equal(zcr, lm.generate_zcr(sig))
    img = ofg.generate_zcr_image(path, str(tmp_path / "png") + "/")
    assert img.shape == (128, 151, 3) and img.dtype == np.float64
    ref = lm.generate_zcr_image(sig)
    np.testing.assert_array_equal(img[:, :, 0], ref[:, :, 0])
    assert np.abs(img - ref).max() <= 3e-4
    assert ofg.generate_zcr_image(path, str(tmp_path / "png") + "/", "1.png") is None
    rgba = np.asarray(Image.open(str(tmp_path / "png" / "1.png")))
    assert rgba.shape == (128, 151, 4) and (rgba[..., 3] == 255).all()
    d = np.abs(rgba[..., :3].astype(int) - lm.imsave_rgb_uint8(ref).astype(int))
    assert d.max() <= 1 and (d > 0).mean() <= MAX_DIFF_FRACTION
    m = np.arange(12, dtype=np.float32).reshape(3, 4)
    np.testing.assert_allclose(ofg.normalize_matrix(m), lm.normalize_matrix(m), rtol=1e-6)
    assert np.isnan(ofg.normalize_matrix(np.ones((2, 2), np.float32))).all()


def test_tensor_core_dft_matches_cuda_core_kernel_and_oracle(cuda, monkeypatch):
    """r02: the folded 400-point DFT runs on tcgen05 (fp16 hi + lo split, three products, fp32 accumulation in TMEM) in
    `overlap_features_tc_kernel`; `MMLA_OVERLAP_KERNEL=fp32` selects the r01 CUDA-core contraction.  Both against the oracle
    (same tolerances), exact ZCR on both, and against each other on a ragged batch incl. quiet and all-zero clips."""
    from mmla_audio_b200 import _lib
    from mmla_audio_b200.overlap_features_generator import OverlapFeaturesGenerator
    ofg = OverlapFeaturesGenerator(wl=25, hl=10)
    lib = _lib.load()
    pcm = synth.synth_clips(60, 40, 24000)
    pcm[3] = pcm[3] // 300                                # quiet clip: ~100 LSB peak (lo halves near the fp16 subnormals)
    pcm[4] = 0
    pcm[5, 5000:] = 0
    names0 = []
    feats = {}
    for mode in ("tc", "fp32"):
        if mode == "fp32":
            monkeypatch.setenv("MMLA_OVERLAP_KERNEL", "fp32")
        else:
            monkeypatch.delenv("MMLA_OVERLAP_KERNEL", raising=False)
        tr = _lib.trace_launches(lambda: feats.__setitem__(mode, ofg.features_batch(pcm, want=("s_db", "s_db_norm", "zcr", "image"))), cuda)
        names0.append([n for n, _ in tr])
    assert names0[0] == ["overlap_features_tc_kernel"] and names0[1] == ["overlap_features_kernel"]
    for i in list(range(8)) + [39]:
        if i == 4:
            continue                                      # constant clip: 0/0 in normalize_matrix (the caller discards it)
        _check_clip(pcm[i], feats["tc"], i, f"tc clip{i}")
    assert cuda.equal(feats["tc"]["zcr"], feats["fp32"]["zcr"])
    live = [i for i in range(40) if i != 4]
    d_db = (feats["tc"]["s_db"][live] - feats["fp32"]["s_db"][live]).abs().max().item()
    d_img = (feats["tc"]["image"][live].int() - feats["fp32"]["image"][live].int()).abs()
    print("tensor-core vs CUDA-core DFT: max |d s_db|", d_db, "pixels differing", (d_img > 0).float().mean().item())
    assert d_db <= 0.02 and d_img.max().item() <= 1 and (d_img > 0).float().mean().item() <= 0.01
