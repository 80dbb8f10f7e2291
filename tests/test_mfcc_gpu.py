"""GPU parity: fused sm_100a MFCC kernel (through the C-ABI) vs the float64 psf oracle.

Tolerance (stated per BASELINE.json's north_star, "MFCCs within a stated relative tolerance
(e.g. 1e-4 rel on fp32)"): the kernel computes in fp32, the reference in float64 at int16 scale.
Cepstra c1..c12 are signed sums that cross zero, so the bound is
    |gpu - ref| <= 1e-4 * |ref| + 1e-4 * max|ref over the clip|
(rtol 1e-4 plus an absolute floor of 1e-4 of the clip's largest coefficient).
"""
import numpy as np
import pytest

from oracle import psf, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def assert_mfcc_close(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    atol = RTOL * max(np.abs(ref).max(), 1.0)
    err = np.abs(got - ref)
    bound = RTOL * np.abs(ref) + atol
    bad = err > bound
    assert not bad.any(), f"{bad.sum()} of {bad.size} beyond tolerance; max err {err.max():.3e} (atol {atol:.3e})"


def ref_mfcc(sig, nfilt=26):
    return psf.mfcc(sig, 16000, winlen=0.025, winstep=0.01, nfft=512, nfilt=nfilt)


@pytest.mark.parametrize("clip_len", [40000, 24000, 40960])
def test_mfcc13_matches_oracle(cuda, clip_len):
    from mmla_audio_b200 import speaker_identification as si
    pcm = synth.synth_clips(0, 6, clip_len)
    out = si.mfcc_batch(pcm).cpu().numpy()
    for i in range(pcm.shape[0]):
        assert_mfcc_close(out[i], ref_mfcc(pcm[i]))


def test_mfcc_nfilt40_config3(cuda):
    from mmla_audio_b200 import speaker_identification as si
    pcm = synth.synth_clips(100, 4, 40000)
    out = si.mfcc_batch(pcm, si.MfccConfig(nfilt=40)).cpu().numpy()
    assert out.shape == (4, 249, 13)
    for i in range(4):
        assert_mfcc_close(out[i], ref_mfcc(pcm[i], nfilt=40))


def test_reference_signature_mfcc(cuda):
    from mmla_audio_b200 import speaker_identification as si
    sig = synth.synth_clips(7, 1, 40000)[0]
    got = si.mfcc(sig, 16000, winlen=0.025, winstep=0.01, nfft=512)
    assert got.dtype == np.float64 and got.shape == (249, 13)
    assert_mfcc_close(got, ref_mfcc(sig))


def test_speaker_features_256x39(cuda):
    from mmla_audio_b200 import speaker_identification as si
    for clip_len in (24000, 40960, 50000):          # pad (149, 255 frames) and truncate (312 frames)
        pcm = synth.synth_clips(20, 3, clip_len)
        out = si.speaker_features_batch(pcm).cpu().numpy()
        assert out.shape == (3, 256, 39)
        for i in range(3):
            ref = psf.input_feature_gen(pcm[i])[0]
            assert_mfcc_close(out[i], ref)
            T = psf.num_frames(clip_len)
            if T < 256:
                assert np.all(out[i, T:] == 0.0)


def test_input_feature_gen_signature_and_silent(cuda, tmp_path):
    from mmla_audio_b200 import speaker_identification as si
    from mmla_audio_b200.audio_io import write_wav_int16
    sig = synth.synth_clips(3, 1, 40960)[0]
    assert si.input_feature_gen(sig[:3999]) == "silent"
    path = str(tmp_path / "clip.wav")
    write_wav_int16(path, sig)
    got = si.input_feature_gen(path)
    assert got.shape == (1, 256, 39) and got.dtype == np.float64
    assert_mfcc_close(got[0], psf.input_feature_gen(sig)[0])


def test_edge_clips(cuda):
    """all-zero, DC, full-scale square, shorter than a frame, length not a multiple of hop."""
    from mmla_audio_b200 import speaker_identification as si
    rng = np.random.default_rng(5)
    cases = {
        "zero": np.zeros(8000, np.int16),
        "dc": np.full(8000, 1234, np.int16),
        "square": (np.where((np.arange(8000) // 40) % 2 == 0, 32767, -32768)).astype(np.int16),
        "short": rng.integers(-3000, 3000, 137).astype(np.int16),
        "one_frame": rng.integers(-3000, 3000, 400).astype(np.int16),
        "ragged": rng.integers(-20000, 20000, 12345).astype(np.int16),
        "noise_fs": rng.integers(-32768, 32767, 16000).astype(np.int16),
    }
    for name, sig in cases.items():
        got = si.mfcc_batch(sig)[0].cpu().numpy()
        ref = ref_mfcc(sig)
        assert got.shape == ref.shape, name
        if name == "zero":
            np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-4, err_msg=name)
        else:
            assert_mfcc_close(got, ref)


def test_ragged_batch_and_long_file(cuda):
    """Ragged clips in one buffer; a 60 s file crosses several 256-frame tiles (delta halo)."""
    from mmla_audio_b200 import speaker_identification as si
    lens = [40000, 4001, 24000, 163, 33333]
    clips = [synth.synth_clips(50 + i, 1, n)[0] for i, n in enumerate(lens)]
    offs, flat, pos = [], [], 0
    for c in clips:
        pos = (pos + 7) // 8 * 8
        offs.append(pos)
        flat.append((pos, c))
        pos += len(c)
    buf = np.zeros(pos + 8, np.int16)
    for o, c in flat:
        buf[o:o + len(c)] = c
    out, rows = si.mfcc_ragged(buf, offs, lens, with_deltas=True)
    out = out.cpu().numpy()
    for i, c in enumerate(clips):
        assert_mfcc_close(out[i, :rows[i]], psf.mfcc39(c))
    long_sig = synth.synth_clips(900, 24, 40000).reshape(-1)       # 60 s
    chunks = si.whole_file_chunks(long_sig).cpu().numpy()
    ref = psf.chunked_features(long_sig)
    assert chunks.shape == ref.shape == (24, 256, 39)
    for i in range(chunks.shape[0]):
        assert_mfcc_close(chunks[i], ref[i])


def test_delta_signature(cuda):
    from mmla_audio_b200 import speaker_identification as si
    rng = np.random.default_rng(0)
    feat = rng.normal(size=(57, 13))
    np.testing.assert_allclose(si.delta(feat, 2), psf.delta(feat, 2), rtol=1e-5, atol=1e-5)


def test_synth_bit_exact(cuda):
    from mmla_audio_b200 import synth as dsynth
    for first, n, L in ((0, 5, 40000), (123456789012, 3, 24000), (7, 2, 1001)):
        got = dsynth.synth_clips(first, n, L).cpu().numpy()
        np.testing.assert_array_equal(got, synth.synth_clips(first, n, L))


def test_bulk_property_full_size(cuda):
    """Size-independent checks at bench scale: every clip of a large batch equals the same
    clip computed alone (no cross-clip leakage), and output is finite."""
    from mmla_audio_b200 import speaker_identification as si, synth as dsynth
    torch = cuda
    B = 8192
    pcm = dsynth.synth_clips(0, B, 40000)
    out = si.mfcc_batch(pcm, si.MfccConfig(nfilt=40))
    assert torch.isfinite(out).all()
    idx = [0, 1, 4095, 8191]
    solo = si.mfcc_batch(pcm[idx].contiguous(), si.MfccConfig(nfilt=40))
    assert torch.equal(out[idx], solo)
    ref = ref_mfcc(pcm[8191].cpu().numpy(), nfilt=40)
    assert_mfcc_close(out[8191].cpu().numpy(), ref)
