"""GPU parity: stationary spectral-gating noise reduction (`mmla_noise_profile` / `mmla_noise_gate`) against the
oracle's restatement of noisereduce's stationary gate over scipy.signal.stft / istft / fftconvolve.

Floating point (fp32 FFTs on the device, float64 scipy in the oracle); tolerances, stated per check:
  * gate thresholds (dB):                    |d| <= 0.02 dB
  * rewritten PCM_16 samples:                relative L2 error <= 1e-3 (-60 dB) per clip, >= 99 % of samples within 2 LSB
    (a spectrogram cell that sits within float rounding of its threshold may flip its 0/1 mask; after the 33 x 7
    smoothing that moves a handful of samples by a few LSB, which is what the 1 % allowance is for)
"""
import numpy as np
import pytest

from oracle import noisereduce_stationary as onr, synth

pytestmark = pytest.mark.gpu


def _noise(seed, n, amp):
    return (np.random.default_rng(seed).standard_normal(n) * amp).astype(np.int16)


def _noisy_clips(first, n, L, amp, seed):
    rng = np.random.default_rng(seed)
    clean = synth.synth_clips(first, n, L).astype(np.int32) // 2
    return (clean + (rng.standard_normal((n, L)) * amp).astype(np.int32)).clip(-32768, 32767).astype(np.int16)


def _check(got, ref):
    d = got.astype(np.float64) - ref.astype(np.float64)
    denom = max(np.linalg.norm(ref.astype(np.float64)), 1.0)
    rel = np.linalg.norm(d) / denom
    within = (np.abs(d) <= 2).mean()
    assert rel <= 1e-3 and within >= 0.99, (rel, within, np.abs(d).max())
    return rel, within


def test_noise_profile_threshold_matches_oracle(cuda):
    from mmla_audio_b200.noise_reduction import NoiseProfile
    for seed, n, amp in ((1, 160000, 60.0), (2, 20000, 300.0), (3, 700000, 25.0), (4, 1500, 100.0)):
        noise = _noise(seed, n, amp)
        got = NoiseProfile(noise).thresh.cpu().numpy()
        ref = onr.noise_threshold(noise.astype(np.float32) / np.float32(32768.0))
        assert got.shape == ref.shape == (513,)
        assert np.abs(got - ref).max() <= 0.02, (n, np.abs(got - ref).max())


@pytest.mark.parametrize("L", [40960, 24000])
def test_noise_gate_matches_oracle(cuda, L):
    """record_on_pc.py:208-212 on a batch: 10 s ambient-noise recording, 2.56 s / 1.5 s clips."""
    from mmla_audio_b200.noise_reduction import NoiseProfile, reduce_noise_batch
    noise = _noise(11, 160000, 60.0)
    clips = _noisy_clips(50, 10, L, 60.0, 12)
    clips[7] = 0                                                     # an all-zero clip stays all-zero
    clips[8] = _noise(13, L, 60.0)                                   # noise only: mostly gated away
    out = reduce_noise_batch(clips, NoiseProfile(noise)).cpu().numpy()
    assert out.shape == clips.shape and out.dtype == np.int16
    for i in range(len(clips)):
        ref = onr.reduce_noise_wav(noise, clips[i])
        if i == 7:
            assert not out[i].any() and not ref.any()
            continue
        rel, within = _check(out[i], ref)
    # the gate does something: the noise-only clip loses most of its energy, a speech clip keeps most of its
    e = lambda a: float((a.astype(np.float64) ** 2).sum())
    assert e(out[8]) < 0.2 * e(clips[8]) and e(out[0]) > 0.5 * e(clips[0])


def test_noise_gate_quiet_noise_edge_mask_and_ragged(cuda):
    """Very quiet ambient noise + loud clips: the floor (row max - 80 dB) of loud bins lies ABOVE the gate threshold, so
    even the all-zero frames of the chunk padding carry mask 1 (the kernel's "edge mask" path); ragged clip lengths."""
    from mmla_audio_b200.noise_reduction import NoiseProfile, reduce_noise_batch
    noise = _noise(21, 48000, 1.2)
    clips = (synth.synth_clips(70, 5, 30000).astype(np.int32) * 3).clip(-32768, 32767).astype(np.int16)
    lengths = np.array([30000, 29999, 5000, 1024, 17], np.int32)
    prof = NoiseProfile(noise)
    out = reduce_noise_batch(clips, prof, lengths=lengths).cpu().numpy()
    for i, n in enumerate(lengths):
        ref = onr.reduce_noise_wav(noise, clips[i, :n])
        _check(out[i, :n], ref)
        assert not out[i, n:].any()


def test_reduce_noise_reference_signature(cuda):
    """`nr.reduce_noise(y_noise=noise, y=y, sr=sr, stationary=True)` on librosa-style float audio."""
    from mmla_audio_b200 import noise_reduction as nr
    noise = _noise(31, 160000, 60.0)
    clip = _noisy_clips(90, 1, 40960, 60.0, 32)[0]
    y, yn = clip.astype(np.float32) / 32768, noise.astype(np.float32) / 32768
    got = nr.reduce_noise(y_noise=yn, y=y, sr=16000, stationary=True)
    assert got.dtype == np.float32 and got.shape == y.shape
    _check(np.rint(got * 32768).astype(np.int16), onr.reduce_noise_wav(noise, clip))
    with pytest.raises(nr._lib.MmlaError):
        nr.reduce_noise(y=y, sr=16000, stationary=False, y_noise=yn)
