"""Oracle: ``noisereduce.reduce_noise(y_noise=noise, y=y, sr=sr, stationary=True)`` followed by
``sf.write(path, out, 16000)`` — TEST INFRASTRUCTURE ONLY (see oracle/__init__).

PARITY UNPINNED.  ``noisereduce`` is an un-pinned, un-vendored dependency (imported as ``nr`` at
OverlapDetection/scripts/record_on_pc.py:10; called at :211 and overlap_detection_post_processing.py:131 and in the
SpeakerIdentification copies).  Its stationary gate (class SpectralGateStationary of release 2.x) is a thin layer over
``scipy.signal.stft`` / ``istft`` / ``fftconvolve`` — all importable here — so this file restates only that layer, with
noisereduce's defaults for every argument the reference does not pass:

    n_fft = win_length = 1024, hop_length = win_length // 4 = 256, n_std_thresh_stationary = 1.5, prop_decrease = 1.0,
    freq_mask_smooth_hz = 500, time_mask_smooth_ms = 50, chunk_size = 600000, padding = 30000, clip_noise_stationary = True

Assumptions that cannot be checked here (no copy of the package): the dB conversion ``20 log10(|x| + eps)`` floored at
``max over time - 80``; statistics of the noise taken over time per frequency (mean + 1.5 * population std); the
smoothing filter = normalised outer product of two triangles built with ``np.linspace`` as in ``_smoothing_filter``;
the clip processed as ONE chunk, zero-padded by ``padding`` samples on both sides; ``soundfile`` writing float data as
PCM_16 through libsndfile's un-clipped path ``lrint(x * 32767)``.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import fftconvolve, istft, stft

N_FFT, HOP, PADDING, CHUNK = 1024, 256, 30000, 600000


def _amp_to_db(x, top_db=80.0, eps=np.finfo(np.float64).eps):
    x_db = 20 * np.log10(np.abs(x) + eps)
    return np.maximum(x_db, np.max(x_db, axis=-1, keepdims=True) - top_db)


def _smoothing_filter(n_grad_freq, n_grad_time):
    f = np.outer(
        np.concatenate([np.linspace(0, 1, n_grad_freq + 1, endpoint=False), np.linspace(1, 0, n_grad_freq + 2)])[1:-1],
        np.concatenate([np.linspace(0, 1, n_grad_time + 1, endpoint=False), np.linspace(1, 0, n_grad_time + 2)])[1:-1])
    return f / np.sum(f)


def noise_threshold(y_noise, n_std_thresh_stationary=1.5):
    """Per-bin gate threshold from the noise recording (float audio in [-1, 1))."""
    y_noise = np.asarray(y_noise)[:CHUNK]
    _, _, noise_stft = stft(y_noise, nfft=N_FFT, noverlap=N_FFT - HOP, nperseg=N_FFT, padded=False)
    noise_stft_db = _amp_to_db(np.abs(noise_stft))
    return np.mean(noise_stft_db, axis=1) + np.std(noise_stft_db, axis=1) * n_std_thresh_stationary


def reduce_noise_stationary(y, sr, y_noise, prop_decrease=1.0, freq_mask_smooth_hz=500, time_mask_smooth_ms=50):
    """float audio in, float32 audio out (same length)."""
    y = np.asarray(y)
    assert y.ndim == 1 and len(y) <= CHUNK, "single-chunk clips only (the hot path's clips are <= 2.56 s)"
    thresh = noise_threshold(y_noise)
    n_grad_freq = int(freq_mask_smooth_hz / (sr / (N_FFT / 2)))
    n_grad_time = int(time_mask_smooth_ms / ((HOP / sr) * 1000))
    smoothing = _smoothing_filter(n_grad_freq, n_grad_time)
    chunk = np.zeros(len(y) + 2 * PADDING)                    # _read_chunk: float64 zeros, the clip at `padding`
    chunk[PADDING:PADDING + len(y)] = y
    _, _, sig_stft = stft(chunk, nfft=N_FFT, noverlap=N_FFT - HOP, nperseg=N_FFT, padded=False)
    sig_stft_db = _amp_to_db(np.abs(sig_stft))
    db_thresh = np.repeat(np.reshape(thresh, [1, len(thresh)]), np.shape(sig_stft_db)[1], axis=0).T
    sig_mask = sig_stft_db > db_thresh
    sig_mask = sig_mask * prop_decrease + np.ones(np.shape(sig_mask)) * (1.0 - prop_decrease)
    sig_mask = fftconvolve(sig_mask, smoothing, mode="same")
    sig_stft_denoised = sig_stft * sig_mask
    _, denoised = istft(sig_stft_denoised, nfft=N_FFT, noverlap=N_FFT - HOP, nperseg=N_FFT)
    out = np.zeros(len(chunk))
    out[:len(denoised)] = denoised[:len(chunk)]
    return out[PADDING:PADDING + len(y)].astype(np.float32)


def sf_write_pcm16(x) -> np.ndarray:
    """``soundfile.write(path, x, 16000)`` of float data to a WAV: PCM_16 via libsndfile's normalised, un-clipped
    conversion lrint(x * 0x7FFF) (values beyond int16 saturate here; libsndfile would wrap)."""
    return np.clip(np.rint(np.asarray(x, np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def reduce_noise_wav(noise_int16, sig_int16, sr=16000) -> np.ndarray:
    """The file the reference leaves on disk: librosa.load (int16 / 32768) -> reduce_noise -> sf.write."""
    y = np.asarray(sig_int16, np.float32) / np.float32(32768.0)
    n = np.asarray(noise_int16, np.float32) / np.float32(32768.0)
    return sf_write_pcm16(reduce_noise_stationary(y, sr, n))
