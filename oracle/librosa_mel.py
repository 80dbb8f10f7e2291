"""Oracle: the librosa (0.8/0.9) + matplotlib ``imsave`` arithmetic behind the reference's
``OverlapFeaturesGenerator`` (OverlapDetection/scripts/overlap_features_generator.py:29-151).
TEST INFRASTRUCTURE ONLY (see oracle/__init__).

PARITY UNPINNED — librosa and matplotlib are un-pinned, un-vendored dependencies.  Version
evidence: the positional call ``melspectrogram(y, sr, ...)`` at
``overlap_features_generator.py:81`` is only accepted by librosa < 0.10, whose ``stft`` centres
with ``pad_mode='reflect'``, computes the FFT in float64 and stores complex64, and whose mel
basis is float32 slaney-normalised.  ``plt.imsave`` quantises floats with
``(x*255).astype(uint8)`` (truncation) after flipping rows for ``origin='lower'``.

Input here is an int16 array instead of a WAV path; ``librosa.load(sr=None)`` on a 16-bit WAV
yields float32 = int16/32768.
"""
from __future__ import annotations

import numpy as np

SR = 16000
N_FFT = 400
HOP = 160
TIME_DIM = 150
N_MELS = 128
CLIP_SAMPLES = HOP * TIME_DIM      # 24 000  (overlap_features_generator.py:73-80)


def load_pcm(sig_int16) -> np.ndarray:
    """librosa.load(sr=None) on PCM16: float32 in [-1, 1)."""
    return (np.asarray(sig_int16, dtype=np.int16).astype(np.float32) / np.float32(32768.0))


def pad_or_truncate(y: np.ndarray) -> np.ndarray:
    """overlap_features_generator.py:73-80,94-98."""
    if len(y) < CLIP_SAMPLES:
        y = np.pad(y, (0, CLIP_SAMPLES - len(y)), "constant")
    return y[:CLIP_SAMPLES]


def hann_periodic(n: int) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True), float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft(y: np.ndarray, n_fft=N_FFT, hop=HOP) -> np.ndarray:
    """librosa.stft(center=True, pad_mode='reflect', window='hann') → complex64 [1+n_fft/2, T]."""
    win = hann_periodic(n_fft).reshape(-1, 1)
    yp = np.pad(y, n_fft // 2, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx]                                  # float32 [n_fft, T]
    spec = np.fft.rfft(win * frames, axis=0)          # float64 maths
    return spec.astype(np.complex64)


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        m = f >= min_log_hz
        mels[m] = min_log_mel + np.log(f[m] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(mels):
    mels = np.asanyarray(mels, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        m = mels >= min_log_mel
        freqs[m] = min_log_hz * np.exp(logstep * (mels[m] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_frequencies(n_mels, fmin=0.0, fmax=8000.0):
    return mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels))


def mel_basis(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32) → float32 [n_mels, 1+n_fft/2]."""
    if fmax is None:
        fmax = sr / 2.0
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def melspectrogram(y: np.ndarray, n_mels=N_MELS) -> np.ndarray:
    """librosa.feature.melspectrogram(y, sr, hop_length=160, n_fft=400, n_mels) → float32."""
    S = np.abs(stft(y)) ** 2.0                        # float32 [201, T]
    return np.dot(mel_basis(n_mels=n_mels), S)        # float32 [n_mels, T]


def power_to_db_refmax(S: np.ndarray, amin=1e-10, top_db=80.0) -> np.ndarray:
    """librosa.power_to_db(S, ref=np.max)."""
    S = np.asarray(S)
    ref_value = np.max(S)
    log_spec = 10.0 * np.log10(np.maximum(np.float32(amin), S))            # float32 array
    # numpy-1.21 semantics: scalar-with-scalar maximum/log10 run in float64, and the float64
    # scalar is then applied to the float32 array in float32.
    ref_db = 10.0 * np.log10(max(float(amin), float(ref_value)))
    log_spec -= np.float32(ref_db)
    return np.maximum(log_spec, log_spec.max() - np.float32(top_db))


def normalize_matrix(m: np.ndarray) -> np.ndarray:
    """overlap_features_generator.py:103-117 (vectorised; same per-element float32 ops)."""
    max_val = np.max(m)
    min_val = np.min(m)
    diff = max_val - min_val
    with np.errstate(invalid="ignore", divide="ignore"):
        return ((m - min_val) / diff).astype(m.dtype)


def generate_mels(sig_int16, n_mels=N_MELS):
    """``generate_mels`` (overlap_features_generator.py:65-85) → (s_db, s_db_norm) float32."""
    y = pad_or_truncate(load_pcm(sig_int16))
    s = melspectrogram(y, n_mels)
    s_db = power_to_db_refmax(s)
    return s_db, normalize_matrix(s_db)


def generate_zcr(sig_int16) -> np.ndarray:
    """``generate_zcr`` (overlap_features_generator.py:87-101): librosa
    zero_crossing_rate(frame_length=400, hop_length=160, center=True) → float64 [1,151]."""
    y = pad_or_truncate(load_pcm(sig_int16))
    yp = np.pad(y, N_FFT // 2, mode="edge")
    n_frames = 1 + (len(yp) - N_FFT) // HOP
    idx = np.arange(N_FFT)[:, None] + HOP * np.arange(n_frames)[None, :]
    fr = yp[idx].copy()
    fr[np.abs(fr) <= 1e-10] = 0
    sign = np.signbit(fr)
    cross = np.zeros(fr.shape, dtype=bool)
    cross[1:] = sign[1:] != sign[:-1]
    return np.mean(cross, axis=0, keepdims=True)


def generate_zcr_image(sig_int16) -> np.ndarray:
    """``generate_zcr_image(..., out_name=None)`` (overlap_features_generator.py:133-149) →
    float64 [128,151,3] = (zcr, 1-norm, 1-norm)."""
    _, norm = generate_mels(sig_int16)
    zcr = generate_zcr(sig_int16)
    img = np.empty((norm.shape[0], norm.shape[1], 3))
    img[:, :, 0] = zcr[0][None, :]
    img[:, :, 1] = 1 - norm          # float32 arithmetic, widened on store
    img[:, :, 2] = 1 - norm
    return img


def imsave_rgb_uint8(img: np.ndarray) -> np.ndarray:
    """``plt.imsave(origin='lower')`` then ``tf.image.decode_png(., 3)``: flip rows,
    ``(x*255).astype(uint8)`` (truncation), drop alpha → uint8 [128,151,3]."""
    arr = img[::-1]
    if arr.max() > 1 or arr.min() < 0:
        raise ValueError("Floating point image RGB values must be in the 0..1 range.")
    return (arr * 255).astype(np.uint8)


def classifier_input(sig_int16) -> np.ndarray:
    """Full feature path of OverlapDetection/scripts/record_on_pc.py:139,156-158 →
    float32 [128,151,3] with values 0..255 (NOT rescaled)."""
    return imsave_rgb_uint8(generate_zcr_image(sig_int16)).astype(np.float32)
