"""Stationary spectral-gating noise reduction on the device — the ``noise_reduce`` branch of the reference's
``save_wave_file`` (OverlapDetection/scripts/record_on_pc.py:208-212):

    noise, sr = librosa.load(NOISE_PATH, sr=None); y, sr = librosa.load(filepath, sr=None)
    noise_reduced_wav = nr.reduce_noise(y_noise=noise, y=y, sr=sr, stationary=True)
    sf.write(filepath, noise_reduced_wav, 16000)

and ``standardize_audio``'s repeated passes (overlap_detection_post_processing.py:128-133).  ``reduce_noise`` keeps
noisereduce's keyword signature for the arguments the reference passes; every other parameter is at its default
(n_fft 1024, hop 256, n_std_thresh_stationary 1.5, prop_decrease 1.0, freq_mask_smooth_hz 500, time_mask_smooth_ms 50).
"""
from __future__ import annotations

import numpy as np

from . import _lib


class NoiseProfile:
    """Per-bin gate threshold (float32 CUDA [513]) of one ambient-noise recording; reusable across clips."""

    def __init__(self, y_noise, n_std_thresh_stationary: float = 1.5):
        from .speaker_identification import _to_device_pcm
        torch = _lib.require_cuda()
        lib = _lib.load()
        x = _to_device_pcm(torch, _as_int16(y_noise)).reshape(-1)
        self.thresh = torch.empty((513,), dtype=torch.float32, device=x.device)
        _lib.check(lib.mmla_noise_profile(x.data_ptr(), x.numel(), float(n_std_thresh_stationary), self.thresh.data_ptr(),
                                          _lib.stream_ptr(torch)), "mmla_noise_profile")


def _as_int16(y):
    """int16 samples from int16 input or from librosa-style float audio in [-1, 1) (= int16 / 32768 exactly)."""
    a = y if hasattr(y, "dtype") and str(y.dtype).endswith("int16") else np.asarray(y)
    if hasattr(a, "is_cuda"):
        return a
    if a.dtype == np.int16:
        return a
    if a.dtype.kind == "f":
        return np.clip(np.rint(a.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    raise TypeError("audio must be int16 or float in [-1, 1)")


def reduce_noise_batch(pcm, profile: NoiseProfile, lengths=None):
    """pcm: int16 [B, L] (numpy / torch).  Returns int16 CUDA [B, L]: every clip gated against ``profile`` and
    quantised as ``sf.write`` does.  ``lengths``: optional int32 [B] for ragged clips."""
    from .speaker_identification import _to_device_pcm
    torch = _lib.require_cuda()
    lib = _lib.load()
    x = _to_device_pcm(torch, pcm)
    if x.dim() == 1:
        x = x[None, :]
    B, L = x.shape
    stride0 = x.stride(0) if B > 1 else L
    out = torch.zeros((B, L), dtype=torch.int16, device=x.device)
    len_ptr = None
    if lengths is not None:
        lengths = torch.as_tensor(lengths, dtype=torch.int32).to(x.device).contiguous()
        len_ptr = lengths.data_ptr()
    _lib.check(lib.mmla_noise_gate(x.data_ptr(), B, L, stride0, len_ptr, profile.thresh.data_ptr(), out.data_ptr(), L,
                                   _lib.stream_ptr(torch)), "mmla_noise_gate")
    return out


def reduce_noise_int16(y_noise, y, sr=16000):
    """One clip, int16 in → int16 numpy out (the WAV the reference leaves on disk after ``sf.write``)."""
    if sr != 16000:
        raise _lib.MmlaError("the device gate is built for sr = 16000 (the reference's only rate)")
    return reduce_noise_batch(_as_int16(y), NoiseProfile(y_noise))[0].cpu().numpy()


def reduce_noise(y=None, sr=16000, stationary=True, y_noise=None, **kwargs):
    """``noisereduce.reduce_noise(y_noise=noise, y=y, sr=sr, stationary=True)`` → float32 array like the input
    (values are the PCM_16 samples ``sf.write`` would store, divided by 32768 as ``librosa.load`` reads them back)."""
    if not stationary or y_noise is None or kwargs:
        raise _lib.MmlaError("only the reference's call is built: reduce_noise(y_noise=noise, y=y, sr=sr, stationary=True)")
    return reduce_noise_int16(y_noise, y, sr).astype(np.float32) / np.float32(32768.0)
