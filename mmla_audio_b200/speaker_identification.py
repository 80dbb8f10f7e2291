"""Drop-in for the feature half of the reference's ``speaker_identification`` module
(SpeakerIdentification/scripts/speaker_identification.py) — same names, argument meaning and
return types, computed by the fused sm_100a MFCC kernel through ``mmla_psf_mfcc``.

Reference signatures kept (file:line in SpeakerIdentification/scripts/):
  * ``mfcc(signal, samplerate, winlen, winstep, numcep, nfilt, nfft, ...)`` — psf keyword
    signature as called at speaker_identification.py:89,285,341,386
  * ``delta(feat, N)``                          speaker_identification.py:141-151
  * ``input_feature_gen(wav_path)``             speaker_identification.py:372-398
  * ``make_feature_experiment(wav_files)``      speaker_identification.py:317-369
  * ``binarizer(str_list, dim)``                speaker_identification.py:122-138

In addition every function accepts in-memory int16 buffers (numpy / torch) where the reference
takes a WAV path, and ``mfcc_batch`` / ``speaker_features_batch`` expose the batched device
path the benchmark uses.  Results are float32-computed; the reference-signature functions
return float64 numpy arrays like the reference does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .audio_io import as_int16_signal
from .params import MfccConfig, SILENT_MIN_SAMPLES, SPEAKER_FRAMES, WINDOW_IDS

__all__ = ["mfcc", "delta", "input_feature_gen", "make_feature_experiment", "binarizer",
           "mfcc_batch", "speaker_features_batch", "whole_file_chunks", "MfccConfig"]


def _c_params(cfg: MfccConfig, with_deltas: bool, pad_frames: int) -> _lib.MfccParams:
    return _lib.MfccParams(
        samplerate=cfg.samplerate, frame_len=cfg.frame_len, frame_step=cfg.frame_step,
        nfft=cfg.nfft, nfilt=cfg.nfilt, numcep=cfg.numcep, ceplifter=cfg.ceplifter,
        append_energy=1 if cfg.appendEnergy else 0, window=WINDOW_IDS[cfg.window],
        with_deltas=1 if with_deltas else 0, pad_frames=int(pad_frames),
        preemph=cfg.preemph, lowfreq=cfg.lowfreq,
        highfreq=cfg.highfreq if cfg.highfreq else cfg.samplerate / 2.0)


def _to_device_pcm(torch, pcm):
    """int16 CUDA tensor (contiguous, 16-byte aligned) from numpy / torch / list input."""
    if isinstance(pcm, torch.Tensor):
        t = pcm
        if t.dtype != torch.int16:
            raise TypeError(f"PCM must be int16 (as scipy.io.wavfile.read returns it), got {t.dtype}")
        if not t.is_cuda:
            t = t.pin_memory().cuda(non_blocking=True) if t.numel() else t.cuda()
    else:
        a = np.asarray(pcm)
        if a.dtype != np.int16:
            raise TypeError(f"PCM must be int16 (as scipy.io.wavfile.read returns it), got {a.dtype}")
        a = np.ascontiguousarray(a)
        if not a.flags.writeable:                      # e.g. np.frombuffer views: torch wants a writable source
            a = a.copy()
        t = torch.from_numpy(a).cuda()
    if t.dim() == 0 or t.stride(-1) != 1 or t.data_ptr() % 16:
        t = t.contiguous().clone()
    return t


def mfcc_batch(pcm, cfg: MfccConfig = MfccConfig(), with_deltas: bool = False, pad_frames: int = 0,
               out=None, row_stride: int = 0, cmvn=False):
    """Batched device path.  ``pcm``: int16 [B, L] (numpy or torch; CUDA tensors are used in
    place).  Returns a float32 CUDA tensor [B, rows, dim] with rows = pad_frames or T(L).
    ``row_stride`` > dim widens the rows (zero filled): [B, rows, row_stride].
    ``cmvn``: False (the reference: none), ``"mean"`` or True (mean and variance) — per clip and column
    over the clip's real frames (BASELINE north_star option; applied to all ``dim`` columns)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    x = _to_device_pcm(torch, pcm)
    if x.dim() == 1:
        x = x[None, :]
    if x.dim() != 2:
        raise ValueError("pcm must be [B, L] or [L]")
    B, L = x.shape
    T = cfg.num_frames(L)
    rows = pad_frames if pad_frames > 0 else T
    dim = cfg.numcep * (3 if with_deltas else 1)
    width = row_stride if row_stride else dim
    if width < dim:
        raise ValueError(f"row_stride {row_stride} < row width {dim}")
    if out is None:
        out = torch.empty((B, rows, width), dtype=torch.float32, device=x.device)
    elif tuple(out.shape) != (B, rows, width) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 tensor of shape {(B, rows, width)}")
    p = _c_params(cfg, with_deltas, pad_frames)
    stride0 = x.stride(0) if B > 1 else L
    _lib.check(lib.mmla_psf_mfcc_rows(x.data_ptr(), (B - 1) * stride0 + L, None, None, B, L, stride0, C.byref(p),
                                      out.data_ptr(), rows * width, width, _lib.stream_ptr(torch)), "mmla_psf_mfcc_rows")
    if cmvn:
        if cmvn not in (True, "mean", "mvn"):
            raise ValueError("cmvn must be False, True / 'mvn' or 'mean'")
        _lib.check(lib.mmla_cmvn(out.data_ptr(), B, rows * width, width, dim, min(T, rows), None,
                                 0 if cmvn == "mean" else 1, _lib.stream_ptr(torch)), "mmla_cmvn")
    return out


def mfcc_ragged(pcm_flat, clip_off: Sequence[int], clip_len: Sequence[int], cfg: MfccConfig = MfccConfig(),
                with_deltas: bool = False, pad_frames: int = 0):
    """Ragged clips packed in one int16 buffer.  Returns (out_flat float32 CUDA, row_offsets)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    x = _to_device_pcm(torch, pcm_flat).reshape(-1)
    n = len(clip_off)
    off = np.ascontiguousarray(clip_off, dtype=np.int64)
    ln = np.ascontiguousarray(clip_len, dtype=np.int32)
    dim = cfg.numcep * (3 if with_deltas else 1)
    rows = [pad_frames if pad_frames > 0 else cfg.num_frames(int(l)) for l in ln]
    max_rows = max(rows) if rows else 0
    out = torch.empty((n, max_rows, dim), dtype=torch.float32, device=x.device)
    p = _c_params(cfg, with_deltas, pad_frames)
    _lib.check(lib.mmla_psf_mfcc(x.data_ptr(), x.numel(), off.ctypes.data, ln.ctypes.data, n, 0, 0,
                                 C.byref(p), out.data_ptr(), max_rows * dim, _lib.stream_ptr(torch)),
               "mmla_psf_mfcc")
    return out, rows


def speaker_features_batch(pcm, cfg: MfccConfig = MfccConfig(), out=None, row_stride: int = 0):
    """[B, L] int16 → float32 CUDA [B, 256, 39]: MFCC ‖ Δ ‖ ΔΔ padded / truncated to 256 rows —
    the batched equivalent of ``input_feature_gen`` (speaker_identification.py:386-395).
    ``row_stride=40`` gives the channel-padded [B, 256, 40] layout (column 39 = 0) the tensor-core
    classifier stem consumes without a pad pass."""
    return mfcc_batch(pcm, cfg, with_deltas=True, pad_frames=SPEAKER_FRAMES, out=out, row_stride=row_stride)


# ---------------------------------------------------------------------------------------------
# reference signatures
# ---------------------------------------------------------------------------------------------
def mfcc(signal, samplerate=16000, winlen=0.025, winstep=0.01, numcep=13, nfilt=26, nfft=None,
         lowfreq=0, highfreq=None, preemph=0.97, ceplifter=22, appendEnergy=True, winfunc=None):
    """python_speech_features.mfcc keyword signature.  ``winfunc``: None / 'rect' (psf default,
    what the reference uses), 'hann' or 'hamming'.  Returns float64 [T, numcep]."""
    if nfft is None:
        nfft = 1
        while nfft < winlen * samplerate:
            nfft *= 2
    window = "rect" if winfunc is None else winfunc
    if window not in WINDOW_IDS:
        raise ValueError("winfunc must be None, 'rect', 'hann' or 'hamming' (device windows)")
    cfg = MfccConfig(samplerate=samplerate, winlen=winlen, winstep=winstep, numcep=numcep, nfilt=nfilt,
                     nfft=nfft, lowfreq=lowfreq, highfreq=highfreq, preemph=preemph,
                     ceplifter=ceplifter, appendEnergy=appendEnergy, window=window)
    sig = np.asarray(signal)
    if sig.ndim != 1:
        raise ValueError("mfcc expects a mono signal")
    if sig.dtype != np.int16:
        raise TypeError("signal must be int16 at int16 scale (scipy.io.wavfile.read output)")
    out = mfcc_batch(sig, cfg)
    return out[0].cpu().numpy().astype(np.float64)


def delta(feat, N):
    """Reference ``delta`` (speaker_identification.py:141-151) on the device."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    a = np.asarray(feat)
    if a.ndim != 2:
        raise ValueError("delta expects [frames, dim]")
    x = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    y = torch.empty_like(x)
    _lib.check(lib.mmla_delta(x.data_ptr(), x.shape[0], x.shape[1], int(N), y.data_ptr(),
                              _lib.stream_ptr(torch)), "mmla_delta")
    return y.cpu().numpy().astype(a.dtype if a.dtype.kind == "f" else np.float64)


def input_feature_gen(wav_path):
    """``'silent'`` if the clip has < 4000 samples, else float64 ``[1, 256, 39]``
    (speaker_identification.py:372-398).  Accepts a WAV path or an int16 array."""
    _, sig = as_int16_signal(wav_path)
    sig = np.asarray(sig)
    if len(sig) < SILENT_MIN_SAMPLES:
        return "silent"
    out = speaker_features_batch(sig)
    return out.cpu().numpy().astype(np.float64)


def whole_file_chunks(sig, cfg: MfccConfig = MfccConfig()):
    """MFCC-39 over a whole recording, zero-padded to a multiple of 256 frames and viewed as
    chunks — speaker_identification.py:341-353 and
    speaker_identification_post_processing.py:255-269.  Returns float32 CUDA [chunks,256,39]."""
    torch = _lib.require_cuda()
    x = _to_device_pcm(torch, sig).reshape(1, -1)
    T = cfg.num_frames(x.shape[1])
    segs = max(1, math.ceil(T / SPEAKER_FRAMES))
    out = mfcc_batch(x, cfg, with_deltas=True, pad_frames=segs * SPEAKER_FRAMES)
    return out.view(segs, SPEAKER_FRAMES, 3 * cfg.numcep)


def binarizer(str_list, dim, speakers_count_dict=None):
    """One-hot labels in order of first appearance (speaker_identification.py:122-138)."""
    d = {} if speakers_count_dict is None else speakers_count_dict
    rows = np.zeros((len(str_list), dim))
    count = 0
    for i, s in enumerate(str_list):
        if s not in d:
            d[s] = count
            count += 1
        rows[i, d[s]] = 1
    return rows


def make_feature_experiment(wav_files):
    """(x [M,256,39], y one-hot [M,n], {str(idx): name}) — speaker_identification.py:317-369.
    ``wav_files``: WAV paths, or (label, int16 array) pairs for in-memory corpora."""
    train_x, train_y = [], []
    for item in wav_files:
        if isinstance(item, (tuple, list)):
            label, sig = item
        else:
            label = os.path.basename(os.fspath(item))[:-4]
            _, sig = as_int16_signal(item)
        chunks = whole_file_chunks(np.asarray(sig)).cpu().numpy().astype(np.float64)
        for i in range(chunks.shape[0]):
            train_x.append(chunks[i])
            train_y.append(label)
    dimension = len(set(train_y))
    yy = binarizer(train_y, dim=dimension)
    x = np.asarray(train_x)
    speaker_id = {str(int(np.argmax(yy[i]))): train_y[i] for i in range(len(train_y))}
    return x, yy, speaker_id
