"""PCM ingest: WAV paths (stdlib ``wave``; the reference uses scipy.io.wavfile / librosa.load)
or in-memory int16 buffers → int16 arrays.  Host-side only; no arithmetic on samples."""
from __future__ import annotations

import os
import wave

import numpy as np


def read_wav_int16(path: str):
    """(rate, int16 mono array) like ``scipy.io.wavfile.read`` for the reference's files
    (mono, 16-bit — the format guards of ``read_wave_file``,
    OverlapDetection/scripts/record_on_pc.py:188-197)."""
    with wave.open(path, "rb") as wf:
        if wf.getnchannels() != 1:
            raise ValueError(f"{path}: expected mono audio")
        if wf.getsampwidth() != 2:
            raise ValueError(f"{path}: expected 16-bit PCM")
        rate = wf.getframerate()
        data = wf.readframes(wf.getnframes())
    return rate, np.frombuffer(data, dtype="<i2").astype(np.int16, copy=False)


def write_wav_int16(path: str, sig, rate: int = 16000) -> None:
    with wave.open(path, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(rate)
        wf.writeframes(np.asarray(sig, dtype="<i2").tobytes())


def as_int16_signal(x):
    """Accept a WAV path, a numpy array or a torch tensor; return (rate|None, array-like int16)."""
    if isinstance(x, (str, os.PathLike)):
        return read_wav_int16(os.fspath(x))
    return None, x
