"""Oracle: ``webrtcvad.Vad(3)`` + the reference's ``frame_generator`` / ``vad_collector``.
TEST INFRASTRUCTURE ONLY (see oracle/__init__).  PARITY UNPINNED (webrtcvad is an un-vendored dependency).

The detector itself is the C restatement in ``oracle/webrtc_vad.c`` (built by ``make -C oracle`` into
``oracle/_build/liboracle_c.so``); this module binds it and restates, line by line, the reference's own Python:

  * ``frame_generator(30, audio, 16000)``                 OverlapDetection/scripts/record_on_pc.py:229-244
  * ``vad_collector(16000, 30, 300, vad, frames)``        record_on_pc.py:247-295
  * the rewrite of the WAV from the yielded segments and the ``len(sig) < 4000 => 'silent'`` rule
                                                          record_on_pc.py:214-226,141-154
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_c.so")
FRAME = 480                      # 30 ms at 16 kHz
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "webrtc_vad.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.vad_oracle_state_bytes.restype = C.c_int
        _lib.vad_oracle_reset.argtypes = [C.c_void_p]
        _lib.vad_oracle_is_speech.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.vad_oracle_is_speech.restype = C.c_int
        _lib.vad_oracle_num_frames.argtypes = [C.c_int]
        _lib.vad_oracle_num_frames.restype = C.c_int
        _lib.vad_oracle_clip.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.vad_oracle_clip.restype = C.c_int
    return _lib


class Vad:
    """``webrtcvad.Vad(3)`` at 16 kHz / 30 ms frames.  Stateful, like the reference's module-global object."""

    def __init__(self, mode: int = 3):
        if mode != 3:
            raise ValueError("the reference only uses Vad(3)")
        self._l = lib()
        self._state = C.create_string_buffer(self._l.vad_oracle_state_bytes())
        self._l.vad_oracle_reset(self._state)

    def reset(self):
        self._l.vad_oracle_reset(self._state)

    def is_speech(self, buf, sample_rate: int = 16000, features=None) -> bool:
        if sample_rate != 16000:
            raise ValueError("16 kHz only")
        a = np.frombuffer(buf, dtype=np.int16) if isinstance(buf, (bytes, bytearray, memoryview)) else np.ascontiguousarray(buf, np.int16)
        if a.size != FRAME:
            raise ValueError("30 ms frames (480 samples) only")
        a = np.ascontiguousarray(a)
        fp = features.ctypes.data if features is not None else None
        return bool(self._l.vad_oracle_is_speech(self._state, a.ctypes.data, fp))

    def clip_flags(self, sig) -> np.ndarray:
        """is_speech of every frame ``frame_generator`` yields for this clip (uint8 [n_frames])."""
        a = np.ascontiguousarray(sig, np.int16)
        out = np.zeros(max(1, a.size // FRAME + 1), np.uint8)
        n = self._l.vad_oracle_clip(self._state, a.ctypes.data, a.size, out.ctypes.data)
        return out[:n]


def num_frames(n_samples: int) -> int:
    """Frames ``frame_generator(30, audio, 16000)`` yields: ``while offset + n < len(audio)`` on BYTE offsets with
    n = 960, so the last frame of a clip that is an exact multiple of 480 samples is dropped (:239)."""
    n, offset, count = 960, 0, 0
    while offset + n < 2 * n_samples:
        count += 1
        offset += n
    return count


def vad_collector_mask(flags, num_padding_frames: int = 10) -> np.ndarray:
    """Which frames ``vad_collector`` ends up yielding (uint8 [n_frames]) given the per-frame decisions —
    record_on_pc.py:247-295 with the deque of maxlen ``padding_duration_ms / frame_duration_ms`` = 10."""
    ring = collections.deque(maxlen=num_padding_frames)
    triggered = False
    keep = np.zeros(len(flags), np.uint8)
    voiced = []
    for i, is_speech in enumerate(flags):
        is_speech = bool(is_speech)
        if not triggered:
            ring.append((i, is_speech))
            num_voiced = len([f for f, speech in ring if speech])
            if num_voiced > 0.9 * ring.maxlen:
                triggered = True
                for f, s in ring:
                    voiced.append(f)
                ring.clear()
        else:
            voiced.append(i)
            ring.append((i, is_speech))
            num_unvoiced = len([f for f, speech in ring if not speech])
            if num_unvoiced > 0.9 * ring.maxlen:
                triggered = False
                keep[voiced] = 1                   # yield b''.join(voiced_frames)
                ring.clear()
                voiced = []
    if voiced:
        keep[voiced] = 1
    return keep


def remove_silence(sig, vad: Vad):
    """The ``silence_remove`` branch of ``save_wave_file`` (record_on_pc.py:214-226): returns
    (trimmed int16 signal, is_speech flags, kept-frame mask).  ``vad`` keeps its state across calls."""
    a = np.ascontiguousarray(sig, np.int16)
    flags = vad.clip_flags(a)
    keep = vad_collector_mask(flags)
    frames = [a[FRAME * i: FRAME * (i + 1)] for i in np.nonzero(keep)[0]]
    out = np.concatenate(frames) if frames else np.zeros(0, np.int16)
    return out, flags, keep


def is_silent(trimmed) -> bool:
    """``len(sig) < 4000`` (record_on_pc.py:142; speaker_identification.py:375)."""
    return len(trimmed) < 4000
