"""Oracle twin of the on-device synthetic PCM generator (mmla_audio_b200/csrc/synth.cu).
TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Not a reference function: the reference records from a microphone
(OverlapDetection/scripts/record_on_pc.py:115-124); BASELINE.json's north_star replaces that
with synthetic buffers.  The generator is integer-only so that CPU and GPU produce
bit-identical int16 PCM for any (seed, clip index, sample index):

  clip = sum over 1..3 "speakers" of an 8-harmonic stack (f0 85..255 Hz, 1/h roll-off) with a
  3..6 Hz raised-sine amplitude envelope and random onset/offset, plus +-512 uniform noise.
"""
from __future__ import annotations

import numpy as np

SEED = 0x6D6D6C61
N_HARM = 8
INC_PER_HZ = 268435          # floor(2^32 / 16000)


def sine_table() -> np.ndarray:
    """int16[1024] = round(32767*sin(2*pi*i/1024)); the product builds the identical table."""
    i = np.arange(1024, dtype=np.float64)
    return np.round(32767.0 * np.sin(2.0 * np.pi * i / 1024.0)).astype(np.int16)


def _hash32(x):
    x = np.asarray(x, dtype=np.uint64) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def _clip_key(seed: int, clip):
    clip = np.asarray(clip, dtype=np.uint64)
    lo = clip & np.uint64(0xFFFFFFFF)
    hi = clip >> np.uint64(32)
    k = _hash32((np.uint64(seed) + lo * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF))
    return _hash32(k ^ ((hi * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)))


def _param(key, slot: int):
    return _hash32((key + np.uint64((slot * 0x9E3779B9) & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF))


def synth_clips(first_clip: int, n_clips: int, clip_len: int, seed: int = SEED) -> np.ndarray:
    """int16 [n_clips, clip_len]; clip i is a pure function of (seed, first_clip+i)."""
    tab = sine_table().astype(np.int64)
    clips = np.arange(first_clip, first_clip + n_clips, dtype=np.uint64)
    key = _clip_key(seed, clips)[:, None]                       # [B,1]
    n = np.arange(clip_len, dtype=np.uint64)[None, :]           # [1,L]
    nspk = (1 + (_param(key, 0) % np.uint64(3))).astype(np.int64)
    acc = np.zeros((n_clips, clip_len), dtype=np.int64)
    half = np.uint64(max(clip_len // 2, 1))
    for s in range(3):
        b = 1 + 6 * s
        f0 = np.uint64(85) + _param(key, b + 0) % np.uint64(171)
        inc = (f0 * np.uint64(INC_PER_HZ)) & np.uint64(0xFFFFFFFF)
        am_inc = ((np.uint64(3) + _param(key, b + 1) % np.uint64(4)) * np.uint64(INC_PER_HZ))
        am_ph0 = _param(key, b + 2)
        onset = _param(key, b + 3) % half
        offset = half + _param(key, b + 4) % half
        if s == 0:
            onset = onset // np.uint64(4)
        amp = (np.uint64(4000) + _param(key, b + 5) % np.uint64(6000)).astype(np.int64)
        ph = (inc * n) & np.uint64(0xFFFFFFFF)
        hsum = np.zeros((n_clips, clip_len), dtype=np.int64)
        for h in range(1, N_HARM + 1):
            idx = (((ph * np.uint64(h)) & np.uint64(0xFFFFFFFF)) >> np.uint64(22)).astype(np.int64)
            hsum += (tab[idx] * (32768 // h)) >> 15
        am_ph = (am_ph0 + am_inc * n) & np.uint64(0xFFFFFFFF)
        env = (tab[(am_ph >> np.uint64(22)).astype(np.int64)] + 32768) >> 1
        v = (hsum * env) >> 15
        v = (v * amp) >> 17
        active = (n >= onset) & (n < offset) & (s < nspk)
        acc += np.where(active, v, 0)
    noise_key = _param(key, 31)
    noise = (_hash32((noise_key + n * np.uint64(0x9E3779B9)) & np.uint64(0xFFFFFFFF))
             & np.uint64(0x3FF)).astype(np.int64) - 512
    acc += noise
    return np.clip(acc, -32768, 32767).astype(np.int16)
