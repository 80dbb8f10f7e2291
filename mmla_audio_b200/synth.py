"""Synthetic PCM on the device (replaces PyAudio capture per BASELINE.json's north_star).
Integer-only generator, bit-identical to the CPU twin used by the parity tests."""
from __future__ import annotations

import numpy as np

from . import _lib

SEED = 0x6D6D6C61


def sine_table() -> np.ndarray:
    i = np.arange(1024, dtype=np.float64)
    return np.round(32767.0 * np.sin(2.0 * np.pi * i / 1024.0)).astype(np.int16)


_table_cache = {}


def synth_clips(first_clip: int, n_clips: int, clip_len: int, seed: int = SEED, out=None, clip_stride=None):
    """int16 CUDA tensor [n_clips, clip_len] (or fills ``out`` with row stride ``clip_stride``)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.cuda.current_device()
    if dev not in _table_cache:
        _table_cache[dev] = torch.from_numpy(sine_table()).cuda()
    if out is None:
        clip_stride = clip_len if clip_stride is None else clip_stride
        out = torch.empty((n_clips, clip_stride), dtype=torch.int16, device="cuda")
    elif clip_stride is None:
        clip_stride = out.stride(0)
    _lib.check(lib.mmla_synth_pcm(out.data_ptr(), first_clip, n_clips, clip_len, clip_stride,
                                  seed & 0xFFFFFFFF, _table_cache[dev].data_ptr(), _lib.stream_ptr(torch)),
               "mmla_synth_pcm")
    return out if out.shape[1] == clip_len else out[:, :clip_len]
