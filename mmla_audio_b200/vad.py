"""Silence removal on the device — the ``silence_remove`` branch of the reference's ``save_wave_file``
(OverlapDetection/scripts/record_on_pc.py:214-226): ``frame_generator(30, ...)`` → ``webrtcvad.Vad(3).is_speech`` →
``vad_collector(.., 30, 300, ..)`` (:229-295) → the clip rewritten as the concatenation of the yielded segments.
Same code in SpeakerIdentification/scripts/record_on_pc.py:207-273 and both ``*_post_processing.py`` files.

``vad_trim`` works on batches of clips resident in HBM through ``mmla_vad_trim``; ``frame_generator`` /
``vad_collector`` / ``Vad`` keep the reference's call signatures for single clips (bytes in, bytes out).
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np

from . import _lib
from .params import SILENT_MIN_SAMPLES

FRAME_SAMPLES = 480            # 30 ms at 16 kHz


class VadResult(NamedTuple):
    pcm: object                # int16 CUDA [B, L]: kept frames of every clip, concatenated from sample 0
    voiced_len: object         # int32 CUDA [B]: rewritten clip length (480 * kept frames)
    speech: object             # uint8 CUDA [B, max_frames]: is_speech per frame
    keep: object               # uint8 CUDA [B, max_frames]: frames vad_collector yields


def num_frames(n_samples: int) -> int:
    """Frames ``frame_generator(30, audio, 16000)`` yields (record_on_pc.py:239: ``while offset + n < len(audio)``)."""
    return int(_lib.load().mmla_vad_num_frames(int(n_samples)))


def vad_trim(pcm, lengths=None, clips_per_stream: int = 1, compact: bool = True) -> VadResult:
    """pcm: int16 [B, L] (numpy / torch; CUDA tensors are used in place).  ``lengths``: optional int32 [B] samples per
    clip.  ``clips_per_stream``: 1 = every clip starts from a fresh ``Vad(3)`` (independent clips, one GPU thread per
    clip); B = the clips are one session in temporal order and the detector state carries across them, exactly as the
    reference's module-global ``vad`` object does (one thread: sequential)."""
    from .speaker_identification import _to_device_pcm
    torch = _lib.require_cuda()
    lib = _lib.load()
    x = _to_device_pcm(torch, pcm)
    if x.dim() == 1:
        x = x[None, :]
    B, L = x.shape
    stride0 = x.stride(0) if B > 1 else L
    if stride0 % 4 or x.data_ptr() % 8:                      # clips must start on 8-byte boundaries: padded copy
        xp = torch.zeros((B, (L + 7) & ~7), dtype=torch.int16, device=x.device)
        xp[:, :L] = x
        x, stride0 = xp, xp.stride(0)
    max_frames = max(1, num_frames(L))
    len_ptr = None
    if lengths is not None:
        lengths = torch.as_tensor(lengths, dtype=torch.int32).to(x.device).contiguous()
        if lengths.numel() != B:
            raise ValueError("lengths must have one entry per clip")
        len_ptr = lengths.data_ptr()
    speech = torch.empty((B, max_frames), dtype=torch.uint8, device=x.device)
    keep = torch.empty((B, max_frames), dtype=torch.uint8, device=x.device)
    voiced = torch.empty((B,), dtype=torch.int32, device=x.device)
    out = None
    out_stride = 0
    if compact:
        out_stride = (max_frames * FRAME_SAMPLES + 7) & ~7
        out = torch.zeros((B, out_stride), dtype=torch.int16, device=x.device)
    _lib.check(lib.mmla_vad_trim(x.data_ptr(), B, L, stride0, len_ptr, int(clips_per_stream), speech.data_ptr(),
                                 keep.data_ptr(), max_frames, voiced.data_ptr(), out.data_ptr() if compact else None,
                                 out_stride, _lib.stream_ptr(torch)), "mmla_vad_trim")
    return VadResult(out, voiced, speech, keep)


# ---------------------------------------------------------------------------------------------
# reference signatures (single clip, bytes)
# ---------------------------------------------------------------------------------------------
class Frame(object):
    """record_on_pc.py:37-43."""

    def __init__(self, bytes, timestamp, duration):
        self.bytes = bytes
        self.timestamp = timestamp
        self.duration = duration


def frame_generator(frame_duration_ms, audio, sample_rate):
    """record_on_pc.py:229-244 (host-side slicing of a bytes object; yields ``Frame``)."""
    n = int(sample_rate * (frame_duration_ms / 1000.0) * 2)
    offset = 0
    timestamp = 0.0
    duration = (float(n) / sample_rate) / 2.0
    while offset + n < len(audio):
        yield Frame(audio[offset:offset + n], timestamp, duration)
        timestamp += duration
        offset += n


class Vad:
    """``webrtcvad.Vad(3)`` stand-in whose work happens in :func:`vad_collector` (the device evaluates all frames of a
    clip in one launch).  ``carry_state=True`` chains clips like the reference's module-global object: the clips seen
    so far are replayed as one stream (sequential on the device)."""

    def __init__(self, mode: int = 3, carry_state: bool = False):
        if mode != 3:
            raise _lib.MmlaError("only aggressiveness 3 is built (the reference uses webrtcvad.Vad(3))")
        self.carry_state = carry_state
        self._history = []


def vad_collector(sample_rate, frame_duration_ms, padding_duration_ms, vad: Optional[Vad], frames):
    """record_on_pc.py:247-295: yields the voiced segments (bytes) of the clip made of ``frames``."""
    if sample_rate != 16000 or frame_duration_ms != 30 or padding_duration_ms != 300:
        raise _lib.MmlaError("vad_collector is built for (16000, 30, 300), the reference's only call")
    frames = list(frames)
    if not frames:
        return
    clip = np.frombuffer(b"".join(f.bytes for f in frames) + b"\x00\x00", dtype=np.int16)   # +1 sample: keep the last frame
    if vad is not None and vad.carry_state and vad._history:
        L = max(len(clip), max(len(h) for h in vad._history))
        batch = np.zeros((len(vad._history) + 1, L), np.int16)
        lens = []
        for i, h in enumerate(vad._history + [clip]):
            batch[i, :len(h)] = h
            lens.append(len(h))
        res = vad_trim(batch, lengths=np.asarray(lens, np.int32), clips_per_stream=len(lens), compact=False)
        keep = res.keep[-1].cpu().numpy()
    else:
        res = vad_trim(clip, compact=False)
        keep = res.keep[0].cpu().numpy()
    if vad is not None and vad.carry_state:
        vad._history.append(clip)
    # segments = maximal runs the collector yields; the reference joins them all into one file, run boundaries are where
    # the collector left the TRIGGERED state, which the keep mask alone does not mark — yield one joined segment
    idx = np.nonzero(keep[:len(frames)])[0]
    if len(idx):
        yield b"".join(frames[i].bytes for i in idx)


def is_silent_length(n_samples: int) -> bool:
    """``len(sig) < 4000`` => 'silent' (record_on_pc.py:142; speaker_identification.py:375)."""
    return n_samples < SILENT_MIN_SAMPLES
