"""Pieces both offline drivers share (the reference duplicates them in
OverlapDetection/scripts/overlap_detection_post_processing.py and
SpeakerIdentification/scripts/speaker_identification_post_processing.py)."""
from __future__ import annotations

import math
import os
import wave
from typing import List, Optional

import numpy as np

from . import _lib
from .audio_io import read_wav_int16, write_wav_int16


def segmentation(src_dir, dst_dir, win_time_stride, step_time) -> List[str]:
    """``segmentation(src_dir, dst_dir, win, step)`` — overlap_detection_post_processing.py:23-85,
    speaker_identification_post_processing.py:58-120.  Every ``*.wav`` under ``src_dir`` is cut into
    ``cut_num = int((nframes - win)/step + 1)`` windows ``[j*step, j*step + win)`` written as
    ``dst_dir/<name>/<name>_<j>_<framerate>_split.wav``.  Returns the written paths (the reference returns None).

    The reference joins ``src_dir + "\\\\" + f`` (a Windows path); ``os.path.join`` is used here.  Windows are
    views of the recording (``pipeline.window_view`` is the same index math on the device); only the WAV writing
    touches the samples."""
    written = []
    files = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith(".wav")]
    for filename in files:
        with wave.open(filename, "rb") as f:
            nchannels, sampwidth, framerate, nframes = f.getparams()[:4]
            data = f.readframes(nframes)
        wave_data = np.frombuffer(data, dtype=np.short)
        if nchannels > 1:
            wave_data = wave_data.reshape(-1, nchannels)
        win = int(framerate * win_time_stride)
        step = int(framerate * step_time)
        cut_num = int(((nframes - win) / step) + 1)
        name = os.path.splitext(os.path.split(filename)[-1])[0]
        save_dir = os.path.join(dst_dir, name)
        for j in range(cut_num):
            os.makedirs(save_dir, exist_ok=True)
            out_file = os.path.join(save_dir, name + "_%d_%s_split.wav" % (j, framerate))
            seg = wave_data[step * j: step * j + win]
            with wave.open(out_file, "wb") as f:
                f.setnchannels(nchannels)
                f.setsampwidth(sampwidth)
                f.setframerate(framerate)
                f.writeframes(np.ascontiguousarray(seg, dtype=np.short).tobytes())
            written.append(out_file)
    return written


def read_wave_file(filepath):
    """(bytes, sample_rate) with the reference's format guards (…post_processing.py:88-98 / :123-133)."""
    with wave.open(filepath, "rb") as wf:
        assert wf.getnchannels() == 1
        assert wf.getsampwidth() == 2
        sample_rate = wf.getframerate()
        assert sample_rate in (8000, 16000, 32000, 48000)
        data = wf.readframes(wf.getnframes())
    return data, sample_rate


def segment_index(path: str) -> int:
    """``<name>_<j>_<framerate>_split.wav`` → j (the sort key of speaker_identification_post_processing.py:219)."""
    return int(os.path.basename(path).split("_")[-3])


def apply_dbfs_gain(sig: np.ndarray, dbfs: float) -> np.ndarray:
    """pydub's ``sound.apply_gain(dbfs - sound.dBFS)`` on 16-bit mono samples (standardize_audio,
    overlap_detection_post_processing.py:122-126): dBFS = 20 log10(rms / 32768) with audioop's truncated integer
    rms; samples are multiplied by 10**(change/20), floored and clipped to int16 (audioop.mul)."""
    a = np.asarray(sig, dtype=np.int64)
    if a.size == 0:
        return np.asarray(sig, dtype=np.int16)
    rms = int(math.sqrt(float((a * a).sum()) / a.size))
    if rms == 0:
        return np.asarray(sig, dtype=np.int16)
    change = dbfs - 20.0 * math.log10(rms / 32768.0)
    factor = 10.0 ** (change / 20.0)
    return np.clip(np.floor(a.astype(np.float64) * factor), -32768, 32767).astype(np.int16)


def standardize_audio(source_path, target_path=None, format=None, dbfs=None, channels=1, sampwidth=2, sample_rate=16000,
                      noise_reduced=0, silence_remove=False, noise_path: Optional[str] = None):
    """``standardize_audio`` (overlap_detection_post_processing.py:101-148; SI :136-188) for the case the hot path
    needs: a mono 16-bit WAV already at ``sample_rate``.  The steps that act on the samples are kept — gain to
    ``dbfs``, ``noise_reduced`` passes of stationary spectral gating against ``noise_path`` (device,
    ``noise_reduction.reduce_noise``), optional silence removal (device VAD).  Container / sample-rate conversion
    (pydub / ffmpeg) is file-format conditioning outside the path: other inputs raise."""
    rate, sig = read_wav_int16(source_path)
    if rate != sample_rate:
        raise _lib.MmlaError(f"{source_path}: {rate} Hz; resampling to {sample_rate} Hz is outside the hot path "
                             "(convert the file first)")
    if not target_path:
        target_path = source_path[:-4] + ".wav"
    if dbfs is not None:
        sig = apply_dbfs_gain(sig, dbfs)
    if noise_reduced > 0:
        from .noise_reduction import reduce_noise_int16
        if noise_path is None:
            raise ValueError("noise_reduced > 0 needs noise_path (the reference's experiment/Ambient_Noise.wav)")
        _, noise = read_wav_int16(noise_path)
        for _ in range(noise_reduced):
            sig = reduce_noise_int16(noise, sig, sample_rate)
    if silence_remove:
        from .vad import vad_trim
        res = vad_trim(np.ascontiguousarray(sig))
        n = int(res.voiced_len[0].item())
        sig = res.pcm[0, :n].cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(target_path)), exist_ok=True)
    write_wav_int16(target_path, sig, sample_rate)
    return target_path
