"""Offline speaker-identification session driver —
SpeakerIdentification/scripts/speaker_identification_post_processing.py.

``post_analysing()`` (:191-312): the registered speakers come from ``experiment/corpus`` (:193-197); for every
segmented recording the per-segment VAD pass finds the silent segments (:221-251; the detector state carries from
segment to segment, the reference uses one module-global ``Vad``), the WHOLE standardised file goes through
MFCC + delta + delta-delta in one call and is chunked by 256 frames (:255-269), one ``predict`` labels all chunks
(:272), and the TSV log is written with 2.56 s steps (:275-312).  ``std.visualization()`` tallies the logs."""
from __future__ import annotations

import os
from datetime import datetime
from typing import Dict, List, Optional

import numpy as np

from . import speaker_time_distribution as std
from .audio_io import read_wav_int16
from .offline_common import read_wave_file, segment_index, segmentation, standardize_audio  # noqa: F401 (re-exported)
from .params import SILENT_MIN_SAMPLES
from .speaker_identification import delta  # noqa: F401  (the reference keeps its own copy, :32-42)

Root_Dir = os.getcwd()


def silent_segments(segment_paths: List[str]) -> List[int]:
    """Indices of the segments whose VAD-trimmed audio has fewer than 4000 samples (:225-251)."""
    from .vad import vad_trim
    clips = [read_wav_int16(p)[1] for p in segment_paths]
    if not clips:
        return []
    L = max(len(c) for c in clips)
    batch = np.zeros((len(clips), L), np.int16)
    for i, c in enumerate(clips):
        batch[i, :len(c)] = c
    res = vad_trim(batch, lengths=np.asarray([len(c) for c in clips], np.int32), clips_per_stream=len(clips), compact=False)
    voiced = res.voiced_len.cpu().numpy()
    return [i for i, n in enumerate(voiced) if n < SILENT_MIN_SAMPLES]


def post_analysing(root_dir: Optional[str] = None, model=None, precision: str = "tf32",
                   t0: Optional[datetime] = None) -> Dict[str, List[str]]:
    """Returns {log path: rows}.  ``model``: default ``load_model(Root_Dir/experiment/model)`` (:205-206)."""
    from .models import load_model
    from .pipeline import SpeakerPipeline
    root = root_dir or Root_Dir
    files = os.listdir(os.path.join(root, "experiment/corpus"))
    speaker_id_dict = {str(i): files[i][:-4] for i in range(len(files))}
    if model is None:
        model = load_model(os.path.join(root, "experiment/model"), kind="speaker", precision=precision)
    pipe = SpeakerPipeline(model)
    logs = {}
    seg_root = os.path.join(root, "experiment/recordings/post-time/segments")
    for directory_name in os.listdir(seg_root):
        seg_dir = os.path.join(seg_root, directory_name)
        whole_wav_path = os.path.join(root, "experiment/recordings/post-time/standardized", directory_name + ".wav")
        log_path = os.path.join(root, "experiment/logs", directory_name + ".txt")
        seg_paths = sorted((os.path.join(seg_dir, f) for f in os.listdir(seg_dir)), key=segment_index)
        silent_index = silent_segments(seg_paths)
        _rate, sig = read_wav_int16(whole_wav_path)
        # the file has ceil(T/256) chunks; a silent segment index past the last chunk is ignored, as `if i in silent_index` is
        os.makedirs(os.path.dirname(log_path), exist_ok=True)
        n_chunks_probe = -(-pipe.cfg.num_frames(len(sig)) // 256)
        labels, _ = pipe.run_session(sig, speaker_id_dict, t0=t0 or datetime.today(),
                                     silent_index=[i for i in silent_index if i < n_chunks_probe], log_path=log_path)
        with open(log_path) as f:
            logs[log_path] = f.read().splitlines()
    return logs


if __name__ == "__main__":
    post_analysing()
    std.Root_Dir = Root_Dir
    std.visualization()
