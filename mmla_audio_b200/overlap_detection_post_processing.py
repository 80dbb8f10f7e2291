"""Offline overlap-detection session driver — OverlapDetection/scripts/overlap_detection_post_processing.py.

``post_anlysing()`` (sic, :151-226) with the per-segment loop batched: every conversation under
``experiment/recordings/post-time/whole`` is standardised, cut into 1.5 s segments (``segmentation``, :23-85), all
segments of a conversation go through ``OverlapPipeline`` in one batch (features + classifier on the device), and
the TSV log ``experiment/logs/<name>.txt`` is written row for row as :213-224 does.  ``odd.visualization()`` then
tallies the logs (``overlap_degree_distribution``)."""
from __future__ import annotations

import os
from datetime import datetime
from typing import Dict, List, Optional

import numpy as np

from . import tally
from . import overlap_degree_distribution as odd
from .audio_io import read_wav_int16
from .distributions import write_log
from .offline_common import read_wave_file, segment_index, segmentation, standardize_audio  # noqa: F401 (re-exported)
from .overlap_features_generator import OverlapFeaturesGenerator

Root_Dir = os.getcwd()
overlap_degree_dict = {"0": "non-overlapped", "1": "overlapped"}           # :18


def post_anlysing(root_dir: Optional[str] = None, model=None, precision: str = "tf32", write_png: bool = False,
                  t0: Optional[datetime] = None, silence_removed: bool = False) -> Dict[str, List[str]]:
    """Returns {log path: rows}.  ``model``: a loaded ``models.Model`` (default: ``load_model(Root_Dir/timit/models/
    timit2.0)``, :153-154).  ``write_png=True`` also writes every segment's feature image as the reference does (:203).
    Segments are processed in temporal order (the reference iterates ``os.listdir`` unsorted, :199; the tallies do not
    depend on the order).  ``silence_removed`` applies the VAD + 4000-sample rule per segment."""
    from .models import load_model
    from .pipeline import OverlapPipeline
    import torch
    root = root_dir or Root_Dir
    if model is None:
        model = load_model(os.path.join(root, "timit/models/timit2.0"), kind="overlap", precision=precision)
    pipe = OverlapPipeline(model)
    whole = os.path.join(root, "experiment/recordings/post-time/whole")
    std_dir = os.path.join(root, "experiment/recordings/post-time/standardized")
    seg_root = os.path.join(root, "experiment/recordings/post-time/segments")
    feat_root = os.path.join(root, "experiment/recordings/post-time/features")
    names = []
    for dirpath, _dirnames, filenames in os.walk(whole):
        for filename in filenames:
            src = os.sep.join([dirpath, filename])
            name = filename[:-4]
            dst = os.path.join(std_dir, name + ".wav")
            noise = os.path.join(root, "experiment/Ambient_Noise.wav")
            if filename.startswith("zoom"):
                standardize_audio(src, dst, dbfs=0, noise_reduced=0, silence_remove=False)
            elif filename.startswith("audio"):
                standardize_audio(src, dst, dbfs=0, noise_reduced=3, silence_remove=False, noise_path=noise)
            else:
                continue                                       # the reference standardises only these two prefixes (:181-187)
            names.append(name)
    os.makedirs(seg_root, exist_ok=True)
    segmentation(std_dir, seg_root, 1.5, 1.5)
    logs = {}
    for name in names:
        seg_dir = os.path.join(seg_root, name)
        files = sorted((os.path.join(seg_dir, f) for f in os.listdir(seg_dir)), key=segment_index) if os.path.isdir(seg_dir) else []
        if not files:
            continue
        clips = np.stack([read_wav_int16(f)[1] for f in files])              # equal-length windows
        labels, _prob = pipe.run_device(torch.from_numpy(clips).cuda(), silence_removed=silence_removed)
        if write_png:
            feat_dir = os.path.join(feat_root, name) + "/"
            os.makedirs(feat_dir, exist_ok=True)
            ofg = OverlapFeaturesGenerator(wl=25, hl=10)
            for count, f in enumerate(files):
                ofg.generate_zcr_image(f, feat_dir, str(count) + ".png")
        rows = [overlap_degree_dict.get(str(int(l)), "silent") for l in labels.cpu().tolist()]
        lines = tally.log_rows(rows, t0 or datetime.today(), 1.5, "overlapped degree", add_before_first=False)
        log_path = os.path.join(root, "experiment/logs", name + ".txt")
        write_log(log_path, lines)
        logs[log_path] = lines
    return logs


if __name__ == "__main__":
    post_anlysing()
    odd.Root_Dir = Root_Dir
    odd.visualization()
