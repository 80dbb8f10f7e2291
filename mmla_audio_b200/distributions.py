"""File-based tallies: the counting half of the reference's ``visualization()`` functions.

  * ``overlap_degree_distribution.visualization()``  OverlapDetection/scripts/overlap_degree_distribution.py:14-65
  * ``speaker_time_distribution.visualization()``    SpeakerIdentification/scripts/speaker_time_distribution.py:16-86

Both walk ``experiment/logs/*``, parse the TSV rows the recording / post-processing scripts
append (``segment \\t label \\t timestamp``), count rows per label, and turn the counts into
seconds with the reference's own expressions (``round(c/sum, 4)``, ``int(frac*total_seconds)``,
timestamps truncated by ``[:-7]``).  The pyecharts HTML rendering that follows in the reference
is out of scope; instead every log gets a ``<log>.tally.json`` next to it (or
in ``out_dir``) holding exactly the series the charts are built from: labels, counts, normalised
shares, seconds, the x axis of elapsed times and the per-label 0/1 (speaker) or 1/None (overlap)
bars.  Counting runs on the device (``mmla_tally``) over the label ids.
"""
from __future__ import annotations

import json
import os
from datetime import datetime
from typing import Dict, List, Optional, Sequence

from . import tally

OVERLAP_DEGREE_DICT = {"0": "non-overlapped", "1": "overlapped", "2": "silent"}   # overlap_degree_distribution.py:11


def write_log(path: str, lines: Sequence[str]) -> None:
    """Write the TSV log the reference scripts append row by row (``tally.log_rows`` builds the lines)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        for line in lines:
            f.write(line)
            f.write("\n")


def _parse_time(field: str) -> datetime:
    # the reference drops the ".ffffff" tail by position, so a timestamp without microseconds fails there too
    return datetime.strptime(field[:-7], "%Y-%m-%d %H:%M:%S")


def _count_on_device(ids: List[int], n_classes: int) -> List[int]:
    from . import _lib
    torch = _lib.require_cuda()
    labels = torch.tensor(ids, dtype=torch.int32, device="cuda")
    return tally.device_counts(labels, n_classes).cpu().tolist()[:n_classes]


def tally_log_file(log_path: str, initial_labels: Optional[Sequence[str]] = None, bar_fill=0) -> Dict:
    """One log → the chart series.  ``initial_labels`` pre-seeds the label order (overlap script);
    without it labels appear in order of first occurrence (speaker script).  ``bar_fill`` is the
    value a bar holds where another label is active (``None`` in the overlap script, ``0`` in the
    speaker script)."""
    with open(log_path, "r") as f:
        lines = f.readlines()
    n = len(lines)
    if n < 2:
        raise ValueError(f"{log_path}: no rows")
    start = _parse_time(lines[1].strip().split("\t")[2])
    end = _parse_time(lines[n - 1].strip().split("\t")[2])
    total_seconds = (end - start).total_seconds()
    order: List[str] = list(initial_labels) if initial_labels is not None else []
    fixed = initial_labels is not None
    ids, x_bar, row_labels = [], [], []
    for i in range(1, n):
        parts = lines[i].strip().split("\t")
        lab = parts[1]
        if lab not in order:
            if fixed:
                raise ValueError(f"{log_path}: label {lab!r} is not one of {order}")   # val_list.index() raises there too
            order.append(lab)
        ids.append(order.index(lab))
        row_labels.append(lab)
        x_bar.append(str(_parse_time(parts[2]) - start))
    counts = _count_on_device(ids, len(order))
    total = sum(counts)
    norm = [round(float(c) / total, 4) for c in counts]
    seconds = [int(x * total_seconds) for x in norm]
    bars = {}
    first_seen = {}
    for r, lab in enumerate(row_labels):
        first_seen.setdefault(lab, r)
    for lab in order:
        if fixed:
            bars[lab] = [1 if rl == lab else bar_fill for rl in row_labels]
        else:   # the speaker script pads a newly seen speaker's bar with None up to its first row
            f0 = first_seen[lab]
            bars[lab] = [None] * f0 + [1 if rl == lab else 0 for rl in row_labels[f0:]]
    return {"log": os.path.basename(log_path), "labels": order, "counts": counts, "norm": norm,
            "seconds": seconds, "total_seconds": total_seconds, "x_bar": x_bar, "bars": bars}


def visualization(log_dir: str, initial_labels: Optional[Sequence[str]], bar_fill, out_dir: Optional[str] = None) -> Dict[str, Dict]:
    results = {}
    for log_file in os.listdir(log_dir):
        path = os.path.join(log_dir, log_file)
        if not os.path.isfile(path) or log_file.endswith(".tally.json"):
            continue
        res = tally_log_file(path, initial_labels, bar_fill)
        results[log_file] = res
        dst = os.path.join(out_dir or log_dir, log_file + ".tally.json")
        os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
        with open(dst, "w") as f:
            json.dump(res, f, indent=1)
    return results
