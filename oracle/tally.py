"""Oracle: the counting part of the reference's ``visualization()`` functions and the TSV log
rows they parse.  TEST INFRASTRUCTURE ONLY (see oracle/__init__).

Follows, line by line:
  * OverlapDetection/scripts/overlap_degree_distribution.py:41-65
  * SpeakerIdentification/scripts/speaker_time_distribution.py:45-86
  * log rows: OverlapDetection/scripts/record_on_pc.py:164-171,
    overlap_detection_post_processing.py:213-224 (dt = 1.5 s, first row at t0),
    speaker_identification_post_processing.py:278-312 (dt = 2.56 s, added BEFORE each row).

These are pure-Python integer/rounding operations, so this oracle is pinned by the reference
source itself (no third-party arithmetic).
"""
from __future__ import annotations

from datetime import datetime, timedelta
from typing import Dict, List, Sequence, Tuple


def log_rows(labels: Sequence[str], t0: datetime, dt_seconds: float, header: str,
             add_before_first: bool) -> List[str]:
    """Rows ``seg\\tlabel\\ttimestamp`` exactly as the offline scripts write them."""
    lines = ["segment\t" + header + "\ttimestamp"]
    t = t0
    for i, lab in enumerate(labels):
        if add_before_first or i > 0:
            t = t + timedelta(seconds=dt_seconds)
        lines.append(f"{i}\t{lab}\t{t}")
    return lines


def _parse_time(field: str) -> datetime:
    return datetime.strptime(field[:-7], "%Y-%m-%d %H:%M:%S")


def tally_from_log(lines: Sequence[str], initial_labels: Sequence[str] = ()) \
        -> Tuple[Dict[str, int], Dict[str, int], float]:
    """Returns ({label: count}, {label: seconds}, total_seconds) from log lines (header first).

    ``initial_labels`` pre-seeds zero counts in dict order (overlap script, :32-34); the speaker
    script discovers labels in order of first appearance (:59-63)."""
    n = len(lines)
    start = _parse_time(lines[1].strip().split("\t")[2])
    end = _parse_time(lines[n - 1].strip().split("\t")[2])
    total_seconds = (end - start).total_seconds()
    dist: Dict[str, int] = {lab: 0 for lab in initial_labels}
    for i in range(1, n):
        lab = lines[i].strip().split("\t")[1]
        if lab not in dist:
            dist[lab] = 0
        dist[lab] += 1
    counts = list(dist.values())
    norm = [round(float(c) / sum(counts), 4) for c in counts]
    seconds = [int(x * total_seconds) for x in norm]
    return dist, dict(zip(dist.keys(), seconds)), total_seconds


def num_windows(n_samples: int, win: int, step: int) -> int:
    """``segmentation`` window count (overlap_detection_post_processing.py:55-59)."""
    return int(((n_samples - win) / step) + 1)
