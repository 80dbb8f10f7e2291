"""Label → log row → tally, the last leg of the hot path.

Mirrors (a) the TSV rows the reference appends per clip
(OverlapDetection/scripts/record_on_pc.py:164-171, overlap_detection_post_processing.py:213-224,
speaker_identification_post_processing.py:278-312) and (b) the counting done by
``visualization()`` (overlap_degree_distribution.py:41-65, speaker_time_distribution.py:45-86).
Counting runs on the device (``mmla_tally``); the final ``round(c/sum, 4)`` and
``int(frac * total_seconds)`` are the reference's own Python expressions.
"""
from __future__ import annotations

from datetime import datetime, timedelta
from typing import Dict, List, Optional, Sequence, Tuple

from . import _lib

OVERLAP_DEGREE_DICT = {"0": "non-overlapped", "1": "overlapped"}   # record_on_pc.py:34
SILENT = -1                                                          # label id of the 'silent' sentinel


def normalize_names(id_to_name) -> Dict[int, str]:
    """{int id: name} from a dict keyed by ints or by the reference's ``str(idx)`` keys."""
    return {int(k): v for k, v in (id_to_name or {}).items()}


def device_counts(labels, n_classes: int):
    """int64 CUDA tensor [n_classes+1]; the last bin collects out-of-range ids (e.g. SILENT)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    labels = labels.contiguous()
    if labels.dtype != torch.int32:
        labels = labels.to(torch.int32)
    counts = torch.zeros(n_classes + 1, dtype=torch.int64, device=labels.device)
    _lib.check(lib.mmla_tally(labels.data_ptr(), labels.numel(), n_classes, counts.data_ptr(),
                              _lib.stream_ptr(torch)), "mmla_tally")
    return counts


def seconds_from_counts(counts: Sequence[int], total_seconds: float) -> Tuple[List[float], List[int]]:
    """``norm = round(c/sum, 4)``; ``seconds = int(norm * total_seconds)``
    (overlap_degree_distribution.py:63-65)."""
    total = sum(counts)
    norm = [round(float(c) / total, 4) for c in counts]
    return norm, [int(x * total_seconds) for x in norm]


def session_total_seconds(n_rows: int, t0: datetime, dt_seconds: float, add_before_first: bool) -> float:
    """``t_last - t_first`` after the reference's ``[:-7]`` truncation to whole seconds."""
    def trunc(t: datetime) -> datetime:      # the reference's str(t)[:-7] (drops ".ffffff")
        return t.replace(microsecond=0)
    first = t0 + timedelta(seconds=dt_seconds) if add_before_first else t0
    last = first + timedelta(seconds=dt_seconds * (n_rows - 1))
    # repeated float additions in the reference accumulate exactly like this only approximately;
    # log_rows() below reproduces them step by step when exactness matters.
    return (trunc(last) - trunc(first)).total_seconds()


def log_rows(label_names: Sequence[str], t0: datetime, dt_seconds: float, header: str,
             add_before_first: bool) -> List[str]:
    """The TSV lines of an offline session log (header first)."""
    lines = ["segment\t" + header + "\ttimestamp"]
    t = t0
    for i, lab in enumerate(label_names):
        if add_before_first or i > 0:
            t = t + timedelta(seconds=dt_seconds)
        lines.append(f"{i}\t{lab}\t{t}")
    return lines


def tally_session(labels, id_to_name: Dict[int, str], t0: datetime, dt_seconds: float,
                  add_before_first: bool, initial_order: Optional[Sequence[str]] = None):
    """labels: int32 CUDA tensor [N] (SILENT = -1).  Returns ({name: count}, {name: seconds},
    total_seconds) with dict order as the reference builds it: ``initial_order`` first (overlap
    script) or order of first appearance (speaker script)."""
    torch = _lib.require_cuda()
    # the reference's own dicts are keyed by str(idx) (speaker_identification.py:360-369 `speaker_id`,
    # record_on_pc.py:34 `overlap_degree_dict`): accept those as well as int keys
    names = normalize_names(id_to_name)
    n_classes = max(names) + 1 if names else 1
    counts = device_counts(labels, n_classes).cpu().tolist()
    host = labels.cpu().tolist()
    order: List[str] = list(initial_order) if initial_order else []
    seen = set(order)
    for l in host:                       # order of first appearance, as the speaker script does
        nm = names.get(l, "silent")
        if nm not in seen:
            seen.add(nm)
            order.append(nm)
    by_name = {nm: 0 for nm in order}
    for cid, c in enumerate(counts[:-1]):
        if c:
            nm = names.get(cid, "silent")          # an id nobody registered is logged like the sentinel
            by_name[nm] = by_name.get(nm, 0) + c
    if counts[-1]:
        by_name["silent"] = by_name.get("silent", 0) + counts[-1]
    n = len(host)
    first = t0 + timedelta(seconds=dt_seconds) if add_before_first else t0
    t = first
    for _ in range(n - 1):
        t = t + timedelta(seconds=dt_seconds)
    trunc = lambda x: x.replace(microsecond=0)      # the reference's str(t)[:-7]
    total_seconds = (trunc(t) - trunc(first)).total_seconds()
    _, secs = seconds_from_counts(list(by_name.values()), total_seconds)
    return by_name, dict(zip(by_name.keys(), secs)), total_seconds
