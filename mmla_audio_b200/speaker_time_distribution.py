"""``speaker_time_distribution.visualization()`` — SpeakerIdentification/scripts/speaker_time_distribution.py:16-86.
Counts rows per speaker (labels in order of first appearance) in every ``experiment/logs/*`` file
and converts them to seconds; the chart rendering is replaced by a ``<log>.tally.json`` export."""
from __future__ import annotations

import os
from typing import Dict, Optional

from . import distributions

Root_Dir = os.getcwd()


def visualization(log_dir: Optional[str] = None, out_dir: Optional[str] = None) -> Dict[str, Dict]:
    log_dir = log_dir or os.path.join(Root_Dir, "experiment", "logs")
    return distributions.visualization(log_dir, None, 0, out_dir)
