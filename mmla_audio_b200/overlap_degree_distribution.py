"""``overlap_degree_distribution.visualization()`` — OverlapDetection/scripts/overlap_degree_distribution.py:14-65.
Counts rows per overlap degree in every ``experiment/logs/*`` file and converts them to seconds;
the chart rendering is replaced by a ``<log>.tally.json`` export (see ``distributions``)."""
from __future__ import annotations

import os
from typing import Dict, Optional

from . import distributions

overlap_degree_dict = dict(distributions.OVERLAP_DEGREE_DICT)
Root_Dir = os.getcwd()


def visualization(log_dir: Optional[str] = None, out_dir: Optional[str] = None) -> Dict[str, Dict]:
    """The reference takes no arguments and reads ``Root_Dir + '/experiment/logs/'``; that is the default here."""
    log_dir = log_dir or os.path.join(Root_Dir, "experiment", "logs")
    return distributions.visualization(log_dir, list(overlap_degree_dict.values()), None, out_dir)
