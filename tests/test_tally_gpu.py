"""GPU parity: label histogram + tallies vs the oracle restatement of the reference's
visualization() counting (bit-exact integers required)."""
from datetime import datetime

import numpy as np
import pytest

from oracle import tally as otally

pytestmark = pytest.mark.gpu


def test_device_counts_exact(cuda):
    from mmla_audio_b200 import tally
    torch = cuda
    rng = np.random.default_rng(0)
    for n, c in ((1, 2), (1000, 2), (100003, 10), (1 << 20, 630)):
        lab = rng.integers(-1, c, n).astype(np.int32)
        got = tally.device_counts(torch.from_numpy(lab).cuda(), c).cpu().numpy()
        ref = np.bincount(np.where(lab < 0, c, lab), minlength=c + 1)
        np.testing.assert_array_equal(got, ref)
    assert tally.device_counts(torch.empty(0, dtype=torch.int32, device="cuda"), 3).sum().item() == 0


def test_overlap_session_tally_matches_reference_parser(cuda):
    from mmla_audio_b200 import tally
    torch = cuda
    rng = np.random.default_rng(1)
    labels = rng.integers(0, 2, 19200).astype(np.int32)            # 8 h of 1.5 s windows
    t0 = datetime(2021, 11, 3, 14, 25, 36, 123456)
    names = [tally.OVERLAP_DEGREE_DICT[str(l)] for l in labels]
    lines = otally.log_rows(names, t0, 1.5, "overlapped degree", add_before_first=False)
    ref_counts, ref_secs, ref_total = otally.tally_from_log(lines, list(tally.OVERLAP_DEGREE_DICT.values()))
    got_counts, got_secs, got_total = tally.tally_session(
        torch.from_numpy(labels).cuda(), {0: "non-overlapped", 1: "overlapped"}, t0, 1.5, False,
        initial_order=list(tally.OVERLAP_DEGREE_DICT.values()))
    assert got_total == ref_total
    assert got_counts == ref_counts and list(got_counts) == list(ref_counts)
    assert got_secs == ref_secs
    assert tally.log_rows(names, t0, 1.5, "overlapped degree", False) == lines


def test_speaker_session_tally_with_silent(cuda):
    from mmla_audio_b200 import tally
    torch = cuda
    rng = np.random.default_rng(2)
    labels = rng.integers(-1, 10, 11250).astype(np.int32)          # 8 h of 2.56 s chunks, -1 = silent
    id_to_name = {i: f"spk{i}" for i in range(10)}
    t0 = datetime(2022, 1, 9, 9, 0, 0, 500001)
    names = [id_to_name.get(int(l), "silent") for l in labels]
    lines = otally.log_rows(names, t0, 2.56, "speaker", add_before_first=True)
    ref_counts, ref_secs, ref_total = otally.tally_from_log(lines)
    got_counts, got_secs, got_total = tally.tally_session(torch.from_numpy(labels).cuda(), id_to_name, t0, 2.56, True)
    assert got_total == ref_total
    assert got_counts == ref_counts and list(got_counts) == list(ref_counts)
    assert got_secs == ref_secs
