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


def test_visualization_reads_log_files(cuda, tmp_path):
    """File-based visualization() (overlap_degree_distribution.py:14-65, speaker_time_distribution.py:16-86):
    logs written by the session drivers, counted on the device, equal to the oracle's line-by-line
    restatement; the chart series are exported as JSON."""
    import json
    from datetime import datetime
    import numpy as np
    from mmla_audio_b200 import distributions, overlap_degree_distribution as odd, speaker_time_distribution as std, tally
    from oracle import tally as otally

    rng = np.random.default_rng(3)
    t0 = datetime(2021, 11, 5, 10, 0, 0, 250000)
    logs = tmp_path / "experiment" / "logs"
    # overlap session: 1.5 s rows, first row at t0
    ov = [["non-overlapped", "overlapped", "silent"][i] for i in rng.integers(0, 3, 257)]
    lines_o = tally.log_rows(ov, t0, 1.5, "overlapped degree", add_before_first=False)
    assert lines_o == otally.log_rows(ov, t0, 1.5, "overlapped degree", False)
    distributions.write_log(str(logs / "ov" / "a.txt"), lines_o)
    res = odd.visualization(str(logs / "ov"))["a.txt"]
    dist, secs, total = otally.tally_from_log(lines_o, ["non-overlapped", "overlapped", "silent"])
    assert res["labels"] == list(dist.keys()) and res["counts"] == list(dist.values())
    assert res["seconds"] == list(secs.values()) and res["total_seconds"] == total
    assert res["bars"]["overlapped"][:5] == [1 if l == "overlapped" else None for l in ov[:5]]
    exported = json.load(open(logs / "ov" / "a.txt.tally.json"))
    assert exported["counts"] == res["counts"] and len(exported["x_bar"]) == 257
    # speaker session: 2.56 s added before every row, labels in order of first appearance
    sp = [["amy", "bo", "silent", "cy"][i] for i in rng.integers(0, 4, 300)]
    lines_s = tally.log_rows(sp, t0, 2.56, "speaker", add_before_first=True)
    distributions.write_log(str(logs / "sp" / "b.txt"), lines_s)
    res = std.visualization(str(logs / "sp"))["b.txt"]
    dist, secs, total = otally.tally_from_log(lines_s)
    assert res["labels"] == list(dist.keys()) and res["counts"] == list(dist.values())
    assert res["seconds"] == list(secs.values()) and res["total_seconds"] == total
    second = res["labels"][1]
    f0 = sp.index(second)
    assert res["bars"][second][:f0] == [None] * f0 and res["bars"][second][f0] == 1
