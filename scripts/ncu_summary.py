"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + executed-SASS opcode mix.
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01/x_summary.txt [frames]"""
import collections, csv, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
units_n = float(sys.argv[3]) if len(sys.argv) > 3 else None
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "sm__cycles_elapsed.max"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
lines = [f"# ncu summary of {rep} (ncu --set full --clock-control none; cold-cache replay, not a bench number)"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    lines.append(f"## kernel: {d.get('Kernel Name','?')}  grid={d.get('Grid Size','?')} block={d.get('Block Size','?')}")
    for k in KEYS:
        if k in d:
            lines.append(f"{k:95s} {d[k]:>16s} {u[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = None
byop = collections.Counter(); tot = 0
for r in srows:
    if "Source" in r and "Instructions Executed" in r:
        h = r; continue
    if h is None or len(r) != len(h): continue
    n = int(r[h.index("Instructions Executed")] or 0); tot += n
    t = r[h.index("Source")].split()
    op = t[1] if t[0].startswith("@") else t[0]
    byop[op.split(".")[0]] += n
lines.append(f"## executed warp-instructions by SASS opcode (first profiled launch): total {tot}" + (f" = {tot/units_n:.1f} per unit" if units_n else ""))
for op, n in byop.most_common(24):
    lines.append(f"{op:12s} {n:14d} {100*n/max(tot,1):5.1f}%" + (f"  {n/units_n:8.1f}/unit" if units_n else ""))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
