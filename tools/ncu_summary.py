#!/usr/bin/env python
"""Print the headline metrics of an .ncu-rep (raw page) and the hottest source lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
for w in want:
    for i, h in enumerate(H):
        if h == w:
            print("%-90s %s %s" % (h, V[i], U[i]))
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = None
    for i, r in enumerate(rows):
        if "Source" in r and any("Sampling" in x for x in r):
            hdr = i; break
    if hdr is not None:
        Hs = rows[hdr]
        si = Hs.index("Source")
        ci = [i for i, x in enumerate(Hs) if x.startswith("# Warp Stall Sampling (All")][0] if any(x.startswith("# Warp Stall Sampling (All") for x in Hs) else None
        if ci is None:
            ci = [i for i, x in enumerate(Hs) if "Sampling" in x][0]
        data = []
        for r in rows[hdr + 1:]:
            try:
                data.append((float(r[ci]), r[si]))
            except Exception:
                pass
        tot = sum(d[0] for d in data) or 1
        print("--- hottest lines (%s), total samples %d" % (Hs[ci], tot))
        for v, s in sorted(data, reverse=True)[:int(sys.argv[2])]:
            print("%6.2f%%  %s" % (100 * v / tot, s.strip()[:150]))
