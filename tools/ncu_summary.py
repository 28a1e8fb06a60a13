"""Summarise an .ncu-rep: headline metrics per kernel + hottest SASS lines by stall samples.
usage: python tools/ncu_summary.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
for r in rows[2:]:
    print("=" * 100)
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:90s} {r[i][:60]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
for blk in blocks[1:2]:
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    h = rows[1]
    ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) > isamp]
    tot = sum(int(r[isamp] or 0) for r in data) or 1
    print("-" * 100)
    print("hottest SASS (share of stall samples) for", rows[0][1][:80], "| total samples", tot)
    ranked = sorted(data, key=lambda r: -int(r[isamp] or 0))[:topn]
    keep = {id(r) for r in ranked}
    for r in data:
        if id(r) in keep:
            print(f"{r[ia][-5:]} {int(r[isamp] or 0) * 100 / tot:5.1f}%  ex={r[iex]:>10}  {r[isrc][:100]}")
