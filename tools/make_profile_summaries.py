"""Turn gpurun_out/{launches_r1.csv, prof_all_r1.ncu-rep} into the tracked summaries under profiles/."""
import collections
import csv
import io
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rows = list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[ik], float(r[iv].replace(",", "")) / 1000.0) for r in rows[start:] if len(r) > iv]
agg = collections.OrderedDict()
for k, v in seq:
    a = agg.setdefault(k[:100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in seq)
with open(f"profiles/launches_{tag}_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 python bench.py --steps 3 --warmup 3 --no-cpu --no-ref-cuda --repeats 1\n"
            if tag != "r1" else "ncu --metrics gpu__time_duration.sum --clock-control none -c 900 python bench.py --steps 3 --warmup 3 --no-cpu\n")
    f.write("(cold-cache, serialised launches: compare SHARES, not absolutes)  total %.1f us over %d launches\n\n" % (tot, len(seq)))
    f.write("%8s %10s %9s %7s  kernel\n" % ("count", "sum_us", "avg_us", "share"))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%8d %10.1f %9.1f %6.1f%%  %s\n" % (c, v, v / c, 100 * v / tot, k))
subprocess.run(["cp", f"gpurun_out/launches_{tag}.csv", f"profiles/launches_{tag}.csv"])

raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_all_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
keys = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("smsp__inst_executed.sum", "inst")]
out = []
for r in rows[2:]:
    line = []
    for k, n in keys:
        i = h.index(k)
        v = r[i]
        if n in ("dram_rd", "dram_wr", "time"):
            v = f"{float(v):.3f}{u[i]}"
        elif n == "kernel":
            v = v[:46]
        elif n == "inst":
            v = f"{float(v) / 1e6:.1f}M"
        else:
            try:
                v = f"{float(v):.1f}"
            except ValueError:
                pass
        line.append(f"{n}={v}")
    out.append("  ".join(line))
open(f"profiles/ncu_full_{tag}_table.txt", "w").write(
    "ncu --set full --clock-control none --import-source on -k regex:... -c 20 python tools/prof_target.py all 2\n"
    "one line per captured launch (B200; C1 Chamfer, C2 FPS, C3 kNN / group fwd / group bwd shapes; cold-cache replays)\n" + "\n".join(out) + "\n")
print("\n".join(out))
