"""Stall samples and executed instructions per CUDA source line of one kernel in an .ncu-rep (needs -lineinfo):
    python tools/ncu_lines.py report.ncu-rep kernel_regex [top]
All file sections of the first matching launch are merged (inlined helpers live in their own file section)."""
import csv, io, os.path as osp, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
KS = ("stall_math", "stall_wait", "stall_dispatch", "stall_short_sb", "stall_long_sb", "stall_barrier")


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


lines, hdr, ix, fname, func = [], None, None, "?", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = osp.basename(r[1]); hdr = None
        continue
    if r[0] == "Function Name":
        func = func or r[1]
        continue
    if r[0] == "Line No":
        hdr = r; ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == "":
        continue
    lines.append((fname, num(r[0]), r[1].strip(), num(r[ix["# Samples"]]), num(r[ix["Instructions Executed"]]), {k: num(r[ix[k]]) for k in KS}))
print(func)
tot_s = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print("source lines: %d | samples %d | warp instructions %d" % (len(lines), tot_s, tot_i))
print("%-16s %5s %7s %7s | %5s %5s %5s %5s %5s %5s | source" % ("file", "line", "smpl%", "inst%", "math", "wait", "disp", "ssb", "lsb", "bar"))
for fn, ln, src, smp, ins, st in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%-16s %5d %6.1f%% %6.1f%% | %5.1f %5.1f %5.1f %5.1f %5.1f %5.1f | %s" % (
        fn[:16], ln, 100 * smp / tot_s, 100 * ins / tot_i, *(100 * st[k] / tot_s for k in KS), src[:84]))
