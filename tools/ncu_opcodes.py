"""Opcode mix + stall attribution per opcode from `ncu -i X.ncu-rep --page source --csv` (first kernel section).
usage: ncu -i rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_opcodes.py src.csv [top]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 28
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
sec = sections[0]
hdr, data = sec["rows"][0], [r for r in sec["rows"][1:] if len(r) > 10]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
totinst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(sec["name"], "| sass lines", len(data), "| samples", tot, "| warp instructions", totinst)
op, ops, stall = collections.Counter(), collections.Counter(), collections.Counter()
KS = ("stall_math", "stall_wait", "stall_dispatch", "stall_short_sb", "stall_long_sb", "stall_barrier", "stall_not_selected", "stall_selected")
for r in data:
    s = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    o = s.split()[0].split(".")[0] if s else "?"
    op[o] += int(r[ix["Instructions Executed"]])
    ops[o] += int(r[ix["# Samples"]])
    for k in KS:
        stall[(o, k)] += int(r[ix[k]])
print("%-10s %7s %8s | " % ("opcode", "inst%", "samples%") + " ".join("%8s" % k[6:14] for k in KS))
for o, c in op.most_common(top):
    print("%-10s %6.1f%% %7.1f%% | " % (o, c / totinst * 100, ops[o] / tot * 100) + " ".join("%8.1f" % (stall[(o, k)] / tot * 100) for k in KS))
