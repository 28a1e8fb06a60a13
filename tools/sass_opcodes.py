"""Static SASS evidence per kernel of libpointsea_b200.so: python tools/sass_opcodes.py > profiles/sass_opcodes_rN.txt

For every kernel in the shared library (cuobjdump -sass of the sm_100a cubin) prints the instruction count and the
counts of the mnemonics that show how it maps to the hardware: packed FP32 (FFMA2/FADD2/FMUL2), 3-input min/max
(FMNMX3), warp reductions (REDUX / CREDUX), votes, bulk async copies (UBLKCP = cp.async.bulk, TMA engine), async
remote stores (STAS = st.async), cluster barriers (UCGABAR), mbarrier ops (SYNCS), cp.async (LDGSTS), shared-memory
atomics (ATOMS), global reductions (RED), and — to show what is NOT there — tensor-core / TMA-tensor ops
(UTCMMA / tcgen05, UTMALDG) and legacy HMMA."""
import collections
import os.path as osp
import re
import subprocess
import sys

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else osp.join(ROOT, "svdformer_pointsea_b200", "lib", "libpointsea_b200.so")
KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FMNMX3", "FMNMX", "REDUX", "CREDUX", "VOTE", "SHFL", "UBLKCP", "STAS", "UCGABAR", "SYNCS",
        "LDGSTS", "ATOMS", "RED", "ATOMG", "LDS", "STS", "LDG", "STG", "UTCMMA", "UTMALDG", "HMMA", "BAR"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, counts, total = None, None, 0
rows = []


def flush():
    if name is not None:
        rows.append((name, total, counts))


for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and name is not None:
        total += 1
        counts[m.group(1)] += 1
flush()
demangle = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print(f"# {osp.relpath(lib, ROOT)}: {len(rows)} kernels, cuobjdump -sass, sm_100a")
print(f"{'kernel':78s} {'instr':>6s}  " + " ".join(f"{k:>7s}" for k in KEYS))
tot = collections.Counter()
for (n, t, c), d in sorted(zip(rows, demangle), key=lambda x: x[1]):
    short = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "").replace("void ", "").replace("ps::", ""))[:78]
    by = {k: sum(v for op, v in c.items() if op == k or op.startswith(k + "_")) for k in KEYS}
    print(f"{short:78s} {t:6d}  " + " ".join(f"{by[k]:7d}" for k in KEYS))
    tot.update(by)
print(f"{'TOTAL':78s} {sum(r[1] for r in rows):6d}  " + " ".join(f"{tot[k]:7d}" for k in KEYS))
