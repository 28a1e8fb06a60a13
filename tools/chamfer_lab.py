"""A/B of the forward Chamfer kernel variants on one GPU: python tools/chamfer_lab.py [variant ...]

Every variant (PS_CHAMFER_SYM value, optionally "v:ENV=val,ENV=val") runs the same shapes; outputs are compared
bit-for-bit against variant 1 and the median device time of 15 L2-flushed runs is printed as one JSON line each."""
import json
import os
import os.path as osp
import sys

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch  # noqa: E402
import svdformer_pointsea_b200 as ps  # noqa: E402

SHAPES = [(32, 2048, 16384), (32, 16384, 16384), (32, 2048, 2048), (8, 1000, 5000)]
variants = sys.argv[1:] or ["1", "-1"]
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
data = [((torch.rand(B, N, 3, generator=g) - 0.5).cuda(), (torch.rand(B, M, 3, generator=g) - 0.5).cuda()) for B, N, M in SHAPES]
os.environ["PS_CHAMFER_SYM"] = "1"
ref = [tuple(t.clone() for t in ps.chamfer_forward(a, b)) for a, b in data]
KNOWN = ("PS_CHAMFER_SYM", "PS_CHAMFER_SPLIT", "PS_CHAMFER_TAIL", "PS_CHAMFER_SYM_Q")
for v in variants:
    for k in KNOWN:
        os.environ.pop(k, None)
    name, _, extra = v.partition(":")
    os.environ["PS_CHAMFER_SYM"] = name
    for kv in filter(None, extra.split(",")):
        k, _, val = kv.partition("=")
        os.environ[k] = val
    for (B, N, M), (a, b), r in zip(SHAPES, data, ref):
        for _ in range(3):
            out = ps.chamfer_forward(a, b)
        same = all(torch.equal(x, y) for x, y in zip(out, r))
        ts = []
        for _ in range(15):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ps.chamfer_forward(a, b); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(json.dumps({"variant": v, "shape": [B, N, M], "us": round(t * 1e3, 1), "min_us": round(min(ts) * 1e3, 1),
                          "tpair_alg": round(2 * B * N * M / t / 1e9, 2), "bit_identical": same}), flush=True)
