"""Bucket-pruned FPS against the cluster kernel: python tools/fps_prune_time.py
Times ps_fps under PS_FPS_PRUNE=0 (cluster kernel), 1 (pruned kernel forced) and unset (pruned kernel + give-up path)
on uniform cubes, surfaces and a cloud on which the pruning cannot bite; checks every result against the oracle."""
import json
import os
import os.path as osp
import sys

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import svdformer_pointsea_b200 as ps  # noqa: E402
from oracle import oracle as O  # noqa: E402

g = torch.Generator().manual_seed(0)
MODES = [m if m != "auto" else None for m in sys.argv[1:]] or ["0", "1", None]  # "1:161" = forced, kernel variant 161


def cube(B, N):
    return torch.rand(B, N, 3, generator=g) - 0.5


def sphere(B, N):
    x = torch.randn(B, N, 3, generator=g)
    return x / x.norm(dim=2, keepdim=True) * 0.45


def squeezed(B, N):
    x = torch.randn(B, N, 3, generator=g) * 1e-3 + 0.3
    x[:, 5] = torch.tensor([900.0, -700.0, 800.0])
    return x


CASES = [("cube", cube, 32, 16384, 2048), ("sphere", sphere, 32, 16384, 2048), ("cube", cube, 4, 16384, 2048),
         ("cube", cube, 148, 16384, 2048), ("cube", cube, 32, 8192, 1024), ("cube", cube, 32, 4096, 512),
         ("cube", cube, 32, 2048, 512), ("squeezed", squeezed, 32, 16384, 2048)]
for name, fn, B, N, m in CASES:
    x = fn(B, N).cuda()
    want = O.fps(x[:2].cpu().numpy(), m)
    for mode in MODES:
        os.environ.pop("PS_FPS_PRUNE_VARIANT", None)
        if mode is None:
            os.environ.pop("PS_FPS_PRUNE", None)
        elif ":" in mode:
            os.environ["PS_FPS_PRUNE"], os.environ["PS_FPS_PRUNE_VARIANT"] = mode.split(":")
        else:
            os.environ["PS_FPS_PRUNE"] = mode
        for _ in range(2):
            idx = ps.furthest_point_sample(x, m)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx = ps.furthest_point_sample(x, m); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ok = bool(np.array_equal(idx[:2].cpu().numpy(), want))
        print(json.dumps({"cloud": name, "B": B, "N": N, "npoint": m, "PS_FPS_PRUNE": mode, "ms": round(min(ts), 4),
                          "us_per_iter": round(min(ts) * 1e3 / (m - 1), 4), "exact": ok}), flush=True)
