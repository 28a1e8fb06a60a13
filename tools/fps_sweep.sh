# forced-plan sweep of the FPS cluster kernels: bash tools/fps_sweep.sh  (B = 32, 16 and 4 clouds of 16384 points -> 2048)
python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import svdformer_pointsea_b200 as ps
g = torch.Generator().manual_seed(0)
for B in (32, 16, 4):
    x = (torch.rand(B, 16384, 3, generator=g) - 0.5).cuda()
    os.environ.pop("PS_FPS_CLUSTER", None); os.environ.pop("PS_FPS_THREADS", None)
    ref = ps.furthest_point_sample(x, 2048).clone()
    for c, t in ((0, 0), (4, 128), (4, 256), (8, 128), (8, 256), (16, 128)):
        if c:
            os.environ["PS_FPS_CLUSTER"], os.environ["PS_FPS_THREADS"] = str(c), str(t)
        else:
            os.environ.pop("PS_FPS_CLUSTER", None); os.environ.pop("PS_FPS_THREADS", None)
        for _ in range(2): idx = ps.furthest_point_sample(x, 2048)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); idx = ps.furthest_point_sample(x, 2048); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        print(json.dumps({"B": B, "cluster": c or "auto", "threads": t or "auto", "ms": round(min(ts), 4), "us_per_iter": round(min(ts) * 1e3 / 2047, 4),
                          "same": bool(torch.equal(idx, ref))}), flush=True)
PY
