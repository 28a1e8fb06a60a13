"""group/edge backward at the models' shapes: inverse-index (CSR) path vs shared-memory accumulators.
usage: PS_SCATTER_CSR={0,1} python tools/sweep_scatter.py"""
import json, os, sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
from svdformer_pointsea_b200 import pointnet2_utils as pu, model_ops as mo
dev = "cuda:0"
fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        fl.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
g = torch.Generator().manual_seed(0)
for (B, C, N, S, K) in ((32, 128, 512, 128, 16), (32, 3, 2048, 512, 16), (32, 64, 512, 512, 8), (32, 256, 512, 512, 4), (32, 64, 2048, 2048, 16),
                        (32, 128, 2048, 2048, 16), (32, 6, 2048, 2048, 16), (32, 64, 1024, 1024, 8), (4, 128, 2048, 2048, 16), (32, 512, 128, 128, 16)):
    idx = torch.randint(0, N, (B, S, K), generator=g, dtype=torch.int32).to(dev)
    go = torch.randn(B, C, S, K, generator=g).to(dev)
    t = timed(lambda: pu.group_grad_raw(go, idx, N))
    byts = 4 * (B * S * K + B * C * N + B * C * S * K)
    print(json.dumps({"csr_env": os.environ.get("PS_SCATTER_CSR", "default"), "shape": [B, C, N, S, K], "ms": round(t, 4), "gbs": round(byts / t / 1e6, 1)}), flush=True)
