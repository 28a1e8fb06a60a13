"""Grouping backward at the C3 shape (B=32, C=128, N=2048, S=2048, K=16) and two EdgeConv shapes: python tools/group_bwd_time.py"""
import json, sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
from svdformer_pointsea_b200 import pointnet2_utils as pu
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator().manual_seed(3)
for (B, C, N, S, K) in ((32, 128, 2048, 2048, 16), (32, 64, 512, 512, 8), (32, 256, 2048, 512, 16), (4, 128, 2048, 2048, 16)):
    p = (torch.rand(B, N, 3, generator=g) - 0.5).cuda()
    q = p[:, :S].contiguous()
    idx = ps.query_knn(K, p, q)
    go = torch.randn(B, C, S, K, generator=g).cuda()
    for _ in range(3):
        out = pu.group_grad_raw(go, idx, N)
    ts = []
    for _ in range(15):
        fl.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pu.group_grad_raw(go, idx, N); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    nbytes = 4 * (B * S * K + B * C * N + B * C * S * K)
    want = torch.zeros(B, C, N, device="cuda").scatter_add_(2, idx.long().view(B, 1, S * K).expand(B, C, S * K), go.view(B, C, S * K))
    err = float((out - want).abs().max() / want.abs().max())
    print(json.dumps({"shape": [B, C, N, S, K], "ms": round(t, 4), "min_ms": round(min(ts), 4), "gbs": round(nbytes / t / 1e6, 1),
                      "frac_hbm": round(nbytes / t / 1e6 / 6555.8, 4), "max_rel_err_vs_scatter_add": err}), flush=True)
