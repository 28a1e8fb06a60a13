"""Small driver for ncu captures: runs one hot-path op at its BASELINE config a few times.
usage: python tools/prof_target.py {chamfer|fps|knn|group|group_bwd|all} [iters]"""
import sys
import os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
from svdformer_pointsea_b200 import pointnet2_utils as pu

what = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
mk = lambda B, N: (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
if what in ("chamfer", "all"):
    a, b = mk(32, 2048), mk(32, 16384)
    for _ in range(iters):
        out = ps.chamfer_forward(a, b)
    gd1, gd2 = torch.randn(32, 2048, device=dev), torch.randn(32, 16384, device=dev)
    ps.chamfer_backward(a, b, gd1, gd2, out[2], out[3])
if what in ("fps", "all"):
    x = mk(32, 16384)
    for _ in range(iters):
        idx = ps.furthest_point_sample(x, 2048)
    ps.gather_operation(x.transpose(1, 2).contiguous(), idx)
if what in ("knn", "group", "group_bwd", "all"):
    p = mk(32, 2048)
    feat = torch.randn(32, 128, 2048, device=dev)
    for _ in range(iters):
        k = ps.query_knn(16, p, p)
    if what != "knn":
        for _ in range(iters):
            o = pu.group_raw(feat, k)
        if what in ("group_bwd", "all"):
            go = torch.randn_like(o)
            for _ in range(iters):
                pu.group_grad_raw(go, k, 2048)
torch.cuda.synchronize()
print("done", what)
