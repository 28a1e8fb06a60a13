"""ncu target for the 8(f) kernels at the models' shapes. usage: python tools/prof_next.py"""
import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
from svdformer_pointsea_b200 import model_ops as mo
dev = "cuda:0"
g = torch.Generator().manual_seed(1)
for (B, C, N, k) in ((32, 3, 2048, 16), (32, 64, 512, 8), (32, 256, 512, 4)):
    x = ((torch.rand(B, N, 3, generator=g) - 0.5).permute(0, 2, 1).contiguous() if C == 3 else torch.randn(B, C, N, generator=g)).to(dev)
    for _ in range(2):
        idx = mo.knn_self(x, k)
        f = mo.edge_features_raw(x, idx)
    go = torch.randn_like(f)
    mo.edge_features_grad_raw(go, idx)
xyz = (torch.rand(32, 2048, 3, generator=g) - 0.5).to(dev)
ctr = xyz[:, :512].contiguous()
for _ in range(2):
    mo.knn_group_xyz_raw(xyz, ctr, 16)
gt = (torch.rand(32, 16384, 3, generator=g) - 0.5).to(dev)
xx = gt + 0.004 * torch.randn(32, 16384, 3, generator=g).to(dev)
d1, d2, i1, i2 = ps.chamfer_forward(gt, xx)
for _ in range(2):
    ps.chamfer_metrics_raw(d1, d2, i1, i2)
pts = torch.randn(32, 2048, 64, generator=g).to(dev)
ii = torch.randint(0, 2048, (32, 2048 * 16), generator=g, dtype=torch.int32).to(dev)
for _ in range(2):
    mo.index_points_raw(pts, ii)
torch.cuda.synchronize()
print("done")
