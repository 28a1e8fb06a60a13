"""ncu target: the bucket-pruned FPS kernel forced at the C2 shape (B=32, 16384 -> 2048). usage: python tools/prof_fps_pruned.py"""
import os, sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
os.environ["PS_FPS_PRUNE"] = "1"
import torch
import svdformer_pointsea_b200 as ps
g = torch.Generator().manual_seed(1)
x = (torch.rand(32, 16384, 3, generator=g) - 0.5).cuda()
for _ in range(2):
    ps.furthest_point_sample(x, 2048)
torch.cuda.synchronize()
print("done")
