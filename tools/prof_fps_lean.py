"""ncu target: the C2 FPS call (B=32, 16384 -> 2048; fps_lean_kernel) three times."""
import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
g = torch.Generator().manual_seed(0)
x = (torch.rand(32, 16384, 3, generator=g) - 0.5).cuda()
for _ in range(3):
    ps.furthest_point_sample(x, 2048)
torch.cuda.synchronize()
