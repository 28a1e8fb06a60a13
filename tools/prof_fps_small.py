import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
g = torch.Generator().manual_seed(0)
x = (torch.rand(32, 2048, 3, generator=g) - 0.5).cuda()
for _ in range(3):
    ps.furthest_point_sample(x, 512)
torch.cuda.synchronize()
