"""A/B of two builds of the library on one box: python tools/fps_ab.py  (POINTSEA_B200_LIB selects the build)"""
import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
g = torch.Generator().manual_seed(0)
for B, N, m in ((32, 16384, 2048), (4, 16384, 2048), (32, 8192, 1024), (16, 16384, 2048)):
    x = (torch.rand(B, N, 3, generator=g) - 0.5).cuda()
    for _ in range(3): ps.furthest_point_sample(x, m)
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ps.furthest_point_sample(x, m); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(B, N, m, "min %.4f median %.4f ms" % (ts[0], ts[4]), flush=True)
