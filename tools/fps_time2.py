import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
g = torch.Generator().manual_seed(0)
for B, N, m in ((32, 2048, 512), (32, 2304, 512), (32, 512, 128), (32, 4096, 512), (32, 8192, 1024), (32, 1024, 256)):
    x = (torch.rand(B, N, 3, generator=g) - 0.5).cuda()
    try:
        for _ in range(2): idx = ps.furthest_point_sample(x, m)
    except Exception as e:
        print(B, N, m, "unsupported"); continue
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx = ps.furthest_point_sample(x, m); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    print(B, N, m, "%.4f ms" % min(ts), "%.3f us/iter" % (min(ts) * 1e3 / (m - 1)))
