"""FPS timing at the models' shapes: python tools/fps_time.py   (PS_FPS_EXCHANGE=poll|async selects the small-cluster exchange)"""
import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import numpy as np, torch
import svdformer_pointsea_b200 as ps
from oracle import oracle as O
g = torch.Generator().manual_seed(0)
for B, N, m in ((32, 16384, 2048), (32, 8192, 1024), (32, 4096, 512), (32, 2048, 512), (32, 2304, 512), (32, 1024, 256), (32, 512, 128), (4, 16384, 2048),
                (8, 131072, 2048)):
    x = (torch.rand(B, N, 3, generator=g) - 0.5).cuda()
    for _ in range(2): idx = ps.furthest_point_sample(x, m)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); idx = ps.furthest_point_sample(x, m); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ok = np.array_equal(idx[:2].cpu().numpy(), O.fps(x[:2].cpu().numpy(), m)) if N <= 16384 else None
    print(B, N, m, "%.4f ms" % min(ts), "%.3f us/iter" % (min(ts) * 1e3 / (m - 1)), "exact" if ok else ok)
