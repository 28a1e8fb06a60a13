"""Forward-Chamfer timing probe: python tools/chamfer_probe.py [B N M]...  (env PS_CHAMFER_* select the kernel)."""
import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
shapes = [(32, 2048, 16384), (32, 16384, 16384), (32, 2048, 2048), (32, 512, 2048)]
if len(sys.argv) > 3:
    v = list(map(int, sys.argv[1:]))
    shapes = [tuple(v[i:i + 3]) for i in range(0, len(v), 3)]
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, N, M in shapes:
    a = (torch.rand(B, N, 3, generator=g) - 0.5).cuda()
    b = (torch.rand(B, M, 3, generator=g) - 0.5).cuda()
    for _ in range(3):
        ps.chamfer_forward(a, b)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ps.chamfer_forward(a, b); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"B={B} N={N} M={M}: {t*1e3:.1f} us  {2*B*N*M/t/1e9:.2f} Tpair/s algorithmic")
