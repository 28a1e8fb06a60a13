"""Where does the host time of one bench step go?  python tools/host_profile.py  (1 GPU, cProfile over 300 steps)"""
import cProfile, pstats, sys, time, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
dev = torch.device("cuda", 0)
B, N, M = 32, 2048, 16384
g = torch.Generator().manual_seed(0)
x1 = (torch.rand(B, N, 3, generator=g) - 0.5).to(dev); x2 = (torch.rand(B, M, 3, generator=g) - 0.5).to(dev)
gd1 = torch.randn(B, N, generator=g).to(dev); gd2 = torch.randn(B, M, generator=g).to(dev)
fwd_out = (torch.empty(B, N, device=dev), torch.empty(B, M, device=dev), torch.empty(B, N, device=dev, dtype=torch.int32), torch.empty(B, M, device=dev, dtype=torch.int32))
bwd_out = (torch.empty(B, N, 3, device=dev), torch.empty(B, M, 3, device=dev))
sums = torch.empty(6, device=dev, dtype=torch.float64)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def step():
    flush_buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d1, d2, i1, i2 = ps.chamfer_forward(x1, x2, out=fwd_out)
    v = ps.chamfer_sums(d1, d2, out=sums)
    ps.chamfer_backward(x1, x2, gd1, gd2, i1, i2, out=bwd_out)
    e1.record()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host ms per step (no profiler): %.4f" % ((t1 - t0) * 1e3 / 300))
pr = cProfile.Profile(); pr.enable()
for _ in range(300): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
