import torch, sys
sys.path.insert(0, ".")
import svdformer_pointsea_b200 as ps
g = torch.Generator().manual_seed(3)
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        fl.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for (B, N, S, k) in ((32, 2048, 2048, 16), (32, 2048, 512, 16), (32, 512, 128, 16)):
    x = (torch.rand(B, N, 3, generator=g) - 0.5).cuda(); q = x[:, :S].contiguous()
    print(B, N, S, k, "sort %.4f ms" % timed(lambda: ps.query_knn(k, x, q)))
