"""Timings of the small call shapes inside SVDFormer's forward, ours vs the reference ops."""
import importlib.util, os.path as osp, sys
ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import svdformer_pointsea_b200 as ps
from svdformer_pointsea_b200 import pointnet2_utils as pu

def load_ext(name):
    p = osp.join(ROOT, "oracle", "_ref", name + ".so")
    if not osp.exists(p):
        return None
    spec = importlib.util.spec_from_file_location(name, p); m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m

def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

dev = "cuda:0"; g = torch.Generator().manual_seed(0)
mk = lambda B, N: (torch.rand(B, N, 3, generator=g) - 0.5).to(dev)
ref_pn, ref_ch = load_ext("ref_pointnet2_ext"), load_ext("ref_chamfer_3D")
B = 32
for N, m in ((2048, 512), (512, 128), (2304, 512), (2048, 256), (16384, 2048)):
    x = mk(B, N)
    ours = t(lambda: ps.furthest_point_sample(x, m))
    ref = t(lambda: ref_pn.furthest_point_sampling(x, m), 5) if ref_pn else float("nan")
    print(f"fps {N}->{m}: ours {ours:8.1f} us  ref {ref:9.1f} us  x{ref/ours:5.1f}  ({ours/(m-1):.3f} us/iter)")
for N, M in ((512, 2048), (2048, 2048), (256, 256), (16384, 16384), (2048, 16384)):
    a, b = mk(B, N), mk(B, M)
    ours = t(lambda: ps.chamfer_forward(a, b), 10)
    if ref_ch:
        r = [torch.zeros(B, N, device=dev), torch.zeros(B, M, device=dev), torch.zeros(B, N, device=dev, dtype=torch.int32), torch.zeros(B, M, device=dev, dtype=torch.int32)]
        ref = t(lambda: ref_ch.forward(a, b, *r), 3)
    else:
        ref = float("nan")
    print(f"chamfer fwd {N}x{M}: ours {ours:8.1f} us  ref {ref:9.1f} us  x{ref/ours:5.1f}  ({2*B*N*M/ours/1e3:.0f} Gpair/s)")
for N, S, C in ((2048, 512, 3), (512, 128, 128), (2048, 2048, 128)):
    x = mk(B, N); q = x[:, :S].contiguous(); feat = torch.randn(B, C, N, device=dev)
    k = ps.query_knn(16, x, q)
    ours_k = t(lambda: ps.query_knn(16, x, q))
    def torch_knn():
        d = -2 * torch.matmul(q, x.permute(0, 2, 1)); d += torch.sum(q ** 2, -1).view(B, S, 1); d += torch.sum(x ** 2, -1).view(B, 1, N)
        return torch.argsort(d, dim=-1)[:, :, :16].int()
    ref_k = t(torch_knn, 5)
    ours_g = t(lambda: pu.group_raw(feat, k))
    ref_g = t(lambda: ref_pn.group_points(feat, k), 5) if ref_pn else float("nan")
    print(f"knn N={N} S={S}: ours {ours_k:7.1f} us  torch {ref_k:9.1f} us x{ref_k/ours_k:5.1f} | group C={C}: ours {ours_g:7.1f} us ref {ref_g:9.1f} us x{ref_g/ours_g:5.1f}")
