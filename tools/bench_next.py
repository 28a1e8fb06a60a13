"""Device timings of the SURVEY 8(f) ops at the models' shapes, next to the reference's torch expressions
on the same GPU (results -> profiles/next_ops_r1.jsonl).  usage: python tools/bench_next.py [iters]"""
import json
import os.path as osp
import statistics
import sys

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch  # noqa: E402

import svdformer_pointsea_b200 as ps  # noqa: E402
from svdformer_pointsea_b200 import model_ops as mo, pointnet2_utils as pu  # noqa: E402
from oracle import oracle as O  # noqa: E402  (torch re-expressions of the reference call sites: the GPU baseline)

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = "cuda:0"
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
HBM = json.load(open(osp.join(osp.dirname(osp.dirname(osp.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if osp.exists(osp.join(osp.dirname(osp.dirname(osp.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, n=iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), statistics.median(ts)


def emit(**kw):
    print(json.dumps(kw), flush=True)


g = torch.Generator().manual_seed(7)
# ---- EdgeConv front at the three EdgeConv shapes of the models (B=32) ------------------------------
for (B, C, N, k) in ((32, 3, 2048, 16), (32, 64, 512, 8), (32, 256, 512, 4), (32, 64, 1024, 8)):
    x = ((torch.rand(B, N, 3, generator=g) - 0.5).permute(0, 2, 1).contiguous() if C == 3 else torch.randn(B, C, N, generator=g)).to(dev)
    idx = mo.knn_self(x, k)
    want, widx = O.torch_edge_features(x, k)
    assert torch.equal(idx.long(), widx) and torch.equal(mo.edge_features_raw(x, idx), want)
    t_knn = timed(lambda: mo.knn_self(x, k))
    t_edge = timed(lambda: mo.edge_features_raw(x, idx))
    t_all = timed(lambda: ps.edge_features(x, k))
    t_ref = timed(lambda: O.torch_edge_features(x, k), n=max(3, iters // 4))
    byts = 4 * (B * N * k + B * C * N + 2 * B * C * N * k)
    emit(op="edge_features (EdgeConv front)", shape=f"B={B} C={C} N={N} k={k}", knn_ms=round(t_knn[0], 4), gather_ms=round(t_edge[0], 4),
         total_ms=round(t_all[0], 4), torch_expr_ms=round(t_ref[0], 4), speedup=round(t_ref[0] / t_all[0], 1),
         gather_gbs=round(byts / t_edge[1] / 1e6, 1), gather_frac_hbm=round(byts / t_edge[1] / 1e6 / HBM, 3),
         knn_gpair_per_s=round(B * N * N / t_knn[0] / 1e6, 1), parity="idx and features identical to the torch expression")
    go = torch.randn_like(want)
    t_bwd = timed(lambda: mo.edge_features_grad_raw(go, idx))
    emit(op="edge_features backward", shape=f"B={B} C={C} N={N} k={k}", ms=round(t_bwd[0], 4), gbs=round(byts / t_bwd[1] / 1e6, 1))
    del want, go

# ---- sample_and_group_knn at the SA-module shapes of the PCN model -----------------------------------
for (B, N, npoint, k, f) in ((32, 2048, 512, 16, 0), (32, 512, 128, 16, 128)):
    xyz = (torch.rand(B, N, 3, generator=g) - 0.5).permute(0, 2, 1).contiguous().to(dev)
    pts = torch.randn(B, f, N, generator=g).to(dev) if f else None

    def ref_expr():
        xf = xyz.permute(0, 2, 1).contiguous()
        nx = pu.gather_operation(xyz, pu.furthest_point_sample(xf, npoint))
        ii = O.torch_knn(k, xf, nx.permute(0, 2, 1).contiguous())
        gx = pu.grouping_operation(xyz, ii)
        gx -= nx.unsqueeze(3).repeat(1, 1, 1, k)
        return nx, (torch.cat([gx, pu.grouping_operation(pts, ii)], 1) if f else gx), ii, gx

    a, b = ps.sample_and_group_knn(xyz, pts, npoint, k), ref_expr()
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    t_f = timed(lambda: ps.sample_and_group_knn(xyz, pts, npoint, k))
    t_r = timed(ref_expr, n=max(3, iters // 4))
    xf = xyz.permute(0, 2, 1).contiguous()
    nx = a[0].permute(0, 2, 1).contiguous()
    t_k = timed(lambda: mo.knn_group_xyz_raw(xf, nx, k))
    emit(op="sample_and_group_knn", shape=f"B={B} N={N} npoint={npoint} k={k} f={f}", ms=round(t_f[0], 4),
         knn_group_kernel_ms=round(t_k[0], 4), reference_expr_on_our_ops_ms=round(t_r[0], 4), speedup=round(t_r[0] / t_f[0], 1))

# ---- evaluation metrics at the PCN evaluation shape ---------------------------------------------------
for (B, n) in ((32, 16384), (32, 8192)):
    gt = (torch.rand(B, n, 3, generator=g) - 0.5).to(dev)
    x = gt + 0.004 * torch.randn(B, n, 3, generator=g).to(dev)
    d1, d2, i1, i2 = ps.chamfer_forward(gt, x)

    def ref_metrics():
        return O.torch_dcd_from_raw(d1, d2, i1, i2, n, n), O.torch_cd_terms(d1, d2), O.torch_fscore(d1, d2)

    t_o = timed(lambda: ps.chamfer_metrics_raw(d1, d2, i1, i2))
    t_r = timed(ref_metrics)
    emit(op="calc_dcd + calc_cd + fscore epilogue", shape=f"B={B} n={n}", ms=round(t_o[0], 4), torch_expr_ms=round(t_r[0], 4),
         speedup=round(t_r[0] / t_o[0], 1), bytes=16 * B * n, gbs=round(16 * B * n / t_o[1] / 1e6, 1))

# ---- index_points (GeoSpecNet neighbourhood gather) ----------------------------------------------------
for (B, N, C, K) in ((32, 2048, 64, 16), (32, 512, 256, 8)):
    pts = torch.randn(B, N, C, generator=g).to(dev)
    idx = torch.randint(0, N, (B, N, K), generator=g).to(dev)
    assert torch.equal(ps.index_points(pts, idx), O.torch_index_points(pts, idx))
    idx32 = idx.reshape(B, -1).int().contiguous()
    t_o = timed(lambda: mo.index_points_raw(pts, idx32))
    t_r = timed(lambda: O.torch_index_points(pts, idx))
    byts = 4 * (B * N * K + B * N * C + B * N * K * C)
    emit(op="index_points", shape=f"B={B} N={N} C={C} K={K}", ms=round(t_o[0], 4), torch_expr_ms=round(t_r[0], 4),
         gbs=round(byts / t_o[1] / 1e6, 1), frac_hbm=round(byts / t_o[1] / 1e6 / HBM, 3))
