"""C4 loss part (get_loss_sharded on one GPU) and its pieces at a given shard size: python tools/c4loss_time.py [B ...]
Environment switches under test: PS_LOSS_OVERLAP, PS_LOSS_CORUN, PS_FPS_PRUNE, PS_FPS_CLUSTER, PS_FPS_THREADS."""
import sys, os, os.path as osp, json
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
import svdformer_pointsea_b200 as ps
from svdformer_pointsea_b200.dist import get_loss_sharded, GraphedLoss
from svdformer_pointsea_b200.pointnet2_utils import fps_sample_raw

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[len(ts) // 2], 4)


env = {k: os.environ[k] for k in ("PS_LOSS_OVERLAP", "PS_LOSS_CORUN", "PS_FPS_PRUNE", "PS_FPS_CLUSTER", "PS_FPS_THREADS") if k in os.environ}
for B in [int(a) for a in sys.argv[1:]] or [32, 4]:
    g = torch.Generator().manual_seed(1238)
    gt = (torch.rand(B, 16384, 3, generator=g) - 0.5).to(dev)
    preds = [((torch.rand(B, n, 3, generator=g) - 0.5).to(dev)).requires_grad_(True) for n in (512, 2048, 16384)]

    def step():
        loss, _ = get_loss_sharded(preds, gt, sqrt=True)
        loss.backward()
        for p in preds: p.grad = None

    def cd2():
        d1, d2, _, _ = ps.chamfer_3DFunction.apply(preds[2], gt)
        (d1.sqrt().mean() + d2.sqrt().mean()).backward()
        preds[2].grad = None

    def chain(corun):
        x = fps_sample_raw(gt, 2048, corun=corun)[1]
        fps_sample_raw(x, 512, corun=corun)[1]

    graphed = GraphedLoss([p.shape for p in preds], gt.shape, sqrt=True)
    print(json.dumps({"B": B, "env": env, "loss_ms": timed(step), "graphed_loss_ms": timed(lambda: graphed(preds, gt)), "cd2_alone_ms": timed(cd2), "fps_chain_ms": timed(lambda: chain(False)),
                      "fps_chain_corun_ms": timed(lambda: chain(True))}), flush=True)
