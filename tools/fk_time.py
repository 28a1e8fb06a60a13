"""Feature-space kNN timing at the EdgeConv shapes: python tools/fk_time.py
(PS_KNN_FEAT8=0 selects the 4x8 kernel, PS_KNN_FEAT_UNION=0 the 8-channel stages beside the distance block)."""
import json, os, sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
from svdformer_pointsea_b200 import model_ops as mo
fl = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = torch.Generator().manual_seed(1)
for (B, C, N, k) in ((32, 64, 512, 8), (32, 256, 512, 4), (32, 128, 1024, 8), (32, 512, 128, 4), (32, 200, 500, 4)):
    x = torch.randn(B, C, N, generator=g).cuda()
    os.environ["PS_KNN_FEAT8"] = "0"
    ref = mo.knn_self(x, k).clone()
    for mode in ("0", "1", "1u"):
        os.environ["PS_KNN_FEAT8"] = mode[0]
        os.environ["PS_KNN_FEAT_UNION"] = "1" if mode.endswith("u") else "0"
        for _ in range(3):
            out = mo.knn_self(x, k)
        ts = []
        for _ in range(15):
            fl.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); mo.knn_self(x, k); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        t = min(ts)
        print(json.dumps({"shape": [B, C, N, k], "feat8": mode, "ms": round(t, 4), "tflops": round(2.0 * C * B * N * N / t / 1e9, 2),
                          "identical_to_4x8": bool(torch.equal(out, ref))}), flush=True)
