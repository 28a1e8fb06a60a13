"""ncu target: feature-space kNN at EdgeConv(256,512,4)'s shape (B=32 C=256 N=512 k=4). usage: python tools/prof_fk.py"""
import sys, os.path as osp
sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))
import torch
from svdformer_pointsea_b200 import model_ops as mo
g = torch.Generator().manual_seed(1)
x = torch.randn(32, 256, 512, generator=g).cuda()
for _ in range(3):
    mo.knn_self(x, 4)
torch.cuda.synchronize()
print("done")
