"""F-score from squared Chamfer distances; mirrors metrics/CD/fscore.py:3-16 (pure torch)."""
import torch


def fscore(dist1, dist2, threshold=0.0001):
    """dist1, dist2: (B, N) squared distances -> (fscore, precision_1, precision_2), each (B,)."""
    p1 = (dist1 < threshold).float().mean(dim=1)
    p2 = (dist2 < threshold).float().mean(dim=1)
    f = 2 * p1 * p2 / (p1 + p2)
    f[torch.isnan(f)] = 0
    return f, p1, p2
