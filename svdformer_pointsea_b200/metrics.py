"""Evaluation metrics of utils/loss_utils.py and metrics/CD/fscore.py on the sm_100a kernels.

    calc_cd   utils/loss_utils.py:98-115    fscore   metrics/CD/fscore.py:3-16    calc_dcd   utils/loss_utils.py:117-155

Same names, arguments and return lists as the reference.  One Chamfer forward + ONE epilogue kernel
(ps_chamfer_metrics) per call; per-cloud results, so the batch-sharded multi-GPU path needs no collective.
These are evaluation paths (the reference calls them under torch.no_grad()); no autograd.
"""
import torch

from . import _lib as L
from .chamfer import chamfer_forward

FSCORE_THRESHOLD = 0.0001  # metrics/CD/fscore.py:3


def chamfer_metrics_raw(dist1, dist2, idx1=None, idx2=None, threshold=FSCORE_THRESHOLD, alpha=1000.0, n_lambda=1.0,
                        frac1=1.0, frac2=1.0):
    """(B,8) float32: mean sqrt d1, mean sqrt d2, mean d1, mean d2, precision_1, precision_2, fscore, dcd."""
    L.require(dist1, "dist1", torch.float32, 2)
    L.require(dist2, "dist2", torch.float32, 2)
    ts = [dist1, dist2]
    if (idx1 is None) != (idx2 is None):
        raise L.PointSeaError("idx1 and idx2 must both be given or both be None")
    if idx1 is not None:
        L.require(idx1, "idx1", torch.int32, 2)
        L.require(idx2, "idx2", torch.int32, 2)
        ts += [idx1, idx2]
    dev = L.same_device(*ts)
    B, n1 = dist1.shape
    n2 = dist2.size(1)
    out = torch.empty(B, 8, device=dist1.device, dtype=torch.float32)
    L.check(L.load().ps_chamfer_metrics(L.ptr(dist1), L.ptr(dist2), L.ptr(idx1) if idx1 is not None else None,
                                        L.ptr(idx2) if idx2 is not None else None, L.ptr(out), B, n1, n2,
                                        float(threshold), float(alpha), float(n_lambda), float(frac1), float(frac2),
                                        dev, L.stream_ptr(dev)), "ps_chamfer_metrics")
    return out


def fscore(dist1, dist2, threshold=FSCORE_THRESHOLD):
    """metrics/CD/fscore.py:3-16 -> (fscore, precision_1, precision_2), each (B,)."""
    m = chamfer_metrics_raw(dist1.detach().contiguous(), dist2.detach().contiguous(), threshold=threshold)
    return m[:, 6], m[:, 4], m[:, 5]


def calc_cd(output, gt, calc_f1=False, return_raw=False, normalize=False, separate=False):
    """utils/loss_utils.py:98-115."""
    dist1, dist2, idx1, idx2 = chamfer_forward(gt.detach().contiguous(), output.detach().contiguous())
    m = chamfer_metrics_raw(dist1, dist2)
    cd_p = (m[:, 0] + m[:, 1]) / 2
    cd_t = m[:, 2] + m[:, 3]
    if separate:
        res = [torch.cat([m[:, 0].unsqueeze(0), m[:, 1].unsqueeze(0)]), torch.cat([m[:, 2].unsqueeze(0), m[:, 3].unsqueeze(0)])]
    else:
        res = [cd_p, cd_t]
    if calc_f1:
        res.append(m[:, 6])
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


def calc_dcd(x, gt, alpha=1000, n_lambda=1, return_raw=False, non_reg=False):
    """utils/loss_utils.py:117-155: density-aware Chamfer distance -> [loss (B,), cd_p, cd_t (, raw...)]."""
    x = x.float()
    gt = gt.float()
    n_x, n_gt = x.shape[1], gt.shape[1]
    assert x.shape[0] == gt.shape[0]
    if non_reg:
        frac_12 = max(1, n_x / n_gt)
        frac_21 = max(1, n_gt / n_x)
    else:
        frac_12 = n_x / n_gt
        frac_21 = n_gt / n_x
    # calc_cd(x, gt) -> cham_loss(gt, x): dist1/idx1 per gt point, dist2/idx2 per x point (:131-139)
    dist1, dist2, idx1, idx2 = chamfer_forward(gt.detach().contiguous(), x.detach().contiguous())
    m = chamfer_metrics_raw(dist1, dist2, idx1, idx2, alpha=alpha, n_lambda=n_lambda, frac1=frac_21, frac2=frac_12)
    res = [m[:, 7], (m[:, 0] + m[:, 1]) / 2, m[:, 2] + m[:, 3]]
    if return_raw:
        res.extend([dist1, dist2, idx1, idx2])
    return res


def patch_loss_utils(module):
    """Rebind calc_cd / calc_dcd / fscore inside an imported reference `utils.loss_utils` (evaluation loops call
    them through that module: core/test_pcn.py:64, core/eval_55.py:74)."""
    module.calc_cd = calc_cd
    module.calc_dcd = calc_dcd
    module.fscore = fscore
    return module
