"""numpy front-end of oracle/pointsea_oracle.c (libpointsea_oracle.so) + pure-torch CPU baselines.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Two things live here:
  * bit-exact checker: thin ctypes wrappers over the C restatement (each C function cites the
    reference file:line it follows);
  * timed CPU baseline: the pure-PyTorch re-expression of each op the north star names as the
    CPU baseline — torch_chamfer, torch_fps, torch_knn, torch_group.
"""
import ctypes
import os
import os.path as osp
import subprocess

import numpy as np

_HERE = osp.dirname(osp.abspath(__file__))
_SO = osp.join(_HERE, "libpointsea_oracle.so")
_lib = None

_F = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i = ctypes.c_int


def build(force=False):
    src = osp.join(_HERE, "pointsea_oracle.c")
    if force or not osp.exists(_SO) or osp.getmtime(_SO) < osp.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpointsea_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not osp.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.or_chamfer_fwd.argtypes = [_F, _F, _F, _F, _I, _I, _i, _i, _i]
        L.or_chamfer_bwd.argtypes = [_F, _F, _F, _F, _I, _I, _F, _F, _i, _i, _i]
        L.or_fps.argtypes = [_F, _I, _i, _i, _i]
        L.or_opt_n_threads.argtypes = [_i]
        L.or_opt_n_threads.restype = _i
        L.or_gather.argtypes = [_F, _I, _F, _i, _i, _i, _i]
        L.or_gather_grad.argtypes = [_F, _I, _F, _i, _i, _i, _i]
        L.or_group.argtypes = [_F, _I, _F, _i, _i, _i, _i, _i]
        L.or_group_grad.argtypes = [_F, _I, _F, _i, _i, _i, _i, _i]
        L.or_ball_query.argtypes = [_F, _F, _I, _i, _i, _i, ctypes.c_float, _i]
        L.or_three_nn.argtypes = [_F, _F, _F, _I, _i, _i, _i]
        L.or_three_interpolate.argtypes = [_F, _I, _F, _F, _i, _i, _i, _i]
        L.or_three_interpolate_grad.argtypes = [_F, _I, _F, _F, _i, _i, _i, _i]
        L.or_knn.argtypes = [_F, _F, _I, _i, _i, _i, _i, _i, _i]
        for n in ("or_chamfer_fwd", "or_chamfer_bwd", "or_fps", "or_gather", "or_gather_grad", "or_group",
                  "or_group_grad", "or_ball_query", "or_three_nn", "or_three_interpolate",
                  "or_three_interpolate_grad", "or_knn"):
            getattr(L, n).restype = None
        _lib = L
    return _lib


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _n(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


# ---- bit-exact checker ---------------------------------------------------------------------------
def chamfer_fwd(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    d1, d2 = np.empty((B, N), np.float32), np.empty((B, M), np.float32)
    i1, i2 = np.empty((B, N), np.int32), np.empty((B, M), np.int32)
    lib().or_chamfer_fwd(xyz1, xyz2, d1, d2, i1, i2, B, N, M)
    return d1, d2, i1, i2


def chamfer_bwd(xyz1, xyz2, gd1, gd2, idx1, idx2):
    xyz1, xyz2, gd1, gd2, idx1, idx2 = _f(xyz1), _f(xyz2), _f(gd1), _f(gd2), _n(idx1), _n(idx2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    g1, g2 = np.empty_like(xyz1), np.empty_like(xyz2)
    lib().or_chamfer_bwd(xyz1, xyz2, gd1, gd2, idx1, idx2, g1, g2, B, N, M)
    return g1, g2


def fps(xyz, npoint):
    xyz = _f(xyz)
    B, N, _ = xyz.shape
    out = np.zeros((B, npoint), np.int32)
    lib().or_fps(xyz, out, B, N, npoint)
    return out


def opt_n_threads(n):
    return lib().or_opt_n_threads(int(n))


def gather(points, idx):
    points, idx = _f(points), _n(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), np.float32)
    lib().or_gather(points, idx, out, B, C, N, M)
    return out


def gather_grad(grad_out, idx, N):
    grad_out, idx = _f(grad_out), _n(idx)
    B, C, M = grad_out.shape
    out = np.empty((B, C, N), np.float32)
    lib().or_gather_grad(grad_out, idx, out, B, C, N, M)
    return out


def group(points, idx):
    points, idx = _f(points), _n(idx)
    B, C, N = points.shape
    _, S, K = idx.shape
    out = np.empty((B, C, S, K), np.float32)
    lib().or_group(points, idx, out, B, C, N, S, K)
    return out


def group_grad(grad_out, idx, N):
    grad_out, idx = _f(grad_out), _n(idx)
    B, C, S, K = grad_out.shape
    out = np.empty((B, C, N), np.float32)
    lib().or_group_grad(grad_out, idx, out, B, C, N, S, K)
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, xyz = _f(new_xyz), _f(xyz)
    B, S, _ = new_xyz.shape
    N = xyz.shape[1]
    out = np.empty((B, S, nsample), np.int32)
    lib().or_ball_query(new_xyz, xyz, out, B, N, S, float(radius), nsample)
    return out


def three_nn(unknown, known):
    unknown, known = _f(unknown), _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d, i = np.empty((B, n, 3), np.float32), np.empty((B, n, 3), np.int32)
    lib().or_three_nn(unknown, known, d, i, B, n, m)
    return d, i


def three_interpolate(points, idx, weight):
    points, idx, weight = _f(points), _n(idx), _f(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), np.float32)
    lib().or_three_interpolate(points, idx, weight, out, B, C, m, n)
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f(grad_out), _n(idx), _f(weight)
    B, C, n = grad_out.shape
    out = np.empty((B, C, m), np.float32)
    lib().or_three_interpolate_grad(grad_out, idx, weight, out, B, C, n, m)
    return out


KNN_VARIANT = int(os.environ.get("PS_KNN_VARIANT", "0"))


def knn(xyz, new_xyz, k, skip=0, variant=None):
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    out = np.empty((B, S, k), np.int32)
    lib().or_knn(xyz, new_xyz, out, B, N, S, k, skip, KNN_VARIANT if variant is None else variant)
    return out


# ---- pure-PyTorch CPU baselines (timed, not bit-exact: torch has no fused multiply-add) ----------
def torch_chamfer(xyz1, xyz2):
    """Direct-difference Chamfer, chunked per cloud."""
    import torch
    d1s, d2s, i1s, i2s = [], [], [], []
    for a, b in zip(xyz1, xyz2):
        diff = a[:, None, :] - b[None, :, :]
        P = (diff * diff).sum(-1)
        d1, i1 = P.min(1)
        d2, i2 = P.min(0)
        d1s.append(d1); d2s.append(d2); i1s.append(i1.int()); i2s.append(i2.int())
    return torch.stack(d1s), torch.stack(d2s), torch.stack(i1s), torch.stack(i2s)


def torch_chamfer_fwd_bwd(xyz1, xyz2, gd1, gd2):
    """Forward + backward through autograd on the direct form (the timed CPU baseline step)."""
    import torch
    a = xyz1.detach().clone().requires_grad_(True)
    b = xyz2.detach().clone().requires_grad_(True)
    d1, d2, i1, i2 = torch_chamfer(a, b)
    ((d1 * gd1).sum() + (d2 * gd2).sum()).backward()
    return d1.detach(), d2.detach(), i1, i2, a.grad, b.grad


def torch_fps(xyz, npoint):
    """npoint-step loop of batched torch ops incl. the origin-skip rule (sampling_gpu.cu:100-101)."""
    import torch
    B, N, _ = xyz.shape
    mag = (xyz * xyz).sum(-1)
    elig = ~(mag.double() <= 1e-3)
    temp = torch.full((B, N), 1e10, dtype=xyz.dtype)
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    old = torch.zeros(B, dtype=torch.long)
    ar = torch.arange(B)
    for j in range(1, npoint):
        last = xyz[ar, old][:, None, :]
        d = ((xyz - last) ** 2).sum(-1)
        temp = torch.where(elig, torch.minimum(temp, d), temp)
        old = torch.where(elig, temp, torch.full_like(temp, -1.0)).argmax(1)
        idx[:, j] = old.int()
    return idx


def torch_knn(nsample, xyz, new_xyz, include_self=True):
    """query_knn verbatim (models/model_utils.py:258-286)."""
    import torch
    pad = 0 if include_self else 1
    B, S, _ = new_xyz.shape
    N = xyz.shape[1]
    dist = -2 * torch.matmul(new_xyz, xyz.permute(0, 2, 1))
    dist += torch.sum(new_xyz ** 2, -1).view(B, S, 1)
    dist += torch.sum(xyz ** 2, -1).view(B, 1, N)
    idx = torch.argsort(dist, dim=-1, descending=False)[:, :, pad: nsample + pad]
    return idx.int()


def torch_group(features, idx):
    """Advanced-indexing gather as in index_points (models/model_utils.py:828-845)."""
    import torch
    B, C, N = features.shape
    _, S, K = idx.shape
    flat = idx.long().reshape(B, 1, S * K).expand(B, C, S * K)
    return torch.gather(features, 2, flat).reshape(B, C, S, K)


def torch_gather(features, idx):
    import torch
    B, C, N = features.shape
    return torch.gather(features, 2, idx.long()[:, None, :].expand(B, C, idx.shape[1]))


# ---- "next" rows: torch re-expressions of the reference call sites -----------------
def torch_square_distance(src, dst):
    """models/model_utils.py:258-279 verbatim (expanded form through matmul)."""
    import torch
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def torch_query_knn_point(k, xyz, new_xyz):
    """models/model_utils.py:807-810: topk(k, largest=False) of square_distance; (B,S,k) int64."""
    dist = torch_square_distance(new_xyz, xyz)
    _, group_idx = dist.topk(k, largest=False)
    return group_idx


def torch_index_points(points, idx):
    """models/model_utils.py:828-845: points (B,N,C), idx (B,S[,K]) -> (B,S[,K],C)."""
    import torch
    B = points.shape[0]
    view_shape = list(idx.shape)
    view_shape[1:] = [1] * (len(view_shape) - 1)
    repeat_shape = list(idx.shape)
    repeat_shape[0] = 1
    batch_indices = torch.arange(B, dtype=torch.long, device=points.device).view(view_shape).repeat(repeat_shape)
    return points[batch_indices, idx, :]


def torch_edge_features(x, k, idx=None):
    """EdgeConv up to the convolution (models/model_utils.py:812-826, 869-877): x (B,C,N) ->
    (feature (B,2C,N,k) = cat(central - neighbour, central), idx (B,N,k) int64)."""
    import torch
    xt = x.transpose(2, 1).contiguous()
    if idx is None:
        idx = torch_query_knn_point(k, xt, xt)
    neigh = torch_index_points(xt, idx).permute(0, 3, 1, 2).contiguous()
    central = x.unsqueeze(dim=3).repeat(1, 1, 1, k)
    edge = central - neigh
    return torch.cat((edge, central), dim=1), idx


def torch_fscore(dist1, dist2, threshold=0.0001):
    """metrics/CD/fscore.py:3-16 verbatim."""
    import torch
    precision_1 = torch.mean((dist1 < threshold).float(), dim=1)
    precision_2 = torch.mean((dist2 < threshold).float(), dim=1)
    fscore = 2 * precision_1 * precision_2 / (precision_1 + precision_2)
    fscore[torch.isnan(fscore)] = 0
    return fscore, precision_1, precision_2


def torch_cd_terms(dist1, dist2):
    """calc_cd's reductions (utils/loss_utils.py:102-103): per-cloud cd_p, cd_t."""
    import torch
    cd_p = (torch.sqrt(dist1).mean(1) + torch.sqrt(dist2).mean(1)) / 2
    cd_t = dist1.mean(1) + dist2.mean(1)
    return cd_p, cd_t


def torch_dcd_from_raw(dist1, dist2, idx1, idx2, n_x, n_gt, alpha=1000, n_lambda=1, non_reg=False):
    """calc_dcd after its Chamfer call (utils/loss_utils.py:117-155): dist1/idx1 are per gt point
    (nearest x), dist2/idx2 per x point (nearest gt).  Returns the per-cloud loss (B,)."""
    import torch
    if non_reg:
        frac_12 = max(1, n_x / n_gt)
        frac_21 = max(1, n_gt / n_x)
    else:
        frac_12 = n_x / n_gt
        frac_21 = n_gt / n_x
    exp_dist1, exp_dist2 = torch.exp(-dist1 * alpha), torch.exp(-dist2 * alpha)
    count1 = torch.zeros_like(idx2)
    count1.scatter_add_(1, idx1.long(), torch.ones_like(idx1))
    weight1 = count1.gather(1, idx1.long()).float().detach() ** n_lambda
    weight1 = (weight1 + 1e-6) ** (-1) * frac_21
    loss1 = (1 - exp_dist1 * weight1).mean(dim=1)
    count2 = torch.zeros_like(idx1)
    count2.scatter_add_(1, idx2.long(), torch.ones_like(idx2))
    weight2 = count2.gather(1, idx2.long()).float().detach() ** n_lambda
    weight2 = (weight2 + 1e-6) ** (-1) * frac_12
    loss2 = (1 - exp_dist2 * weight2).mean(dim=1)
    return (loss1 + loss2) / 2


# ---- bit-exact checker for the "next" rows (C restatement) ------------------------------------------
def _next_lib():
    L = lib()
    if not getattr(L, "_next_ready", False):
        _fo = ctypes.c_float
        L.or_knn_feat.argtypes = [_F, _F, _I, _i, _i, _i, _i, _i, _i]
        L.or_edge_features.argtypes = [_F, _I, _F, _i, _i, _i, _i]
        L.or_edge_features_grad.argtypes = [_F, _I, _F, _i, _i, _i, _i]
        L.or_index_points.argtypes = [_F, _I, _F, _i, _i, _i, _i]
        L.or_index_points_grad.argtypes = [_F, _I, _F, _i, _i, _i, _i]
        L.or_chamfer_metrics.argtypes = [_F, _F, ctypes.c_void_p, ctypes.c_void_p, _F, _i, _i, _i, _fo, _fo, _fo, _fo, _fo]
        for n in ("or_knn_feat", "or_edge_features", "or_edge_features_grad", "or_index_points", "or_index_points_grad",
                  "or_chamfer_metrics"):
            getattr(L, n).restype = None
        L._next_ready = True
    return L


def knn_feat(xr, xq, k, order=0):
    """xr (B,N,C), xq (B,S,C) point-major -> (B,S,k) int32; order 0 = (dist, index), 1 = torch.topk order."""
    xr, xq = _f(xr), _f(xq)
    B, N, C = xr.shape
    S = xq.shape[1]
    out = np.empty((B, S, k), np.int32)
    _next_lib().or_knn_feat(xr, xq, out, B, C, N, S, k, order)
    return out


def knn_point(xyz, new_xyz, k):
    """query_knn_point on coordinates: torch.topk order."""
    return knn_feat(xyz, new_xyz, k, order=1)


def knn_group_xyz(xyz, new_xyz, k):
    """sample_and_group_knn's kNN + grouping + centre subtraction (models/model_utils.py:342-345)."""
    xyz, new_xyz = _f(xyz), _f(new_xyz)
    idx = knn(xyz, new_xyz, k)
    g = group(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx)  # (B,3,S,k)
    g = g - np.ascontiguousarray(new_xyz.transpose(0, 2, 1))[:, :, :, None]
    return idx, g.astype(np.float32)


def edge_features(x, idx):
    x, idx = _f(x), _n(idx)
    B, C, N = x.shape
    K = idx.shape[2]
    out = np.empty((B, 2 * C, N, K), np.float32)
    _next_lib().or_edge_features(x, idx, out, B, C, N, K)
    return out


def edge_features_grad(gout, idx):
    gout, idx = _f(gout), _n(idx)
    B, C2, N, K = gout.shape
    out = np.empty((B, C2 // 2, N), np.float32)
    _next_lib().or_edge_features_grad(gout, idx, out, B, C2 // 2, N, K)
    return out


def index_points(points, idx):
    points, idx = _f(points), _n(idx)
    B, N, C = points.shape
    flat = np.ascontiguousarray(idx.reshape(B, -1))
    out = np.empty((B, flat.shape[1], C), np.float32)
    _next_lib().or_index_points(points, flat, out, B, N, flat.shape[1], C)
    return out.reshape(*idx.shape, C)


def index_points_grad(gout, idx, N):
    gout, idx = _f(gout), _n(idx)
    B = gout.shape[0]
    C = gout.shape[-1]
    flat = np.ascontiguousarray(idx.reshape(B, -1))
    g = np.ascontiguousarray(gout.reshape(B, -1, C))
    out = np.empty((B, N, C), np.float32)
    _next_lib().or_index_points_grad(g, flat, out, B, N, flat.shape[1], C)
    return out


def chamfer_metrics(dist1, dist2, idx1=None, idx2=None, threshold=0.0001, alpha=1000.0, n_lambda=1.0, frac1=1.0, frac2=1.0):
    dist1, dist2 = _f(dist1), _f(dist2)
    B, n1 = dist1.shape
    n2 = dist2.shape[1]
    out = np.empty((B, 8), np.float32)
    p1 = p2 = None
    if idx1 is not None:
        idx1, idx2 = _n(idx1), _n(idx2)
        p1, p2 = idx1.ctypes.data, idx2.ctypes.data
    _next_lib().or_chamfer_metrics(dist1, dist2, p1, p2, out, B, n1, n2, threshold, alpha, n_lambda, frac1, frac2)
    return out
