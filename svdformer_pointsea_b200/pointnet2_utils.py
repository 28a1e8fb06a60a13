"""pointnet2_ops.pointnet2_utils, B200-native: same Function classes, names and signatures as
pointnet2_ops_lib/pointnet2_ops/pointnet2_utils.py:34-379 over the C ABI.

Behaviour kept from the reference: index outputs are int32 and non-differentiable
(:56, :267); gather/group outputs are fresh contiguous tensors that callers may modify in
place (models/model_utils.py:345); backward for `idx` returns zeros_like(idx) for grouping
(:237) and None for gather (:98).  Behaviour changed on purpose: argument errors raise
RuntimeError (PointSeaError) instead of AT_ASSERT / exit(-1).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L


# ---- raw ops (no autograd) -------------------------------------------------------------------
def fps_raw(xyz, npoint):
    L.require(xyz, "xyz", torch.float32, 3)
    if xyz.size(2) != 3:
        raise L.PointSeaError(f"xyz must be (B,N,3), got {tuple(xyz.shape)}")
    npoint = int(npoint)
    dev = L.same_device(xyz)
    B, N, _ = xyz.shape
    out = torch.empty(B, npoint, device=xyz.device, dtype=torch.int32)
    L.check(L.load().ps_fps(L.ptr(xyz), L.ptr(out), B, N, npoint, dev, L.stream_ptr(dev)), "ps_fps")
    return out


def gather_raw(features, idx):
    L.require(features, "features", torch.float32, 3)
    L.require(idx, "idx", torch.int32, 2)
    dev = L.same_device(features, idx)
    B, C, N = features.shape
    M = idx.size(1)
    out = torch.empty(B, C, M, device=features.device, dtype=torch.float32)
    L.check(L.load().ps_gather_fwd(L.ptr(features), L.ptr(idx), L.ptr(out), B, C, N, M, dev, L.stream_ptr(dev)),
            "ps_gather_fwd")
    return out


def gather_grad_raw(grad_out, idx, N):
    L.require(grad_out, "grad_out", torch.float32, 3)
    L.require(idx, "idx", torch.int32, 2)
    dev = L.same_device(grad_out, idx)
    B, C, M = grad_out.shape
    out = torch.empty(B, C, N, device=grad_out.device, dtype=torch.float32)
    L.check(L.load().ps_gather_bwd(L.ptr(grad_out), L.ptr(idx), L.ptr(out), B, C, N, M, dev, L.stream_ptr(dev)),
            "ps_gather_bwd")
    return out


def group_raw(features, idx):
    L.require(features, "features", torch.float32, 3)
    L.require(idx, "idx", torch.int32, 3)
    dev = L.same_device(features, idx)
    B, C, N = features.shape
    _, S, K = idx.shape
    out = torch.empty(B, C, S, K, device=features.device, dtype=torch.float32)
    L.check(L.load().ps_group_fwd(L.ptr(features), L.ptr(idx), L.ptr(out), B, C, N, S, K, dev, L.stream_ptr(dev)),
            "ps_group_fwd")
    return out


def group_grad_raw(grad_out, idx, N):
    L.require(grad_out, "grad_out", torch.float32, 4)
    L.require(idx, "idx", torch.int32, 3)
    dev = L.same_device(grad_out, idx)
    B, C, S, K = grad_out.shape
    out = torch.empty(B, C, N, device=grad_out.device, dtype=torch.float32)
    L.check(L.load().ps_group_bwd(L.ptr(grad_out), L.ptr(idx), L.ptr(out), B, C, N, S, K, dev, L.stream_ptr(dev)),
            "ps_group_bwd")
    return out


def ball_query_raw(new_xyz, xyz, radius, nsample):
    """Argument order of `_ext.ball_query` (ball_query.cpp:8-9)."""
    L.require(new_xyz, "new_xyz", torch.float32, 3)
    L.require(xyz, "xyz", torch.float32, 3)
    dev = L.same_device(new_xyz, xyz)
    B, S, _ = new_xyz.shape
    N = xyz.size(1)
    out = torch.empty(B, S, int(nsample), device=xyz.device, dtype=torch.int32)
    L.check(L.load().ps_ball_query(L.ptr(new_xyz), L.ptr(xyz), L.ptr(out), B, N, S, float(radius), int(nsample),
                                   dev, L.stream_ptr(dev)), "ps_ball_query")
    return out


def knn_raw(xyz, new_xyz, k, skip=0):
    L.require(xyz, "xyz", torch.float32, 3)
    L.require(new_xyz, "new_xyz", torch.float32, 3)
    dev = L.same_device(xyz, new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.size(1)
    out = torch.empty(B, S, int(k), device=xyz.device, dtype=torch.int32)
    L.check(L.load().ps_knn(L.ptr(xyz), L.ptr(new_xyz), L.ptr(out), B, N, S, int(k), int(skip), dev,
                            L.stream_ptr(dev)), "ps_knn")
    return out


def three_nn_raw(unknown, known):
    L.require(unknown, "unknown", torch.float32, 3)
    L.require(known, "known", torch.float32, 3)
    dev = L.same_device(unknown, known)
    B, n, _ = unknown.shape
    m = known.size(1)
    dist2 = torch.empty(B, n, 3, device=unknown.device, dtype=torch.float32)
    idx = torch.empty(B, n, 3, device=unknown.device, dtype=torch.int32)
    L.check(L.load().ps_three_nn(L.ptr(unknown), L.ptr(known), L.ptr(dist2), L.ptr(idx), B, n, m, dev,
                                 L.stream_ptr(dev)), "ps_three_nn")
    return dist2, idx


def three_interpolate_raw(features, idx, weight):
    L.require(features, "features", torch.float32, 3)
    L.require(idx, "idx", torch.int32, 3)
    L.require(weight, "weight", torch.float32, 3)
    dev = L.same_device(features, idx, weight)
    B, C, m = features.shape
    n = idx.size(1)
    out = torch.empty(B, C, n, device=features.device, dtype=torch.float32)
    L.check(L.load().ps_three_interpolate_fwd(L.ptr(features), L.ptr(idx), L.ptr(weight), L.ptr(out), B, C, m, n,
                                              dev, L.stream_ptr(dev)), "ps_three_interpolate_fwd")
    return out


def three_interpolate_grad_raw(grad_out, idx, weight, m):
    L.require(grad_out, "grad_out", torch.float32, 3)
    dev = L.same_device(grad_out, idx, weight)
    B, C, n = grad_out.shape
    out = torch.empty(B, C, m, device=grad_out.device, dtype=torch.float32)
    L.check(L.load().ps_three_interpolate_bwd(L.ptr(grad_out), L.ptr(idx), L.ptr(weight), L.ptr(out), B, C, n, m,
                                              dev, L.stream_ptr(dev)), "ps_three_interpolate_bwd")
    return out


# ---- autograd Functions: the reference's public surface ----------------------------------------
class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        out = fps_raw(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.N = features.size(2)
        return gather_raw(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return gather_grad_raw(grad_out.contiguous(), idx, ctx.N), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = three_nn_raw(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, grad_dist, grad_idx):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.m = features.size(2)
        return three_interpolate_raw(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        grad_features = three_interpolate_grad_raw(grad_out.contiguous(), idx, weight, ctx.m)
        return grad_features, torch.zeros_like(idx), torch.zeros_like(weight)


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        # only idx and N are needed for backward: the output is NOT saved, callers modify it
        # in place (models/model_utils.py:345)
        ctx.save_for_backward(idx)
        ctx.N = features.size(2)
        return group_raw(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return group_grad_raw(grad_out.contiguous(), idx, ctx.N), torch.zeros_like(idx)


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        output = ball_query_raw(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """pointnet2_utils.py:279-335: ball query + grouping (+ centre subtraction)."""

    def __init__(self, radius, nsample, use_xyz=True):
        super(QueryAndGroup, self).__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, nsample)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
            else:
                new_features = grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        return new_features


class GroupAll(nn.Module):
    """pointnet2_utils.py:338-379."""

    def __init__(self, use_xyz=True):
        super(GroupAll, self).__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
            else:
                new_features = grouped_features
        else:
            new_features = grouped_xyz
        return new_features


# ---- kNN: the torch expression of models/model_utils.py:281-286 as one kernel -------------------
def query_knn(nsample, xyz, new_xyz, include_self=True):
    """Find k-NN of new_xyz in xyz; same signature and result as models/model_utils.query_knn."""
    pad = 0 if include_self else 1
    return knn_raw(xyz.contiguous(), new_xyz.contiguous(), nsample, pad)


PS_FPS_CORUN = 1  # include/pointsea_b200.h


def fps_sample_raw(xyz, npoint, corun=False):
    """(idx (B,npoint) int32, new_xyz (B,npoint,3)) from ONE kernel: the FPS kernel writes the
    coordinates of every selected point as it goes (ps_fps_sample).  `corun`: another kernel runs next to this one
    (PS_FPS_CORUN: the launcher leaves most of each SM's shared memory to it); the samples are the same."""
    L.require(xyz, "xyz", torch.float32, 3)
    if xyz.size(2) != 3:
        raise L.PointSeaError(f"xyz must be (B,N,3), got {tuple(xyz.shape)}")
    npoint = int(npoint)
    dev = L.same_device(xyz)
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, device=xyz.device, dtype=torch.int32)
    new_xyz = torch.empty(B, npoint, 3, device=xyz.device, dtype=torch.float32)
    L.check(L.load().ps_fps_sample_ex(L.ptr(xyz), L.ptr(idx), L.ptr(new_xyz), B, N, npoint, PS_FPS_CORUN if corun else 0, dev,
                                      L.stream_ptr(dev)), "ps_fps_sample_ex")
    return idx, new_xyz


class _FpsSubsample(Function):
    """Fused fps_subsample with the reference's gradient (gather backward = scatter-add by idx)."""

    @staticmethod
    def forward(ctx, pcd, n_points):
        idx, new_pcd = fps_sample_raw(pcd.contiguous(), n_points)
        ctx.save_for_backward(idx)
        ctx.N = pcd.size(1)
        return new_pcd

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = gather_grad_raw(grad_out.permute(0, 2, 1).contiguous(), idx, ctx.N)  # (B,3,N)
        return g.permute(0, 2, 1).contiguous(), None


def fps_subsample(pcd, n_points=2048):
    """models/model_utils.py:489-499 (FPS + gather + two transposes) as one kernel launch."""
    return _FpsSubsample.apply(pcd, n_points)
