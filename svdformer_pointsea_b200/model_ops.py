"""Call-site functions of models/model_utils.py (and models_PointSea/model_utils.py) on the sm_100a
kernels: same names, arguments and results as the reference's torch expressions, each one kernel.

    query_knn_point        models/model_utils.py:807-810   square_distance + topk(k, largest=False)
    index_points           :828-845                        advanced-indexing row gather
    group_local            :812-826                        kNN in feature space + index_points + permute
    edge_features          :869-877 (top of EdgeConv.forward): cat(central - neighbour, central)
    EdgeConv               :846-881                        the module, with the fused front
    sample_and_group_knn   :322-358                        FPS + gather + kNN + grouping + centre subtraction

`patch_model_utils(module)` rebinds these names (plus query_knn / fps_subsample) inside an imported
reference `models.model_utils`, so the models pick them up without edits.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib as L
from . import pointnet2_utils as pu

ORDER_SORT, ORDER_TOPK = 0, 1


# ---- raw ops ---------------------------------------------------------------------------------------
def knn_point_raw(xyz, new_xyz, k):
    """3-D coordinates, torch.topk result order; int32 (B,S,k)."""
    L.require(xyz, "xyz", torch.float32, 3)
    L.require(new_xyz, "new_xyz", torch.float32, 3)
    dev = L.same_device(xyz, new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.size(1)
    out = torch.empty(B, S, int(k), device=xyz.device, dtype=torch.int32)
    L.check(L.load().ps_knn_point(L.ptr(xyz), L.ptr(new_xyz), L.ptr(out), B, N, S, int(k), dev, L.stream_ptr(dev)),
            "ps_knn_point")
    return out


def knn_feat_raw(xr, xq, k, channel_major, order=ORDER_TOPK):
    """Feature-space kNN; xr/xq are (B,C,N)/(B,C,S) when channel_major else (B,N,C)/(B,S,C)."""
    L.require(xr, "xr", torch.float32, 3)
    L.require(xq, "xq", torch.float32, 3)
    dev = L.same_device(xr, xq)
    if channel_major:
        B, C, N = xr.shape
        S = xq.size(2)
        cq = xq.size(1)
    else:
        B, N, C = xr.shape
        S = xq.size(1)
        cq = xq.size(2)
    if cq != C or xq.size(0) != B:
        raise L.PointSeaError(f"feature kNN: mismatched shapes {tuple(xr.shape)} vs {tuple(xq.shape)}")
    out = torch.empty(B, S, int(k), device=xr.device, dtype=torch.int32)
    L.check(L.load().ps_knn_feat(L.ptr(xr), L.ptr(xq), L.ptr(out), B, C, N, S, int(k), 1 if channel_major else 0,
                                 int(order), dev, L.stream_ptr(dev)), "ps_knn_feat")
    return out


def knn_group_xyz_raw(xyz, new_xyz, k):
    """(idx (B,S,k) int32, grouped_xyz (B,3,S,k) = xyz[idx] - new_xyz) from one kernel."""
    L.require(xyz, "xyz", torch.float32, 3)
    L.require(new_xyz, "new_xyz", torch.float32, 3)
    dev = L.same_device(xyz, new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.size(1)
    idx = torch.empty(B, S, int(k), device=xyz.device, dtype=torch.int32)
    gxyz = torch.empty(B, 3, S, int(k), device=xyz.device, dtype=torch.float32)
    L.check(L.load().ps_knn_group_xyz(L.ptr(xyz), L.ptr(new_xyz), L.ptr(idx), L.ptr(gxyz), B, N, S, int(k), dev,
                                      L.stream_ptr(dev)), "ps_knn_group_xyz")
    return idx, gxyz


def edge_features_raw(x, idx):
    L.require(x, "x", torch.float32, 3)
    L.require(idx, "idx", torch.int32, 3)
    dev = L.same_device(x, idx)
    B, C, N = x.shape
    if idx.size(0) != B or idx.size(1) != N:
        raise L.PointSeaError(f"edge features: idx must be (B,N,K), got {tuple(idx.shape)} for x {tuple(x.shape)}")
    K = idx.size(2)
    out = torch.empty(B, 2 * C, N, K, device=x.device, dtype=torch.float32)
    L.check(L.load().ps_edge_features_fwd(L.ptr(x), L.ptr(idx), L.ptr(out), B, C, N, K, dev, L.stream_ptr(dev)),
            "ps_edge_features_fwd")
    return out


def edge_features_grad_raw(grad_out, idx):
    L.require(grad_out, "grad_out", torch.float32, 4)
    L.require(idx, "idx", torch.int32, 3)
    dev = L.same_device(grad_out, idx)
    B, C2, N, K = grad_out.shape
    C = C2 // 2
    out = torch.empty(B, C, N, device=grad_out.device, dtype=torch.float32)
    L.check(L.load().ps_edge_features_bwd(L.ptr(grad_out), L.ptr(idx), L.ptr(out), B, C, N, K, dev, L.stream_ptr(dev)),
            "ps_edge_features_bwd")
    return out


def index_points_raw(points, idx32):
    """points (B,N,C), idx32 (B,M) int32 -> (B,M,C)."""
    L.require(points, "points", torch.float32, 3)
    L.require(idx32, "idx", torch.int32, 2)
    dev = L.same_device(points, idx32)
    B, N, C = points.shape
    M = idx32.size(1)
    out = torch.empty(B, M, C, device=points.device, dtype=torch.float32)
    L.check(L.load().ps_index_points_fwd(L.ptr(points), L.ptr(idx32), L.ptr(out), B, N, M, C, dev, L.stream_ptr(dev)),
            "ps_index_points_fwd")
    return out


def index_points_grad_raw(grad_out, idx32, N):
    L.require(grad_out, "grad_out", torch.float32, 3)
    dev = L.same_device(grad_out, idx32)
    B, M, C = grad_out.shape
    out = torch.empty(B, N, C, device=grad_out.device, dtype=torch.float32)
    L.check(L.load().ps_index_points_bwd(L.ptr(grad_out), L.ptr(idx32), L.ptr(out), B, N, M, C, dev, L.stream_ptr(dev)),
            "ps_index_points_bwd")
    return out


# ---- the reference's functions -----------------------------------------------------------------------
def query_knn_point(k, xyz, new_xyz):
    """models/model_utils.py:807-810.  xyz (B,N,C) references, new_xyz (B,S,C) queries -> (B,S,k) int64
    (torch.topk returns int64 indices; callers index with them directly)."""
    xyz, new_xyz = xyz.detach().contiguous(), new_xyz.detach().contiguous()
    if xyz.size(2) == 3:
        idx = knn_point_raw(xyz, new_xyz, k)
    else:
        idx = knn_feat_raw(xyz, new_xyz, k, channel_major=False, order=ORDER_TOPK)
    return idx.long()


class _IndexPoints(Function):
    @staticmethod
    def forward(ctx, points, idx):
        idx32 = idx.reshape(idx.size(0), -1).to(torch.int32).contiguous()
        ctx.save_for_backward(idx32)
        ctx.N = points.size(1)
        out = index_points_raw(points.contiguous(), idx32)
        return out.view(*idx.shape, points.size(2))

    @staticmethod
    def backward(ctx, grad_out):
        (idx32,) = ctx.saved_tensors
        g = grad_out.contiguous().view(grad_out.size(0), -1, grad_out.size(-1))
        return index_points_grad_raw(g, idx32, ctx.N), None


def index_points(points, idx):
    """models/model_utils.py:828-845: points (B,N,C), idx (B,S) or (B,S,K) -> (B,S[,K],C)."""
    return _IndexPoints.apply(points, idx)


class _EdgeFeatures(Function):
    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        return edge_features_raw(x.contiguous(), idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return edge_features_grad_raw(grad_out.contiguous(), idx), None


def knn_self(x, k):
    """Neighbour indices (B,N,k) int32 of every point of x (B,C,N) among x itself, in the order
    query_knn_point / torch.topk gives (what group_local feeds EdgeConv)."""
    x = x.detach()
    if x.size(1) == 3:
        pts = x.transpose(1, 2).contiguous()
        return knn_point_raw(pts, pts, k)
    x = x.contiguous()
    return knn_feat_raw(x, x, k, channel_major=True, order=ORDER_TOPK)


def edge_features(x, k, idx=None):
    """Top of EdgeConv.forward (models/model_utils.py:869-877): x (B,C,N) -> (B,2C,N,k) =
    cat(central - neighbour, central) over the k nearest neighbours in feature space.  Two kernels
    (kNN, fused gather/subtract/concat) instead of the (B,N,N) matrix, topk and six tensor passes."""
    if idx is None:
        idx = knn_self(x, k)
    return _EdgeFeatures.apply(x, idx)


def group_local(xyz, k=20, return_idx=False):
    """models/model_utils.py:812-826: xyz (B,C,N) -> group_xyz (B,C,N,K) (+ idx (B,N,K) int64)."""
    idx = knn_self(xyz, k)
    group_xyz = pu.grouping_operation(xyz.contiguous(), idx)
    if return_idx:
        return group_xyz, idx.long()
    return group_xyz


class EdgeConv(nn.Module):
    """models/model_utils.py:846-881 with the fused front; parameters and state_dict keys are the reference's."""

    def __init__(self, input_channel, output_channel, k):
        super(EdgeConv, self).__init__()
        self.num_neigh = k
        self.conv = nn.Sequential(
            nn.Conv2d(2 * input_channel, output_channel // 2, kernel_size=1),
            nn.BatchNorm2d(output_channel // 2),
            nn.LeakyReLU(negative_slope=0.2),
            nn.Conv2d(output_channel // 2, output_channel // 2, kernel_size=1),
            nn.BatchNorm2d(output_channel // 2),
            nn.LeakyReLU(negative_slope=0.2),
            nn.Conv2d(output_channel // 2, output_channel, kernel_size=1)
        )

    def forward(self, inputs):
        if self.num_neigh is not None:
            feature = edge_features(inputs, self.num_neigh)
        else:
            batch_size, dims, num_points = inputs.shape
            central_feat = torch.zeros(batch_size, dims, num_points, 1, device=inputs.device)
            feature = torch.cat((central_feat - inputs.unsqueeze(-1), central_feat), dim=1)
        feature = self.conv(feature)
        return feature.max(dim=-1, keepdim=False)[0]


class _KnnGroupXyz(Function):
    """idx + (xyz[idx] - new_xyz) with the gradients of grouping_operation and the subtraction."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, k):
        # xyz (B,3,N), new_xyz (B,3,S) as sample_and_group_knn holds them
        pts = xyz.detach().permute(0, 2, 1).contiguous()
        ctr = new_xyz.detach().permute(0, 2, 1).contiguous()
        idx, gxyz = knn_group_xyz_raw(pts, ctr, k)
        ctx.save_for_backward(idx)
        ctx.N = xyz.size(2)
        ctx.mark_non_differentiable(idx)
        return idx, gxyz

    @staticmethod
    def backward(ctx, grad_idx, grad_gxyz):
        (idx,) = ctx.saved_tensors
        g = grad_gxyz.contiguous()
        return pu.group_grad_raw(g, idx, ctx.N), -g.sum(dim=3), None


def sample_and_group_knn(xyz, points, npoint, k, use_xyz=True, idx=None):
    """models/model_utils.py:322-358.  xyz (B,3,N), points (B,f,N) ->
    (new_xyz (B,3,npoint), new_points, idx (B,npoint,k) int32, grouped_xyz (B,3,npoint,k))."""
    xyz = xyz.contiguous()
    xyz_flipped = xyz.permute(0, 2, 1).contiguous()
    new_xyz = pu.gather_operation(xyz, pu.furthest_point_sample(xyz_flipped, npoint))
    if idx is None:
        idx, grouped_xyz = _KnnGroupXyz.apply(xyz, new_xyz, k)
    else:
        grouped_xyz = pu.grouping_operation(xyz, idx)
        grouped_xyz -= new_xyz.unsqueeze(3).repeat(1, 1, 1, k)
    if points is not None:
        grouped_points = pu.grouping_operation(points.contiguous(), idx)
        new_points = torch.cat([grouped_xyz, grouped_points], 1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def _with_reference_fallback(ours, reference):
    """`ours`, and the reference module's ORIGINAL torch expression when the C ABI answers PS_ERR_UNSUPPORTED (a
    shape outside what the kernels cover: k > 32, feature kNN beyond the shared-memory distance block, ...).  The
    reference's models accept those shapes, so a patched model must not crash on them; every other error, and a
    missing library, still raises."""
    if reference is None:
        return ours

    def call(*args, **kwargs):
        try:
            return ours(*args, **kwargs)
        except L.PointSeaError as e:
            if e.code != L.PS_ERR_UNSUPPORTED:
                raise
            return reference(*args, **kwargs)

    call.__name__ = getattr(ours, "__name__", "call")
    call.__doc__ = ours.__doc__
    call.pointsea_kernel, call.reference = ours, reference
    return call


def patch_model_utils(module):
    """Rebind the call-site functions inside an imported reference `models.model_utils` (or
    `models_PointSea.model_utils`).  Classes defined there (PointNet_SA_Module_KNN, EdgeConv users) look the
    names up in the module globals at call time, so no reference source changes.  Shapes the kernels do not cover
    (PS_ERR_UNSUPPORTED) are answered by the module's own original functions."""
    if getattr(module, "_pointsea_patched", False):
        return module
    for name, ours in (("query_knn", pu.query_knn), ("fps_subsample", pu.fps_subsample), ("query_knn_point", query_knn_point),
                       ("index_points", index_points), ("group_local", group_local), ("sample_and_group_knn", sample_and_group_knn)):
        setattr(module, name, _with_reference_fallback(ours, getattr(module, name, None)))
    if hasattr(module, "EdgeConv"):
        module.EdgeConv.forward = _with_reference_fallback(EdgeConv.forward, module.EdgeConv.forward)
    module._pointsea_patched = True
    return module
