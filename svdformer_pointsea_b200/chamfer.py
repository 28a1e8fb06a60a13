"""Chamfer distance: host-side mirror of metrics/CD/chamfer3D/dist_chamfer_3D.py.

Same names, argument meaning and outputs as the reference's `chamfer_3DFunction` /
`chamfer_3DDist` (dist_chamfer_3D.py:26-74); underneath, one C-ABI call per direction pair.
Differences that are deliberate: outputs are allocated on the device directly
(the reference allocates on the CPU and copies, :33-42), work is enqueued on the caller's
current stream (the reference uses the legacy default stream), errors raise.
"""
import ctypes

import torch
from torch import nn
from torch.autograd import Function

from . import _lib as L


def _check_out(t, name, dtype, shape, like):
    L.require(t, name, dtype, len(shape))
    if tuple(t.shape) != tuple(shape) or t.device != like.device:
        raise L.PointSeaError(f"{name} must be a {tuple(shape)} tensor on {like.device}, got {tuple(t.shape)} on {t.device}")
    return t


def chamfer_forward(xyz1, xyz2, out=None, sums=None):
    """(dist1, dist2, idx1, idx2) for xyz1 (B,N,3), xyz2 (B,M,3); no autograd.  `out` may carry the four
    preallocated outputs, as the reference's pybind `chamfer_3D.forward` takes them (chamfer_cuda.cpp:17-19).
    `sums`: a (6,) float64 CUDA tensor that receives the loss sums of `chamfer_sums` from the same launch that
    writes dist/idx (ps_chamfer_fwd_sums)."""
    L.require(xyz1, "xyz1", torch.float32, 3)
    L.require(xyz2, "xyz2", torch.float32, 3)
    if xyz1.size(2) != 3 or xyz2.size(2) != 3 or xyz1.size(0) != xyz2.size(0):
        raise L.PointSeaError(f"chamfer expects (B,N,3) and (B,M,3), got {tuple(xyz1.shape)} {tuple(xyz2.shape)}")
    dev = L.same_device(xyz1, xyz2)
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    if out is not None:
        dist1, dist2, idx1, idx2 = out
        _check_out(dist1, "dist1", torch.float32, (B, N), xyz1)
        _check_out(dist2, "dist2", torch.float32, (B, M), xyz1)
        _check_out(idx1, "idx1", torch.int32, (B, N), xyz1)
        _check_out(idx2, "idx2", torch.int32, (B, M), xyz1)
    else:
        dist1 = torch.empty(B, N, device=xyz1.device, dtype=torch.float32)
        dist2 = torch.empty(B, M, device=xyz1.device, dtype=torch.float32)
        idx1 = torch.empty(B, N, device=xyz1.device, dtype=torch.int32)
        idx2 = torch.empty(B, M, device=xyz1.device, dtype=torch.int32)
    if sums is not None:
        _check_out(sums, "sums", torch.float64, (6,), xyz1)
        rc = L.load().ps_chamfer_fwd_sums(L.ptr(xyz1), L.ptr(xyz2), L.ptr(dist1), L.ptr(dist2), L.ptr(idx1), L.ptr(idx2),
                                          L.ptr(sums), B, N, M, dev, L.stream_ptr(dev))
        L.check(rc, "ps_chamfer_fwd_sums")
        return dist1, dist2, idx1, idx2
    rc = L.load().ps_chamfer_fwd(L.ptr(xyz1), L.ptr(xyz2), L.ptr(dist1), L.ptr(dist2), L.ptr(idx1), L.ptr(idx2),
                                 B, N, M, dev, L.stream_ptr(dev))
    L.check(rc, "ps_chamfer_fwd")
    return dist1, dist2, idx1, idx2


class ChamferStep:
    """One training-style Chamfer step on DEVICE clouds as ONE launch (ps_chamfer_step): forward + loss sums
    [+ publish to the peers] + backward [+ world-wide sums], replayed from a CUDA graph.

    Buffers are allocated once for a shape; `__call__(xyz1, xyz2, graddist1, graddist2)` returns
    `(sums_local, sums_global, gradxyz1, gradxyz2)` — device tensors owned by this object (overwritten by the next
    call); `dist1/dist2/idx1/idx2` are attributes.  `comm`: a `dist.PeerComm` (then `sums_global` holds the sums over
    all ranks, identical bits on every rank) or None (`sums_global` is `sums_local`)."""

    def __init__(self, B, N, M, device, comm=None):
        dev = torch.device(device)
        self.B, self.N, self.M, self.device, self.comm = B, N, M, dev, comm
        self.dist1 = torch.empty(B, N, device=dev)
        self.dist2 = torch.empty(B, M, device=dev)
        self.idx1 = torch.empty(B, N, device=dev, dtype=torch.int32)
        self.idx2 = torch.empty(B, M, device=dev, dtype=torch.int32)
        self.gradxyz1 = torch.empty(B, N, 3, device=dev)
        self.gradxyz2 = torch.empty(B, M, 3, device=dev)
        self.sums_local = torch.zeros(6, device=dev, dtype=torch.float64)
        self.sums_global = torch.zeros(6, device=dev, dtype=torch.float64) if comm is not None else self.sums_local

    def _args(self, xyz1, xyz2, graddist1, graddist2):
        B, N, M = self.B, self.N, self.M
        L.require(xyz1, "xyz1", torch.float32, 3)
        L.require(xyz2, "xyz2", torch.float32, 3)
        if tuple(xyz1.shape) != (B, N, 3) or tuple(xyz2.shape) != (B, M, 3):
            raise L.PointSeaError(f"ChamferStep was built for ({B},{N},3) / ({B},{M},3), got {tuple(xyz1.shape)} {tuple(xyz2.shape)}")
        with_bwd = graddist1 is not None or graddist2 is not None
        if with_bwd:
            if graddist1 is None or graddist2 is None:
                raise L.PointSeaError("ChamferStep: backward needs both graddist1 and graddist2")
            L.require(graddist1, "graddist1", torch.float32, 2)
            L.require(graddist2, "graddist2", torch.float32, 2)
            dev = L.same_device(xyz1, xyz2, graddist1, graddist2, self.dist1)
        else:
            dev = L.same_device(xyz1, xyz2, self.dist1)
        args = (L.ptr(xyz1), L.ptr(xyz2), L.ptr(graddist1) if with_bwd else None, L.ptr(graddist2) if with_bwd else None,
                L.ptr(self.dist1), L.ptr(self.dist2), L.ptr(self.idx1), L.ptr(self.idx2),
                L.ptr(self.gradxyz1) if with_bwd else None, L.ptr(self.gradxyz2) if with_bwd else None,
                L.ptr(self.sums_local), L.ptr(self.sums_global) if self.comm is not None else None,
                self.comm.handle if self.comm is not None else None, B, N, M, dev)
        return args, dev, with_bwd

    def prepare(self, xyz1, xyz2, graddist1=None, graddist2=None):
        """Capture + instantiate the graph for these exact buffers without launching anything (ps_chamfer_step_prepare)."""
        args, _, _ = self._args(xyz1, xyz2, graddist1, graddist2)
        L.check(L.load().ps_chamfer_step_prepare(*args), "ps_chamfer_step_prepare")

    def __call__(self, xyz1, xyz2, graddist1=None, graddist2=None):
        args, dev, with_bwd = self._args(xyz1, xyz2, graddist1, graddist2)
        L.check(L.load().ps_chamfer_step(*args, L.stream_ptr(dev)), "ps_chamfer_step")
        return self.sums_local, self.sums_global, (self.gradxyz1 if with_bwd else None), (self.gradxyz2 if with_bwd else None)


def chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2, out=None):
    """(gradxyz1, gradxyz2); `out` may carry the two preallocated gradient buffers (they are overwritten, no
    zero-filling needed — the reference's `chamfer_3D.backward` needs them pre-zeroed, dist_chamfer_3D.py:56-60)."""
    L.require(graddist1, "graddist1", torch.float32, 2)
    L.require(graddist2, "graddist2", torch.float32, 2)
    L.require(idx1, "idx1", torch.int32, 2)
    L.require(idx2, "idx2", torch.int32, 2)
    dev = L.same_device(xyz1, xyz2, graddist1, graddist2, idx1, idx2)
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    if out is not None:
        gradxyz1, gradxyz2 = out
        _check_out(gradxyz1, "gradxyz1", torch.float32, (B, N, 3), xyz1)
        _check_out(gradxyz2, "gradxyz2", torch.float32, (B, M, 3), xyz1)
    else:
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.empty_like(xyz2)
    rc = L.load().ps_chamfer_bwd(L.ptr(xyz1), L.ptr(xyz2), L.ptr(graddist1), L.ptr(graddist2), L.ptr(idx1),
                                 L.ptr(idx2), L.ptr(gradxyz1), L.ptr(gradxyz2), B, N, M, dev, L.stream_ptr(dev))
    L.check(rc, "ps_chamfer_bwd")
    return gradxyz1, gradxyz2


def chamfer_sums(dist1, dist2, out=None):
    """One-launch reduction of Chamfer outputs: float64 tensor [sum sqrt(d1), sum sqrt(d2), sum d1,
    sum d2, numel(d1), numel(d2)] on the device (no host sync, no autograd) — the partial sums behind
    utils/loss_utils.py:10-31 and the only data the multi-GPU path all-reduces."""
    L.require(dist1, "dist1", torch.float32, dist1.dim())
    L.require(dist2, "dist2", torch.float32, dist2.dim())
    dev = L.same_device(dist1, dist2)
    if out is None:
        out = torch.empty(6, device=dist1.device, dtype=torch.float64)
    else:
        _check_out(out, "out", torch.float64, (6,), dist1)
    rc = L.load().ps_chamfer_sums(L.ptr(dist1), L.ptr(dist2), L.ptr(out), dist1.numel(), dist2.numel(), dev,
                                  L.stream_ptr(dev))
    L.check(rc, "ps_chamfer_sums")
    return out


def _require_host(t, name, dtype, shape):
    if not isinstance(t, torch.Tensor) or t.is_cuda:
        raise L.PointSeaError(f"{name} must be a CPU torch.Tensor (this is the host-buffer entry point)")
    if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise L.PointSeaError(f"{name} must be a contiguous {dtype} tensor of shape {tuple(shape)}, got {t.dtype} {tuple(t.shape)}")
    return t


def _host_args(who, xyz1, xyz2, graddist1, graddist2, out, device, sums_out, comm):
    """Shared validation of the host-buffer calls: returns (B, N, M, device index, out list, gradient pointers)."""
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.size(2) != 3 or xyz2.size(2) != 3 or xyz1.size(0) != xyz2.size(0):
        raise L.PointSeaError(f"chamfer expects (B,N,3) and (B,M,3), got {tuple(xyz1.shape)} {tuple(xyz2.shape)}")
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    _require_host(xyz1, "xyz1", torch.float32, (B, N, 3))
    _require_host(xyz2, "xyz2", torch.float32, (B, M, 3))
    with_bwd = graddist1 is not None or graddist2 is not None
    if with_bwd:
        if graddist1 is None or graddist2 is None:
            raise L.PointSeaError(f"{who}: backward needs both graddist1 and graddist2")
        _require_host(graddist1, "graddist1", torch.float32, (B, N))
        _require_host(graddist2, "graddist2", torch.float32, (B, M))
    if not torch.cuda.is_available():
        raise L.PointSeaError(f"{who} needs a CUDA device (there is no CPU implementation)")
    index = torch.cuda.current_device() if device is None else torch.device(device).index
    L._check_device(index)
    if out is None:
        out = [torch.empty(B, N, dtype=torch.float32).pin_memory(), torch.empty(B, M, dtype=torch.float32).pin_memory(),
               torch.empty(B, N, dtype=torch.int32).pin_memory(), torch.empty(B, M, dtype=torch.int32).pin_memory()]
        if with_bwd:
            out += [torch.empty(B, N, 3, dtype=torch.float32).pin_memory(), torch.empty(B, M, 3, dtype=torch.float32).pin_memory()]
    shapes = [((B, N), torch.float32), ((B, M), torch.float32), ((B, N), torch.int32), ((B, M), torch.int32),
              ((B, N, 3), torch.float32), ((B, M, 3), torch.float32)]
    if len(out) != (6 if with_bwd else 4):
        raise L.PointSeaError(f"{who}: `out` must hold {6 if with_bwd else 4} tensors")
    for t, (shape, dt), nm in zip(out, shapes, ("dist1", "dist2", "idx1", "idx2", "gradxyz1", "gradxyz2")):
        _require_host(t, nm, dt, shape)
    gp = [L.ptr(graddist1), L.ptr(graddist2), L.ptr(out[4]), L.ptr(out[5])] if with_bwd else [None] * 4
    if sums_out is not None or comm is not None:
        if sums_out is None:
            raise L.PointSeaError(f"{who}: `comm` needs `sums_out`")
        _require_host(sums_out, "sums_out", torch.float64, (6,))
    return B, N, M, index, out, gp


def chamfer_host(xyz1, xyz2, graddist1=None, graddist2=None, out=None, chunk=0, device=None, blocking=True, sums_out=None,
                 comm=None):
    """Chamfer forward (+ backward when graddist1/2 are given) on HOST tensors, pipelined through the GPU.

    The batch is cut into chunks of `chunk` clouds (0: library default); upload, kernels and download
    of consecutive chunks overlap on three streams (csrc/host_pipeline.cu).  Inputs and outputs are
    CPU tensors — pinned (`.pin_memory()`) for the overlap to happen; `out` may carry preallocated
    pinned outputs `(dist1, dist2, idx1, idx2[, gradxyz1, gradxyz2])`, otherwise they are allocated
    pinned here.  With `blocking=False` the outputs are valid once the current CUDA stream of
    `device` has been synchronised.  Results are bit-identical to chamfer_forward / chamfer_backward.
    `sums_out`: a pinned (6,) float64 CPU tensor that also receives the loss sums of `chamfer_sums`; with `comm` (a
    `dist.PeerComm`) they are the sums over all ranks' shards, exchanged by a kernel inside the same graph.
    """
    B, N, M, index, out, gp = _host_args("chamfer_host", xyz1, xyz2, graddist1, graddist2, out, device, sums_out, comm)
    if sums_out is not None:
        rc = L.load().ps_chamfer_host_full(L.ptr(xyz1), L.ptr(xyz2), L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(out[3]),
                                           gp[0], gp[1], gp[2], gp[3], L.ptr(sums_out), comm.handle if comm is not None else None,
                                           B, N, M, int(chunk), index, L.stream_ptr(index))
        L.check(rc, "ps_chamfer_host_full")
    else:
        rc = L.load().ps_chamfer_host(L.ptr(xyz1), L.ptr(xyz2), L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(out[3]),
                                      gp[0], gp[1], gp[2], gp[3], B, N, M, int(chunk), index, L.stream_ptr(index))
        L.check(rc, "ps_chamfer_host")
    if blocking:
        torch.cuda.current_stream(index).synchronize()
    return tuple(out)


class HostStep(object):
    """A step submitted by `chamfer_host_async`: `out` (and `sums`) are valid after `synchronize()` (host) or, for
    work queued on a CUDA stream, after `wait()`.  Keeps its buffers alive until then."""

    def __init__(self, ticket, index, out, sums, keep):
        self.ticket, self.index, self.out, self.sums, self._keep = ticket, index, out, sums, keep

    def wait(self, stream=None):
        """Makes `stream` (default: the current stream of the step's device) wait for the step's last output byte."""
        sp = L.stream_ptr(self.index) if stream is None else stream.cuda_stream
        L.check(L.load().ps_chamfer_host_wait(self.ticket, self.index, sp, 1, 0), "ps_chamfer_host_wait")
        return self

    def synchronize(self):
        """Blocks the calling thread until the step has written its last output byte; returns `out`."""
        L.check(L.load().ps_chamfer_host_wait(self.ticket, self.index, None, 0, 1), "ps_chamfer_host_wait")
        self._keep = None
        return self.out


def chamfer_host_async(xyz1, xyz2, graddist1=None, graddist2=None, out=None, chunk=0, device=None, sums_out=None, comm=None):
    """`chamfer_host` for loops that keep more than one step in flight: returns a `HostStep` at once.

    The step is ordered behind the current stream of `device` but runs on one of the library's four lanes: uploads
    and downloads on the lane's copy streams, the kernels (one replayed graph) on the stream that serves all lanes in
    submission order — so the copies of the neighbouring steps overlap the kernels of this one (a loader that
    prefetches the next batches while the results of an earlier one are read back; up to four steps in flight):

        pending = collections.deque()
        for i, batch in enumerate(loader):         # pinned host tensors, DEPTH + 1 buffer sets in rotation
            pending.append(chamfer_host_async(*batch, out=outs[i % (DEPTH + 1)], sums_out=sums[i % (DEPTH + 1)]))
            if len(pending) == DEPTH:              # DEPTH = 4: the loop runs at the speed of the busiest resource
                consume(pending.popleft().synchronize())

    Same arguments and results as `chamfer_host` (bit-identical); the buffers of a step must not be reused before
    its `synchronize()` / `wait()`.  With `comm` every lane exchanges on its own channel of the communicator: all ranks
    must submit the same sequence of steps."""
    B, N, M, index, out, gp = _host_args("chamfer_host_async", xyz1, xyz2, graddist1, graddist2, out, device, sums_out, comm)
    ticket = ctypes.c_longlong(0)
    rc = L.load().ps_chamfer_host_submit(L.ptr(xyz1), L.ptr(xyz2), L.ptr(out[0]), L.ptr(out[1]), L.ptr(out[2]), L.ptr(out[3]),
                                         gp[0], gp[1], gp[2], gp[3], L.ptr(sums_out) if sums_out is not None else None,
                                         comm.handle if comm is not None else None, B, N, M, int(chunk), index,
                                         L.stream_ptr(index), ctypes.byref(ticket))
    L.check(rc, "ps_chamfer_host_submit")
    return HostStep(ticket.value, index, tuple(out), sums_out, (xyz1, xyz2, graddist1, graddist2, out, sums_out))


def chamfer_host_step(xyz1, xyz2, graddist1=None, graddist2=None, grad_out=None, sums_out=None, chunk=0, device=None,
                      blocking=True, comm=None):
    """One training-style Chamfer step on HOST clouds: upload in chunks, forward, loss sums, backward — and only
    the six loss sums come back over PCIe.

    xyz1/xyz2 (and graddist1/2 for the backward) are CPU tensors, pinned for the overlap to happen.  Returns
    `(sums, gradxyz1, gradxyz2)`: `sums` a pinned float64 CPU tensor [sum sqrt d1, sum sqrt d2, sum d1, sum d2, n1, n2]
    (`sums_out` to reuse one), the gradients CUDA tensors on `device` (`grad_out=(g1, g2)` to reuse buffers; None
    without graddist).  With `blocking=False` everything is valid once the current stream of `device` has been
    synchronised.  Values equal chamfer_forward + chamfer_sums + chamfer_backward on the same clouds.  `comm`: a
    `dist.PeerComm`; `sums` then holds the sums over ALL ranks' shards (exchanged by a kernel inside the same graph)."""
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.size(2) != 3 or xyz2.size(2) != 3 or xyz1.size(0) != xyz2.size(0):
        raise L.PointSeaError(f"chamfer expects (B,N,3) and (B,M,3), got {tuple(xyz1.shape)} {tuple(xyz2.shape)}")
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    _require_host(xyz1, "xyz1", torch.float32, (B, N, 3))
    _require_host(xyz2, "xyz2", torch.float32, (B, M, 3))
    with_bwd = graddist1 is not None or graddist2 is not None
    if with_bwd:
        if graddist1 is None or graddist2 is None:
            raise L.PointSeaError("chamfer_host_step: backward needs both graddist1 and graddist2")
        _require_host(graddist1, "graddist1", torch.float32, (B, N))
        _require_host(graddist2, "graddist2", torch.float32, (B, M))
    if not torch.cuda.is_available():
        raise L.PointSeaError("chamfer_host_step needs a CUDA device (there is no CPU implementation)")
    index = torch.cuda.current_device() if device is None else torch.device(device).index
    L._check_device(index)
    dev = torch.device("cuda", index)
    if sums_out is None:
        sums_out = torch.empty(6, dtype=torch.float64).pin_memory()
    _require_host(sums_out, "sums_out", torch.float64, (6,))
    g1 = g2 = None
    if with_bwd:
        if grad_out is None:
            grad_out = (torch.empty(B, N, 3, device=dev), torch.empty(B, M, 3, device=dev))
        g1, g2 = grad_out
        for t, nm, shape in ((g1, "gradxyz1", (B, N, 3)), (g2, "gradxyz2", (B, M, 3))):
            L.require(t, nm, torch.float32, 3)
            if tuple(t.shape) != shape or t.device != dev:
                raise L.PointSeaError(f"{nm} must be a {shape} tensor on {dev}")
    if comm is not None:
        rc = L.load().ps_chamfer_host_step_dist(L.ptr(xyz1), L.ptr(xyz2), L.ptr(graddist1) if with_bwd else None,
                                                L.ptr(graddist2) if with_bwd else None, L.ptr(g1) if with_bwd else None,
                                                L.ptr(g2) if with_bwd else None, L.ptr(sums_out), comm.handle, B, N, M,
                                                int(chunk), index, L.stream_ptr(index))
    else:
        rc = L.load().ps_chamfer_host_step(L.ptr(xyz1), L.ptr(xyz2), L.ptr(graddist1) if with_bwd else None,
                                           L.ptr(graddist2) if with_bwd else None, L.ptr(g1) if with_bwd else None,
                                           L.ptr(g2) if with_bwd else None, L.ptr(sums_out), B, N, M, int(chunk), index,
                                           L.stream_ptr(index))
    L.check(rc, "ps_chamfer_host_step")
    if blocking:
        torch.cuda.current_stream(index).synchronize()
    return sums_out, g1, g2


class chamfer_3DFunction(Function):
    """Drop-in for dist_chamfer_3D.chamfer_3DFunction (dist_chamfer_3D.py:26-64)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        if graddist1 is None:
            graddist1 = torch.zeros(idx1.shape, device=xyz1.device, dtype=torch.float32)
        if graddist2 is None:
            graddist2 = torch.zeros(idx2.shape, device=xyz1.device, dtype=torch.float32)
        gradxyz1, gradxyz2 = chamfer_backward(xyz1, xyz2, graddist1.contiguous(), graddist2.contiguous(), idx1, idx2)
        return gradxyz1, gradxyz2


class chamfer_3DDist(nn.Module):
    """Drop-in for dist_chamfer_3D.chamfer_3DDist (dist_chamfer_3D.py:67-74)."""

    def __init__(self):
        super(chamfer_3DDist, self).__init__()

    def forward(self, input1, input2):
        input1 = input1.contiguous()
        input2 = input2.contiguous()
        return chamfer_3DFunction.apply(input1, input2)
