"""Batch sharding + the single loss/metric reduction.

Every op on the hot path is independent per cloud, so ranks never exchange points: rank r owns
clouds [B*r//G, B*(r+1)//G).  The only collective is ONE all-reduce(sum) of a small vector of
partial sums and element counts (utils/loss_utils.py:10-19,50-57 define what is summed); the
division happens after the reduce so uneven shards stay exact.  Backend: NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests.

Gradients.  The global mean is a function of every rank's shard, and each rank back-propagates only through its
own shard, so a rank's parameter gradient is its SHARE of the full-batch gradient: the shares must be SUMMED
across ranks (`grad_reduce="sum"`, the default; e.g. a manual all-reduce(sum) of the gradients).  Stock
`DistributedDataParallel` AVERAGES gradients instead, which would give 1/world of the reference's
(`nn.DataParallel`, full batch on one process) gradient; pass `grad_reduce="mean"` to the sharded losses under
DDP: the backward of the reduction is then scaled by the world size, so the DDP average equals the full-batch
gradient.  INTEGRATION.md section "Multi-GPU" says the same.
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib as L


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to GPU `device_index` (its NUMA node).
    With one process per GPU on a multi-socket host, pinned host buffers are then first-touched on the
    socket the GPU hangs off, so host<->device copies of different ranks do not cross the inter-socket link.
    Call before allocating pinned memory.  Returns the CPU set, or None when NVML / affinity is unavailable
    (containers may forbid it) — never raises."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            try:
                phys = int(vis.split(",")[device_index])
            except Exception:
                phys = device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or None
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_bounds(B, rank, world):
    """Clouds [lo, hi) of a batch of B owned by `rank` out of `world` (uneven shards allowed)."""
    return (B * rank) // world, (B * (rank + 1)) // world


def shard_batch(t, rank=None, world=None):
    """Contiguous batch shard of a tensor whose dim 0 is the cloud index."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(t.size(0), rank, world)
    return t[lo:hi].contiguous()


class PeerComm:
    """The path's single collective over PEER MEMORY (csrc/comm.cuh): every rank owns a mailbox in its HBM that the
    peers store into over NVLink; the kernel that produces the loss sums publishes them, a one-warp kernel adds the
    world's contributions in rank order (identical bits on every rank).  No host call per step, graph-replayable.

    `PeerComm(group)`: one rank per process; the CUDA IPC handles of the mailboxes are exchanged once through
    `torch.distributed.all_gather_object` on `group` (any backend).  `PeerComm.local(devices)`: all ranks in this
    process (threads / tests), one per entry of `devices`."""

    def __init__(self, group=None, device=None, _handle=None, _rank=0, _world=1):
        lib = L.load()
        if _handle is not None:
            self.handle, self.rank, self.world, self.device = _handle, _rank, _world, device
            return
        if not torch.cuda.is_available():
            raise L.PointSeaError("PeerComm needs a CUDA device (the exchange is a CUDA kernel over peer memory)")
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        index = torch.cuda.current_device() if device is None else torch.device(device).index
        L._check_device(index)
        self.device = torch.device("cuda", index)
        h = ctypes.c_void_p()
        L.check(lib.ps_comm_create(self.rank, self.world, index, ctypes.byref(h)), "ps_comm_create")
        self.handle = h
        if self.world > 1:
            nb = lib.ps_comm_handle_bytes()
            mine = ctypes.create_string_buffer(nb)
            L.check(lib.ps_comm_export(self.handle, mine), "ps_comm_export")
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine.raw), group=group)
            blob = ctypes.create_string_buffer(b"".join(gathered), nb * self.world)
            L.check(lib.ps_comm_connect(self.handle, blob), "ps_comm_connect")
            dist.barrier(group=group)  # nobody publishes before every mailbox is mapped everywhere

    @classmethod
    def local(cls, devices):
        lib = L.load()
        world = len(devices)
        handles = (ctypes.c_void_p * world)()
        comms = []
        for r, d in enumerate(devices):
            index = torch.device(d).index
            L._check_device(index)
            h = ctypes.c_void_p()
            L.check(lib.ps_comm_create(r, world, index, ctypes.byref(h)), "ps_comm_create")
            handles[r] = h
            comms.append(cls(device=torch.device("cuda", index), _handle=h, _rank=r, _world=world))
        L.check(lib.ps_comm_connect_local(handles, world), "ps_comm_connect_local")
        return comms

    def all_reduce(self, vec, out=None):
        """Sum of `vec` (float64 CUDA tensor, <= 30 elements) over all ranks, on the current stream."""
        L.require(vec, "vec", torch.float64, 1)
        if out is None:
            out = torch.empty_like(vec)
        index = L.same_device(vec, out)
        L.check(L.load().ps_comm_allreduce(self.handle, L.ptr(vec), L.ptr(out), vec.numel(), L.stream_ptr(index)), "ps_comm_allreduce")
        return out

    def status(self):
        p, c, t = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_int()
        L.check(L.load().ps_comm_status(self.handle, ctypes.byref(p), ctypes.byref(c), ctypes.byref(t)), "ps_comm_status")
        return {"published": p.value, "consumed": c.value, "timed_out": bool(t.value)}

    def close(self):
        if self.handle is not None:
            L.load().ps_comm_destroy(self.handle)
            self.handle = None


def _world(group, comm):
    if comm is not None:
        return comm.world
    return dist.get_world_size(group) if dist.is_initialized() else 1


class _AllReduceSum(torch.autograd.Function):
    """all-reduce(sum) that autograd can cross.  Every rank holds the same replicated loss and back-propagates
    through its own shard only, so the backward is the identity times `scale`: 1 when the caller SUMS parameter
    gradients across ranks, world_size when they are AVERAGED (stock DistributedDataParallel) — see the module
    docstring."""

    @staticmethod
    def forward(ctx, vec, group, comm, scale):
        ctx.scale = scale
        if comm is not None:
            return comm.all_reduce(vec.detach().to(torch.float64).contiguous()).to(vec.dtype)
        out = vec.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        return (grad if ctx.scale == 1 else grad * ctx.scale), None, None, None


def _grad_scale(grad_reduce, world):
    if grad_reduce == "sum":
        return 1
    if grad_reduce == "mean":
        return world
    raise ValueError(f"grad_reduce must be 'sum' or 'mean', got {grad_reduce!r}")


class LossSums:
    """Accumulates named (sum, count) pairs on the device and reduces them with one all-reduce.

    add("cd2.d1", values) records sum(values) and values.numel(); reduce() all-reduces the whole
    vector once (enqueued on the current stream, no host sync) and returns {name: global mean}.
    `comm`: a PeerComm — the exchange then runs over peer memory instead of torch.distributed.
    """

    def __init__(self, device, dtype=torch.float32):
        self.device, self.dtype = device, dtype
        self.names, self.parts = [], []

    def add(self, name, values):
        self.names.append(name)
        self.parts.append(values.sum(dtype=self.dtype).reshape(1))
        self.parts.append(torch.full((1,), float(values.numel()), device=self.device, dtype=self.dtype))

    def reduce(self, group=None, comm=None, grad_reduce="sum"):
        vec = torch.cat(self.parts) if self.parts else torch.zeros(0, device=self.device, dtype=self.dtype)
        world = _world(group, comm)
        scale = _grad_scale(grad_reduce, world)
        if world > 1 and vec.numel():
            vec = _AllReduceSum.apply(vec, group, comm, scale)
        out = {}
        for i, name in enumerate(self.names):
            out[name] = vec[2 * i] / vec[2 * i + 1]
        return out


def chamfer_metric_means(d1, d2, group=None, comm=None):
    """Global means {sqrt_d1, sqrt_d2, d1, d2} of Chamfer outputs over all ranks' shards: one fused
    reduction kernel (ps_chamfer_sums) + ONE all-reduce of 6 doubles.  Metric path (no autograd)."""
    from .chamfer import chamfer_sums
    vec = chamfer_sums(d1, d2)  # [sum sqrt d1, sum sqrt d2, sum d1, sum d2, n1, n2]
    if comm is not None and comm.world > 1:
        vec = comm.all_reduce(vec)
    elif dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return {"sqrt_d1": vec[0] / vec[4], "sqrt_d2": vec[1] / vec[5], "d1": vec[2] / vec[4], "d2": vec[3] / vec[5]}


class PipelinedSums:
    """The same single all-reduce, taken off the critical path of the step.

    The reduced loss/metric sums of step i are an OUTPUT of the step (logging, schedulers); nothing the
    GPU runs in step i consumes them — the backward's upstream gradients are constants.  So the
    all-reduce is issued asynchronously (NCCL's own stream, ordered after the producing kernel) and
    joined one step later: submit(vec) enqueues the reduction of this step and returns the reduced
    vector of the PREVIOUS step (None on the first call); flush() joins the outstanding one.  Joining
    never blocks the host: it makes the caller's stream wait on the collective's completion event.
    With G ranks this removes the per-step rendezvous from the compute stream: a rank that arrives
    early keeps computing instead of idling inside the collective (measured on 8 B200s: the
    in-line all-reduce cost 0.15 ms per 0.38 ms step in rank skew).
    """

    def __init__(self, group=None):
        self.group = group
        self.pending = None  # (work, vec)

    def _join(self):
        if self.pending is None:
            return None
        work, vec = self.pending
        self.pending = None
        if work is not None:
            work.wait()
        return vec

    def submit(self, vec):
        prev = self._join()
        work = None
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            work = dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending = (work, vec)
        return prev

    def flush(self):
        return self._join()


def chamfer_loss_terms(sums, name, d1, d2, sqrt=True):
    """Registers the two directional means of one Chamfer term (chamfer / chamfer_sqrt,
    utils/loss_utils.py:10-19)."""
    sums.add(name + ".d1", torch.sqrt(d1) if sqrt else d1)
    sums.add(name + ".d2", torch.sqrt(d2) if sqrt else d2)


def combine_chamfer(means, name, sqrt=True):
    m1, m2 = means[name + ".d1"], means[name + ".d2"]
    return (m1 + m2) / 2 if sqrt else m1 + m2


_SIDE_STREAMS = {}


def _side_stream(device):
    """One cached side stream per device for work that is independent of the main stream's next kernels."""
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def _sharded_terms(pcds_pred, gt, sqrt, partial=None, overlap_fps=None):
    """The Chamfer terms of get_loss / get_loss_PM on this rank's shard, as named (sum, count) pairs.

    The two FPS calls that thin the ground truth (gt -> |P1| -> |Pc| points, utils/loss_utils.py:40-41) are a serial
    chain of ~2500 latency-bound iterations on a quarter of the SMs' issue slots, and the largest Chamfer term
    (P2 vs the full gt) does not depend on them: with `overlap_fps` (default: on CUDA) the FPS chain runs on a side
    stream while the main stream computes that term.  Values are unchanged."""
    from .chamfer import chamfer_3DFunction
    from .pointnet2_utils import fps_subsample, fps_sample_raw

    Pc, P1, P2 = pcds_pred
    if overlap_fps is None:  # PS_LOSS_OVERLAP=0 keeps everything on the current stream (A/B measurements)
        overlap_fps = gt.is_cuda and os.environ.get("PS_LOSS_OVERLAP", "1") != "0"
    overlap_fps = bool(overlap_fps) and gt.is_cuda and not gt.requires_grad  # the side-stream chain runs without autograd
    capturing = gt.is_cuda and torch.cuda.is_current_stream_capturing()
    sums = LossSums(gt.device)

    def term(name, p, q):
        d1, d2, _, _ = chamfer_3DFunction.apply(p.contiguous(), q.contiguous())
        chamfer_loss_terms(sums, name, d1, d2, sqrt)

    if overlap_fps:
        cur, side = torch.cuda.current_stream(gt.device), _side_stream(gt.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            # corun: leave the SMs' shared memory to the Chamfer term that runs next to the chain
            corun = os.environ.get("PS_LOSS_CORUN", "1") != "0"
            gt_1 = fps_sample_raw(gt.contiguous(), P1.shape[1], corun=corun)[1]
            have_1 = torch.cuda.Event()
            have_1.record(side)
            gt_c = fps_sample_raw(gt_1, Pc.shape[1], corun=corun)[1]
        term("cd2", P2, gt)  # meanwhile, on the main stream
        if os.environ.get("PS_LOSS_CD1_EARLY", "1") == "0":  # A/B: join the whole chain first
            cur.wait_stream(side)
        cur.wait_event(have_1)  # (inside a stream capture the fork and the joins become edges of the graph)
        if not capturing:  # a capture's private pool never hands the blocks to anyone else
            gt_1.record_stream(cur)
        term("cd1", P1, gt_1)  # under the second FPS call
        cur.wait_stream(side)
        if not capturing:
            gt_c.record_stream(cur)
        term("cdc", Pc, gt_c)
    else:
        gt_1 = fps_subsample(gt, P1.shape[1])
        gt_c = fps_subsample(gt_1, Pc.shape[1])
        term("cdc", Pc, gt_c)
        term("cd1", P1, gt_1)
        term("cd2", P2, gt)
    if partial is not None:  # chamfer_single_side(_sqrt)(partial, P2): the dist1 side only (utils/loss_utils.py:22-31)
        d1, _, _, _ = chamfer_3DFunction.apply(partial.contiguous(), P2.contiguous())
        sums.add("pm.d1", torch.sqrt(d1) if sqrt else d1)
    return sums


def get_loss_sharded(pcds_pred, gt, sqrt=True, alpha1=1, alpha2=1, group=None, comm=None, grad_reduce="sum"):
    """utils/loss_utils.get_loss (:33-58) on a batch shard: identical value on every rank, equal
    to the single-process loss over the concatenated batch.  `grad_reduce`: how the caller combines parameter
    gradients across ranks — "sum" (default) or "mean" (stock DistributedDataParallel); see the module docstring.
    `comm`: a PeerComm to run the one collective over peer memory instead of torch.distributed."""
    means = _sharded_terms(pcds_pred, gt, sqrt).reduce(group, comm, grad_reduce)
    cdc, cd1, cd2 = (combine_chamfer(means, n, sqrt) for n in ("cdc", "cd1", "cd2"))
    return cdc + alpha1 * cd1 + alpha2 * cd2, [cdc, cd1, cd2]


def get_loss_PM_sharded(pcds_pred, partial, gt, sqrt=True, group=None, comm=None, grad_reduce="sum"):
    """utils/loss_utils.get_loss_PM (:60-85; core/train_55.py:154, core/train_geospec.py:108) on a batch shard: the
    three Chamfer terms plus the single-sided partial-matching term, still ONE all-reduce (14 numbers)."""
    means = _sharded_terms(pcds_pred, gt, sqrt, partial=partial).reduce(group, comm, grad_reduce)
    cdc, cd1, cd2 = (combine_chamfer(means, n, sqrt) for n in ("cdc", "cd1", "cd2"))
    return cdc + cd1 + cd2 + means["pm.d1"], [cdc, cd1, cd2]


class GraphedLoss:
    """get_loss_sharded / get_loss_PM_sharded, forward AND backward, as one replayed CUDA graph.

    The eager loss is ~100 small launches around the five big kernels (three Chamfer terms, two FPS calls): square
    roots, sums, the concatenation for the single all-reduce, and the same again backwards — 0.45 ms of host time per
    step, which is what a rank of an 8-GPU run waits for once its shard is down to 4 clouds.  Shapes are fixed per
    training run, so the step is captured once (the FPS chain still forks onto the side stream: fork and join become
    graph edges; with `comm`, the peer-memory exchange is one more kernel node) and replayed:

        step = GraphedLoss([Pc.shape, P1.shape, P2.shape], gt.shape, sqrt=True, comm=comm)
        loss, (cdc, cd1, cd2), (gPc, gP1, gP2) = step([Pc, P1, P2], gt)     # copies in, one graph launch
        torch.autograd.backward([Pc, P1, P2], [gPc, gP1, gP2])              # continue into the model

    The returned tensors are the graph's static outputs: consume (or clone) them before the next call.  Every rank of
    `comm` must call the same number of times (the exchange is part of the graph).  Reference: utils/loss_utils.py:33-85.
    """

    def __init__(self, pred_shapes, gt_shape, sqrt=True, alpha1=1, alpha2=1, comm=None, grad_reduce="sum",
                 partial_shape=None, device=None, warmup=2):
        if not torch.cuda.is_available():
            raise L.PointSeaError("GraphedLoss needs a CUDA device (the eager get_loss_sharded runs anywhere)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        mk = lambda shape: torch.zeros(tuple(shape), device=self.device, dtype=torch.float32)
        self.preds = [mk(sh).requires_grad_(True) for sh in pred_shapes]
        self.gt = mk(gt_shape)
        self.partial = mk(partial_shape) if partial_shape is not None else None
        g = torch.Generator().manual_seed(7)  # warm-up on spread-out points: all-zero clouds are one big tie
        with torch.no_grad():
            for t in self.preds + [self.gt] + ([self.partial] if self.partial is not None else []):
                t.copy_(torch.rand(t.shape, generator=g) - 0.5)

        def run():
            if self.partial is None:
                loss, terms = get_loss_sharded(self.preds, self.gt, sqrt=sqrt, alpha1=alpha1, alpha2=alpha2, comm=comm,
                                               grad_reduce=grad_reduce)
            else:
                loss, terms = get_loss_PM_sharded(self.preds, self.partial, self.gt, sqrt=sqrt, comm=comm, grad_reduce=grad_reduce)
            grads = torch.autograd.grad(loss, self.preds)
            return loss, terms, grads

        with torch.cuda.device(self.device):
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):  # function attributes, pools, plans: outside the capture
                    run()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.terms, self.grads = run()
        self.calls = 0

    def __call__(self, pcds_pred, gt, partial=None):
        with torch.no_grad():
            for dst, src in zip(self.preds, pcds_pred):
                dst.copy_(src, non_blocking=True)
            self.gt.copy_(gt, non_blocking=True)
            if self.partial is not None:
                if partial is None:
                    raise L.PointSeaError("this GraphedLoss was captured with a partial cloud (get_loss_PM)")
                self.partial.copy_(partial, non_blocking=True)
        with torch.cuda.device(self.device):
            self.graph.replay()
        self.calls += 1
        return self.loss, self.terms, self.grads
