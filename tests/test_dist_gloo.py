"""world_size-2 gloo test of the N>1 path: batch sharding + ONE all-reduce of loss partial sums
reproduces the single-process loss exactly.  Chamfer distances come from the CPU
oracle here (the CUDA path needs a GPU); the sharding / reduction logic is the product's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svdformer_pointsea_b200.dist import shard_bounds, shard_batch, LossSums, chamfer_loss_terms, combine_chamfer


def test_shard_bounds_cover_the_batch_exactly():
    for B in (1, 5, 8, 32, 33):
        for G in (1, 2, 3, 4, 8):
            spans = [shard_bounds(B, r, G) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    from oracle import oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(42)  # every rank builds the same global batch, then shards it
    a = torch.rand(B, 96, 3, generator=g) - 0.5
    b = torch.rand(B, 160, 3, generator=g) - 0.5
    a_s, b_s = shard_batch(a), shard_batch(b)
    d1, d2, _, _ = O.chamfer_fwd(a_s.numpy(), b_s.numpy())
    sums = LossSums(torch.device("cpu"), dtype=torch.float64)
    chamfer_loss_terms(sums, "cd", torch.from_numpy(d1), torch.from_numpy(d2), sqrt=True)
    chamfer_loss_terms(sums, "cd_l2", torch.from_numpy(d1), torch.from_numpy(d2), sqrt=False)
    means = sums.reduce()
    q.put((rank, a_s.size(0), float(combine_chamfer(means, "cd", True)), float(combine_chamfer(means, "cd_l2", False))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_loss_equals_single_process_loss():
    from oracle import oracle as O
    B, world = 5, 2  # uneven shards: 2 + 3 clouds
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert [r[1] for r in res] == [2, 3]
    g = torch.Generator().manual_seed(42)
    a = torch.rand(B, 96, 3, generator=g) - 0.5
    b = torch.rand(B, 160, 3, generator=g) - 0.5
    d1, d2, _, _ = O.chamfer_fwd(a.numpy(), b.numpy())
    t1, t2 = torch.sqrt(torch.from_numpy(d1)).double(), torch.sqrt(torch.from_numpy(d2)).double()  # fp32 sqrt, f64 sums
    want_sqrt = float((t1.mean() + t2.mean()) / 2)
    want_l2 = d1.astype(np.float64).mean() + d2.astype(np.float64).mean()
    for _, _, got_sqrt, got_l2 in res:  # identical on every rank and equal to the global value
        assert abs(got_sqrt - want_sqrt) < 1e-12 and abs(got_l2 - want_l2) < 1e-12


def test_loss_sums_single_process_and_autograd():
    x = torch.rand(3, 7, dtype=torch.float64, requires_grad=True)
    sums = LossSums(torch.device("cpu"), dtype=torch.float64)
    sums.add("t", x * 2)
    m = sums.reduce()["t"]
    m.backward()
    assert torch.allclose(m, (x * 2).mean()) and torch.allclose(x.grad, torch.full_like(x, 2 / 21))


def _pipelined_worker(rank, world, port, q):
    import torch.distributed as dist
    from svdformer_pointsea_b200.dist import PipelinedSums
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    red = PipelinedSums()
    outs = []
    for step in range(4):
        vec = torch.tensor([float(rank + 1) * (step + 1), 1.0], dtype=torch.float64)
        prev = red.submit(vec)  # result of the PREVIOUS step
        outs.append(None if prev is None else prev.tolist())
    outs.append(red.flush().tolist())
    assert red.flush() is None
    q.put((rank, outs))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_sums_return_the_previous_steps_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipelined_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    for _, outs in res:
        assert outs[0] is None
        # step s reduces (1 + 2) * (s + 1) and the two counts
        assert outs[1:] == [[3.0 * (s + 1), 2.0] for s in range(4)]


def test_pipelined_sums_without_a_process_group():
    from svdformer_pointsea_b200.dist import PipelinedSums
    red = PipelinedSums()
    a, b = torch.tensor([1.0, 2.0]), torch.tensor([3.0, 4.0])
    assert red.submit(a) is None and red.submit(b) is a and red.flush() is b


# ---- gradients of the sharded losses (identity backward of the all-reduce vs DDP's averaging) ----------
class _TorchChamfer:
    """CPU stand-in for chamfer_3DFunction in these host-logic tests: same outputs, differentiable torch expression."""

    @staticmethod
    def apply(a, b):
        d = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1)
        d1, i1 = d.min(2)
        d2, i2 = d.min(1)
        return d1, d2, i1.int(), i2.int()


def _cpu_fps_subsample(pcd, n):
    from oracle import oracle as O
    idx = torch.from_numpy(O.fps(pcd.detach().numpy().astype(np.float32), n)).long()
    return torch.gather(pcd, 1, idx[:, :, None].expand(-1, -1, 3))


def _toy_batch(B):
    g = torch.Generator().manual_seed(7)
    partial = torch.rand(B, 24, 3, generator=g) - 0.5
    gt = torch.rand(B, 64, 3, generator=g) - 0.5
    w = torch.randn(3, 3, generator=g) * 0.1 + torch.eye(3)
    return partial, gt, w


def _toy_loss(fn_name, w, partial, gt, **kw):
    """A one-parameter 'model': the three predictions are linear maps of fixed clouds."""
    import svdformer_pointsea_b200.dist as D
    Pc, P1, P2 = gt[:, :16] @ w, gt[:, :32] @ w, gt @ w
    if fn_name == "get_loss":
        return D.get_loss_sharded((Pc, P1, P2), gt, sqrt=True, **kw)[0]
    return D.get_loss_PM_sharded((Pc, P1, P2), partial, gt, sqrt=False, **kw)[0]


def _patch_cpu_ops(monkeypatch=None):
    """In a worker process: plain assignment.  In the pytest process: through `monkeypatch`, so later tests see the real ops."""
    import svdformer_pointsea_b200.chamfer as C
    import svdformer_pointsea_b200.pointnet2_utils as P
    if monkeypatch is None:
        C.chamfer_3DFunction = _TorchChamfer
        P.fps_subsample = _cpu_fps_subsample
    else:
        monkeypatch.setattr(C, "chamfer_3DFunction", _TorchChamfer)
        monkeypatch.setattr(P, "fps_subsample", _cpu_fps_subsample)


def _grad_worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _patch_cpu_ops()
    out = {}
    for fn in ("get_loss", "get_loss_PM"):
        for mode in ("sum", "mean"):
            partial, gt, w = _toy_batch(B)
            w = w.double().requires_grad_(True)
            loss = _toy_loss(fn, w, shard_batch(partial).double(), shard_batch(gt).double(), grad_reduce=mode)
            loss.backward()
            g = w.grad.clone()
            # what the caller does with parameter gradients: all-reduce(sum), or DDP's average
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            if mode == "mean":
                g /= world
            out[(fn, mode)] = (float(loss.detach()), g.numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_match_the_single_process_gradient(monkeypatch):
    B, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, B, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    _patch_cpu_ops(monkeypatch)
    for fn in ("get_loss", "get_loss_PM"):
        partial, gt, w = _toy_batch(B)
        w = w.double().requires_grad_(True)
        loss = _toy_loss(fn, w, partial.double(), gt.double())  # no process group: the reference's full-batch loss
        loss.backward()
        for mode in ("sum", "mean"):
            for _, out in res:
                got_loss, got_grad = out[(fn, mode)]
                # the partial sums travel as fp32 (LossSums' default, the reference's own precision): 1e-6 relative
                assert abs(got_loss - float(loss.detach())) < 1e-6 * abs(float(loss.detach()))
                assert np.allclose(got_grad, w.grad.numpy(), rtol=1e-5, atol=1e-8), (fn, mode)


def test_get_loss_pm_matches_the_reference_expression(monkeypatch):
    """get_loss_PM_sharded without a process group == utils/loss_utils.get_loss_PM (:60-85) written out in torch."""
    _patch_cpu_ops(monkeypatch)
    import svdformer_pointsea_b200.dist as D
    partial, gt, w = _toy_batch(3)
    Pc, P1, P2 = gt[:, :16] @ w, gt[:, :32] @ w, gt @ w
    for sqrt in (True, False):
        got, parts = D.get_loss_PM_sharded((Pc, P1, P2), partial, gt, sqrt=sqrt)
        gt_1 = _cpu_fps_subsample(gt, 32)
        gt_c = _cpu_fps_subsample(gt_1, 16)

        def cd(p, q):
            d1, d2, _, _ = _TorchChamfer.apply(p, q)
            return (torch.sqrt(d1).mean() + torch.sqrt(d2).mean()) / 2 if sqrt else d1.mean() + d2.mean()

        d1 = _TorchChamfer.apply(partial, P2)[0]
        pm = torch.sqrt(d1).mean() if sqrt else d1.mean()
        want = cd(Pc, gt_c) + cd(P1, gt_1) + cd(P2, gt) + pm
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)
        assert len(parts) == 3
