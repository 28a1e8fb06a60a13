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
