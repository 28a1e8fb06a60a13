"""GPU tests of the one-launch step and the peer-memory collective:
fused loss sums in the forward's epilogue, ps_chamfer_step (graph replay, retargeting on fresh buffers),
ps_comm_* over local peers (two ranks on one device / on two devices) and over CUDA IPC between processes."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import make_cloud

pytestmark = pytest.mark.gpu


def _ps():
    import svdformer_pointsea_b200 as ps
    return ps


def _clouds(B, N, M, seed=0, dev="cuda:0"):
    g = torch.Generator().manual_seed(seed)
    return make_cloud(g, B, N).to(dev), make_cloud(g, B, M).to(dev), torch.randn(B, N, generator=g).to(dev), torch.randn(B, M, generator=g).to(dev)


@pytest.mark.parametrize("shape", [(3, 2048, 16384), (2, 1000, 700), (4, 300, 90), (1, 4096, 4096)])
def test_fused_sums_equal_the_separate_reduction_and_are_reproducible(shape):
    """ps_chamfer_fwd_sums: same dist/idx bits as ps_chamfer_fwd; sums equal ps_chamfer_sums to 1e-13 relative
    (fp64 accumulation in another, fixed order) and are bit-identical run to run."""
    ps = _ps()
    B, N, M = shape
    a, b, _, _ = _clouds(B, N, M, seed=11)
    ref = ps.chamfer_forward(a, b)
    want = ps.chamfer_sums(ref[0], ref[1])
    sums = torch.full((6,), -1.0, device="cuda:0", dtype=torch.float64)
    got = ps.chamfer_forward(a, b, sums=sums)
    for x, y in zip(ref, got):
        assert torch.equal(x, y)
    assert torch.allclose(sums, want, rtol=1e-13, atol=0)
    assert sums[4].item() == B * N and sums[5].item() == B * M
    first = sums.clone()
    for _ in range(5):
        ps.chamfer_forward(a, b, sums=sums)
        assert torch.equal(sums, first)


def test_step_equals_forward_sums_backward_and_replays():
    ps = _ps()
    from svdformer_pointsea_b200 import _lib as L
    B, N, M = 4, 2048, 4096
    a, b, ga, gb = _clouds(B, N, M, seed=3)
    d1, d2, i1, i2 = ps.chamfer_forward(a, b)
    g1, g2 = ps.chamfer_backward(a, b, ga, gb, i1, i2)
    want = ps.chamfer_sums(d1, d2)
    step = ps.ChamferStep(B, N, M, "cuda:0")
    before = L.graph_stats(0, "step")
    for it in range(4):
        sl, sg, s1, s2 = step(a, b, ga, gb)
        assert torch.equal(step.dist1, d1) and torch.equal(step.idx1, i1) and torch.equal(step.dist2, d2) and torch.equal(step.idx2, i2)
        assert torch.allclose(sl, want, rtol=1e-13, atol=0) and sg is sl
        assert torch.allclose(s1, g1, rtol=1e-5, atol=1e-7) and torch.allclose(s2, g2, rtol=1e-5, atol=1e-7)
    after = L.graph_stats(0, "step")
    assert after["hits"] - before["hits"] >= 3  # replayed, not re-captured
    # forward-only form
    sl, _, n1, n2 = step(a, b)
    assert n1 is None and n2 is None and torch.allclose(sl, want, rtol=1e-13, atol=0)


def test_step_retargets_the_cached_graph_on_fresh_buffers():
    """A loader handing out fresh device buffers every step: more address sets than cache entries.  Results stay
    exact and the cache updates executables in place instead of instantiating one per step."""
    ps = _ps()
    from svdformer_pointsea_b200 import _lib as L
    B, N, M = 2, 1024, 2048
    step = ps.ChamferStep(B, N, M, "cuda:0")
    keep = []
    before = L.graph_stats(0, "step")
    for it in range(20):
        a, b, ga, gb = _clouds(B, N, M, seed=100 + it)
        keep.append((a, b, ga, gb))  # keep them alive: every step sees new addresses
        sl, _, s1, s2 = step(a, b, ga, gb)
        d1, d2, i1, i2 = ps.chamfer_forward(a, b)
        g1, g2 = ps.chamfer_backward(a, b, ga, gb, i1, i2)
        assert torch.equal(step.dist1, d1) and torch.equal(step.idx2, i2)
        assert torch.allclose(s1, g1, rtol=1e-5, atol=1e-7) and torch.allclose(s2, g2, rtol=1e-5, atol=1e-7)
        assert torch.allclose(sl, ps.chamfer_sums(d1, d2), rtol=1e-13, atol=0)
    after = L.graph_stats(0, "step")
    assert after["instantiations"] - before["instantiations"] <= 9
    assert after["updates"] - before["updates"] >= 10


def test_step_inside_a_torch_cuda_graph():
    """The entry points stay capturable by the caller's own graph (stream-ordered scratch becomes graph memory)."""
    ps = _ps()
    B, N, M = 2, 2048, 2048
    a, b, ga, gb = _clouds(B, N, M, seed=5)
    step = ps.ChamferStep(B, N, M, "cuda:0")
    step(a, b, ga, gb)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            step(a, b, ga, gb)
    want = step.sums_local.clone()
    step.sums_local.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(step.sums_local, want)


@pytest.mark.parametrize("with_partial", [False, True])
def test_graphed_loss_replays_get_loss_forward_and_backward(with_partial):
    """GraphedLoss = get_loss_sharded / get_loss_PM_sharded + autograd.grad captured once (FPS chain forked onto the
    side stream inside the capture) and replayed on NEW clouds: loss, terms and gradients match the eager call."""
    _ps()
    from svdformer_pointsea_b200.dist import GraphedLoss, get_loss_sharded, get_loss_PM_sharded
    dev = torch.device("cuda:0")
    B = 3
    g = torch.Generator().manual_seed(77)
    shapes = [(B, 128, 3), (B, 512, 3), (B, 4096, 3)]
    step = GraphedLoss(shapes, (B, 4096, 3), sqrt=True, partial_shape=(B, 1024, 3) if with_partial else None)
    for it in range(3):
        preds = [make_cloud(g, B, sh[1]).to(dev).requires_grad_(True) for sh in shapes]
        gt = make_cloud(g, B, 4096, dup=300, near_origin=4).to(dev)
        part = make_cloud(g, B, 1024).to(dev) if with_partial else None
        if with_partial:
            want, wt = get_loss_PM_sharded(preds, part, gt, sqrt=True)
        else:
            want, wt = get_loss_sharded(preds, gt, sqrt=True)
        wg = torch.autograd.grad(want, preds)
        loss, terms, grads = step(preds, gt, part)
        torch.cuda.synchronize()
        assert abs(float(loss.detach()) - float(want.detach())) <= 1e-6 * abs(float(want.detach()))
        for a, b in zip(terms, wt):
            assert abs(float(a.detach()) - float(b.detach())) <= 1e-6 * abs(float(b.detach()))
        for a, b in zip(grads, wg):
            assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
    assert step.calls == 3


def _local_comms(devs):
    from svdformer_pointsea_b200.dist import PeerComm
    return PeerComm.local(devs)


def _run_local_allreduce(devs, rounds=50):
    comms = _local_comms(devs)
    world = len(devs)
    streams = [torch.cuda.Stream(device=d) for d in devs]
    ins = [torch.zeros(7, device=d, dtype=torch.float64) for d in devs]
    outs = [[] for _ in devs]
    for it in range(rounds):
        for r, d in enumerate(devs):
            with torch.cuda.device(d), torch.cuda.stream(streams[r]):
                ins[r].copy_(torch.arange(7, dtype=torch.float64) * (r + 1) + it, non_blocking=False)
                outs[r].append(comms[r].all_reduce(ins[r]).clone())
    for d in devs:
        torch.cuda.synchronize(d)
    base = torch.arange(7, dtype=torch.float64)
    for it in range(rounds):
        want = sum(base * (r + 1) + it for r in range(world))
        for r in range(world):
            assert torch.equal(outs[r][it].cpu(), want), (it, r)
    for c in comms:
        st = c.status()
        assert st["published"] == rounds and st["consumed"] == rounds and not st["timed_out"]
        c.close()


def test_peer_allreduce_two_ranks_on_one_device():
    """The mailbox protocol (publish / acquire-wait / rank-ordered sum, slot reuse every 4 steps) with both ranks on
    cuda:0, each on its own stream: runs on a single-GPU box."""
    _run_local_allreduce(["cuda:0", "cuda:0"])
    _run_local_allreduce(["cuda:0"] * 5, rounds=13)


def test_peer_allreduce_across_devices():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_local_allreduce(["cuda:0", "cuda:1"])


def test_step_with_comm_two_ranks_on_one_device():
    """ps_chamfer_step with a communicator: the epilogue's last block publishes, the trailing wait kernel returns the
    world-wide sums; both 'ranks' on cuda:0 on separate streams, several replays."""
    ps = _ps()
    B, N, M = 2, 2048, 4096
    comms = _local_comms(["cuda:0", "cuda:0"])
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    data = [_clouds(B, N, M, seed=40 + r) for r in range(2)]
    steps = [ps.ChamferStep(B, N, M, "cuda:0", comm=comms[r]) for r in range(2)]
    want_local = []
    for r in range(2):
        d1, d2, _, _ = ps.chamfer_forward(data[r][0], data[r][1])
        want_local.append(ps.chamfer_sums(d1, d2))
    for r in range(2):
        steps[r].prepare(*data[r])  # graph instantiation may synchronise the device: not while a peer's wait kernel spins
    torch.cuda.synchronize()
    for it in range(6):
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                steps[r](*data[r])
        torch.cuda.synchronize()
        for r in range(2):
            assert torch.allclose(steps[r].sums_local, want_local[r], rtol=1e-13, atol=0)
        assert torch.equal(steps[0].sums_global, steps[1].sums_global)  # bit-identical on every rank
        assert torch.equal(steps[0].sums_global, steps[0].sums_local + steps[1].sums_local)  # rank-ordered sum
    for c in comms:
        assert not c.status()["timed_out"]
        c.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, q):
    import torch.distributed as dist
    import svdformer_pointsea_b200 as ps
    from svdformer_pointsea_b200.dist import PeerComm, shard_batch, get_loss_sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    comm = PeerComm()
    B, N, M = 6, 1024, 2048
    g = torch.Generator().manual_seed(9)
    a, b = make_cloud(g, B, N), make_cloud(g, B, M)
    ga, gb = torch.randn(B, N, generator=g), torch.randn(B, M, generator=g)
    sa, sb, sga, sgb = (shard_batch(t).to(dev) for t in (a, b, ga, gb))
    step = ps.ChamferStep(sa.size(0), N, M, dev, comm=comm)
    outs = []
    for it in range(5):
        _, sg, _, _ = step(sa, sb, sga, sgb)
        outs.append(sg.clone())
    # host-buffer step with the collective inside its graph
    hs = torch.empty(6, dtype=torch.float64).pin_memory()
    ps.chamfer_host_step(sa.cpu().pin_memory(), sb.cpu().pin_memory(), sga.cpu().pin_memory(), sgb.cpu().pin_memory(),
                         sums_out=hs, comm=comm)
    # asynchronous host steps with the collective: three in flight, each lane on its own channel of the communicator
    import collections
    h_in = [t.cpu().pin_memory() for t in (sa, sb, sga, sgb)]
    sets = [([torch.empty(sa.size(0), N).pin_memory(), torch.empty(sa.size(0), M).pin_memory(),
              torch.empty(sa.size(0), N, dtype=torch.int32).pin_memory(), torch.empty(sa.size(0), M, dtype=torch.int32).pin_memory(),
              torch.empty(sa.size(0), N, 3).pin_memory(), torch.empty(sa.size(0), M, 3).pin_memory()],
             torch.empty(6, dtype=torch.float64).pin_memory()) for _ in range(4)]
    pend, async_sums = collections.deque(), []
    for i in range(11):
        pend.append(ps.chamfer_host_async(*h_in, out=sets[i % 4][0], sums_out=sets[i % 4][1], comm=comm))
        if len(pend) == 3:
            st = pend.popleft()
            st.synchronize()
            async_sums.append(st.sums.numpy().copy())
    while pend:
        st = pend.popleft()
        st.synchronize()
        async_sums.append(st.sums.numpy().copy())
    # the autograd loss over the peer-memory collective vs the NCCL one
    P = [(shard_batch(a)[:, :256].to(dev)).requires_grad_(True), shard_batch(a)[:, :512].to(dev), shard_batch(a).to(dev)]
    gt = shard_batch(b).to(dev)
    l_peer, _ = get_loss_sharded(P, gt, comm=comm)
    l_nccl, _ = get_loss_sharded(P, gt)
    # the same loss, forward + backward, captured once with the exchange inside and replayed on new clouds
    from svdformer_pointsea_b200.dist import GraphedLoss
    P3 = [p.detach().clone().requires_grad_(True) for p in P]
    l_eager, _ = get_loss_sharded(P3, gt, comm=comm)
    g_eager = torch.autograd.grad(l_eager, P3)
    graphed = GraphedLoss([p.shape for p in P3], gt.shape, comm=comm)
    for _ in range(3):
        l_graph, _, g_graph = graphed(P3, gt)
    gerr = max(float((a_ - b_).abs().max() / b_.abs().max()) for a_, b_ in zip(g_graph, g_eager))
    torch.cuda.synchronize()
    q.put((rank, [o.cpu().numpy() for o in outs], hs.numpy().copy(), float(l_peer), float(l_nccl), comm.status(), async_sums,
           float(l_eager), float(l_graph), gerr))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_peer_comm_over_cuda_ipc_between_processes():
    """Two processes, one GPU each: mailboxes mapped through CUDA IPC handles exchanged over torch.distributed."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ps = _ps()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    [p.join(timeout=120) for p in procs]
    B, N, M = 6, 1024, 2048
    g = torch.Generator().manual_seed(9)
    a, b = make_cloud(g, B, N).cuda(), make_cloud(g, B, M).cuda()
    d1, d2, _, _ = ps.chamfer_forward(a, b)
    want = ps.chamfer_sums(d1, d2).cpu().numpy()
    for rank, outs, hs, l_peer, l_nccl, st, async_sums, l_eager, l_graph, gerr in res:
        assert not st["timed_out"]
        assert abs(l_graph - l_eager) <= 1e-6 * abs(l_eager) and l_graph == res[0][8] and gerr < 1e-5
        assert len(async_sums) == 11
        for v in async_sums:
            assert np.array_equal(v, res[0][6][0]) and np.allclose(v, want, rtol=1e-12, atol=0)
        for o in outs:
            assert np.array_equal(o, res[0][1][0])  # identical bits on both ranks, every replay
            assert np.allclose(o, want, rtol=1e-12, atol=0)
        assert np.allclose(hs, want, rtol=1e-12, atol=0)
        assert abs(l_peer - l_nccl) <= 1e-6 * abs(l_nccl)
