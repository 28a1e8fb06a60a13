"""Size-independent properties of the CPU oracle (hypothesis): the checker itself is checked against plain numpy
float64 restatements of WHAT each op computes, independent of the fp32 operation order it pins."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import oracle as O

SET = dict(max_examples=25, deadline=None)


def cloud(rng, B, N, lo=-0.5, hi=0.5):
    return rng.uniform(lo, hi, size=(B, N, 3)).astype(np.float32)


@settings(**SET)
@given(st.integers(0, 2**31 - 1), st.integers(1, 3), st.integers(1, 70), st.integers(1, 90))
def test_chamfer_is_the_nearest_neighbour_distance_and_swaps_with_its_arguments(seed, B, N, M):
    rng = np.random.default_rng(seed)
    a, b = cloud(rng, B, N), cloud(rng, B, M)
    d1, d2, i1, i2 = O.chamfer_fwd(a, b)
    D = ((a[:, :, None, :].astype(np.float64) - b[:, None, :, :].astype(np.float64)) ** 2).sum(-1)
    assert np.allclose(d1, D.min(2), rtol=1e-5, atol=1e-9) and np.allclose(d2, D.min(1), rtol=1e-5, atol=1e-9)
    # the reported index realises the reported distance
    assert np.allclose(np.take_along_axis(D, i1[:, :, None].astype(np.int64), 2)[..., 0], d1, rtol=1e-5, atol=1e-9)
    assert np.allclose(np.take_along_axis(D, i2[:, None, :].astype(np.int64), 1)[:, 0], d2, rtol=1e-5, atol=1e-9)
    # swapping the clouds swaps the outputs bit-for-bit ((b - a)^2 == (a - b)^2 exactly)
    e1, e2, j1, j2 = O.chamfer_fwd(b, a)
    assert np.array_equal(e1, d2) and np.array_equal(e2, d1) and np.array_equal(j1, i2) and np.array_equal(j2, i1)


@settings(**SET)
@given(st.integers(0, 2**31 - 1), st.integers(1, 2), st.integers(2, 60), st.integers(1, 40))
def test_chamfer_gradient_is_the_derivative_of_the_selected_pairs(seed, B, N, M):
    rng = np.random.default_rng(seed)
    a, b = cloud(rng, B, N), cloud(rng, B, M)
    g1, g2 = rng.standard_normal((B, N)).astype(np.float32), rng.standard_normal((B, M)).astype(np.float32)
    _, _, i1, i2 = O.chamfer_fwd(a, b)
    ga, gb = O.chamfer_bwd(a, b, g1, g2, i1, i2)
    wa, wb = np.zeros((B, N, 3)), np.zeros((B, M, 3))
    for bb in range(B):
        for i in range(N):
            d = 2.0 * g1[bb, i] * (a[bb, i].astype(np.float64) - b[bb, i1[bb, i]])
            wa[bb, i] += d; wb[bb, i1[bb, i]] -= d
        for j in range(M):
            d = 2.0 * g2[bb, j] * (b[bb, j].astype(np.float64) - a[bb, i2[bb, j]])
            wb[bb, j] += d; wa[bb, i2[bb, j]] -= d
    assert np.allclose(ga, wa, rtol=1e-4, atol=1e-5) and np.allclose(gb, wb, rtol=1e-4, atol=1e-5)


@settings(**SET)
@given(st.integers(0, 2**31 - 1), st.integers(1, 2), st.integers(1, 200), st.integers(1, 40))
def test_fps_is_greedy_farthest_point_selection(seed, B, N, npoint):
    rng = np.random.default_rng(seed)
    x = cloud(rng, B, N, 0.2, 1.2)  # away from the origin: the skip rule (|p|^2 <= 1e-3) stays out of the way
    idx = O.fps(x, npoint)
    assert (idx[:, 0] == 0).all() and idx.min() >= 0 and idx.max() < N
    for b in range(B):
        xs = x[b].astype(np.float64)
        mind = np.full(N, 1e10)
        for j in range(1, npoint):
            mind = np.minimum(mind, ((xs - xs[idx[b, j - 1]]) ** 2).sum(-1))
            # the chosen point is (one of) the farthest from everything chosen so far
            assert mind[idx[b, j]] >= mind.max() * (1 - 1e-5) - 1e-12
        if npoint <= N:
            assert len(set(idx[b].tolist())) == min(npoint, N)


@settings(**SET)
@given(st.integers(0, 2**31 - 1), st.integers(1, 2), st.integers(2, 9), st.integers(4, 80), st.integers(1, 8))
def test_feature_knn_returns_the_k_smallest_in_both_orders(seed, B, C, N, k):
    rng = np.random.default_rng(seed)
    k = min(k, N)
    x = rng.standard_normal((B, N, C)).astype(np.float32)
    x[:, N // 2] = x[:, 0]  # one exact duplicate -> an exact tie at the front of row 0
    D = ((x[:, :, None, :].astype(np.float64) - x[:, None, :, :].astype(np.float64)) ** 2).sum(-1)
    for order in (0, 1):
        idx = O.knn_feat(x, x, k, order=order)
        got = np.take_along_axis(D, idx.astype(np.int64), 2)
        want = np.sort(D, 2)[:, :, :k]
        # the selected distances are the k smallest (the fp32 expanded form may reorder near-equal ones)
        assert np.allclose(np.sort(got, 2), want, rtol=1e-3, atol=2e-5)
        assert all(len(set(r.tolist())) == k for r in idx.reshape(-1, k))
    a, b = O.knn_feat(x, x, k, order=0), O.knn_feat(x, x, k, order=1)
    assert np.array_equal(np.sort(a, -1), np.sort(b, -1))  # torch.topk's order only permutes ties


@settings(**SET)
@given(st.integers(0, 2**31 - 1), st.integers(1, 2), st.integers(1, 6), st.integers(1, 50), st.integers(1, 20), st.integers(1, 5))
def test_group_edge_and_index_points_are_plain_indexing(seed, B, C, N, S, K):
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, S, K)).astype(np.int32)
    want = np.stack([f[b][:, idx[b]] for b in range(B)])
    assert np.array_equal(O.group(f, idx), want)
    go = rng.standard_normal(want.shape).astype(np.float32)
    acc = np.zeros((B, C, N))
    for b in range(B):
        for c in range(C):
            np.add.at(acc[b, c], idx[b].reshape(-1), go[b, c].reshape(-1).astype(np.float64))
    assert np.allclose(O.group_grad(go, idx, N), acc, rtol=1e-5, atol=1e-6)
    eidx = rng.integers(0, N, (B, N, K)).astype(np.int32)
    e = O.edge_features(f, eidx)
    cen = np.repeat(f[:, :, :, None], K, 3)
    nb = np.stack([f[b][:, eidx[b]] for b in range(B)])
    assert np.array_equal(e, np.concatenate([cen - nb, cen], 1))
    pts = np.ascontiguousarray(f.transpose(0, 2, 1))
    assert np.array_equal(O.index_points(pts, idx), np.stack([pts[b][idx[b]] for b in range(B)]))


@settings(**SET)
@given(st.integers(0, 2**31 - 1), st.integers(1, 3), st.integers(1, 60), st.integers(1, 60))
def test_metrics_epilogue_against_numpy(seed, B, n1, n2):
    rng = np.random.default_rng(seed)
    d1 = (rng.random((B, n1)) ** 4 * 1e-3).astype(np.float32)
    d2 = (rng.random((B, n2)) ** 4 * 1e-3).astype(np.float32)
    i1 = rng.integers(0, n2, (B, n1)).astype(np.int32)
    i2 = rng.integers(0, n1, (B, n2)).astype(np.int32)
    m = O.chamfer_metrics(d1, d2, i1, i2, threshold=1e-4, alpha=1000.0, n_lambda=1.0, frac1=n1 / n2, frac2=n2 / n1)
    D1, D2 = d1.astype(np.float64), d2.astype(np.float64)
    assert np.allclose(m[:, 0], np.sqrt(D1).mean(1), rtol=1e-5) and np.allclose(m[:, 3], D2.mean(1), rtol=1e-5)
    p1, p2 = (d1 < np.float32(1e-4)).mean(1), (d2 < np.float32(1e-4)).mean(1)
    assert np.allclose(m[:, 4], p1, rtol=1e-6) and np.allclose(m[:, 5], p2, rtol=1e-6)
    f = np.where(p1 + p2 > 0, 2 * p1 * p2 / np.maximum(p1 + p2, 1e-300), 0.0)
    assert np.allclose(m[:, 6], f, rtol=1e-5, atol=1e-7)
    loss = np.zeros(B)
    for b in range(B):
        c1 = np.bincount(i1[b], minlength=n2); c2 = np.bincount(i2[b], minlength=n1)
        l1 = (1 - np.exp(-D1[b] * 1000.0) / (c1[i1[b]] + 1e-6) * (n1 / n2)).mean()
        l2 = (1 - np.exp(-D2[b] * 1000.0) / (c2[i2[b]] + 1e-6) * (n2 / n1)).mean()
        loss[b] = (l1 + l2) / 2
    assert np.allclose(m[:, 7], loss, rtol=1e-4, atol=1e-6)
