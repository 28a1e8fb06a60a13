"""Pins the oracle's "next"-row restatements to tests/golden/next.npz: outputs of the
reference's torch expressions + its own CUDA ops on a B200 (tests/golden/make_golden_next.py)."""
import numpy as np

from conftest import load_golden
from oracle import oracle as O


def rel(a, b):
    return np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30)


def pm(x):
    """(B,C,N) -> point-major (B,N,C)"""
    return np.ascontiguousarray(np.asarray(x).transpose(0, 2, 1))


def test_feature_knn_matches_torch_topk_bit_exactly():
    z = load_golden("next")
    # C = 3 / 64 / 256, with exact ties (duplicated points: torch.topk's unstable order) and post-ReLU features
    for name in ("edge3", "edge3dup", "edge64", "edge256", "edge64post"):
        x, k = pm(z[f"{name}.x"]), int(z[f"{name}.k"])
        got = O.knn_feat(x, x, k, order=1)
        assert np.array_equal(got, z[f"{name}.idx"]), name


def test_topk_order_differs_from_sorted_order_only_inside_ties():
    z = load_golden("next")
    x, k = pm(z["edge3dup.x"]), int(z["edge3dup.k"])
    topk, srt = O.knn_feat(x, x, k, order=1), O.knn_feat(x, x, k, order=0)
    assert not np.array_equal(topk, srt)  # the case exists to exercise the tie order
    assert np.array_equal(np.sort(topk, -1), np.sort(srt, -1))  # same neighbour sets
    assert np.array_equal(O.knn_point(x, x, k), topk)


def test_edge_features_exact():
    z = load_golden("next")
    got = O.edge_features(z["edge3dup.x"], z["edge3dup.idx"])
    assert np.array_equal(got, z["edge3dup.feat"])


def test_sample_and_group_knn_exact():
    z = load_golden("next")
    for name in ("sg", "sgdup"):
        xyz = z[f"{name}.xyz"]                      # (B,3,N)
        npoint, k = int(z[f"{name}.npoint"]), int(z[f"{name}.k"])
        fidx = O.fps(pm(xyz), npoint)
        new_xyz = O.gather(xyz, fidx)
        assert np.array_equal(new_xyz, z[f"{name}.new_xyz"]), name
        idx, gxyz = O.knn_group_xyz(pm(xyz), pm(new_xyz), k)
        assert np.array_equal(idx, z[f"{name}.idx"]), name
        assert np.array_equal(gxyz, z[f"{name}.grouped_xyz"]), name
        if f"{name}.points" in z.files:
            new_points = np.concatenate([gxyz, O.group(z[f"{name}.points"], idx)], 1)
            assert np.array_equal(new_points, z[f"{name}.new_points"]), name


def test_metrics_against_reference_expressions():
    z = load_golden("next")
    for name in ("dcd", "dcddup", "dcdnear"):
        d1, d2, i1, i2 = (z[f"{name}.{k}"] for k in ("dist1", "dist2", "idx1", "idx2"))
        n_gt, n_x = d1.shape[1], d2.shape[1]
        # the stored dist/idx are the reference kernel's: the oracle's Chamfer must reproduce them
        od1, od2, oi1, oi2 = O.chamfer_fwd(z[f"{name}.gt"], z[f"{name}.x"])
        assert np.array_equal(od1, d1) and np.array_equal(od2, d2) and np.array_equal(oi1, i1) and np.array_equal(oi2, i2)
        m = O.chamfer_metrics(d1, d2, i1, i2, frac1=n_gt / n_x, frac2=n_x / n_gt)
        assert rel((m[:, 0] + m[:, 1]) / 2, z[f"{name}.cd_p"]) < 1e-5, name
        assert rel(m[:, 2] + m[:, 3], z[f"{name}.cd_t"]) < 1e-5, name
        assert np.allclose(m[:, 4], z[f"{name}.p1"], rtol=1e-6, atol=0) and np.allclose(m[:, 5], z[f"{name}.p2"], rtol=1e-6, atol=0)
        assert np.allclose(m[:, 6], z[f"{name}.f1"], rtol=1e-5, atol=1e-7), name
        assert np.allclose(m[:, 7], z[f"{name}.dcd"], rtol=1e-5, atol=1e-7), name
        m2 = O.chamfer_metrics(d1, d2, i1, i2, threshold=0.01, alpha=40, n_lambda=0.5,
                               frac1=max(1, n_gt / n_x), frac2=max(1, n_x / n_gt))
        assert np.allclose(m2[:, 6], z[f"{name}.f1_t01"], rtol=1e-5, atol=1e-7), name
        # terms are 1 - exp(..) * w in [0, 1]: torch sums them in fp32, so allow 1e-7 absolute on the mean
        assert np.allclose(m2[:, 7], z[f"{name}.dcd_nonreg"], rtol=1e-5, atol=1e-7), name
    # the near-duplicate case has a non-trivial F-score
    assert (z["dcdnear.f1"] > 0).any() and (z["dcdnear.f1_t01"] > 0.5).all()


def test_index_points_and_gradients_small():
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((2, 11, 5)).astype(np.float32)
    idx = rng.integers(0, 11, (2, 7, 3)).astype(np.int32)
    out = O.index_points(pts, idx)
    assert np.array_equal(out, np.stack([pts[b][idx[b]] for b in range(2)]))
    g = rng.standard_normal(out.shape).astype(np.float32)
    want = np.zeros_like(pts, dtype=np.float64)
    for b in range(2):
        np.add.at(want[b], idx[b].reshape(-1), g[b].reshape(-1, 5).astype(np.float64))
    assert rel(O.index_points_grad(g, idx, 11), want) < 1e-6
    # edge-feature gradient against a finite expression: out = cat(x[n] - x[idx], x[n])
    x = rng.standard_normal((2, 4, 9)).astype(np.float32)
    eidx = rng.integers(0, 9, (2, 9, 3)).astype(np.int32)
    go = rng.standard_normal((2, 8, 9, 3)).astype(np.float32)
    want = (go[:, :4] + go[:, 4:]).astype(np.float64).sum(-1)
    for b in range(2):
        for c in range(4):
            np.subtract.at(want[b, c], eidx[b].reshape(-1), go[b, c].reshape(-1).astype(np.float64))
    assert rel(O.edge_features_grad(go, eidx), want) < 1e-6
