import importlib.util
import os
import os.path as osp
import sys

import numpy as np
import pytest

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = osp.join(ROOT, "tests", "golden")
REF_DIR = osp.join(ROOT, "oracle", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    path = osp.join(GOLDEN, name + ".npz")
    if not osp.exists(path):
        pytest.skip(f"{path} not generated yet")
    return np.load(path)


_ref_cache = {}


def load_ref_ext(name):
    """The reference's own CUDA ops (oracle/_ref/<name>.so, built by oracle/build_ref.py)."""
    if name in _ref_cache:
        return _ref_cache[name]
    path = osp.join(REF_DIR, name + ".so")
    if not osp.exists(path):
        pytest.skip(f"{path} not built (oracle/build_ref.py needs /root/reference)")
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _ref_cache[name] = mod
    return mod


def make_cloud(gen, B, N, dup=0, near_origin=0):
    """Seeded rand-0.5 cloud; `dup` duplicates points like UpSamplePoints does."""
    import torch
    x = torch.rand(B, N, 3, generator=gen) - 0.5
    if dup:
        uniq = N - dup
        for b in range(B):
            src = torch.randint(0, uniq, (dup,), generator=gen)
            x[b, uniq:] = x[b, src]
    if near_origin:
        for b in range(B):
            pos = torch.randperm(N, generator=gen)[:near_origin]
            x[b, pos] = (torch.rand(near_origin, 3, generator=gen) - 0.5) * 0.03
    return x.contiguous()
