"""The C-ABI library loads on a CPU-only box and exports every symbol include/pointsea_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os.path as osp
import re

from conftest import ROOT


def declared_symbols():
    hdr = open(osp.join(ROOT, "include", "pointsea_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for need in ("ps_chamfer_fwd", "ps_chamfer_bwd", "ps_fps", "ps_gather_fwd", "ps_gather_bwd", "ps_group_fwd",
                 "ps_group_bwd", "ps_ball_query", "ps_knn", "ps_three_nn", "ps_three_interpolate_fwd",
                 "ps_three_interpolate_bwd", "ps_last_error", "ps_version"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from svdformer_pointsea_b200 import _lib as L
    assert osp.exists(L.LIB_PATH), "run `python __graft_entry__.py build` first"
    lib = ctypes.CDLL(L.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in pointsea_b200.h but not exported"
    # and the Python binding table covers the header exactly
    assert sorted(L.EXPORTED_SYMBOLS) == declared_symbols()


def test_version_and_error_string_without_gpu():
    from svdformer_pointsea_b200 import _lib as L
    lib = L.load()
    assert lib.ps_version() >= 100
    assert isinstance(lib.ps_last_error(), bytes)
    assert lib.ps_launch_count(1) >= 0


def test_argument_validation_happens_before_any_cuda_call():
    """Bad sizes / null pointers are rejected with PS_ERR_INVALID_ARG and a message, GPU or not."""
    from svdformer_pointsea_b200 import _lib as L
    lib = L.load()
    rc = lib.ps_chamfer_fwd(None, None, None, None, None, None, 2, 0, 5, 0, None)
    assert rc == -1 and b"bad sizes" in lib.ps_last_error()
    rc = lib.ps_chamfer_fwd(None, None, None, None, None, None, 2, 4, 5, 0, None)
    assert rc == -1 and b"null pointer" in lib.ps_last_error()
    rc = lib.ps_knn(None, None, None, 1, 8, 4, 16, 0, 0, None)
    assert rc == -1 and b"exceeds the number of points" in lib.ps_last_error()
    rc = lib.ps_fps(None, None, 1, -3, 4, 0, None)
    assert rc == -1
    # empty work is a successful no-op (the reference's kernels simply do not iterate)
    assert lib.ps_fps(None, None, 0, 16, 4, 0, None) == 0
    assert lib.ps_gather_fwd(None, None, None, 0, 3, 16, 4, 0, None) == 0


def test_only_cuda_sources_in_the_product_and_no_oracle_import():
    """The product package must not import or link the oracle (no CPU fallback)."""
    import glob
    pkg = osp.join(ROOT, "svdformer_pointsea_b200")
    for path in glob.glob(osp.join(pkg, "**", "*.py"), recursive=True):
        src = open(path).read()
        assert "oracle" not in src.replace("# oracle", ""), f"{path} mentions the oracle"
    for path in glob.glob(osp.join(pkg, "csrc", "*.cu")) + glob.glob(osp.join(pkg, "csrc", "*.cuh")):
        assert "pointsea_oracle" not in open(path).read()
