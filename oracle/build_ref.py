"""Build the UNMODIFIED reference CUDA ops into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE — never imported by the product package.  Only tests/,
__graft_entry__.smoke() and bench.py's baseline legs may load what this builds.

The reference's own sources are compiled *where they lie* under /root/reference
(nothing is copied into this repository); only the resulting shared objects land
in oracle/_ref/, which is git-ignored but travels to the GPU box with gpurun.

Sources compiled (reference file list):
  metrics/CD/chamfer3D/{chamfer_cuda.cpp, chamfer3D.cu}           -> ref_chamfer_3D.so
  pointnet2_ops_lib/pointnet2_ops/_ext-src/src/*.{cpp,cu}         -> ref_pointnet2_ext.so

We do NOT run the reference's build system: its setup.py / JIT fallback force
TORCH_CUDA_ARCH_LIST="3.7+PTX;..." (pointnet2_utils.py:23, setup.py:19) which
CUDA 12.9 rejects.  Instead torch.utils.cpp_extension.load is called with
-gencode arch=compute_100a,code=sm_100a and the reference's own -O3.
"""
import glob
import os
import os.path as osp
import sys

HERE = osp.dirname(osp.abspath(__file__))
OUT = osp.join(HERE, "_ref")
REF = os.environ.get("POINTSEA_REFERENCE", "/root/reference")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def build(verbose=False):
    if not osp.isdir(REF):
        print(f"[oracle/_ref] {REF} not present; using prebuilt objects if any")
        return False
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    cd = osp.join(REF, "metrics", "CD", "chamfer3D")
    bdir = osp.join(OUT, "build_chamfer")
    os.makedirs(bdir, exist_ok=True)
    load(
        name="ref_chamfer_3D",
        sources=[osp.join(cd, "chamfer_cuda.cpp"), osp.join(cd, "chamfer3D.cu")],
        extra_cuda_cflags=ARCH,
        build_directory=bdir,
        verbose=verbose,
    )
    pn = osp.join(REF, "pointnet2_ops_lib", "pointnet2_ops", "_ext-src")
    bdir2 = osp.join(OUT, "build_pointnet2")
    os.makedirs(bdir2, exist_ok=True)
    load(
        name="ref_pointnet2_ext",
        sources=sorted(glob.glob(osp.join(pn, "src", "*.cpp")) + glob.glob(osp.join(pn, "src", "*.cu"))),
        extra_include_paths=[osp.join(pn, "include")],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3"] + ARCH,
        build_directory=bdir2,
        verbose=verbose,
    )
    import shutil
    shutil.copy(osp.join(bdir, "ref_chamfer_3D.so"), osp.join(OUT, "ref_chamfer_3D.so"))
    shutil.copy(osp.join(bdir2, "ref_pointnet2_ext.so"), osp.join(OUT, "ref_pointnet2_ext.so"))
    return True


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("built" if ok else "skipped")
