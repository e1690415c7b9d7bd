"""Build recipe for the CUDA library and the per-shape reference-ABI shims (sm_100a only).

Everything is built IN-TREE under nav-slam_b200/_build/ (git-ignored, shipped to the GPU box by
gpurun).  nvcc cross-compiles without a GPU, so this also runs on the CPU-only build host.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SHIM = os.path.join(HERE, "shim")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libnavslam_b200.so")

CU_SOURCES = ["stencil.cu", "rowmap.cu", "kdbuild.cu", "kdtree.cu", "bf_tc.cu", "io.cu", "csvfmt.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # nothing that decides an output may be contracted into an FMA
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-Xptxas", "-v",
]
SHIM_SHAPES = [(8, 8), (5, 33), (16, 1800), (64, 2048)]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _deps():
    out = [os.path.join(ROOT, "include", "navslam_b200.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def build_library(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    if not force and _newer(LIB, _deps()):
        return LIB
    objs = []
    for src in CU_SOURCES:
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        spath = os.path.join(CSRC, src)
        if force or not _newer(obj, _deps()):
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", spath, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            with open(obj + ".ptxas.txt", "w") as f:
                f.write(res.stderr)
        objs.append(obj)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    subprocess.check_call(cmd)
    return LIB


def shim_path(rows: int, cols: int) -> str:
    return os.path.join(BUILD, f"libnavslam_shim_{rows}x{cols}.so")


def build_shims(shapes=SHIM_SHAPES, force: bool = False):
    """Per-shape drop-in libraries exporting the reference's own symbols (headers/slam.h,
    utils/kdtree.h, utils/pointcloud.h) on top of libnavslam_b200.so."""
    src = os.path.join(SHIM, "navslam_shim.c")
    if not os.path.exists(src):
        return []
    cc = os.environ.get("CC") or shutil.which("gcc") or "gcc"
    out = []
    for r, c in shapes:
        so = shim_path(r, c)
        deps = [src, os.path.join(ROOT, "include", "navslam_ref_abi.h"),
                os.path.join(ROOT, "include", "navslam_b200.h"), LIB]
        if force or not _newer(so, deps):
            subprocess.check_call([
                cc, "-std=gnu11", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
                f"-DMAX_ROWS={r}", f"-DMAX_COLS={c}", f"-I{os.path.join(ROOT, 'include')}", src,
                f"-L{BUILD}", "-lnavslam_b200", "-Wl,-rpath,$ORIGIN", "-lm", "-o", so])
        out.append(so)
    return out


def build_all(verbose: bool = False, force: bool = False):
    lib = build_library(verbose=verbose, force=force)
    shims = build_shims(force=force)
    return lib, shims


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
