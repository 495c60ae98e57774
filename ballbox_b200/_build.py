"""Build recipes: nvcc for the sm_100a kernels, gcc for the C host library.

Everything is built IN-TREE (ballbox_b200/libballbox_b200.so) so that the shared
object travels with a snapshot of the repository.  The oracle libraries are built
by oracle/Makefile; building a checker is not using it.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ballbox_b200")
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libballbox_b200.so")

CUDA_SOURCES = ["bbx_state.cu", "bbx_integrate.cu", "bbx_broadphase.cu", "bbx_narrowphase.cu",
                "bbx_solver.cu", "bbx_solver_flow.cu", "bbx_replay.cu", "bbx_colour.cu", "bbx_joints.cu", "bbx_step.cu", "bbx_sort.cu", "bbx_queries.cu"]
HOST_SOURCES = ["ballbox_host.c", "bbx_scenes.c"]

# -fmad=false: no FMA contraction, results must match the reference bit for bit
# (SURVEY.md F5).  IEEE sqrt/div are nvcc defaults; never --use_fast_math.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-Xptxas", "-v"]
GCC_FLAGS = ["-O2", "-std=gnu11", "-ffp-contract=off", "-fPIC", "-DBBX_PRODUCT", "-Wall", "-Wno-unused-function"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    hs = [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    hs += [os.path.join(CSRC, "cuda", f) for f in os.listdir(os.path.join(CSRC, "cuda")) if f.endswith(".cuh")]
    return hs


def _run(cmd, log):
    p = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + p.stdout + p.stderr)


def build_product(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA layer and the host library into libballbox_b200.so."""
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    hdrs = _headers()
    jobs, objs = [], []
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, "cuda", src)
        o = os.path.join(BUILD, src + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append(([nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-I", os.path.join(CSRC, "cuda"), "-c", s, "-o", o],
                         os.path.join(BUILD, src + ".log")))
    for src in HOST_SOURCES:
        s = os.path.join(CSRC, "host", src)
        o = os.path.join(BUILD, src + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append((["gcc"] + GCC_FLAGS + ["-I", INCLUDE, "-c", s, "-o", o], os.path.join(BUILD, src + ".log")))
    if jobs:
        with ThreadPoolExecutor(max_workers=8) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    if jobs or force or _newer(LIB, objs):
        _run([nvcc, "-shared", "-o", LIB] + objs + ["-lm"], os.path.join(BUILD, "link.log"))
    if verbose:
        for src in CUDA_SOURCES:
            log = os.path.join(BUILD, src + ".log")
            if os.path.exists(log):
                print(open(log).read())
    return LIB


def build_oracles(verbose: bool = False) -> None:
    """Compile oracle/ (our C restatement) and, where the reference header is
    present, oracle/_ref (the unmodified reference).  Test infrastructure."""
    odir = os.path.join(ROOT, "oracle")
    for target in ("all", "ref"):
        p = subprocess.run(["make", "-s", "-C", odir, target], capture_output=True, text=True)
        if verbose or p.returncode != 0:
            print(p.stdout + p.stderr)
        if p.returncode != 0:
            raise RuntimeError("oracle build failed (make %s)" % target)


if __name__ == "__main__":
    import sys
    print(build_product(force="--force" in sys.argv, verbose="-v" in sys.argv))
