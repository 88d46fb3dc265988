"""Build libminiamr_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree.

    python -m miniamr_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the
GPU box with the repository snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libminiamr_b200.so")
SRCS = sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))
NO_FMA = {"stencil0.cu"}
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
DEPS = SRCS + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + \
    [os.path.join(ROOT, "include", "miniamr_b200.h")]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    # the image exports CC/CXX pointing at a toolchain without libgomp specs;
    # nvcc only needs a host g++
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    common = [nvcc_path(), "-ccbin", host, "-O3", "-std=c++17",
              "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-O3,-Wall"]
    if verbose:
        common.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # files whose arithmetic must follow the reference's expression trees operation by
    # operation are compiled without FMA contraction (the other kernels only add and use
    # explicit fma() where they mean it)
    # one object per source, compiled in parallel; an object is reused while it is newer than
    # its source and every header
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = max(os.path.getmtime(d) for d in DEPS if not d.endswith(".cu"))
    jobs, objs = [], []
    for src in SRCS:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_t):
            continue
        cmd = common + (["-fmad=false"] if os.path.basename(src) in NO_FMA else []) + ["-c", src, "-o", obj]
        jobs.append(cmd)

    def run(cmd):
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd, env=env)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    cmd = common + ["-shared", "-Xlinker", "-soname=libminiamr_b200.so", "-o", LIB] + objs + ["-ldl"]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
