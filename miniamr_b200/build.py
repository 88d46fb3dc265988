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
    cmd = [nvcc_path(), "-O3", "-std=c++17",
           "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC,-O3,-Wall", "-shared",
           "-Xlinker", "-soname=libminiamr_b200.so", "-o", LIB] + SRCS + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a toolchain without libgomp specs;
    # nvcc only needs a host g++
    host = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    cmd[1:1] = ["-ccbin", host]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
