"""TEST INFRASTRUCTURE — run tests/mp_worker.py at N ranks and collect the result."""
import glob
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MPIRUN = os.path.join(ROOT, "minimpi", "_bin", "minimpirun")


def run_ranks(variant, nranks, args, timeout=600, env=None):
    """-> {"blocks": {number: (rank, level, data)}, "sums": ..., "per_rank": [...]}"""
    with tempfile.TemporaryDirectory(prefix="mamr_mp_") as out:
        cmd = [MPIRUN, "-n", str(nranks), sys.executable, os.path.join(ROOT, "tests", "mp_worker.py"),
               variant, out] + [str(a) for a in args]
        e = dict(os.environ)
        e.setdefault("OMP_NUM_THREADS", "1")
        if env:
            e.update(env)
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e)
        assert r.returncode == 0, f"{' '.join(cmd)}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
        files = sorted(glob.glob(os.path.join(out, "rank*.npz")))
        assert len(files) == nranks, (files, r.stdout[-2000:], r.stderr[-2000:])
        res = dict(blocks={}, per_rank=[])
        for f in files:
            z = np.load(f)
            rank = int(os.path.basename(f)[4:-4])
            res["per_rank"].append(dict(rank=rank, n=len(z["numbers"]), counters=z["counters"].copy(),
                                        fp_adds=float(z["fp_adds"])))
            for a, num in enumerate(z["numbers"]):
                assert int(num) not in res["blocks"], f"block {num} active on two ranks"
                res["blocks"][int(num)] = (rank, int(z["levels"][a]), z["data"][a].copy())
            res["sums"] = z["sums"].copy()           # the same on every rank (allreduced)
            res["global_active"] = int(z["global_active"])
            res["params"] = z["params"].copy()
        return res


def defined_mask(nx, ny, nz, stencil):
    """cells the exchange defines (the 7-point comm never writes ghost edges/corners)"""
    m = np.ones((nx + 2, ny + 2, nz + 2), bool)
    if stencil == 7:
        gi = np.zeros(nx + 2, int); gi[[0, -1]] = 1
        gj = np.zeros(ny + 2, int); gj[[0, -1]] = 1
        gk = np.zeros(nz + 2, int); gk[[0, -1]] = 1
        m = (gi[:, None, None] + gj[None, :, None] + gk[None, None, :]) <= 1
    return m
