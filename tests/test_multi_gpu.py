"""N > 1 on real GPUs: blocks sharded over ranks, ghost faces over NCCL.  Needs
at least 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


CASES = [
    dict(np=[2, 1, 1], n=[4, 6, 8], b=[2, 2, 2], vars=3, stencil=7, stages=3, seed=1, check_comm=1),
    dict(np=[2, 1, 1], n=[4, 6, 8], b=[2, 3, 2], vars=3, stencil=27, stages=3, seed=2, check_comm=1),
    dict(np=[1, 2, 1], n=[6, 4, 4], b=[2, 2, 3], vars=4, stencil=27, stages=4, seed=3, comm_vars=3),
    dict(np=[1, 1, 2], n=[4, 4, 6], b=[3, 2, 2], vars=2, stencil=27, stages=7, seed=4, permute=1),
    dict(np=[1, 1, 2], n=[16, 16, 16], b=[2, 2, 2], vars=2, stencil=27, stages=2, seed=5),
    dict(np=[2, 1, 1], n=[10, 10, 10], b=[2, 2, 2], vars=5, stencil=7, stages=3, seed=6, comm_vars=2),
    # the fixed-size kernel with ghost elision: several comm() groups per stage (stale ghost
    # layers of the other groups must survive the reuse of the receive buffers), 7-point
    dict(np=[2, 1, 1], n=[16, 16, 16], b=[2, 2, 3], vars=5, stencil=27, stages=3, seed=7, comm_vars=2),
    dict(np=[1, 2, 1], n=[16, 16, 16], b=[2, 2, 2], vars=3, stencil=7, stages=3, seed=8),
    # streamed 7-point kernel: off-rank faces in X, Y and Z (cells out of the receive buffers)
    dict(np=[2, 1, 1], n=[32, 32, 32], b=[1, 2, 2], vars=3, stencil=7, stages=3, seed=9, comm_vars=2),
    dict(np=[1, 2, 1], n=[32, 32, 32], b=[2, 1, 2], vars=2, stencil=7, stages=3, seed=10),
    dict(np=[1, 1, 2], n=[32, 32, 32], b=[2, 2, 1], vars=2, stencil=7, stages=3, seed=11),
    # 4 and 8 ranks: partners in two / three directions -- edges and corners resolve through a
    # chain of receive buffers, and the pack of a later phase reads an earlier phase's receive buffer
    dict(np=[2, 2, 1], n=[4, 6, 8], b=[2, 2, 2], vars=3, stencil=27, stages=3, seed=12, comm_vars=2),
    dict(np=[1, 2, 2], n=[16, 16, 16], b=[2, 2, 2], vars=3, stencil=27, stages=3, seed=13, permute=1),
    dict(np=[2, 1, 2], n=[32, 32, 32], b=[2, 2, 2], vars=2, stencil=7, stages=3, seed=14),
    dict(np=[2, 2, 2], n=[16, 16, 16], b=[2, 2, 2], vars=4, stencil=27, stages=4, seed=15, comm_vars=3),
    dict(np=[2, 2, 2], n=[10, 10, 10], b=[2, 3, 2], vars=3, stencil=7, stages=3, seed=16, permute=1),
]


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_two_ranks_match_single_rank_oracle(case, fused):
    cfg = CASES[case]
    world = cfg["np"][0]*cfg["np"][1]*cfg["np"][2]
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    env["MAMR_NO_FUSED"] = "0" if fused else "1"
    port = 29600 + case*2 + fused
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), json.dumps(cfg)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("case", [4, 6, 8, 12, 14])
def test_nccl_transport_matches_single_rank_oracle(case):
    """the same through ncclSend/ncclRecv and ncclAllReduce (MAMR_TRANSPORT=nccl) instead of the
    peer-memory transport"""
    cfg = CASES[case]
    world = cfg["np"][0]*cfg["np"][1]*cfg["np"][2]
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    env["MAMR_TRANSPORT"] = "nccl"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + case),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), json.dumps(cfg)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
