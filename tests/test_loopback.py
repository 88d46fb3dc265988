"""The off-rank path on ONE GPU (runs under the driver's single-GPU `pytest -m gpu`): N ranks that
share the GPU -- threads of one fresh process per case, and for a few cases one process per rank
(CUDA IPC) -- connected through the peer-memory transport (tests/loopback.py).
Same cases and same oracle comparison as tests/test_multi_gpu.py (which needs N GPUs): pack
kernels (boxop_kernel) reading resolved origins and earlier receive buffers, widened Y/Z faces
forwarded through up to three ranks, receive-buffer reads inside fused2 / slab7, staged comm
groups, --permute, and the all-reduce of check_sum (comm.c:254-401,1002-1150; check_sum.c:57)."""
import os

import pytest

from loopback import run_ranks
from test_multi_gpu import CASES

pytestmark = pytest.mark.gpu


def world_of(cfg):
    return cfg["np"][0]*cfg["np"][1]*cfg["np"][2]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_ranks_on_one_gpu_match_single_rank_oracle(case):
    run_ranks("uniform", world_of(CASES[case]), CASES[case])


@pytest.mark.parametrize("case", [0, 1, 3, 5, 9, 12, 15])
def test_split_path_ranks_on_one_gpu(case):
    """MAMR_NO_FUSED=1: ghost cells materialised by the split path (pack / unpack kernels of
    ghost.cu around the same transport)"""
    run_ranks("uniform", world_of(CASES[case]), CASES[case], env={"MAMR_NO_FUSED": "1"})


BASELINE_SHAPES = [
    # BASELINE variable counts on 8 ranks: cfg2 (16^3, 40 vars, 27-pt), cfg3 (32^3, 40 vars, 7-pt),
    # cfg5 (10^3, 160 vars in four comm groups = four receive-buffer sets, 27-pt)
    dict(np=[2, 2, 2], n=[16, 16, 16], b=[2, 2, 2], vars=40, stencil=27, stages=3, seed=31),
    dict(np=[2, 2, 2], n=[32, 32, 32], b=[1, 1, 2], vars=40, stencil=7, stages=3, seed=32),
    dict(np=[2, 2, 2], n=[10, 10, 10], b=[2, 2, 2], vars=160, comm_vars=40, stencil=27, stages=3, seed=33),
    dict(np=[2, 2, 1], n=[10, 10, 10], b=[2, 2, 3], vars=40, stencil=7, stages=3, seed=34, permute=1),
]


@pytest.mark.parametrize("case", range(len(BASELINE_SHAPES)))
def test_baseline_variable_counts_on_eight_ranks(case):
    run_ranks("uniform", world_of(BASELINE_SHAPES[case]), BASELINE_SHAPES[case])


def test_block_migration_between_ranks_on_one_gpu():
    """mamr_stage_send_block / mamr_stage_recv_block / mamr_flush_block_moves over the windows
    (pull): 4 ranks, every rank sends two blocks to the next rank and one to the one after it.
    Half of the variables live in the second pool when the blocks move (rcb.c:207-337 payloads,
    pack.c:66-70 layout); the moved blocks then take part in one more stage."""
    run_ranks("migration", 4, dict(n=[4, 6, 8], vars=5))


@pytest.mark.parametrize("case", [0, 4, 7, 8])
def test_two_processes_share_one_gpu(case):
    """one PROCESS per rank: windows mapped with CUDA IPC (between processes on one device just as
    between devices); the ranks' kernels are time-sliced, which is why only 2-rank cases run so"""
    assert world_of(CASES[case]) == 2
    run_ranks("uniform", 2, CASES[case], processes=True)


def test_block_migration_between_two_processes():
    run_ranks("migration", 3, dict(n=[4, 4, 4], vars=4), processes=True)
