"""The off-rank path on ONE GPU (runs under the driver's single-GPU `pytest -m gpu`): N ranks as
N contexts in this process, connected through the peer-memory transport (tests/loopback.py).
Same cases and same oracle comparison as tests/test_multi_gpu.py (which needs N GPUs): pack
kernels (boxop_kernel) reading resolved origins and earlier receive buffers, widened Y/Z faces
forwarded through up to three ranks, receive-buffer reads inside fused2 / slab7, staged comm
groups, --permute, and the all-reduce of check_sum (comm.c:254-401,1002-1150; check_sum.c:57)."""
import os

import pytest

from loopback import run_uniform_case
from test_multi_gpu import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_ranks_on_one_gpu_match_single_rank_oracle(case, fused):
    old = os.environ.get("MAMR_NO_FUSED")
    os.environ["MAMR_NO_FUSED"] = "0" if fused else "1"
    try:
        out = run_uniform_case(CASES[case])
    finally:
        if old is None:
            os.environ.pop("MAMR_NO_FUSED", None)
        else:
            os.environ["MAMR_NO_FUSED"] = old
    assert len(out) == CASES[case]["np"][0]*CASES[case]["np"][1]*CASES[case]["np"][2]


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
    run_uniform_case(BASELINE_SHAPES[case])
