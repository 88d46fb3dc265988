"""--code 1|2, --send_faces, --blocking_send in the reference itself, at N
ranks on the CPU: which combinations deliver exactly what --code 0 delivers on every cell
the stencil reads (the device path runs its code-0 exchange for those), and the one that
does not (--code 1|2 with --permute and a wide stencil: mamr_create refuses it)."""
import os

import numpy as np
import pytest

from mputil import MPIRUN, defined_mask, run_ranks
from oracle import refharness

needs = pytest.mark.skipif(not (os.path.exists(MPIRUN) and refharness.available("ref_mp")),
                           reason="minimpi/_bin or oracle/_ref/libminiamr_ref_mp.so not built")
MOVING = "--num_objects 1 --object 2 0 0.2 0.2 0.2 0.09 0.07 0.05 0.2 0.2 0.2 0 0 0"
AMR7 = (f"--npx 2 --init_x 1 --init_y 2 --init_z 2 --nx 4 --ny 6 --nz 4 --num_vars 3 --comm_vars 2 --stencil 7 "
        f"--num_refine 2 --max_blocks 2000 --refine_freq 1 --num_tsteps 3 --stages_per_ts 7 --lb_opt 1 {MOVING}")
UNI = ("--npy 2 --init_x 2 --init_y 1 --init_z 2 --nx 4 --ny 6 --nz 4 --num_vars {nv} --stencil {st} "
       "--uniform_refine 1 --num_refine 1 --max_blocks 200 --num_tsteps 2 --stages_per_ts 7")


def same_on_read_cells(a, b, stencil):
    m = defined_mask(4, 6, 4, stencil)[None]
    return all(not ((a["blocks"][k][2].view(np.uint64) != b["blocks"][k][2].view(np.uint64)) & m).any()
               for k in a["blocks"]) and np.array_equal(a["sums"], b["sums"])


@needs
@pytest.mark.parametrize("base,stencil", [
    (AMR7, 7), (AMR7 + " --permute", 7),
    (UNI.format(nv=3, st=27), 27), (UNI.format(nv=9, st=0), 27),
])
def test_modes_that_equal_code0(base, stencil):
    ref = run_ranks("ref_mp", 2, base.split())
    for mode in ("--code 1", "--code 2", "--send_faces", "--blocking_send", "--code 2 --send_faces"):
        got = run_ranks("ref_mp", 2, (base + " " + mode).split())
        assert same_on_read_cells(ref, got, stencil), mode


@needs
@pytest.mark.parametrize("nv,st", [(3, 27), (9, 0)])
def test_the_combination_that_does_not(nv, st):
    base = UNI.format(nv=nv, st=st) + " --permute"
    ref = run_ranks("ref_mp", 2, base.split())
    assert same_on_read_cells(ref, run_ranks("ref_mp", 2, (base + " --send_faces").split()), 27)
    for mode in ("--code 1", "--code 2"):
        got = run_ranks("ref_mp", 2, (base + " " + mode).split())
        assert not same_on_read_cells(ref, got, 27), mode      # hence MAMR_EUNSUPPORTED (api.cu: mamr_create)
