"""Host-side logic of the fused path at N ranks on REFINED meshes, on the CPU:
the reference itself (N ranks on the minimpi back-end) produces the partition,
the topology and the off-rank comm lists — whole faces, fine-to-coarse quarters
and coarse-to-fine quarters across rank boundaries, after RCB/SFC migration —
and the halo planner's pack ops + halo plan, executed with numpy and exchanged
over the host channel, must reproduce the reference's comm() bit for bit."""
import glob
import os
import subprocess
import sys
import tempfile

import pytest

from mputil import MPIRUN, ROOT
from oracle import refharness
from miniamr_b200 import build

build.build()
needs = pytest.mark.skipif(not (os.path.exists(MPIRUN) and refharness.available("ref_mp")),
                           reason="minimpi/_bin or oracle/_ref/libminiamr_ref_mp.so not built")
MOVING = "--num_objects 1 --object 2 0 0.2 0.2 0.2 0.09 0.07 0.05 0.2 0.2 0.2 0 0 0"
CASES = {
    "amr7_rcb_2": (2, 2, f"--npx 2 --init_x 1 --init_y 2 --init_z 2 --nx 4 --ny 4 --nz 4 --num_vars 3 "
                         f"--comm_vars 2 --num_refine 3 --block_change 1 --max_blocks 3000 --refine_freq 1 "
                         f"--num_tsteps 4 --stages_per_ts 2 --lb_opt 1 {MOVING}"),
    "amr7_rcb_8": (8, 1, f"--npx 2 --npy 2 --npz 2 --init_x 1 --init_y 1 --init_z 1 --nx 4 --ny 6 --nz 4 "
                         f"--num_vars 2 --num_refine 3 --max_blocks 3000 --refine_freq 1 --num_tsteps 3 "
                         f"--stages_per_ts 2 --lb_opt 1 {MOVING}"),
    "amr7_hilbert_permute_4": (4, 6, f"--npx 2 --npy 2 --init_x 1 --init_y 1 --init_z 2 --nx 4 --ny 4 --nz 6 "
                                     f"--num_vars 2 --num_refine 2 --max_blocks 2000 --refine_freq 1 "
                                     f"--num_tsteps 3 --stages_per_ts 2 --hilbert --permute {MOVING}"),
    # message modes: --send_faces / --blocking_send change the transport only; --code 1|2 also
    # what a message holds -- the cells the stencil reads must not change
    "amr7_code1_sendfaces_4": (4, 2, f"--npx 2 --npy 2 --init_x 1 --init_y 1 --init_z 2 --nx 4 --ny 6 --nz 4 "
                                     f"--num_vars 3 --comm_vars 2 --num_refine 3 --max_blocks 3000 --refine_freq 1 "
                                     f"--num_tsteps 3 --stages_per_ts 2 --lb_opt 1 --code 1 --send_faces {MOVING}"),
    "amr7_code2_blocking_2": (2, 2, f"--npz 2 --init_x 2 --init_y 2 --init_z 1 --nx 4 --ny 4 --nz 4 --num_vars 2 "
                                    f"--num_refine 2 --max_blocks 2000 --refine_freq 1 --num_tsteps 3 "
                                    f"--stages_per_ts 2 --code 2 --blocking_send {MOVING}"),
    "uni27_code1_2": (2, 2, "--npy 2 --init_x 2 --init_y 1 --init_z 1 --nx 4 --ny 4 --nz 6 --num_vars 3 "
                            "--stencil 27 --uniform_refine 1 --num_refine 1 --max_blocks 100 --num_tsteps 1 "
                            "--stages_per_ts 2 --code 1"),
    "uni27_4": (4, 2, "--npx 2 --npz 2 --init_x 1 --init_y 2 --init_z 1 --nx 4 --ny 4 --nz 6 --num_vars 3 "
                      "--comm_vars 2 --stencil 27 --uniform_refine 1 --num_refine 1 --max_blocks 100 "
                      "--num_tsteps 1 --stages_per_ts 2"),
}


@needs
@pytest.mark.parametrize("name", sorted(CASES))
def test_plan_reproduces_reference_comm_at_n_ranks(name):
    n, stages, args = CASES[name]
    with tempfile.TemporaryDirectory(prefix="mamr_plan_") as out:
        cmd = [MPIRUN, "-n", str(n), sys.executable, os.path.join(ROOT, "tests", "mp_plan_worker.py"),
               out, str(stages)] + args.split()
        env = dict(os.environ, OMP_NUM_THREADS="1")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        lines = [open(f).read() for f in sorted(glob.glob(os.path.join(out, "rank*.txt")))]
    assert len(lines) == n and all(l.startswith("PLAN_OK") for l in lines), lines
    assert any("offrank_faces=0" not in l for l in lines)
    if name.startswith("amr"):
        assert any("levels=[" in l and "," in l.split("levels=")[1] for l in lines), lines
