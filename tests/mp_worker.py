"""TEST INFRASTRUCTURE — one rank of a multi-rank whole run, started by

    minimpi/_bin/minimpirun -n N python tests/mp_worker.py VARIANT OUTDIR <miniAMR args>

VARIANT "ref_mp": the unmodified reference on the CPU; "int_mp": the same host
code on the CUDA stage path (one rank per GPU).  The rank runs the reference's
own driver() to the end, does one more comm() (so that every ghost face is
defined, cf. tests/test_integration.py), and writes OUTDIR/rank<r>.npz with its
blocks keyed by global block number."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refharness  # noqa: E402


def main():
    variant, outdir, args = sys.argv[1], sys.argv[2], sys.argv[3:]
    r = refharness.RefMiniAMR(args, variant=variant, run_driver=True)
    p = r.p
    cv = p["comm_vars"] if 0 < p["comm_vars"] <= p["num_vars"] else p["num_vars"]
    for start in range(0, p["num_vars"], cv):      # message buffers hold comm_vars variables
        r.comm(start, min(cv, p["num_vars"] - start), 0)
    r.sync_host()
    slots = r.sorted_slots()
    numbers = np.zeros(len(slots), np.int64)
    levels = np.zeros(len(slots), np.int32)
    data = np.zeros((len(slots), p["num_vars"]) + r.tile_shape, np.float64)
    for a, s in enumerate(slots):
        b = r.block(int(s))
        numbers[a], levels[a] = b["number"], b["level"]
        data[a] = r.get_slot(int(s))
    sums = np.array([r.lib.refh_get_grid_sum(v) for v in range(p["num_vars"])])
    c = r.counters()
    np.savez(os.path.join(outdir, f"rank{p['my_pe']}.npz"), numbers=numbers, levels=levels, data=data,
             sums=sums, counters=np.array(c["same"] + c["diff"] + c["bc"]),
             params=np.array([p[k] for k in refharness.P_NAMES]),
             global_active=r.global_active(), fp_adds=r.timers()["fp_adds"])


if __name__ == "__main__":
    main()
