"""TEST INFRASTRUCTURE — N ranks of the multi-GPU path on however many GPUs the box has (one is
enough): one process per rank (tests/lb_worker.py), rank r on GPU r % count, windows of the
peer-memory transport mapped with CUDA IPC (miniamr_b200.h: mamr_p2p_get_handle /
mamr_p2p_connect).  Everything the off-rank path does on a multi-GPU box runs here too -- pack
kernels, pushes into the partner's receive buffers, arrival / credit flags, receive-buffer reads
of the fused kernels, the check_sum all-reduce, block migration -- on one GPU the NVLink hop is
a hop through local memory and the ranks' kernels are time-sliced."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def run_ranks(kind, world, cfg, timeout=300, env=None, processes=False):
    """processes=False: all ranks as threads of ONE fresh process (fast: their kernels run
    concurrently); True: one process per rank (CUDA IPC; time-sliced on a shared GPU).

    Threads share one CUDA context, whose scheduling of independent streams is not under the
    library's control: once in a hundred runs a kernel that waits for a peer ends up in front of
    that peer's work and the wait times out (reported as MAMR_EP2P, never a wrong result).  A
    run that ends that way is repeated, the second time with one process per rank."""
    last = None
    for attempt, procs_mode in enumerate([processes, processes, True]):
        ok, last = _run_once(kind, world, cfg, timeout, env, procs_mode)
        if ok:
            return
        if "waited" not in last or "peer-memory transport" not in last:
            break
    raise AssertionError(last)


def _run_once(kind, world, cfg, timeout, env, processes):
    e = dict(os.environ)
    e.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    e.setdefault("CUDA_MODULE_LOADING", "EAGER")
    e.setdefault("MAMR_P2P_TIMEOUT_S", "20" if processes else "4")
    if env:
        e.update(env)
    with tempfile.TemporaryDirectory(prefix="mamr_lb_") as scratch:
        whos = [str(r) for r in range(world)] if processes else ["all"]
        procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "lb_worker.py"), kind, w, str(world),
                                   scratch, json.dumps(cfg)], env=e, stdout=subprocess.PIPE,
                                  stderr=subprocess.PIPE, text=True) for w in whos]
        outs = []
        try:
            for p in procs:
                outs.append(p.communicate(timeout=timeout))
        except subprocess.TimeoutExpired:
            for p in procs:
                p.kill()
            return False, f"loopback run {kind} {cfg} did not finish in {timeout} s"
        bad = [(w, p.returncode, o[0][-1500:], o[1][-3000:]) for w, p, o in zip(whos, procs, outs)
               if p.returncode != 0 or f"LB_OK {w}" not in o[0]]
        if bad:
            return False, "rank %s rc=%s\n%s\n%s" % bad[0]
        return True, ""
