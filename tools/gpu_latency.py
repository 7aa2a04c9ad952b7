"""Latency probe (not a test): one 64-agent mission, (a) one Jacobi sweep = 64 independent QPs in parallel -> single-QP
latency; (b) the reference's Gauss-Seidel chain at b = 1 and b = 4 -> single-mission latency.  Kernel times (CUDA events,
inputs resident).  RBPE_KERNEL=cta selects the CTA-per-QP kernel for one-agent batches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swarm_simulator_b200 import engine as E, synth
nm = int(os.environ.get("LAT_MISSIONS", 1))
ms = synth.load_pack(os.path.join(ROOT, "tests", "golden", "missions_cfg3.npz"), select=range(nm))
if os.environ.get("LAT_LIB"):            # e.g. a -DRBPE_PROFILE build: per-phase clock64 ticks of warp 0 are printed
    E.LIB_PATH = os.path.join(ROOT, os.environ["LAT_LIB"])
eng = E.Engine()
import ctypes as C
def prof(reset=True):
    lib = E.load_library()
    if not hasattr(lib, "rbpe_prof_read"):
        return None
    buf = (C.c_ulonglong * 16)()
    lib.rbpe_prof_read(buf, 1 if reset else 0)
    return [int(v) for v in buf]
for label, seq, bs, mode in (("jacobi sweep b=1 (64 QPs in parallel)", True, 1, E.MODE_JACOBI), ("gauss-seidel b=1 (chain of 64)", True, 1, E.MODE_GAUSS_SEIDEL),
                             ("gauss-seidel b=4 (chain of 16)", True, 4, E.MODE_GAUSS_SEIDEL)):
    prob = E.PackedProblem(synth.pack(ms), sequential=seq, batch_size=bs)
    eng.upload(prob)
    best = 1e9
    for rep in range(5):
        prof()
        eng.timer_start(); eng.run(mode); best = min(best, eng.timer_stop())
    pr = prof()
    if pr:
        print("   clock64 ticks of thread 0 per phase (last run):", pr[:12], flush=True)
    r = eng.download(prob)
    print("%-42s %d mission(s): %.3f ms  (rc %d, mean iters %.1f)" % (label, nm, best, r.rc, r.qp_iters.mean()), flush=True)
eng.close()
