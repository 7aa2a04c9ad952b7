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
eng = E.Engine()
for label, seq, bs, mode in (("jacobi sweep b=1 (64 QPs in parallel)", True, 1, E.MODE_JACOBI), ("gauss-seidel b=1 (chain of 64)", True, 1, E.MODE_GAUSS_SEIDEL),
                             ("gauss-seidel b=4 (chain of 16)", True, 4, E.MODE_GAUSS_SEIDEL)):
    prob = E.PackedProblem(synth.pack(ms), sequential=seq, batch_size=bs)
    eng.upload(prob)
    best = 1e9
    for rep in range(5):
        eng.timer_start(); eng.run(mode); best = min(best, eng.timer_stop())
    r = eng.download(prob)
    print("%-42s %d mission(s): %.3f ms  (rc %d, mean iters %.1f)" % (label, nm, best, r.rc, r.qp_iters.mean()), flush=True)
eng.close()
