"""A/B of alternative builds of librbpe.so (tools only): python tools/gpu_ab.py lib1.so lib2.so ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swarm_simulator_b200 import engine as E, synth
npool = int(os.environ.get("AB_POOL", 64))
ms = synth.load_pack(os.path.join(ROOT, "tests", "golden", "missions_cfg3.npz"), select=range(npool))
count = int(os.environ.get("AB_COUNT", 2368))
packed = synth.pack([ms[i % npool] for i in range(count)])
ref_ctrl = None
for lib in sys.argv[1:]:
    E._lib = E.load_library(os.path.join(ROOT, lib))
    eng = E.Engine()
    prob = E.PackedProblem(packed, sequential=True, batch_size=1)
    eng.upload(prob)
    best = 1e9
    for rep in range(3):
        eng.timer_start(); eng.run(); best = min(best, eng.timer_stop())
    r = eng.download(prob)
    if ref_ctrl is None:
        ref_ctrl = r.ctrl.copy()
    print("%-40s rc=%d kernel %.1f ms -> %.0f agent-QPs/s iters %.2f  max |ctrl - ctrl of the first library| %.3e" % (lib, r.rc, best, count * 64 / best * 1e3, r.qp_iters.mean(), float(np.abs(r.ctrl - ref_ctrl).max())), flush=True)
    eng.close()
