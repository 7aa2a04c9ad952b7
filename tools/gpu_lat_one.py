"""One 64-agent mission, Gauss-Seidel chain at batch size BS (default 1): kernel time of rbpe_run (for ncu captures and A/B of a
library given by LAT_LIB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swarm_simulator_b200 import engine as E, synth
if os.environ.get("LAT_LIB"):
    E.LIB_PATH = os.path.join(ROOT, os.environ["LAT_LIB"])
ms = synth.load_pack(os.path.join(ROOT, "tests", "golden", "missions_cfg3.npz"), select=range(1))
eng = E.Engine()
prob = E.PackedProblem(synth.pack(ms), sequential=True, batch_size=int(os.environ.get("BS", "1")))
eng.upload(prob)
for rep in range(2):
    eng.timer_start(); eng.run(E.MODE_GAUSS_SEIDEL); print("BS=%s %.3f ms" % (os.environ.get("BS", "1"), eng.timer_stop()))
eng.close()
