"""ncu target (tool): ONE 64-agent mission at the launch default b = 4 (Gauss-Seidel chain of 16 joint QPs), inputs resident."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from swarm_simulator_b200 import engine as E, synth
ms = synth.load_pack(os.path.join(ROOT, "tests", "golden", "missions_cfg3.npz"), select=[0])
eng = E.Engine()
prob = E.PackedProblem(synth.pack(ms), sequential=True, batch_size=4)
eng.upload(prob)
for rep in range(3):
    eng.timer_start(); eng.run(); t = eng.timer_stop()
print("b=4 single mission: %.2f ms" % t)
eng.close()
