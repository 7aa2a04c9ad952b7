import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from swarm_simulator_b200 import engine as E, synth
import oracle_util
ms = synth.load_pack(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "missions_cfg3.npz"), select=range(64))
prob = E.PackedProblem(synth.pack(ms), sequential=True, batch_size=1)
out = {}
for lat in ("1", "0"):
    os.environ["RBPE_LAT"] = lat
    e = E.Engine(device=0)
    out[lat] = e.solve_many(prob)
    e.close()
a, b = out["1"], out["0"]
d = np.abs(a.ctrl - b.ctrl).reshape(64, -1).max(1)
print("max diff per mission (top 5):", np.sort(d)[-5:], "argmax", d.argmax())
print("iters equal:", np.array_equal(a.qp_iters, b.qp_iters))
c = int(d.argmax())
ro = oracle_util.oracle_problem(ms[c], sequential=True, batch_size=1).update()
print("mission", c, "x1 vs oracle", np.abs(a.ctrl[c] - ro["ctrl"]).max(), "pdip1 vs oracle", np.abs(b.ctrl[c] - ro["ctrl"]).max())
dq = np.abs(a.ctrl[c] - b.ctrl[c]).reshape(64, -1).max(1)
print("per agent diff:", np.round(np.log10(dq + 1e-30), 1))
print("res x1:", a.qp_res[c][:, :].max(0), "res pdip1:", b.qp_res[c].max(0))
