"""tests/golden/smoke_map32.npz: planner inputs of ONE map of the reference's smoke loop
(src/swarm_traj_planner_rbp_test_all.cpp L49-L103, launch/plan_rbp_test.launch): mission_64agents_15.json on worlds/map32.bt,
as swarm_simulator_b200/host/swarm_plan_cli produces them on the host (octomap .bt reader, ECBS with w = 1.5, SFC boxes).
That map yields two consecutive corridor boxes that only share a face, i.e. a batch QP without interior (the case the
presolve rule `cp_bounds` exists for).  Run where /root/reference is mounted:  python tests/golden/make_smoke_fixture.py"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, ROOT)
from swarm_simulator_b200 import synth  # noqa: E402

PKG = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/swarm_planner"
MAP = 32
PARAMS = ["ecbs/w=1.5", "grid/xy_res=0.5", "grid/z_res=1.0", "grid/margin=0.2", "world/z_min=0.3"]   # plan_rbp_test.launch

out = subprocess.check_output([os.path.join(ROOT, "swarm_simulator_b200", "host", "swarm_plan_cli"),
                               PKG + "/missions/mission_64agents_15.json", PKG + "/worlds/map%d.bt" % MAP, "/tmp", "stage=sfc"] + PARAMS,
                              text=True).splitlines()
i = [k for k, l in enumerate(out) if l.startswith("dump")][0]
N, M = map(int, out[i].split()[1:])
T = np.array(out[i + 1].split(), float)
i += 2
traj = np.zeros((N, M + 1, 3), np.float32)
sfc = []
for q in range(N):
    traj[q] = np.array(out[i].split(), np.float32).reshape(M + 1, 3)
    nb = int(out[i + 1])
    rows = np.array([out[i + 2 + b].split() for b in range(nb)], float)
    sfc.append((rows[:, :6], rows[:, 6]))
    i += 2 + nb
mj = json.load(open(PKG + "/missions/mission_64agents_15.json"))
start = np.zeros((N, 9)); goal = np.zeros((N, 9))
for q, a in enumerate(mj["agents"]):
    start[q, :3] = a["start"]; goal[q, :3] = a["goal"]
m = dict(N=N, M=M, T=T, start=start, goal=goal, radius=np.full(N, 0.15), sfc=sfc, init_traj=traj, downwash=2.0, seed=MAP)
synth.save_pack([m], [0.2], os.path.join(HERE, "smoke_map32.npz"))
print("wrote smoke_map32.npz: N=%d M=%d boxes=%d" % (N, M, sum(len(t) for _, t in sfc)))
