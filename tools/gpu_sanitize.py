"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of librbpe.so once, small shapes.
usage (GPU box): compute-sanitizer --tool racecheck python tools/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from swarm_simulator_b200 import engine as E, synth
# one-agent batches: the warp-per-QP kernel (RBPE_LAT=0) and the several-warps-per-QP latency kernel (the default for a
# handful of missions), Gauss-Seidel chain and Jacobi sweep each
for lat in ("0", "1"):
    os.environ["RBPE_LAT"] = lat
    e1 = E.Engine(device=0)
    for N, M, seed in ((8, 5, 7), (7, 6, 9)):
        m = synth.synth_mission(N, M, 0.2, seed)
        prob = E.PackedProblem(synth.pack([m, m]), sequential=True, batch_size=1)
        r = e1.solve_many(prob)
        assert r.rc == 0, (lat, N, M, r.rc)
        e1.upload(prob); e1.run(E.MODE_JACOBI); rj = e1.download(prob)
        assert rj.rc == 0
    e1.close()
os.environ.pop("RBPE_LAT", None)
eng = E.Engine(device=0)
for N, M, rho, seq, bs, seed in ((4, 3, 0.0, False, 4, 7), (8, 5, 0.2, True, 1, 7), (8, 4, 0.3, True, 3, 9), (12, 5, 0.2, False, 12, 11)):
    m = synth.synth_mission(N, M, rho, seed)
    prob = E.PackedProblem(synth.pack([m, m]), sequential=seq, batch_size=bs)
    for env in (("RBPE_TMA", "1"), ("RBPE_TMA", "0")) if bs > 1 else ((None, None),):
        if env[0]:
            os.environ[env[0]] = env[1]
        r = eng.solve_many(prob)
        assert r.rc == 0, (N, M, bs, r.rc)
    os.environ.pop("RBPE_TMA", None)
    if seq and bs == 1:
        eng.upload(prob); eng.run(E.MODE_JACOBI); rj = eng.download(prob)
        assert rj.rc == 0
    n, t, col = eng.corridor_rsfc(m["init_traj"][None], m["T"][None], 2.0)
    a, b, l = eng.safety_metrics(r.coef[:1], m["T"][None], m["radius"][None])
print("sanitize workload ok")
eng.close()
