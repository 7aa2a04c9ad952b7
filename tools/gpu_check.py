"""Exploratory GPU check (not a test): engine vs oracle on a few shapes + first timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from swarm_simulator_b200 import engine as E, synth
import oracle_util

def main():
    eng = E.Engine(device=0)
    for (N, M, rho, seq, bs, seed) in ((4, 3, 0.0, False, 4, 1001), (6, 5, 0.2, True, 1, 2000), (6, 4, 0.2, True, 4, 2000), (16, 5, 0.2, True, 1, 3016), (16, 5, 0.2, True, 4, 3016), (16, 5, 0.2, False, 16, 3016)):
        m = synth.synth_mission(N, M, rho, seed)
        prob = E.PackedProblem(synth.pack([m]), sequential=seq, batch_size=bs)
        t = time.time(); r = eng.solve_many(prob); dt = time.time() - t
        ro = oracle_util.oracle_problem(m, sequential=seq, batch_size=bs).update()
        print("N=%d M=%d seq=%d bs=%d: rc=%d oracle=%d iters=%s | err ctrl %.2e coef %.2e | %.1f ms  timing=%s" % (
            N, M, seq, bs, r.rc, ro["status"], r.qp_iters[0][:6], np.abs(r.ctrl[0] - ro["ctrl"]).max(),
            np.abs(r.coef[0] - ro["coef"]).max(), dt * 1e3, eng.timing()), flush=True)
    # throughput probe: many copies of a 16-agent mission, b=1
    ms = [synth.synth_mission(16, 5, 0.2, 3016 + i) for i in range(4)]
    for count in (148, 592):
        prob = E.PackedProblem(synth.pack([ms[i % 4] for i in range(count)]), sequential=True, batch_size=1)
        for rep in range(2):
            t = time.time(); r = eng.solve_many(prob); dt = time.time() - t
        print("count=%d N=16 b=1: rc=%d %.1f ms -> %.0f agent-QPs/s  timing=%s" % (count, r.rc, dt * 1e3, count * 16 / dt, eng.timing()), flush=True)
    eng.close()

if __name__ == "__main__":
    main()
