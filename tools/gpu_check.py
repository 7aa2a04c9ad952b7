"""Exploratory GPU check (not a test): engine vs oracle on a few shapes + first timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from swarm_simulator_b200 import engine as E, synth
import oracle, oracle_util

def main():
    eng = E.Engine(device=0)
    for (N, M, rho, seq, bs, seed) in ((4, 3, 0.0, False, 4, 1001), (6, 5, 0.2, True, 1, 2000), (6, 4, 0.2, True, 4, 2000),
                                       (16, 5, 0.2, True, 1, 3016), (16, 5, 0.2, True, 4, 3016), (16, 5, 0.2, False, 16, 3016),
                                       (64, 5, 0.2, True, 1, 3064), (64, 5, 0.2, True, 4, 3064)):
        m = synth.synth_mission(N, M, rho, seed)
        prob = E.PackedProblem(synth.pack([m]), sequential=seq, batch_size=bs)
        t = time.time(); r = eng.solve_many(prob); dt = time.time() - t
        t = time.time(); ro = oracle_util.oracle_problem(m, sequential=seq, batch_size=bs).update(); dto = time.time() - t
        print("N=%d M=%d seq=%d bs=%d: rc=%d oracle=%d iters=%s same_iters=%s | err ctrl %.2e coef %.2e | gpu %.1f ms oracle %.1f ms timing=%s" % (
            N, M, seq, bs, r.rc, ro["status"], r.qp_iters[0][:6], bool((r.qp_iters[0] == ro["batch_iters"]).all()), np.abs(r.ctrl[0] - ro["ctrl"]).max(),
            np.abs(r.coef[0] - ro["coef"]).max(), dt * 1e3, dto * 1e3, eng.timing()), flush=True)
    eng.close()
    # throughput probe: 64-agent missions, b=1, sweep mission count and smem budget
    ms = [synth.synth_mission(64, 5, 0.2, 3064 + i) for i in range(8)]
    for smem in (100 * 1024, 48 * 1024, 200 * 1024):
        eng = E.Engine(device=0, smem_budget=smem)
        for count in (148, 296, 592):
            prob = E.PackedProblem(synth.pack([ms[i % 8] for i in range(count)]), sequential=True, batch_size=1)
            for rep in range(2):
                t = time.time(); r = eng.solve_many(prob); dt = time.time() - t
            tm = eng.timing()
            print("smem=%dK count=%d N=64 b=1: rc=%d e2e %.1f ms -> %.0f agent-QPs/s ; kernel %.1f ms -> %.0f /s; iters mean %.2f" % (
                smem // 1024, count, r.rc, dt * 1e3, count * 64 / dt, tm["solve_ms"], count * 64 / tm["solve_ms"] * 1e3, r.qp_iters.mean()), flush=True)
        eng.close()
    # CPU oracle throughput on the same missions (all cores)
    ops = [oracle_util.oracle_problem(m, sequential=True, batch_size=1) for m in ms]
    for nt in (1, os.cpu_count()):
        t = time.time(); oracle.update_many(ops * 4, nthreads=nt); dt = time.time() - t
        print("oracle threads=%d: %.0f agent-QPs/s" % (nt, len(ops) * 4 * 64 / dt), flush=True)

if __name__ == "__main__":
    main()
