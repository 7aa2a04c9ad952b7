"""Exploratory GPU probe (not a test): joint-batch (b > 1) workloads of BASELINE configs 1, 2, 4 and the launch default b=4:
parity against the oracle on one mission, then kernel time over `count` tiled missions."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from swarm_simulator_b200 import engine as E, synth
import oracle_util

CASES = ((4, 3, 0.0, False, 4, 1001, 1184), (16, 5, 0.2, False, 16, 2001, 592), (64, 5, 0.2, True, 4, 3001, 296),
         (64, 5, 0.2, True, 16, 3001, 148), (256, 5, 0.4, True, 32, 4001, 8))


def prof_read(reset=True):
    import ctypes as C
    lib = E.load_library()
    if not hasattr(lib, "rbpe_prof_read"):
        return None
    buf = (C.c_ulonglong * 16)()
    lib.rbpe_prof_read(buf, 1 if reset else 0)
    return [int(v) for v in buf]


def main():
    only = os.environ.get("JOINT_CASES")
    if os.environ.get("JOINT_LIB"):       # e.g. a -DRBPE_PROFILE build: swarm_simulator_b200/librbpe_prof.so
        E.LIB_PATH = os.path.join(ROOT, os.environ["JOINT_LIB"])
    eng = E.Engine(device=0)
    for ci, (N, M, rho, seq, bs, seed, count) in enumerate(CASES):
        if only and str(ci) not in only.split(","):
            continue
        pack = {4: "cfg1", 16: "cfg2", 64: "cfg3", 256: "cfg4"}[N]
        m = synth.load_pack(os.path.join(ROOT, "tests", "golden", "missions_%s.npz" % pack), select=[seed % 1000])[0]
        prob1 = E.PackedProblem(synth.pack([m]), sequential=seq, batch_size=bs)
        t = time.time(); r = eng.solve_many(prob1); dt1 = time.time() - t
        err = float("nan"); same = None; dto = 0.0
        if N <= 64:
            t = time.time(); ro = oracle_util.oracle_problem(m, sequential=seq, batch_size=bs).update(); dto = time.time() - t
            err = float(np.abs(r.ctrl[0] - ro["ctrl"]).max()); same = bool((r.qp_iters[0] == ro["batch_iters"]).all())
        prob = E.PackedProblem(synth.pack([m] * count), sequential=seq, batch_size=bs)
        eng.upload(prob)
        best = 1e9
        eng.run(); eng.sync()            # warm (clocks, scratch)
        prof_read()
        for rep in range(2):
            eng.timer_start(); eng.run(); best = min(best, eng.timer_stop())
        pr = prof_read()
        rr = eng.download(prob)
        if pr:
            tot = float(sum(pr[:5])) or 1.0
            print("   phases (thread-0 clocks summed over CTAs, steady state): setup %.1f%% rows %.1f%% factor %.1f%% solves %.1f%% other %.1f%%" % (
                100 * pr[0] / tot, 100 * pr[1] / tot, 100 * pr[2] / tot, 100 * pr[3] / tot, 100 * pr[4] / tot))
            print("   inside factor: build_W %.1f%% | update %.1f%% chol32 %.1f%% trsm %.1f%% (of total)" % tuple(
                100 * pr[i] / tot for i in (5, 6, 7, 8)))
            print("   chol32 pass-1 cycles per 32-step call: min %d max %d mean-all-phases %.0f" % (pr[13], pr[14], (pr[10] + pr[11] + pr[12]) / max(1, pr[15])))
        print("N=%d M=%d b=%d: rc=%d/%d iters=%s same_iters=%s err=%.2e | one mission %.1f ms (oracle %.1f ms) | %d missions: kernel %.1f ms -> %.0f agent-QPs/s" % (
            N, M, bs, r.rc, rr.rc, r.qp_iters[0][:8], same, err, dt1 * 1e3, dto * 1e3, count, best, count * N / best * 1e3), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
