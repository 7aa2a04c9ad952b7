"""Tuning sweep (not a test): PDIP CTA size x shared-memory budget x missions per call, 64-agent b=1 workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from swarm_simulator_b200 import engine as E, synth

def main():
    N, bs = int(os.environ.get("TUNE_N", 64)), int(os.environ.get("TUNE_BS", 1))
    ms = [synth.synth_mission(N, 5, 0.2, 3000 + i) for i in range(8)]
    ref = None
    for threads in [int(t) for t in os.environ.get("TUNE_THREADS", "64,128,256").split(",")]:
        for smem in [int(t) for t in os.environ.get("TUNE_SMEM", "12,48,100").split(",")]:
            eng = E.Engine(device=0, smem_budget=smem * 1024, threads=threads)
            for count in [int(t) for t in os.environ.get("TUNE_COUNT", "1184").split(",")]:
                prob = E.PackedProblem(synth.pack([ms[i % 8] for i in range(count)]), sequential=True, batch_size=bs)
                eng.upload(prob)
                for rep in range(2):
                    eng.timer_start(); eng.run(); ms_ = eng.timer_stop()
                r = eng.download(prob)
                if ref is None: ref = r.ctrl[:8].copy()
                print("threads=%3d smem=%3dK count=%5d: rc=%d kernel %.1f ms -> %.0f agent-QPs/s  (same result: %s, iters %.2f)" % (
                    threads, smem, count, r.rc, ms_, count * N / ms_ * 1e3, bool(np.array_equal(ref, r.ctrl[:8])), r.qp_iters.mean()), flush=True)
            eng.close()

if __name__ == "__main__":
    main()
