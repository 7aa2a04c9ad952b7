"""Out-of-pack robustness sweep (tool, CPU only): fresh seeded missions, the oracle's verdicts judged by HiGHS exactly as
tests/test_feasibility_classifier.py does for the committed packs.
usage: python tools/robustness_sweep.py <label> N M rho first_seed count <batching>...   (batching: s1, s4, s16 = sequential with
that batch size; j16 = one joint batch).  Round-2 runs are listed in DESIGN.md section 2."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from multiprocessing import Pool
from swarm_simulator_b200 import synth
import oracle, oracle_util, feas_util as fu

def job(a):
    N, M, rho, seed = a
    m = synth.synth_mission(N, M, rho, seed); m.pop('edt', None); return m

def run(ms, sequential, bs):
    ps = [oracle_util.oracle_problem(m, sequential=sequential, batch_size=bs) for m in ms]
    _, ctrl, st = oracle.update_many(ps, nthreads=0)
    first_bad = np.full(len(ms), -1)
    for c in np.nonzero(st)[0]:
        r = ps[c].update()
        first_bad[c] = [k for k, s in enumerate(r["batch_status"]) if s not in (0, -1)][0]
    return st, first_bad, ctrl

if __name__ == '__main__':
    name, N, M, rho, s0, cnt = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    batchings = [(x[0]=='s', int(x[1:])) for x in sys.argv[7:]]
    t=time.time()
    with Pool(8) as pool:
        ms = pool.map(job, [(N, M, rho, s0+i) for i in range(cnt)], chunksize=1)
    print(name, 'generated', len(ms), 'in %.0fs'%(time.time()-t), flush=True)
    for seq, bs in batchings:
        t=time.time()
        st, fb, ctrl = run(ms, seq, bs)
        fails, tally = fu.judge(st, fb, ms, seq, bs, ctrl)
        print(name, 'seq' if seq else 'joint', bs, 'tally', tally, 'fails', fails[:10], 'bad seeds', [ms[c]['seed'] for c in np.nonzero(st)[0]][:20], '%.0fs'%(time.time()-t), flush=True)
