"""Feasibility verdicts of the interior-point solver against an INDEPENDENT classifier (HiGHS through
scipy.optimize.linprog, tests/feas_classifier.py) -- the oracle and the CUDA kernels share one algorithm and cannot check
each other here.  Reference behaviour: RBPPlanner::update() fails exactly when cplex.solve() does
(/root/reference/swarm_planner/include/rbp_planner.hpp L158-L161).

For every mission of the committed packs (tests/golden/missions_cfg*.npz, >= 256 seeds of configs 1-3, 32 of config 4):
  * status OK      -> the final control-point table must satisfy every row of the mission (independent numpy check);
  * status != OK   -> the failing batch QP is rebuilt and classified: a strictly feasible QP (LP slack > 1e-6) is a
                      failure; an LP-infeasible one must carry INFEASIBLE.
A second family inflates the radii after the corridor was built, which makes many QPs infeasible.
CPU: the oracle (this file, not gpu).  GPU: the engine through the C ABI (test_gpu_feasibility.py).
"""
import numpy as np
import pytest

import feas_classifier as fc
import feas_util as fu
import oracle
import oracle_util
from swarm_simulator_b200 import synth

CPU_CASES = [c for c in fu.CASES if not (c[0] == "cfg4" and c[2] == 32)]   # 256-agent joint batches of 32: GPU test only (70 s of CPU)


def _run_oracle(ms, sequential, bs):
    ps = [oracle_util.oracle_problem(m, sequential=sequential, batch_size=bs) for m in ms]
    _, ctrl, st = oracle.update_many(ps, nthreads=0)
    first_bad = np.full(len(ms), -1)
    for c in np.nonzero(st)[0]:
        r = ps[c].update()
        first_bad[c] = [k for k, s in enumerate(r["batch_status"]) if s not in (0, -1)][0]
    return st, first_bad, ctrl


@pytest.mark.parametrize("pack,sequential,bs,count", CPU_CASES)
def test_oracle_verdicts_match_highs(pack, sequential, bs, count):
    ms = fu.missions(pack, count)
    st, first_bad, ctrl = _run_oracle(ms, sequential, bs)
    fails, tally = fu.judge(st, first_bad, ms, sequential, bs, ctrl)
    assert not fails, (tally, fails[:5])
    # the committed packs are valid missions: every one plans (as the reference's smoke loop expects of its 50 maps)
    assert tally.get("ok", 0) == len(ms), tally


@pytest.mark.parametrize("pack,sequential,bs,count,radius", [
    ("cfg3", True, 1, 48, 0.25), ("cfg3", True, 4, 32, 0.25), ("cfg2", False, 16, 64, 0.30), ("cfg2", True, 1, 64, 0.30),
])
def test_oracle_reports_infeasible_when_highs_does(pack, sequential, bs, count, radius):
    ms = fu.missions(pack, count, inflate_radius=radius)
    st, first_bad, ctrl = _run_oracle(ms, sequential, bs)
    fails, tally = fu.judge(st, first_bad, ms, sequential, bs, ctrl)
    assert not fails, (tally, fails[:5])
    assert sum(v for k, v in tally.items() if k.startswith("infeasible")) >= 4, tally   # the family does produce infeasible QPs


@pytest.mark.parametrize("N,M,rho,seed0,count,sequential,bs", [
    (4, 3, 0.0, 51000, 48, False, 4), (4, 3, 0.0, 51000, 48, True, 1),
    (16, 5, 0.2, 41000, 12, False, 16), (16, 5, 0.2, 41000, 12, True, 4), (16, 5, 0.5, 61000, 12, True, 1),
])
def test_fresh_seeds_outside_the_packs(N, M, rho, seed0, count, sequential, bs):
    """Missions generated at test time from seeds no pack contains (the packs could have been tuned to): same judge.
    tools/robustness_sweep.py runs this at scale (9.6 k missions, no disagreement)."""
    ms = []
    for i in range(count):
        m = synth.synth_mission(N, M, rho, seed0 + i)
        m.pop("edt", None)
        ms.append(m)
    st, first_bad, ctrl = _run_oracle(ms, sequential, bs)
    fails, tally = fu.judge(st, first_bad, ms, sequential, bs, ctrl)
    assert not fails, (tally, fails[:5])
    assert tally.get("ok", 0) == len(ms), tally


def test_every_qp_of_sampled_missions_classified():
    """Every batch QP of a few missions through HiGHS (not only the failing ones): all are strictly feasible and all solve."""
    for pack, sequential, bs, picks in (("cfg3", True, 1, (29, 194)), ("cfg3", True, 4, (65,)), ("cfg2", False, 16, (0, 1, 2, 3))):
        all_ms = fu.missions(pack)
        for c in picks:
            m = all_ms[c]
            p = oracle_util.oracle_problem(m, sequential=sequential, batch_size=bs)
            r = p.update()
            nb = len(r["batch_status"])
            assert r["status"] == 0
            N, M = m["N"], m["M"]
            oq = 6 * M
            dummy = p.dummy() if sequential else np.zeros((N * oq, 3))
            _, ebs, _ = p.set_batch()
            for k in range(nb):
                q = p.populate(dummy, k)
                verdict, slack, _ = fc.classify(q)
                sol = q.solve()
                assert sol["status"] == 0 and verdict in (fc.STRICT, fc.BORDERLINE), (m["seed"], k, verdict, slack)
                eq, viol = fc.check_point(q, sol["x"])
                assert eq < 1e-8 and viol < 2e-6, (m["seed"], k, eq, viol)
                n_in = min(ebs, N - k * ebs)
                od = n_in * oq
                for kk in range(3):
                    for bi in range(n_in):
                        qa = k * ebs + bi
                        dummy[qa * oq:(qa + 1) * oq, kk] = sol["x"][kk * od + bi * oq: kk * od + (bi + 1) * oq]


def test_seeds_the_round_one_solver_gave_up_on():
    """Regression: seeds 3029, 3194, 3220 (b = 1) and 3065 (b = 4) ended NOT_CONVERGED / INFEASIBLE although HiGHS finds
    0.02-0.08 m of slack on every live row (3029, 3194, 3220) or the rows are consistent with zero-width boxes (3065, 3251)."""
    ms = fu.missions("cfg3")
    for seed, bs in ((3029, 1), (3194, 1), (3220, 1), (3251, 1), (3065, 4), (3029, 4), (3194, 4), (3220, 4)):
        m = ms[seed - 3000]
        assert m["seed"] == seed
        r = oracle_util.oracle_problem(m, sequential=True, batch_size=bs).update()
        assert r["status"] == 0, (seed, bs, r["batch_status"])
        veq, vbox, vrel = fu.joint_violation(m, r["ctrl"], True)
        assert veq < 1e-7 and vbox < 2e-6 and vrel < 2e-6


def test_smoke_loop_map_with_face_sharing_boxes():
    """worlds/map32.bt of the reference's smoke loop (fixture tests/golden/smoke_map32.npz, M = 36, launch default b = 4):
    two consecutive corridor boxes of one agent share only a face, so batch 14 has no interior (HiGHS: slack exactly 0).
    CPLEX fixes such a variable in its presolve and solves; so must we (presolve rule cp_bounds / presolve_dead_rows)."""
    import os
    from swarm_simulator_b200 import synth
    m = synth.load_pack(os.path.join(fu.GOLDEN, "smoke_map32.npz"))[0]
    assert (m["N"], m["M"]) == (64, 36)
    _, q = fu.failing_qp(m, True, 4, 14)
    verdict, slack, _ = fc.classify(q)
    assert verdict == fc.BORDERLINE and abs(slack) < 1e-9
    r = oracle_util.oracle_problem(m, sequential=True, batch_size=4).update()
    assert r["status"] == 0, r["batch_status"]
    veq, vbox, vrel = fu.joint_violation(m, r["ctrl"], True)
    assert veq < 1e-7 and vbox < 2e-6 and vrel < 2e-6
