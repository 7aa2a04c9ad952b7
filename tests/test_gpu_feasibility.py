"""GPU side of tests/test_feasibility_classifier.py (-m gpu): the CUDA engine, through the C ABI, must give the verdicts
an INDEPENDENT classifier (HiGHS) gives -- OK on every strictly feasible batch QP, INFEASIBLE on every LP-infeasible one --
over the committed mission packs: 256 seeds each of BASELINE configs 1-3 at their joint batch, b = 4 and b = 1, and 32
seeds of config 4 (256 agents, rho = 0.4) at b = 1, 4 and 32.  Reference: update() fails exactly when cplex.solve() does
(/root/reference/swarm_planner/include/rbp_planner.hpp L158-L161).
"""
import numpy as np
import pytest

import feas_util as fu
import oracle
import oracle_util
from swarm_simulator_b200 import engine as E, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as G
    G.build()
    e = E.Engine(device=0)
    yield e
    e.close()


def _run_engine(eng, ms, sequential, bs):
    prob = E.PackedProblem(synth.pack(ms), sequential=sequential, batch_size=bs)
    r = eng.solve_many(prob)
    first_bad = np.full(len(ms), -1)
    for c in np.nonzero(r.status)[0]:
        first_bad[c] = [k for k in range(r.nrec) if r.qp_status[c][k] != 0][0]
    return r, first_bad


@pytest.mark.parametrize("pack,sequential,bs,count", fu.CASES)
def test_engine_verdicts_match_highs(eng, pack, sequential, bs, count):
    ms = fu.missions(pack, count)
    r, first_bad = _run_engine(eng, ms, sequential, bs)
    fails, tally = fu.judge(r.status, first_bad, ms, sequential, bs, r.ctrl)
    assert not fails, (tally, fails[:5])
    assert tally.get("ok", 0) == len(ms), tally          # every committed mission plans


@pytest.mark.parametrize("pack,sequential,bs,count,radius", [
    ("cfg3", True, 1, 128, 0.25), ("cfg3", True, 4, 64, 0.25), ("cfg2", False, 16, 128, 0.30), ("cfg2", True, 1, 128, 0.30),
])
def test_engine_reports_infeasible_when_highs_does(eng, pack, sequential, bs, count, radius):
    ms = fu.missions(pack, count, inflate_radius=radius)
    r, first_bad = _run_engine(eng, ms, sequential, bs)
    fails, tally = fu.judge(r.status, first_bad, ms, sequential, bs, r.ctrl)
    assert not fails, (tally, fails[:5])
    assert sum(v for k, v in tally.items() if k.startswith("infeasible")) >= 8, tally
    # and the oracle takes the same decisions (status code and failing batch), mission by mission
    ps = [oracle_util.oracle_problem(m, sequential=sequential, batch_size=bs) for m in ms]
    _, _, st = oracle.update_many(ps, nthreads=0)
    assert np.array_equal(st, r.status), (np.flatnonzero(st != r.status), st[st != r.status], r.status[st != r.status])


def test_engine_solves_the_round_one_failures(eng):
    ms_all = fu.missions("cfg3")
    for bs, seeds in ((1, (3029, 3194, 3220, 3251)), (4, (3065, 3029, 3194, 3220))):
        ms = [ms_all[s - 3000] for s in seeds]
        r, _ = _run_engine(eng, ms, True, bs)
        assert not r.status.any(), (bs, r.status)
        for c, m in enumerate(ms):
            ro = oracle_util.oracle_problem(m, sequential=True, batch_size=bs).update()
            assert ro["status"] == 0
            assert np.abs(r.ctrl[c] - ro["ctrl"]).max() < 1e-6


def test_engine_plans_the_smoke_loop_map_with_face_sharing_boxes(eng):
    """GPU side of test_smoke_loop_map_with_face_sharing_boxes (reference: swarm_traj_planner_rbp_test_all.cpp L93-L94)."""
    import os
    m = synth.load_pack(os.path.join(fu.GOLDEN, "smoke_map32.npz"))[0]
    for bs in (4, 1):
        r, _ = _run_engine(eng, [m], True, bs)
        assert r.status[0] == 0, (bs, r.qp_status[0])
        ro = oracle_util.oracle_problem(m, sequential=True, batch_size=bs).update()
        assert ro["status"] == 0
        assert np.abs(r.ctrl[0] - ro["ctrl"]).max() < 1e-6
        veq, vbox, vrel = fu.joint_violation(m, r.ctrl[0], True)
        assert veq < 1e-7 and vbox < 2e-6 and vrel < 2e-6
