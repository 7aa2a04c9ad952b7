"""Randomised GPU-vs-oracle parity over mission shapes and planner parameters (hypothesis; -m gpu)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle_util
from swarm_simulator_b200 import engine as E, synth

pytestmark = pytest.mark.gpu
_ENG = {}


def _engine():
    if "e" not in _ENG:
        import __graft_entry__ as G
        G.build()
        _ENG["e"] = E.Engine()
    return _ENG["e"]


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(N=st.integers(2, 9), M=st.integers(3, 7), rho=st.sampled_from([0.0, 0.2, 0.4]), seed=st.integers(0, 10 ** 6),
       bs=st.integers(1, 5), sequential=st.booleans(), iteration=st.integers(1, 2))
def test_random_missions_match_oracle(N, M, rho, seed, bs, sequential, iteration):
    m = synth.synth_mission(N, M, rho, seed)
    prob = E.PackedProblem(synth.pack([m]), sequential=sequential, batch_size=bs, iteration=iteration)
    r = _engine().solve_many(prob)
    ro = oracle_util.oracle_problem(m, sequential=sequential, batch_size=bs, iteration=iteration).update()
    assert r.status[0] == ro["status"]
    if ro["status"] != 0:
        return
    n = r.nrec
    assert np.array_equal(r.qp_iters[0][:n], ro["batch_iters"][:n])
    assert np.abs(r.ctrl[0] - ro["ctrl"]).max() < 1e-8
    assert np.abs(r.coef[0] - ro["coef"]).max() < 1e-8 * max(1.0, np.abs(ro["coef"]).max())
