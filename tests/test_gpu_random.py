"""Randomised GPU-vs-oracle parity over mission shapes and planner parameters (hypothesis; -m gpu).

Tolerance.  The fixed-seed parity tests (test_gpu_parity.py) hold kernel and oracle to 1e-8 on control points with
identical iteration counts.  Over random shapes hypothesis also finds near-degenerate QPs (17+ interior-point
iterations) where two converged solves legitimately differ at the level tolerance x conditioning: the kernels apply
explicit inverses of the diagonal factor blocks (9 x 9, 32 x 32) where the oracle substitutes, so late Newton
directions differ in the last digits.  Here: control points within 1e-6 absolute (the north-star asks for 1e-4
relative on coefficients), iteration counts within one.  Examples are derandomised so that the test is reproducible."""
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


@settings(max_examples=16, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
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
    assert np.abs(r.qp_iters[0][:n].astype(int) - ro["batch_iters"][:n].astype(int)).max() <= 1
    assert np.abs(r.ctrl[0] - ro["ctrl"]).max() < 1e-6
    assert np.abs(r.coef[0] - ro["coef"]).max() < 1e-6 * max(1.0, np.abs(ro["coef"]).max())


@pytest.mark.parametrize("N,M,rho,seed,bs,sequential", [(5, 7, 0.4, 210, 1, False), (5, 3, 0.4, 210, 1, False)])
def test_near_degenerate_joint_batch_found_by_hypothesis(N, M, rho, seed, bs, sequential):
    """Regression: the hardest joint batches hypothesis has found so far.  The M=7 one has weakly active rows (multiplier
    ~ slack ~ sqrt(mu)): an interior-point iterate at complementarity mu = 1e-10 sits ~1e-5 off such a row, so two runs
    that stop one or two iterations apart (19 vs 17 here: rounding near the stopping threshold) differ by ~3e-5 in a few
    control points although both meet the same KKT tolerances (there the kernel's objective is the lower one).
    Bar here: the north-star's 1e-4, objective agreement to the gap tolerance, and tiny true residuals."""
    m = synth.synth_mission(N, M, rho, seed)
    prob = E.PackedProblem(synth.pack([m]), sequential=sequential, batch_size=bs)
    r = _engine().solve_many(prob)
    ro = oracle_util.oracle_problem(m, sequential=sequential, batch_size=bs).update()
    assert r.status[0] == ro["status"] == 0
    assert np.abs(r.qp_iters[0][:r.nrec].astype(int) - ro["batch_iters"][:r.nrec].astype(int)).max() <= 2
    assert np.abs(r.ctrl[0] - ro["ctrl"]).max() < 1e-4
    assert abs(r.qp_obj[0][0] - ro["batch_obj"][0]) < 1e-7 * max(1.0, abs(ro["batch_obj"][0]))
    gap, rp, rd, rg = r.qp_res[0][0]          # recorded at exit: complementarity, |Ax - b|, dual and inequality residuals
    assert gap < 1e-9 and rp < 1e-9
