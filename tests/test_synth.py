"""Workload generator invariants (swarm_simulator_b200/synth.py follows rbp_corridor.hpp / ecbs_planner.hpp)."""
import numpy as np
import pytest

import oracle_util
from swarm_simulator_b200 import synth


@pytest.fixture(scope="module")
def mission():
    return synth.synth_mission(12, 5, 0.2, 4242)


def test_deterministic(mission):
    m2 = synth.synth_mission(12, 5, 0.2, 4242)
    for k in ("T", "start", "goal", "rsfc_n", "init_traj"):
        assert np.array_equal(mission[k], m2[k])
    assert all(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) for a, b in zip(mission["sfc"], m2["sfc"]))


def test_shapes_and_init_traj(mission):
    N, M = mission["N"], mission["M"]
    assert mission["init_traj"].shape == (N, M + 1, 3) and mission["init_traj"].dtype == np.float32
    assert np.array_equal(mission["T"], np.arange(M + 1.0))
    # ecbs_planner.hpp L49-L70: exact start first, exact goal last (twice: the parked state and the appended goal)
    assert np.array_equal(mission["init_traj"][:, 0], mission["start"][:, :3].astype(np.float32))
    assert np.array_equal(mission["init_traj"][:, -1], mission["goal"][:, :3].astype(np.float32))


def test_rsfc_normals_are_normalised_in_downwash_space(mission):
    """rbp_corridor.hpp L382-L384: normalise, then divide z by the downwash coefficient (float32)."""
    n = mission["rsfc_n"].astype(np.float64)
    nn = np.sqrt(n[..., 0] ** 2 + n[..., 1] ** 2 + (mission["downwash"] * n[..., 2]) ** 2)
    assert np.abs(nn - 1).max() < 5e-7


def test_initial_trajectory_is_feasible_for_every_row(mission):
    """The feasibility guarantee of the method: `dummy` built from initTraj satisfies every SFC and RSFC row."""
    op = oracle_util.oracle_problem(mission, sequential=True, batch_size=3)
    dummy = op.dummy()
    for l in range(4):
        qp = op.populate(dummy, l)
        G, h = qp.csr("g")
        oq = 6 * mission["M"]
        x0 = np.concatenate([dummy.reshape(mission["N"], oq, 3)[3 * l:3 * l + 3, :, k].reshape(-1) for k in range(3)])
        assert (G @ x0 - h).max() < 1e-6
        A, b = qp.csr("a")
        assert np.abs(A @ x0 - b).max() < 1e-9   # piecewise-constant control points are C2 with zero end derivatives


def test_sfc_boxes_contain_their_waypoints(mission):
    for qi, (boxes, tend) in enumerate(mission["sfc"]):
        assert tend[-1] == mission["T"][-1] and np.all(np.diff(tend) > 0)
        traj = mission["init_traj"][qi].astype(np.float64)
        bi = 0
        for m in range(mission["M"]):
            while tend[bi] < mission["T"][m + 1]:
                bi += 1
            for p in (traj[m], traj[m + 1]):
                assert np.all(p >= boxes[bi][:3] - 1e-6) and np.all(p <= boxes[bi][3:] + 1e-6)
