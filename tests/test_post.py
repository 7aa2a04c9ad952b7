"""The two pure-arithmetic neighbours of the path: RSFC generation (Corridor::updateRelBox) and the
post-hoc collision / length checks of RBPPublisher.  CPU: oracle vs golden values.  GPU: kernels vs oracle, bit for bit."""
import numpy as np
import pytest

import fixture_lp as F
import oracle
from swarm_simulator_b200 import synth


def _fixture_coef(golden):
    low = golden["csv"]["coef"]                                   # [N, M, 3, 6] lowest power first (CSV order)
    return np.ascontiguousarray(low[..., ::-1].transpose(0, 2, 1, 3).reshape(F.N, 3, 6 * F.M))   # [N][3][6M] highest first


def test_oracle_metrics_reproduce_reference_run(golden):
    """safety margin ratio 1.00189 at t = 23.4 s and 809.807 m of flight for the committed 64-agent run."""
    ratio, t_min, length, nt = oracle.safety_metrics(_fixture_coef(golden), np.arange(F.M + 1.0), golden["mission"]["radius"])
    assert nt == 360
    assert abs(ratio - 1.00189) < 5e-6 and abs(t_min - 23.4) < 1e-9 and abs(length - 809.807) < 1e-3
    assert ratio >= 1.0                                           # the collision check of the reference (eyeballed there)


def test_oracle_rsfc_matches_generator_and_invariants():
    m = synth.synth_mission(12, 5, 0.2, 4242)
    n, t, collided = oracle.rsfc(m["init_traj"], m["T"], m["downwash"])
    assert not collided
    assert np.array_equal(n, m["rsfc_n"]) and np.array_equal(t, m["rsfc_t"])     # two independent float32 restatements
    # separating-plane property: n . (p_j - p_i) >= |closest point| > r_i + r_j at both ends of every segment
    qi, qj = np.triu_indices(12, 1)
    rel = (m["init_traj"][qj] - m["init_traj"][qi]).astype(np.float64)           # [P, M+1, 3]
    for end in (rel[:, :-1], rel[:, 1:]):
        assert (np.einsum("pmk,pmk->pm", n.astype(np.float64), end) > 0.3 - 1e-6).all()
    # two agents on top of each other -> "initial trajectories are collided" (rbp_corridor.hpp L385-L388)
    tr = m["init_traj"].copy()
    tr[1] = tr[0]
    assert oracle.rsfc(tr, m["T"], 2.0)[2]


@pytest.mark.gpu
def test_gpu_rsfc_bit_exact():
    from swarm_simulator_b200 import engine as E
    eng = E.Engine()
    ms = [synth.synth_mission(64, 5, 0.2, 3000), synth.synth_mission(64, 5, 0.2, 3001)]
    tr = np.stack([m["init_traj"] for m in ms]); T = np.stack([m["T"] for m in ms])
    n, t, col = eng.corridor_rsfc(tr, T, 2.0)
    for c, m in enumerate(ms):
        no, to, co = oracle.rsfc(m["init_traj"], m["T"], 2.0)
        assert np.array_equal(n[c].view(np.uint32), no.view(np.uint32))          # bit for bit, float32
        assert np.array_equal(t[c], to) and col[c] == int(co) == 0
    tr2 = tr.copy(); tr2[1, 5] = tr2[1, 4]
    assert list(eng.corridor_rsfc(tr2, T, 2.0)[2]) == [0, 1]
    # odd shapes: one agent (no pairs), M = 3
    m1 = synth.synth_mission(1, 3, 0.0, 1)
    n1, t1, c1 = eng.corridor_rsfc(m1["init_traj"][None], m1["T"][None], 2.0)
    assert n1.shape == (1, 0, 3, 3) and c1[0] == 0


@pytest.mark.gpu
def test_gpu_safety_metrics_bit_exact(golden):
    from swarm_simulator_b200 import engine as E
    eng = E.Engine()
    coef = _fixture_coef(golden)
    T = np.arange(F.M + 1.0)
    r, tm, ln = eng.safety_metrics(coef[None], T[None], golden["mission"]["radius"][None])
    ro = oracle.safety_metrics(coef, T, golden["mission"]["radius"])
    assert r[0] == ro[0] and tm[0] == ro[1] and ln[0] == ro[2]                   # identical doubles
    assert (r[0] >= 1.0) == (ro[0] >= 1.0)
    # solved synthetic missions of different durations in one call, after a host-side time scaling of one of them
    ms = [synth.synth_mission(16, 5, 0.2, 2000 + i) for i in range(3)]
    res = eng.solve_many(E.PackedProblem(synth.pack(ms), sequential=True, batch_size=4))
    Ts = np.stack([m["T"] for m in ms]); Ts[2] *= 1.21
    cf = res.coef.copy()
    cf[2] = (cf[2].reshape(16, 3, 5, 6) * (1 / 1.21) ** np.arange(5, -1, -1)).reshape(16, 3, 30)
    rad = np.stack([m["radius"] for m in ms])
    r, tm, ln = eng.safety_metrics(cf, Ts, rad)
    for c in range(3):
        ro = oracle.safety_metrics(cf[c], Ts[c], rad[c])
        assert (r[c], tm[c], ln[c]) == ro[:3]
        assert r[c] >= 1.0                                                       # RSFC guarantees collision-free plans
