"""Pins the CPU oracle to the reference's only frozen CPLEX artefacts:
log/QPmodel.lp (the QP of batch 15) and log/coef*.csv (CPLEX's answers), via tests/golden/*.npz.
"""
import numpy as np
import pytest

import fixture_lp as F
import oracle


def _canon(ptr, idx, val):
    rows = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr))
    o = np.lexsort((idx, rows))
    return idx[o], val[o]


@pytest.fixture(scope="module")
def fixture_problem(golden):
    rec = F.recover_inputs(golden["lp"], golden["csv"], golden["mission"])
    ms = golden["mission"]
    T = np.arange(F.M + 1, dtype=float)
    offs, boxes, tend = F.sfc_from_seg_box(rec["seg_box"], T)
    P = F.N * (F.N - 1) // 2
    prob = oracle.Problem(T, ms["start"], ms["goal"], ms["radius"], offs, boxes, tend, rec["rsfc_n"],
                          np.tile(T[1:], (P, 1)), np.zeros((F.N, F.M + 1, 3), np.float32),
                          sequential=True, batch_size=4, batch_iter=-1, iteration=1)
    return prob, rec


def test_lp_counts(golden):
    """Row/variable counts equal the formulas of populatebyrow (rbp_planner.hpp L551-L688)."""
    lp = golden["lp"]
    sense = lp["sense"]
    assert (sense == 0).sum() == 3 * F.NB * (3 * F.M + 3) == 1332
    assert len(sense) - 1332 == 2 * F.NV + (6 + 4 * 60) * 6 * F.M == 5184 + 53136


def test_cost_matrix_matches_lp(golden):
    """LP objective ('[...]/2' => Hessian 2Q) equals Q_base at dt=1 for every (axis, agent, segment)."""
    lp = golden["lp"]
    Q, _ = oracle.Q_base_and_basis()
    H = np.zeros((F.NV, F.NV))
    for i, j, v in zip(lp["obj_i"], lp["obj_j"], lp["obj_v"]):
        if i == j:
            H[i, i] += v
        else:
            H[i, j] += v / 2
            H[j, i] += v / 2
    for blk in range(F.NV // 6):
        s = slice(6 * blk, 6 * blk + 6)
        assert np.array_equal(H[s, s], 2 * Q)
    H2 = H.copy()
    for blk in range(F.NV // 6):
        s = slice(6 * blk, 6 * blk + 6)
        H2[s, s] = 0
    assert not H2.any()


def test_set_batch(fixture_problem):
    prob, _ = fixture_problem
    assert prob.set_batch() == (16, 4, 16)


def test_assembly_reproduces_lp_rows(golden, fixture_problem):
    """populatebyrow restatement vs every row of QPmodel.lp: structure exact, box rhs exact,
    RSFC rhs (r_i+r_j -/+ n.dummy) to the CSVs' 6-digit precision (dummy of agents 0..59 comes from them)."""
    prob, rec = fixture_problem
    qp = prob.populate(rec["dummy"], 15)
    assert (qp.nv, qp.ne, qp.mi, qp.n_box_rows, qp.n_rsfc_rows) == (2592, 1332, 58320, 5184, 53136)
    d = F.lp_as_leq(golden["lp"])
    ap, ai, av, b = qp.csr_raw("a")
    assert np.array_equal(ap, d["a_ptr"])
    i1, v1 = _canon(ap, ai, av)
    i2, v2 = _canon(d["a_ptr"], d["a_idx"], d["a_val"])
    assert np.array_equal(i1, i2) and np.array_equal(v1, v2)
    assert np.array_equal(b, d["b"])
    gp, gi, gv, h = qp.csr_raw("g")
    assert np.array_equal(gp, d["g_ptr"])
    i1, v1 = _canon(gp, gi, gv)
    i2, v2 = _canon(d["g_ptr"], d["g_idx"], d["g_val"])
    assert np.array_equal(i1, i2)
    assert np.abs(v1 - v2).max() < 2e-15  # LP prints 15 digits of float32 normals
    nb = qp.n_box_rows
    assert np.array_equal(h[:nb], d["h"][:nb])
    assert np.abs(h[nb:] - d["h"][nb:]).max() < 1e-5
    both = np.flatnonzero(np.diff(gp)[nb:] > 3)  # rows with both agents in the batch: rhs = -(r_i+r_j)
    assert np.allclose(h[nb:][both], -0.3, rtol=0, atol=1e-15)


def test_solver_known_answer(golden):
    """Mehrotra PDIP on the LP's QP vs CPLEX's own answer (coef61..64.csv).

    CPLEX's barrier stops at a 1e-8 relative gap, which leaves ~1e-4 of play along the flat directions
    of this 36-segment jerk objective; the bar is the north-star's 1e-4 (per-segment vector-relative)."""
    qp = oracle.QP(**F.lp_qp_arrays(golden["lp"]))
    r = qp.solve()
    assert r["status"] == oracle.OK
    assert abs(r["obj"] - 0.0971578) < 2e-7  # CPLEX-convention objective x'Qx
    ctrl = np.transpose(r["x"].reshape(3, F.NB, F.M, 6), (1, 2, 0, 3))  # [NB, M, 3, 6]
    coef = F.ctrl_coef_low(ctrl)
    ref = golden["csv"]["coef"][F.B0:F.B0 + F.NB]
    num = np.linalg.norm((coef - ref).reshape(F.NB, F.M, -1), axis=-1)
    den = np.linalg.norm(ref.reshape(F.NB, F.M, -1), axis=-1)
    assert (num / den).max() < 1e-4
    cref = F.csv_ctrl(ref)
    assert np.abs(ctrl - cref).max() < 1e-4 * max(1.0, np.abs(cref).max())
    # active set size reported in BASELINE.md: 34 of 58320 rows
    d = F.lp_as_leq(golden["lp"])
    rows = np.repeat(np.arange(len(d["h"])), np.diff(d["g_ptr"]))
    gx = np.zeros(len(d["h"]))
    np.add.at(gx, rows, d["g_val"] * r["x"][d["g_idx"]])
    slack = d["h"] - gx
    assert slack.min() > -1e-8
    assert 25 <= (slack < 1e-6).sum() <= 45


def test_assembled_solve_matches_csv(golden, fixture_problem):
    prob, rec = fixture_problem
    r = prob.populate(rec["dummy"], 15).solve()
    assert r["status"] == oracle.OK
    ctrl = np.transpose(r["x"].reshape(3, F.NB, F.M, 6), (1, 2, 0, 3))
    cref = F.csv_ctrl(golden["csv"]["coef"][F.B0:F.B0 + F.NB])
    assert np.abs(ctrl - cref).max() < 1e-4 * max(1.0, np.abs(cref).max())


def test_csv_invariants(golden):
    """Facts the CSVs pin on their own: exact starts, goals, C0..C2 continuity."""
    coef = golden["csv"]["coef"]  # [N, M, 3, 6] lowest power first, dt = 1
    ms = golden["mission"]
    assert np.all(golden["csv"]["duration"] == 1)
    assert np.abs(coef[:, 0, :, 0] - ms["start"][:, :3]).max() < 1e-6
    end = coef.sum(-1)  # p(1)
    assert np.abs(end[:, -1] - ms["goal"][:, :3]).max() < 1e-5
    pw = np.arange(6)
    for d in range(3):
        fac = np.array([np.prod([p - t for t in range(d)]) if p >= d else 0 for p in pw], float)
        at1 = (coef * fac).sum(-1)[:, :-1]
        at0 = coef[:, 1:, :, d] * fac[d]
        assert np.abs(at1 - at0).max() < 2e-4
