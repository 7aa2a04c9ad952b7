"""TEST INFRASTRUCTURE shared by the feasibility-classifier tests (CPU: oracle, GPU: engine through the C ABI)."""
import os

import numpy as np

import feas_classifier as fc
import oracle
import oracle_util
from swarm_simulator_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# (pack, sequential, batch_size, number of missions).  configs[0..3] of BASELINE.json with the batchings the verdict names:
# the joint batch of the config, the launch default b = 4 and the per-agent b = 1.
CASES = [
    ("cfg1", False, 4, 256), ("cfg1", True, 1, 256),
    ("cfg2", False, 16, 256), ("cfg2", True, 4, 256), ("cfg2", True, 1, 256),
    ("cfg3", True, 1, 256), ("cfg3", True, 4, 256),
    ("cfg4", True, 1, 32), ("cfg4", True, 4, 32), ("cfg4", True, 32, 32),
]

_packs = {}


def missions(pack, count=None, inflate_radius=None):
    if pack not in _packs:
        _packs[pack] = synth.load_pack(os.path.join(GOLDEN, "missions_%s.npz" % pack))
    ms = _packs[pack][:count]
    if inflate_radius is not None:
        ms = [dict(m, radius=np.full(m["N"], inflate_radius)) for m in ms]
    return ms


def joint_violation(m, ctrl, sequential):
    """Independent certificate for a whole mission result ctrl[N,3,6M]: after the last batch every row of every batch QP
    holds for the FINAL control points (a pair row was enforced when its later agent was solved, with the earlier one
    already final), so (max |Ax-b|, max box violation, max RSFC violation) of the final table bounds the infeasibility of
    every accepted QP.  Built from the assembly pieces pinned to the reference's LP (tests/test_oracle_golden.py)."""
    p = oracle_util.oracle_problem(m, sequential=sequential, batch_size=1)
    N, M = m["N"], m["M"]
    x = np.transpose(ctrl, (0, 2, 1))                       # [N, 6M, 3]
    rc, ub, lbn, rel = p.dlq()
    assert rc == 0
    vbox = max(float((x - ub).max()), float((-x - lbn).max()))
    A = oracle.Aeq_base(m["T"])                             # (3M+3) x 6M
    deq = p.deq().reshape(N, 3 * M + 3, 3)
    veq = float(np.abs(np.einsum("rj,njk->nrk", A, x) - deq).max())
    vrel = -np.inf
    if sequential and N > 1:
        qi, qj = np.triu_indices(N, 1)
        r = m["radius"][qi] + m["radius"][qj]
        lhs = np.einsum("pjk,pjk->pj", rel, x[qj] - x[qi])  # n.(x_qj - x_qi) >= r
        vrel = float((r[:, None] - lhs).max())
    elif N > 1:                                             # one joint batch: same rows
        qi, qj = np.triu_indices(N, 1)
        r = m["radius"][qi] + m["radius"][qj]
        lhs = np.einsum("pjk,pjk->pj", rel, x[qj] - x[qi])
        vrel = float((r[:, None] - lhs).max())
    return veq, vbox, vrel


def failing_qp(m, sequential, batch_size, bad_batch):
    """The batch QP (oracle.QP) that failed, rebuilt by walking the Gauss-Seidel chain with the oracle up to it."""
    p = oracle_util.oracle_problem(m, sequential=sequential, batch_size=batch_size)
    N, M = m["N"], m["M"]
    oq = 6 * M
    dummy = p.dummy() if sequential else np.zeros((N * oq, 3))
    _, ebs, _ = p.set_batch()
    for k in range(bad_batch):
        q = p.populate(dummy, k)
        r = q.solve()
        assert r["status"] == 0
        nb = min(ebs, N - k * ebs)
        od = nb * oq
        for kk in range(3):
            for bi in range(nb):
                qa = k * ebs + bi
                dummy[qa * oq:(qa + 1) * oq, kk] = r["x"][kk * od + bi * oq: kk * od + (bi + 1) * oq]
    return p, p.populate(dummy, bad_batch)


def judge(statuses, first_bad, ms, sequential, batch_size, ctrls):
    """statuses[c]: mission status (0 OK, 1 INFEASIBLE, 2 NOT_CONVERGED); first_bad[c]: index of the failing batch.
    Returns a list of failure strings (empty = pass) and a tally."""
    fails, tally = [], {}
    for c, m in enumerate(ms):
        st = int(statuses[c])
        if st == 0:
            veq, vbox, vrel = joint_violation(m, ctrls[c], sequential)
            ok = veq < 1e-7 and vbox < 2e-6 and vrel < 2e-6
            tally["ok"] = tally.get("ok", 0) + 1
            if not ok:
                fails.append("seed %d: OK but final table violates rows: eq %.2e box %.2e rsfc %.2e" % (m["seed"], veq, vbox, vrel))
            continue
        _, q = failing_qp(m, sequential, batch_size, int(first_bad[c]))
        verdict, slack, _ = fc.classify(q)
        key = "%s->%d" % (verdict, st)
        tally[key] = tally.get(key, 0) + 1
        if verdict == fc.STRICT:
            fails.append("seed %d batch %d: strictly feasible (slack %.3g) but status %d" % (m["seed"], first_bad[c], slack, st))
        if verdict == fc.INFEASIBLE and st != 1:
            fails.append("seed %d batch %d: LP-infeasible (slack %.3g) but status %d" % (m["seed"], first_bad[c], slack, st))
    return fails, tally
