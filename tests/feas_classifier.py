"""TEST INFRASTRUCTURE: independent feasibility classifier for one batch QP of populatebyrow
(/root/reference/swarm_planner/include/rbp_planner.hpp L551-L688).

The interior-point solvers (oracle and CUDA) share one algorithm, so they cannot check each other's feasibility
verdicts.  This module asks HiGHS (scipy.optimize.linprog, dual simplex / IPM of an unrelated code base) for the
largest uniform slack of the QP's inequality rows:

    max t   s.t.  A x = b,   G_live x + t <= h_live,   t <= 1

``live`` rows are the rows that still depend on x once A x = b holds; rows whose variables are all pinned by the
start / goal equalities are constants and are judged separately against CPLEX's feasibility tolerance (1e-6), exactly as
any presolve would.  Verdicts:

    slack >  1e-6           strictly feasible : cplex.solve() returns true  -> the engine must return OK
    LP infeasible / t < -1e-6  infeasible     : cplex.solve() returns false -> the engine must return INFEASIBLE
    otherwise               borderline (no interior at tolerance): either answer is accepted
"""
import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog

STRICT, BORDERLINE, INFEASIBLE = "strict", "borderline", "infeasible"
FEAS_TOL = 1e-6


def _csr(qp, which):
    ptr, idx, val, rhs = qp.csr_raw(which)
    m = len(rhs)
    return sp.csr_matrix((val, idx, ptr), shape=(m, qp.nv)), rhs


def pinned_variables(A, b):
    """Variables fixed by A x = b alone (start / goal rows pin control points 0..2 and 3M+3..3M+5 of every axis).
    Found structurally from the null space of A restricted to each variable: a variable is pinned iff e_i is in the
    row space of A.  A is small (<= 1728 x 2880); dense QR is fine."""
    Ad = A.toarray()
    # least-squares projection of each unit vector on the row space: pinned iff the residual is ~0
    Q, _ = np.linalg.qr(Ad.T)                      # columns span range(A')
    proj = np.einsum("ij,ij->i", Q, Q)             # |Q' e_i|^2 = sum_j Q[i,j]^2
    return proj > 1 - 1e-9


def classify(qp, cap=1.0):
    """Return (verdict, slack, info) for an oracle.QP."""
    A, b = _csr(qp, "a")
    G, h = _csr(qp, "g")
    nv = qp.nv
    pinned = pinned_variables(A, b)
    # value of the pinned variables: any solution of A x = b
    xfix = np.zeros(nv)
    if pinned.any():
        sol = linprog(np.zeros(nv), A_eq=A, b_eq=b, bounds=[(None, None)] * nv, method="highs")
        if sol.status != 0:
            return INFEASIBLE, -np.inf, dict(reason="equalities inconsistent")
        xfix = sol.x
    free_cols = np.flatnonzero(~pinned)
    nnz_free = np.diff(G[:, free_cols].tocsr().indptr)
    live = nnz_free > 0
    const_viol = 0.0
    if (~live).any():
        gc = G[~live] @ xfix - h[~live]
        const_viol = float(gc.max()) if len(gc) else 0.0
    if const_viol > FEAS_TOL:
        return INFEASIBLE, -const_viol, dict(reason="constant row violated", live=int(live.sum()))
    Gl, hl = G[live], h[live]
    ml = Gl.shape[0]
    # variables (x, t): minimise -t
    c = np.zeros(nv + 1); c[-1] = -1.0
    Aub = sp.hstack([Gl, sp.csr_matrix(np.ones((ml, 1)))], format="csr")
    Aeq = sp.hstack([A, sp.csr_matrix((A.shape[0], 1))], format="csr")
    bounds = [(None, None)] * nv + [(None, cap)]
    sol = linprog(c, A_ub=Aub, b_ub=hl, A_eq=Aeq, b_eq=b, bounds=bounds, method="highs")
    if sol.status == 2:
        return INFEASIBLE, -np.inf, dict(reason="LP infeasible", live=ml)
    if sol.status != 0:
        raise RuntimeError("HiGHS status %d: %s" % (sol.status, sol.message))
    t = float(sol.x[-1])
    verdict = STRICT if t > FEAS_TOL else (INFEASIBLE if t < -FEAS_TOL else BORDERLINE)
    return verdict, t, dict(live=ml, const_viol=const_viol)


def check_point(qp, x, tol=FEAS_TOL):
    """Independent certificate that x is feasible: (max |Ax-b|, max (Gx-h)) computed from the CSR rows."""
    A, b = _csr(qp, "a")
    G, h = _csr(qp, "g")
    return float(np.abs(A @ x - b).max()), float((G @ x - h).max())
