"""Helpers around the golden fixtures distilled from the reference's CPLEX artefacts
(tests/golden/make_golden.py): the QP of batch 15 (agents 60..63 of 64, b=4, M=36) as the LP export
holds it, the CPLEX solution as the coefficient CSVs hold it, and the planner inputs (SFC boxes,
RSFC normals, `dummy` control points) that can be recovered from the two.
"""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

N, M, NB, B0 = 64, 36, 4, 60
OQ = 6 * M
OD = NB * OQ
NV = 3 * OD
NEQ1 = 3 * M + 3

# Bernstein -> monomial, row i = coefficients of B_i^5, highest power first (closed form)
from math import comb

BASIS = np.zeros((6, 6))
for _i in range(6):
    for _l in range(6 - _i):
        BASIS[_i, 5 - (_i + _l)] = comb(5, _i) * comb(5 - _i, _l) * (-1) ** _l


def load_lp():
    return dict(np.load(os.path.join(GOLD, "qpmodel_batch15.npz")))


def load_csv():
    return dict(np.load(os.path.join(GOLD, "coef_csv.npz")))


def load_mission():
    return dict(np.load(os.path.join(GOLD, "mission_64agents_15.npz")))


def csv_ctrl(coef_low_first, dt=1.0):
    """coef[..., 6] lowest power first (CSV order) -> Bernstein control points [..., 6].

    Inverse of rbp_planner.hpp L167-L196: coef_high_first = vals^T (basis * diag(dt^-(5-j)))."""
    high = coef_low_first[..., ::-1] * (dt ** np.arange(5, -1, -1))
    return high @ np.linalg.inv(BASIS)


def ctrl_coef_low(ctrl, dt=1.0):
    """control points [..., 6] -> monomial coefficients lowest power first (CSV order)."""
    high = ctrl @ BASIS / (dt ** np.arange(5, -1, -1))
    return high[..., ::-1]


def lp_as_leq(lp):
    """Split the LP rows into equalities and <=-form inequalities (>= rows are negated).

    Returns dict with CSR pieces: a_ptr,a_idx,a_val,b ; g_ptr,g_idx,g_val,h."""
    ptr, idx, val, sense, rhs = lp["indptr"], lp["indices"], lp["values"], lp["sense"], lp["rhs"]
    out = {}
    eq = np.flatnonzero(sense == 0)
    assert np.array_equal(eq, np.arange(len(eq)))  # equalities come first (populatebyrow order)
    ne = len(eq)
    out["a_ptr"] = ptr[:ne + 1].astype(np.int32)
    out["a_idx"] = idx[:ptr[ne]].astype(np.int32)
    out["a_val"] = val[:ptr[ne]].copy()
    out["b"] = rhs[:ne].copy()
    sgn = np.where(sense[ne:] == 2, -1.0, 1.0)
    out["g_ptr"] = (ptr[ne:] - ptr[ne]).astype(np.int32)
    out["g_idx"] = idx[ptr[ne]:].astype(np.int32)
    out["g_val"] = val[ptr[ne]:] * np.repeat(sgn, np.diff(ptr[ne:]))
    out["h"] = rhs[ne:] * sgn
    return out


def perms():
    """segment-major variable order / knot-major equality order used by the oracle solver."""
    px = np.zeros(NV, np.int32)
    for k in range(3):
        for bi in range(NB):
            for m in range(M):
                for i in range(6):
                    px[k * OD + bi * OQ + m * 6 + i] = ((m * NB + bi) * 3 + k) * 6 + i
    py = np.zeros(3 * NB * NEQ1, np.int32)
    r = 0
    for k in range(3):
        for bi in range(NB):
            for i in range(NEQ1):
                knot = 0 if i < 3 else (M if i < 6 else (i - 6) // 3 + 1)
                d = i % 3 if i < 6 else (i - 6) % 3
                py[r] = knot * 9 * NB + (bi * 3 + k) * 3 + d
                r += 1
    return px, py


def lp_qp_arrays(lp):
    """Arrays for oracle.QP(...) built straight from the LP export (no oracle assembly involved).

    LP objective is printed as [ sum c v w ] / 2, i.e. f = 1/2 sum c v w ; the oracle takes
    f = sum_t qv x_qi x_qj, so qv = c / 2."""
    d = lp_as_leq(lp)
    px, py = perms()
    d.update(nv=NV, ne=len(d["b"]), mi=len(d["h"]), qi=lp["obj_i"].astype(np.int32),
             qj=lp["obj_j"].astype(np.int32), qv=lp["obj_v"] / 2.0, perm_x=px, perm_y=py)
    return d


def recover_inputs(lp, csv, mission):
    """Planner inputs of batch 15 recovered from the artefacts.

    * seg_box[N, M, 6]: from the box rows (agents 60..63); other agents get a huge box (never used).
    * rsfc_n[P, M, 3] float32: from the RSFC rows of every pair touching the batch; other pairs (1,0,0).
    * dummy[N*6M, 3]: Bernstein control points of coef1..60.csv (solved agents 0..59, 6-digit precision);
      agents 60..63 (in batch) are set to NaN on purpose -- the assembly must never read them.
    """
    d = lp_as_leq(lp)
    ne = len(d["b"])
    gp, gi, gv, h = d["g_ptr"], d["g_idx"], d["g_val"], d["h"]
    seg_box = np.zeros((N, M, 6))
    seg_box[:, :, :3] = -100.0
    seg_box[:, :, 3:] = 100.0
    r = 0
    for k in range(3):
        for bi in range(NB):
            for j in range(OQ):
                assert gi[gp[r]] == k * OD + bi * OQ + j and gv[gp[r]] == 1.0
                assert gi[gp[r + 1]] == k * OD + bi * OQ + j and gv[gp[r + 1]] == -1.0
                ub, lb = h[r], -h[r + 1]
                m = j // 6
                if j % 6 == 0:
                    seg_box[B0 + bi, m, 3 + k] = ub
                    seg_box[B0 + bi, m, k] = lb
                else:
                    assert seg_box[B0 + bi, m, 3 + k] == ub and seg_box[B0 + bi, m, k] == lb
                r += 2
    P = N * (N - 1) // 2
    rsfc_n = np.zeros((P, M, 3), np.float32)
    rsfc_n[:, :, 0] = 1.0
    rhs_rsfc = {}
    it = 0
    for qi in range(N):
        for qj in range(qi + 1, N):
            bi, bj = qi - B0, qj - B0
            if bj >= 0:  # qj in batch (qi may be too)
                for j in range(OQ):
                    nvec = np.zeros(3)
                    for t in range(gp[r], gp[r + 1]):
                        v, k = gi[t], gi[t] // OD
                        b_ = (v % OD) // OQ
                        assert v % OQ == j
                        if b_ == bj:
                            nvec[k] = -gv[t]  # <= form carries -n on x_qj
                        else:
                            assert b_ == bi and bi >= 0
                    if j % 6 == 0:
                        rsfc_n[it, j // 6] = nvec.astype(np.float32)
                        # the LP prints 15 significant digits, so float32 -> double round-trips to ~1e-15
                        assert np.allclose(rsfc_n[it, j // 6].astype(np.float64), nvec, rtol=0, atol=2e-15)
                    rhs_rsfc[(qi, qj, j)] = h[r]
                    r += 1
            it += 1
    assert r == len(h)
    ctrl = csv_ctrl(csv["coef"])  # [N, M, 3, 6]
    dummy = np.transpose(ctrl, (0, 1, 3, 2)).reshape(N * OQ, 3).copy()
    dummy[B0 * OQ:] = np.nan
    return dict(seg_box=seg_box, rsfc_n=rsfc_n, dummy=dummy, ne=ne)


def sfc_from_seg_box(seg_box, T):
    """Run-length encode per-segment boxes back into an SFC list (box, t_end) per agent."""
    offs, boxes, tend = [0], [], []
    Nn, Mm = seg_box.shape[:2]
    for q in range(Nn):
        m = 0
        while m < Mm:
            e = m
            while e + 1 < Mm and np.array_equal(seg_box[q, e + 1], seg_box[q, m]):
                e += 1
            boxes.append(seg_box[q, m]); tend.append(T[e + 1])
            m = e + 1
        offs.append(len(boxes))
    return np.array(offs, np.int32), np.array(boxes), np.array(tend)
