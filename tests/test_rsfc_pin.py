"""Pins the float32 operation order of Corridor::updateRelBox (/root/reference/swarm_planner/include/rbp_corridor.hpp
L338-L398, octomap::point3d arithmetic) to OUTPUT OF THE REFERENCE BINARY.

log/QPmodel.lp holds 8 856 RSFC normals (246 pairs touching batch 15 x 36 segments) that the reference computed from the
ECBS lattice paths of mission_64agents_15.json.  The paths are not in the repository, but every relative displacement of
two agents on that lattice is enumerable: positions are multiples of 0.5 m in x, y (starts and goals included) and of
1.0 m in z, and each agent makes one of the seven ECBS moves per step (third_party/ecbs/include/environment.hpp
L467-L524).  So every LP normal must be BIT-EQUAL to updateRelBox(a, b) for some lattice pair (a, b = a + move_j - move_i).
The restatement (oracle.rsfc, and the device kernel behind rbpe_corridor_rsfc) reproduces all 8 856; the same geometry
evaluated in float64 and rounded at the end -- the negative control -- misses about a third of them, so the test does
discriminate between operation orders.
"""
import numpy as np
import pytest

import fixture_lp as F
import oracle

MOVES = [(0, 0, 0), (.5, 0, 0), (-.5, 0, 0), (0, .5, 0), (0, -.5, 0), (0, 0, 1), (0, 0, -1)]


def lp_normals():
    rec = F.recover_inputs(F.load_lp(), F.load_csv(), F.load_mission())
    out, it = [], 0
    for qi in range(F.N):
        for qj in range(qi + 1, F.N):
            if qj >= F.B0:
                out.append(rec["rsfc_n"][it])
            it += 1
    return np.concatenate(out)          # [246 * 36, 3] float32


def lattice_pairs():
    deltas = sorted({(mj[0] - mi[0], mj[1] - mi[1], mj[2] - mi[2]) for mi in MOVES for mj in MOVES})
    A, B = [], []
    for i in range(-16, 17):            # the mission lives on an 8 m square: |dx|, |dy| <= 8
        for j in range(-16, 17):
            for k in (-1, 0, 1):        # z levels 1.0 and 2.0 (world 0.3 .. 2.5, grid 1.0)
                if i == 0 and j == 0 and k == 0:
                    continue
                a = (i * .5, j * .5, k * 1.0)
                for d in deltas:
                    b = (a[0] + d[0], a[1] + d[1], a[2] + d[2])
                    if abs(b[2]) > 1 or b == (0, 0, 0):
                        continue
                    A.append(a); B.append(b)
    return np.asarray(A, np.float32), np.asarray(B, np.float32)


def keyset(n):
    n = np.ascontiguousarray(n, np.float32)
    return set(map(bytes, n.view(np.uint8).reshape(len(n), 12)))


def f64_variant(a, b, dw=2.0):
    a = a.astype(np.float64).copy(); b = b.astype(np.float64).copy()
    a[2] /= dw; b[2] /= dw
    m = a
    if not np.array_equal(a, b):
        dmin, d = np.linalg.norm(a), np.linalg.norm(b)
        if dmin > d:
            m, dmin = b, d
        n = (b - a) / np.linalg.norm(b - a)
        c = a - n * (a @ n)
        if (c - a) @ (c - b) < 0 and dmin > np.linalg.norm(c):
            m = c
    m = m / np.linalg.norm(m)
    m[2] /= dw
    return m.astype(np.float32)


def test_lp_normals_are_bit_equal_to_the_restatement_on_the_ecbs_lattice():
    normals = lp_normals()
    assert normals.shape == (8856, 3)
    A, B = lattice_pairs()
    tr = np.zeros((2, 2, 3), np.float32)
    T = np.array([0.0, 1.0])
    out = np.zeros((len(A), 3), np.float32)
    for n in range(len(A)):
        tr[1, 0] = A[n]; tr[1, 1] = B[n]
        out[n] = oracle.rsfc(tr, T, 2.0)[0][0, 0]
    cand = keyset(out)
    missing = [n for n in normals if bytes(np.ascontiguousarray(n).view(np.uint8)) not in cand]
    assert not missing, (len(missing), missing[:3])
    # negative control: float64 geometry rounded at the end does NOT reproduce the reference's bits
    ctrl = keyset(np.array([f64_variant(A[n], B[n]) for n in range(len(A))]))
    miss64 = sum(1 for n in normals if bytes(np.ascontiguousarray(n).view(np.uint8)) not in ctrl)
    assert miss64 > 1000, miss64


@pytest.mark.gpu
def test_lp_normals_are_bit_equal_to_the_device_kernel_on_the_ecbs_lattice():
    import __graft_entry__ as G
    from swarm_simulator_b200 import engine as E
    G.build()
    eng = E.Engine(device=0)
    normals = lp_normals()
    A, B = lattice_pairs()
    tr = np.zeros((len(A), 2, 2, 3), np.float32)      # one two-agent, one-segment "mission" per lattice pair
    tr[:, 1, 0] = A; tr[:, 1, 1] = B
    T = np.tile(np.array([0.0, 1.0]), (len(A), 1))
    n, _, col = eng.corridor_rsfc(tr, T, 2.0)       # (pairs whose segment passes through the origin are flagged `collided`: fine)
    cand = keyset(n[:, 0, 0])
    missing = [v for v in normals if bytes(np.ascontiguousarray(v).view(np.uint8)) not in cand]
    assert not missing, (len(missing), missing[:3])
    eng.close()
