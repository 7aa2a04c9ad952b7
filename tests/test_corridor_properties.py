"""SFC construction of the host mirror (swarm_simulator_b200/host/rbp_corridor.hpp, updateObsBox) against INDEPENDENT
brute-force properties on random forests -- not against a second restatement of the same lines.
Reference semantics: /root/reference/swarm_planner/include/rbp_corridor.hpp L44-L243.

For every agent of every mission the boxes printed by corridor_cli (stage=sfc, host only) must satisfy, checked here
with plain numpy on the distance grid:
  P1 clearance   every voxel the box's sample lattice can touch (the box plus the 1e-6 nudges of L47-L63) keeps
                 quad_size from the obstacles;
  P2 maximality  every face is blocked: one more resolution step leaves the world or adds a voxel closer than quad_size;
  P3 coverage    both end points of every path edge lie in one common box;
  P4 timing      end times are knots of T, non-decreasing, the last one is the makespan, and at a hand-over knot the path
                 point lies in both boxes;
  P5 feasibility the box build_dlq selects for segment m (first box with t_end >= T[m+1], rbp_planner.hpp L443-L474)
                 holds path points m and m+1 -- what makes the initial trajectory a feasible point of the QP.
"""
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as G
from swarm_simulator_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "swarm_simulator_b200", "host")
EPSF, EPS = 1e-6, 1e-9


@pytest.fixture(scope="module")
def cli():
    G.build()
    return os.path.join(HOST, "corridor_cli")


def run_cli(cli, m, tmp_path, xy_res=0.1, z_res=0.1):
    synth.dump_world_text(m, str(tmp_path / "world.txt"))
    out = subprocess.check_output([cli, str(tmp_path / "world.txt"), "box/xy_res=%g" % xy_res, "box/z_res=%g" % z_res,
                                   "stage=sfc"], text=True).splitlines()
    assert out[0] == "update=true"
    sfc, i = [], 1
    for qi in range(m["N"]):
        tag, q, nb = out[i].split()
        assert tag == "SFC" and int(q) == qi
        rows = np.array([out[i + 1 + b].split() for b in range(int(nb))], float)
        sfc.append((rows[:, :6], rows[:, 6]))
        i += 1 + int(nb)
    return sfc


class Grid:
    def __init__(self, m):
        self.edt, self.k0, self.res = m["edt"], np.asarray(m["edt_k0"]), float(m["resolution"])
        wxy, wz = m["world_xy"], m["world_z"]
        self.wlo = np.array([wxy[0], wxy[1], wz[0]]); self.whi = np.array([wxy[2], wxy[3], wz[1]])

    def vox(self, v, a):
        return int(np.floor(float(np.float32(v)) / self.res)) - int(self.k0[a])

    def touched_min(self, lo, hi):
        """Smallest distance over every voxel a sample of the box [lo, hi] can fall into; -1 if one lies outside the map."""
        sl = []
        for a in range(3):
            first = lo[a] - EPSF if lo[a] > self.wlo[a] + EPSF else lo[a] + EPSF
            k0, k1 = self.vox(first, a), self.vox(hi[a] + EPSF, a)
            if k0 < 0 or k1 >= self.edt.shape[a]:
                return -1.0
            sl.append(slice(k0, k1 + 1))
        return float(self.edt[tuple(sl)].min())


def inside(box, p):
    return bool(np.all(p > box[:3] - EPS) and np.all(p < box[3:] + EPS))


@pytest.mark.parametrize("N,M,rho,seed", [(8, 5, 0.3, 11), (10, 6, 0.5, 12), (6, 7, 0.4, 13), (12, 5, 0.2, 14), (8, 4, 0.6, 15)])
def test_sfc_boxes_satisfy_bruteforce_properties(cli, tmp_path, N, M, rho, seed):
    m = synth.synth_mission(N, M, rho, seed)
    sfc = run_cli(cli, m, tmp_path)
    g = Grid(m)
    res = np.array([0.1, 0.1, 0.1])
    T, makespan = m["T"], m["T"][-1]
    unblocked = 0
    for qi in range(N):
        boxes, tend = sfc[qi]
        r = float(m["radius"][qi])
        path = m["init_traj"][qi].astype(np.float64)
        for b in boxes:
            # P1
            assert g.touched_min(b[:3], b[3:]) >= r - EPSF, (qi, b)
            assert np.all(b[:3] > g.wlo - EPS) and np.all(b[3:] < g.whi + EPS)
            # P2
            for f in range(6):
                a = f % 3
                lo, hi = b[:3].copy(), b[3:].copy()
                if f < 3:
                    hi[a] = lo[a]; lo[a] = lo[a] - res[a]
                    leaves = not lo[a] > g.wlo[a] - EPS
                else:
                    lo[a] = hi[a]; hi[a] = hi[a] + res[a]
                    leaves = not hi[a] < g.whi[a] + EPS
                if not (leaves or g.touched_min(lo, hi) < r - EPSF):
                    unblocked += 1
        # P3
        for j in range(M):
            assert any(inside(b, path[j]) and inside(b, path[j + 1]) for b in boxes), (qi, j)
        # P4
        assert np.all(np.diff(tend) >= 0) and tend[-1] == makespan
        for i, t in enumerate(tend[:-1]):
            j = int(np.flatnonzero(T == t)[0])          # must be a knot
            assert inside(boxes[i], path[j]) and inside(boxes[i + 1], path[j]), (qi, i, j)
        # P5
        bi = 0
        for mm in range(M):
            while bi < len(tend) and tend[bi] < T[mm + 1]:
                bi += 1
            assert bi < len(tend)
            assert inside(boxes[bi], path[mm]) and inside(boxes[bi], path[mm + 1]), (qi, mm, bi)
    # a face can stay unblocked only through the whole-box re-probe corner of the greedy growth (L99-L147); it must be rare
    assert unblocked <= 1, unblocked


def test_coarser_box_resolution_and_world_boundary(cli, tmp_path):
    """An empty world: every box must grow to the world boundary on the lower faces and to one step below it on the
    upper faces (the sample 1e-6 above an upper face on the boundary falls outside the map: L47-L68)."""
    m = synth.synth_mission(4, 4, 0.0, 3)
    sfc = run_cli(cli, m, tmp_path)
    g = Grid(m)
    for boxes, tend in sfc:
        for b in boxes:
            assert np.allclose(b[:3], g.wlo, atol=1e-9)
            assert np.all(b[3:] <= g.whi + 1e-9) and np.all(b[3:] >= g.whi - 0.1 - 1e-9)
