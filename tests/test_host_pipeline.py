"""Host stages in front of the hot path: octomap .bt reader + clamped distance map
(swarm_simulator_b200/host/octree_bt.hpp), ECBS initial trajectories (host/ecbs_planner.hpp) and the whole planner-node
pipeline through swarm_plan_cli.  The reference has no golden vectors for these stages (parity unpinned): the checks are an
independent Python restatement of the .bt format (writer + reader), scipy's exact distance transform, the reference's
conflict rules (third_party/ecbs/include/environment.hpp L656-L681) re-implemented here, and end-to-end invariants."""
import json
import math
import os
import struct
import subprocess

import numpy as np
import pytest

import __graft_entry__ as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "swarm_simulator_b200", "host")
CLI = os.path.join(HOST, "swarm_plan_cli")
REF = "/root/reference/swarm_planner"


@pytest.fixture(scope="module")
def built():
    G.build()
    assert os.path.exists(CLI)
    return True


# ---- independent restatement of the .bt format (octomap OcTree::writeBinaryNode / readBinaryNode) in Python ----
def write_bt(path, occ_cells, free_cells, res=0.1):
    """occ_cells / free_cells: sets of integer voxel indices (ix, iy, iz) = floor(coord / res), written as depth-16 leaves."""
    tree = {}
    def insert(cell, value):
        key = [c + 32768 for c in cell]
        node = tree
        for d in range(15, -1, -1):
            i = ((key[0] >> d) & 1) | (((key[1] >> d) & 1) << 1) | (((key[2] >> d) & 1) << 2)
            if d == 0:
                node[i] = value
            else:
                node = node.setdefault(i, {})
    for c in free_cells:
        insert(c, "free")
    for c in occ_cells:
        insert(c, "occ")
    body = bytearray()
    count = [1]
    def emit(node):
        bits = 0
        for i in range(8):
            ch = node.get(i)
            if ch is None:
                continue
            count[0] += 1
            code = 3 if isinstance(ch, dict) else (2 if ch == "occ" else 1)     # (bit 2i, bit 2i+1): occ = (0,1), free = (1,0)
            bits |= code << (2 * i)
        body.extend(struct.pack("<BB", bits & 0xFF, (bits >> 8) & 0xFF))
        for i in range(8):
            if isinstance(node.get(i), dict):
                emit(node[i])
    emit(tree)
    with open(path, "wb") as f:
        f.write(b"# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n")
        f.write(b"id OcTree\nsize %d\nres %g\ndata\n" % (count[0], res))
        f.write(bytes(body))
    return count[0]


def read_bt_python(path):
    """-> (res, declared nodes, list of (cx, cy, cz, size, occupied))"""
    raw = open(path, "rb").read()
    pos, res, size = 0, None, None
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode()
        pos = end + 1
        if line.startswith("res"):
            res = float(line.split()[1])
        if line.startswith("size"):
            size = int(line.split()[1])
        if line.strip() == "data":
            break
    leaves = []
    def node(p, c, s):
        b = raw[p] | (raw[p + 1] << 8)
        p += 2
        inner = []
        for i in range(8):
            code = (b >> (2 * i)) & 3
            if code == 0:
                continue
            cc = tuple(c[a] + (s / 4 if (i >> a) & 1 else -s / 4) for a in range(3))
            if code == 3:
                inner.append(cc)
            else:
                leaves.append(cc + (s / 2, code == 2))
        for cc in inner:
            p = node(p, cc, s / 2)
        return p
    if size:
        node(pos, (0.0, 0.0, 0.0), res * 65536)
    return res, size, leaves


def forest(seed, n_pillars=12, L=10.0, res=0.1):
    rng = np.random.default_rng(seed)
    occ = set()
    for _ in range(n_pillars):
        cx, cy = rng.uniform(-L / 2 + 1.2, L / 2 - 1.2, 2)
        if abs(abs(cx) - 4) < 0.9 or abs(abs(cy) - 4) < 0.9 or (abs(cx) < 0.9 and abs(cy) < 0.9):
            continue                                    # keep the start / goal ring free
        ix, iy = int(math.floor(cx / res)), int(math.floor(cy / res))
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for iz in range(int(rng.integers(8, 26))):
                    occ.add((ix + dx, iy + dy, iz))
    return occ


def run_cli(*args):
    return subprocess.run([CLI] + [str(a) for a in args], capture_output=True, text=True)


def write_mission(path, n_agents, radius=0.15):
    agents = []
    for i in range(n_agents):
        a = 2 * math.pi * i / n_agents
        s = [round(4 * math.cos(a) * 2) / 2, round(4 * math.sin(a) * 2) / 2, 1.0]
        agents.append({"name": "crazyflie", "start": s, "goal": [-s[0], -s[1], 1.0], "radius": radius, "speed": 1.0})
    json.dump({"quadrotors": {"crazyflie": {"max_vel": [1.7] * 3, "max_acc": [6.2] * 3, "radius": radius, "speed": 1.0}},
               "agents": agents}, open(path, "w"))
    return agents


def test_bt_reader_matches_python_restatement(built, tmp_path):
    occ = forest(1)
    free = {(x, y, 0) for x in range(-3, 3) for y in range(-3, 3)} - occ
    n_nodes = write_bt(tmp_path / "w.bt", occ, free)
    res, size, leaves = read_bt_python(tmp_path / "w.bt")
    assert size == n_nodes and sum(1 for l in leaves if l[4]) == len(occ)
    json.dump({"quadrotors": {}, "agents": []}, open(tmp_path / "m.json", "w"))
    out = run_cli(tmp_path / "m.json", tmp_path / "w.bt", tmp_path, "stage=world", "world/z_min=0").stdout.splitlines()
    head = dict(kv.split("=") for kv in out[0].split()[1:])
    assert int(head["declared_nodes"]) == n_nodes == int(head["inner"]) + int(head["leaves"])
    assert int(head["occupied_leaves"]) == len(occ)
    cols = {}
    for (x, y, z) in occ:
        if -50 <= x <= 50 and -50 <= y <= 50 and 0 <= z <= 25:
            cols[(x, y)] = cols.get((x, y), 0) + 1
    got = {(int(a), int(b)): int(c) for a, b, c in (l.split()[1:] for l in out[1:] if l.startswith("col"))}
    assert got == cols


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference worlds not mounted")
@pytest.mark.parametrize("world", ["empty", "map1", "map23", "map50", "IROS2019", "IROS2019_2", "ICRA2020_64agents_presentation",
                                   "map_reduced_tmp3"])
def test_bt_reader_on_reference_worlds(built, world, tmp_path):
    """The reference's own octomaps: node count of the header = nodes read; C++ and Python readers agree; a random forest
    world holds about obs_num = 20 pillars of 3 x 3 cells (random_map_generator.cpp L61-L104)."""
    path = os.path.join(REF, "worlds", world + ".bt")
    res, size, leaves = read_bt_python(path)
    json.dump({"quadrotors": {}, "agents": []}, open(tmp_path / "m.json", "w"))
    out = run_cli(tmp_path / "m.json", path, tmp_path, "stage=world").stdout.splitlines()
    head = dict(kv.split("=") for kv in out[0].split()[1:])
    assert int(head["declared_nodes"]) == size == int(head["inner"]) + int(head["leaves"])
    assert int(head["leaves"]) == len(leaves) and int(head["occupied_leaves"]) == sum(1 for l in leaves if l[4])
    ncol = sum(1 for l in out if l.startswith("col"))
    if world[3:].isdigit():          # map<k>.bt: the 50 random forests
        assert 9 * 10 <= ncol <= 9 * 20


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference worlds not mounted")
def test_reference_smoke_loop_stage_1_on_all_50_maps(built):
    """Stage 1 of the reference's own smoke loop (swarm_traj_planner_rbp_test_all.cpp L49-L78, plan_rbp_test.launch): the
    64-agent mission must get initial trajectories on every one of worlds/map1..50.bt (the loop returns -1 otherwise)."""
    out = subprocess.run([os.sys.executable, os.path.join(ROOT, "tools", "plan_test_all.py"), REF, "stage=ecbs"], capture_output=True, text=True)
    assert out.returncode == 0 and "50 / 50 maps planned" in out.stdout, out.stdout[-2000:]


def test_distance_map_matches_scipy(built, tmp_path):
    """clamped_edt (Felzenszwalb) behind GridDistanceMap == scipy's exact EDT clamped at 1 m, probed through ECBS's obstacle
    rule: a grid point is an obstacle iff dist < r + grid_margin (ecbs_planner.hpp L99)."""
    from scipy.ndimage import distance_transform_edt
    occ = forest(2)
    write_bt(tmp_path / "w.bt", occ, set())
    grid = np.zeros((101, 101, 26), bool)
    for (x, y, z) in occ:
        if -50 <= x <= 50 and -50 <= y <= 50 and 0 <= z <= 25:
            grid[x + 50, y + 50, z] = True
    edt = np.minimum(distance_transform_edt(~grid) * 0.1, 1.0)
    agents = write_mission(tmp_path / "m.json", 2)
    out = run_cli(tmp_path / "m.json", tmp_path / "w.bt", tmp_path, "stage=ecbs", "world/z_min=0")
    assert "ecbs=true" in out.stdout, out.stdout + out.stderr
    blocked = set()
    for i, x in enumerate(np.arange(-5, 5.0001, 0.5)):
        for j, y in enumerate(np.arange(-5, 5.0001, 0.5)):
            for k, z in enumerate(np.arange(0, 2.0001, 1.0)):
                if edt[int(math.floor(np.float32(x) / 0.1)) + 50, int(math.floor(np.float32(y) / 0.1)) + 50, int(math.floor(np.float32(z) / 0.1))] < 0.15 + 0.2:
                    blocked.add((round(float(x), 3), round(float(y), 3), round(float(z), 3)))
    for line in out.stdout.splitlines():
        if line.startswith("traj"):
            v = [float(t) for t in line.split()[2:]]
            for p in zip(v[0::3], v[1::3], v[2::3]):
                assert (round(p[0], 3), round(p[1], 3), round(p[2], 3)) not in blocked


def _conflicts(paths, radius, grid=0.5):
    """environment.hpp L656-L681 re-implemented: vertex / edge conflicts between grid paths (cells, one per time step)."""
    def min_dist(a, b):
        a, b = np.asarray(a, float), np.asarray(b, float)
        d = np.linalg.norm(a)
        if not np.array_equal(a, b):
            d = min(d, np.linalg.norm(b))
            n = (b - a) / np.linalg.norm(b - a)
            c = a - n * a.dot(n)
            if (c - a).dot(c - b) < 0:
                d = min(d, np.linalg.norm(c))
        return d
    T = max(len(p) for p in paths)
    at = lambda p, t: np.asarray(p[min(t, len(p) - 1)])
    bad = 0
    for t in range(T - 1):
        for i in range(len(paths)):
            for j in range(i + 1, len(paths)):
                if np.linalg.norm(at(paths[j], t) - at(paths[i], t)) * grid < 2 * radius:
                    bad += 1
                if min_dist(at(paths[j], t) - at(paths[i], t), at(paths[j], t + 1) - at(paths[i], t + 1)) * grid <= 2 * radius:
                    bad += 1
    return bad


@pytest.mark.parametrize("n_agents,seed", [(4, 3), (8, 4), (16, 5)])
def test_ecbs_paths_are_valid(built, tmp_path, n_agents, seed):
    occ = forest(seed)
    write_bt(tmp_path / "w.bt", occ, set())
    agents = write_mission(tmp_path / "m.json", n_agents)
    out = run_cli(tmp_path / "m.json", tmp_path / "w.bt", tmp_path, "stage=ecbs", "ecbs/w=1.3", "world/z_min=0")
    assert "ecbs=true" in out.stdout, out.stdout + out.stderr
    M = int(out.stdout.split("M=")[1].split()[0])
    paths, cost = [], 0
    for a, line in zip(agents, [l for l in out.stdout.splitlines() if l.startswith("traj")]):
        v = [float(t) for t in line.split()[2:]]
        pts = list(zip(v[0::3], v[1::3], v[2::3]))
        assert len(pts) == M + 1                                               # ecbs_planner.hpp L66-L71
        assert pts[0] == tuple(a["start"]) and pts[-1] == tuple(a["goal"])
        cells = [(round((p[0] + 5) / 0.5), round((p[1] + 5) / 0.5), round(p[2] / 1.0)) for p in pts[1:]]
        for c0, c1 in zip(cells, cells[1:]):
            assert sum(abs(u - w) for u, w in zip(c0, c1)) <= 1               # wait or one 6-connected move
        steps = max(i for i, c in enumerate(cells) if c != cells[-1]) + 1 if any(c != cells[-1] for c in cells) else 0
        assert steps >= sum(abs(u - w) for u, w in zip(cells[0], cells[-1]))   # never shorter than the Manhattan bound
        cost += steps
        paths.append(cells)
    assert max(len(p) for p in paths) == M
    assert _conflicts(paths, 0.15) == 0


def _optimal_sum_of_costs(starts, goals, blocked, dims, radius=0.15, grid=0.5):
    """Exact minimum sum of costs for two agents on a small 2-D grid under the reference's conflict rules (Dijkstra over joint
    states; an agent that has committed to its goal stays there and stops paying)."""
    import heapq
    moves = [(0, 0), (-1, 0), (1, 0), (0, 1), (0, -1)]
    def ok(c):
        return 0 <= c[0] < dims[0] and 0 <= c[1] < dims[1] and c not in blocked
    def vconf(a, b):
        return a == b                                            # 2 r = 0.3 < grid
    def econf(a0, a1, b0, b1):
        return _conflicts([[a0 + (0,), a1 + (0,)], [b0 + (0,), b1 + (0,)]], radius, grid) > 0 and not vconf(a0, b0)
    start = (tuple(starts[0]), tuple(starts[1]), False, False)
    pq, best = [(0, start)], {start: 0}
    while pq:
        c, st = heapq.heappop(pq)
        if c > best.get(st, 1e9):
            continue
        p, q, fp, fq = st
        if fp and fq:
            return c
        for mp in ([(0, 0)] if fp else moves):
            for mq in ([(0, 0)] if fq else moves):
                np_, nq = (p[0] + mp[0], p[1] + mp[1]), (q[0] + mq[0], q[1] + mq[1])
                if not ok(np_) or not ok(nq) or vconf(np_, nq):
                    continue
                # edge rule of environment.hpp L666-L681 between the two moves
                a = np.array([q[0] - p[0], q[1] - p[1], 0.0]); b = np.array([nq[0] - np_[0], nq[1] - np_[1], 0.0])
                d = np.linalg.norm(a)
                if not np.array_equal(a, b):
                    d = min(d, np.linalg.norm(b))
                    n = (b - a) / np.linalg.norm(b - a)
                    cc = a - n * a.dot(n)
                    if (cc - a).dot(cc - b) < 0:
                        d = min(d, np.linalg.norm(cc))
                if d * grid <= 2 * radius:
                    continue
                step = (0 if fp else 1) + (0 if fq else 1)
                for nfp in ([True] if fp else ([False, True] if np_ == tuple(goals[0]) else [False])):
                    for nfq in ([True] if fq else ([False, True] if nq == tuple(goals[1]) else [False])):
                        ns = (np_, nq, nfp, nfq)
                        if c + step < best.get(ns, 1e9):
                            best[ns] = c + step
                            heapq.heappush(pq, (c + step, ns))
    return None


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_ecbs_is_within_its_suboptimality_bound(built, tmp_path, seed):
    """Two agents on a 7 x 7 grid with random blocked cells: cost of the ECBS solution is between the exact optimum
    (joint-state Dijkstra above) and w times it (w = 1.3, ecbs.hpp focal bound)."""
    rng = np.random.default_rng(100 + seed)
    dims = (7, 7)
    cells = [(x, y) for x in range(7) for y in range(7)]
    pick = rng.permutation(len(cells))
    s0, s1, g0, g1 = (cells[i] for i in pick[:4])
    if seed % 2 == 0:
        g0, g1 = s1, s0                                           # a swap: the hardest two-agent case
    blocked = {cells[i] for i in pick[4:4 + 8]}
    occ = set()
    for (bx, by) in blocked:                                      # a pillar at the grid point (world = -1.5 + 0.5 * index)
        ix, iy = int(math.floor((-1.5 + 0.5 * bx) / 0.1)), int(math.floor((-1.5 + 0.5 * by) / 0.1))
        for z in range(0, 11):
            occ.add((ix, iy, z))
    write_bt(tmp_path / "w.bt", occ, set())
    w2c = lambda c: [-1.5 + 0.5 * c[0], -1.5 + 0.5 * c[1], 0.0]
    json.dump({"quadrotors": {"q": {"max_vel": [1.7] * 3, "max_acc": [6.2] * 3, "radius": 0.15, "speed": 1.0}},
               "agents": [{"name": "q", "start": w2c(s0), "goal": w2c(g0), "radius": 0.15, "speed": 1.0},
                          {"name": "q", "start": w2c(s1), "goal": w2c(g1), "radius": 0.15, "speed": 1.0}]}, open(tmp_path / "m.json", "w"))
    # a pillar cell blocks its own grid point only: margin r + 0.2 = 0.35 m < 0.5 m grid pitch (cell centres are 0.05 m off)
    out = run_cli(tmp_path / "m.json", tmp_path / "w.bt", tmp_path, "stage=ecbs", "ecbs/w=1.3", "world/x_min=-1.5", "world/x_max=1.5",
                  "world/y_min=-1.5", "world/y_max=1.5", "world/z_min=0", "world/z_max=0.9", "grid/z_res=1.0")
    opt = _optimal_sum_of_costs([s0, s1], [g0, g1], blocked, dims)
    if "ecbs=true" not in out.stdout:
        assert opt is None, out.stdout + out.stderr
        return
    cost = 0
    for line in [l for l in out.stdout.splitlines() if l.startswith("traj")]:
        v = [float(t) for t in line.split()[2:]]
        cells_ = [(round((x + 1.5) / 0.5), round((y + 1.5) / 0.5)) for x, y in zip(v[0::3][1:], v[1::3][1:])]
        last = cells_[-1]
        moving = [i for i, c in enumerate(cells_) if c != last]
        cost += (max(moving) + 1) if moving else 0
    print("seed", seed, "optimal", opt, "ecbs", cost)
    assert opt is not None and opt <= cost <= math.floor(1.3 * opt + 1e-9), (opt, cost, out.stdout)


@pytest.mark.gpu
def test_whole_pipeline_on_a_forest_world(built, tmp_path):
    """.bt world -> distance map -> ECBS -> Corridor (SFC on the host, RSFC kernel) -> RBPPlanner (B200 engine) ->
    RBPPublisher's collision check: the plan must be collision free (safety_margin_ratio >= 1, rbp_publisher.hpp L769-L798)."""
    occ = forest(7)
    write_bt(tmp_path / "w.bt", occ, set())
    write_mission(tmp_path / "m.json", 8)
    for extra in (["plan/sequential=true", "plan/batch_size=4"], ["plan/sequential=true", "plan/batch_size=1"]):
        out = run_cli(tmp_path / "m.json", tmp_path / "w.bt", tmp_path, "stage=all", "world/z_min=0", *extra)
        assert "rbp=true" in out.stdout, out.stdout + out.stderr
        ratio = float(out.stdout.split("safety_margin_ratio=")[1].split()[0])
        assert ratio >= 1.0 - 1e-6
