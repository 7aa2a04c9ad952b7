"""Seeded synthetic missions for the RBP trajectory-QP engine.

This is the WORKLOAD GENERATOR: it produces the inputs RBPPlanner::update() consumes -- PlanResult{T, initTraj, SFC,
RSFC} plus Mission -- for a random-forest world, following the reference's upstream stages closely enough that the
QPs have the reference's structure:

  * forest        src/random_map_generator.cpp L61-L104 (0.3 m square pillars, ragged per-cell heights; seeded here)
  * distance map  DynamicEDTOctomap(maxDist=1) semantics (src/swarm_traj_planner_rbp.cpp L76-L80): Euclidean distance
                  between voxel centres on the 0.1 m grid, clamped to 1.0, -1 outside the bounding box
  * initTraj      include/ecbs_planner.hpp L41-L70 ([exact start, grid states, exact goal]); the discrete paths come
                  from prioritised space-time search with the reference's vertex / edge conflict rules
                  (third_party/ecbs/include/environment.hpp L656-L681) as a stand-in for ECBS (Boost is absent)
  * SFC           include/rbp_corridor.hpp L44-L243 (updateObsBox, expand_box, isObstacleInBox, time allocation)
  * RSFC          include/rbp_corridor.hpp L338-L398 (updateRelBox) in float32 with octomap::Vector3 semantics

It feeds tests and bench.py on both sides (engine and oracle consume identical bytes).  It is host-side numpy; the
hot path itself (assembly + QP solve) never runs here.
"""
import math

import numpy as np

SP_EPSILON = 1e-9
SP_EPSILON_FLOAT = 1e-6

DEFAULT_PARAM = dict(
    world_z_min=0.3, world_z_max=2.5, world_resolution=0.1,          # plan_rbp_random_forest.launch L29-L35
    obs_w=0.3, obs_h_min=0.0, obs_h_max=2.5,                         # L38-L43
    grid_xy_res=0.5, grid_z_res=1.0, grid_margin=0.2,                # L46-L49
    box_xy_res=0.1, box_z_res=0.1,                                   # L52-L53
    time_step=1.0, downwash=2.0,                                     # L56-L58
    radius=0.15, max_vel=1.7, max_acc=6.2,                           # missions/mission_64agents_15.json
)


class World:
    """Occupancy + clamped Euclidean distance map on the world_resolution grid."""

    def __init__(self, x_min, y_min, z_min, x_max, y_max, z_max, res):
        self.res = res
        self.lo = np.array([x_min, y_min, z_min], float)
        self.hi = np.array([x_max, y_max, z_max], float)
        self.k0 = np.floor(self.lo / res + 1e-9).astype(int)          # first voxel index inside the bounding box
        self.k1 = np.floor(self.hi / res + 1e-9).astype(int)          # last voxel index (inclusive, as bbxMaxKey)
        self.shape = tuple((self.k1 - self.k0 + 1).tolist())
        self.occ = np.zeros(self.shape, bool)
        self.edt = None

    def add_pillar(self, cx, cy, heights):
        """3x3 cells around the cell holding (cx, cy); heights[3][3] in metres from z = 0."""
        res = self.res
        ix = int(math.floor(cx / res)) - self.k0[0]
        iy = int(math.floor(cy / res)) - self.k0[1]
        for r in range(-1, 2):
            for s in range(-1, 2):
                hn = int(math.ceil(heights[r + 1][s + 1] / res))
                x, y = ix + r, iy + s
                if 0 <= x < self.shape[0] and 0 <= y < self.shape[1]:
                    for t in range(hn):
                        z = t - self.k0[2]
                        if 0 <= z < self.shape[2]:
                            self.occ[x, y, z] = True

    def update(self, max_dist=1.0):
        if self.occ.any():
            from scipy.ndimage import distance_transform_edt
            d = distance_transform_edt(~self.occ) * self.res
        else:
            d = np.full(self.shape, max_dist)
        self.edt = np.minimum(d, max_dist).astype(np.float32)

    def cell(self, x, y, z):
        k = (int(math.floor(x / self.res)) - self.k0[0], int(math.floor(y / self.res)) - self.k0[1],
             int(math.floor(z / self.res)) - self.k0[2])
        return k

    def get_distance(self, x, y, z):
        """DynamicEDTOctomap::getDistance(point3d): float metres, -1 outside the map."""
        x, y, z = np.float32(x), np.float32(y), np.float32(z)       # octomap::point3d is float32
        k = self.cell(float(x), float(y), float(z))
        for a in range(3):
            if k[a] < 0 or k[a] >= self.shape[a]:
                return np.float32(-1.0)
        return self.edt[k]

    def min_distance_lattice(self, xs, ys, zs):
        """min over the lattice xs x ys x zs of get_distance (vectorised)."""
        idx = []
        outside = False
        for a, vals in enumerate((xs, ys, zs)):
            v = np.asarray(vals, np.float32).astype(np.float64)
            k = np.floor(v / self.res).astype(int) - self.k0[a]
            ok = (k >= 0) & (k < self.shape[a])
            if not ok.all():
                outside = True
                k = k[ok]
            idx.append(k)
        if outside:
            return -1.0
        if any(len(k) == 0 for k in idx):
            return 1e9
        return float(self.edt[np.ix_(*idx)].min())


# ---------------------------------------------------------------------------------------------------------------------
# Corridor: SFC (rbp_corridor.hpp L44-L243)
# ---------------------------------------------------------------------------------------------------------------------
class Corridor:
    def __init__(self, world, param, wxy):
        self.w = world
        self.p = param
        self.wmin = (wxy[0], wxy[1], param["world_z_min"])
        self.wmax = (wxy[2], wxy[3], param["world_z_max"])

    def _axis_samples(self, lo, hi, res, wmin):
        out, i, count = [], lo, 0
        while i < hi + SP_EPSILON_FLOAT:                              # double accumulation as L47-L51
            v = i + SP_EPSILON_FLOAT
            if count == 0 and lo > wmin + SP_EPSILON_FLOAT:
                v = lo - SP_EPSILON_FLOAT
            out.append(v)
            i += res
            count += 1
        return out

    def is_obstacle_in_box(self, box, margin):                       # L44-L78
        p = self.p
        xs = self._axis_samples(box[0], box[3], p["box_xy_res"], self.wmin[0])
        ys = self._axis_samples(box[1], box[4], p["box_xy_res"], self.wmin[1])
        zs = self._axis_samples(box[2], box[5], p["box_z_res"], self.wmin[2])
        return self.w.min_distance_lattice(xs, ys, zs) < margin - SP_EPSILON_FLOAT

    def is_box_in_boundary(self, b):                                  # L80-L87
        return (b[0] > self.wmin[0] - SP_EPSILON and b[1] > self.wmin[1] - SP_EPSILON and b[2] > self.wmin[2] - SP_EPSILON
                and b[3] < self.wmax[0] + SP_EPSILON and b[4] < self.wmax[1] + SP_EPSILON and b[5] < self.wmax[2] + SP_EPSILON)

    @staticmethod
    def is_point_in_box(pt, b):                                       # L89-L97
        return (pt[0] > b[0] - SP_EPSILON and pt[1] > b[1] - SP_EPSILON and pt[2] > b[2] - SP_EPSILON
                and pt[0] < b[3] + SP_EPSILON and pt[1] < b[4] + SP_EPSILON and pt[2] < b[5] + SP_EPSILON)

    def expand_box(self, box, margin):                                # L99-L147
        p = self.p
        axis_cand = [0, 1, 2, 3, 4, 5]
        i = -1
        while axis_cand:
            box_cand = list(box)
            box_update = list(box)
            while (not self.is_obstacle_in_box(box_update, margin)) and self.is_box_in_boundary(box_update):
                i += 1
                if i >= len(axis_cand):
                    i = 0
                axis = axis_cand[i]
                box = list(box_cand)
                box_update = list(box_cand)
                if axis < 3:
                    box_update[axis + 3] = box_cand[axis]
                    box_cand[axis] = box_cand[axis] - (p["box_z_res"] if axis == 2 else p["box_xy_res"])
                    box_update[axis] = box_cand[axis]
                else:
                    box_update[axis - 3] = box_cand[axis]
                    box_cand[axis] = box_cand[axis] + (p["box_z_res"] if axis == 5 else p["box_xy_res"])
                    box_update[axis] = box_cand[axis]
            del axis_cand[i]
            if i > 0:
                i -= 1
            else:
                i = len(axis_cand) - 1
        return box

    def sfc_for_agent(self, traj, T, radius):                         # updateObsBox L149-L243, one agent
        p = self.p
        boxes = []
        box_prev = [0.0] * 6
        rxy, rz = p["box_xy_res"], p["box_z_res"]
        for i in range(len(traj) - 1):
            x, y, z = (float(v) for v in traj[i])
            xn, yn, zn = (float(v) for v in traj[i + 1])
            if self.is_point_in_box((xn, yn, zn), box_prev):
                continue
            box = [round(min(x, xn) / rxy) * rxy, round(min(y, yn) / rxy) * rxy, round(min(z, zn) / rz) * rz,
                   round(max(x, xn) / rxy) * rxy, round(max(y, yn) / rxy) * rxy, round(max(z, zn) / rz) * rz]
            if self.is_obstacle_in_box(box, radius):
                return None                                           # "Invalid initial trajectory"
            box = self.expand_box(box, radius)
            boxes.append(box)
            box_prev = box
        box_max, path_max = len(boxes), len(traj)
        t_end = [-1.0] * box_max
        log = np.zeros((box_max, path_max))
        for i in range(box_max):
            for j in range(path_max):
                if self.is_point_in_box([float(v) for v in traj[j]], boxes[i]):
                    log[i, j] = 1 if j == 0 else log[i, j - 1] + 1
        box_iter, path_iter = 0, 0
        while path_iter < path_max:                                   # the for-loop of L209-L235 incl. its ++ / --
            if box_iter == box_max - 1:
                if log[box_iter, path_iter] > 0:
                    path_iter += 1
                    continue
                box_iter -= 1
            if log[box_iter, path_iter] > 0 and log[box_iter + 1, path_iter] > 0:
                count = 1
                while (path_iter + count < path_max and log[box_iter, path_iter + count] > 0
                       and log[box_iter + 1, path_iter + count] > 0):
                    count += 1
                t_end[box_iter] = T[path_iter + count // 2]
                path_iter = path_iter + count // 2
                box_iter += 1
            elif log[box_iter, path_iter] == 0:
                box_iter -= 1
                path_iter -= 1
            if box_iter < 0 or path_iter < -1:
                return None                                           # index -1 in the reference (undefined behaviour)
            path_iter += 1
        t_end[box_max - 1] = T[-1]
        return np.array(boxes, float), np.array(t_end, float)


# ---------------------------------------------------------------------------------------------------------------------
# Corridor: RSFC (rbp_corridor.hpp L338-L398), float32 with octomap::Vector3 operator semantics
# ---------------------------------------------------------------------------------------------------------------------
def _norm(v):      # Vector3::norm(): sqrt of the float32 sum of squares, as double
    f = np.float32
    s = f(f(f(v[..., 0] * v[..., 0]) + f(v[..., 1] * v[..., 1])) + f(v[..., 2] * v[..., 2]))
    return np.sqrt(s.astype(np.float64))


def _dot(a, b):    # Vector3::dot(): float32 products and sums, returned as double
    f = np.float32
    return f(f(f(a[..., 0] * b[..., 0]) + f(a[..., 1] * b[..., 1])) + f(a[..., 2] * b[..., 2])).astype(np.float64)


def _normalize(v):  # len = norm(); if (len > 0) *this /= (float)len
    ln = _norm(v)
    d = np.where(ln > 0, ln, 1.0).astype(np.float32)
    return (v / d[..., None]).astype(np.float32)


def rsfc_from_init_traj(init_traj, T, downwash):
    """init_traj [N, M+1, 3] float32 -> (rsfc_n [P, M, 3] float32, rsfc_t [P, M], ok)."""
    init_traj = np.asarray(init_traj, np.float32)
    N, M1, _ = init_traj.shape
    M = M1 - 1
    qi, qj = np.triu_indices(N, 1)                                    # lexicographic qi < qj
    rel = (init_traj[qj] - init_traj[qi]).astype(np.float32)          # [P, M+1, 3]
    rel[..., 2] = (rel[..., 2].astype(np.float64) / downwash).astype(np.float32)
    a, b = rel[:, :-1], rel[:, 1:]
    same = np.all(a == b, axis=-1)
    m = a.copy()
    dist_min = _norm(a)
    dist = _norm(b)
    pick_b = dist_min > dist
    m = np.where(pick_b[..., None], b, m)
    dist_min = np.where(pick_b, dist, dist_min)
    n = _normalize((b - a).astype(np.float32))
    c = (a - (n * _dot(a, n).astype(np.float32)[..., None]).astype(np.float32)).astype(np.float32)
    dist = _norm(c)
    pick_c = (_dot((c - a).astype(np.float32), (c - b).astype(np.float32)) < 0) & (dist_min > dist)
    m = np.where(pick_c[..., None], c, m)
    m = np.where(same[..., None], a, m).astype(np.float32)
    m = _normalize(m)
    m[..., 2] = (m[..., 2].astype(np.float64) / downwash).astype(np.float32)
    ok = bool(np.all(_norm(m) != 0))
    rsfc_t = np.tile(np.asarray(T, float)[1:], (len(qi), 1))
    return np.ascontiguousarray(m, np.float32), rsfc_t, ok


# ---------------------------------------------------------------------------------------------------------------------
# discrete initial trajectories: prioritised space-time search (stand-in for ECBS)
# ---------------------------------------------------------------------------------------------------------------------
_MOVES = [(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]  # environment.hpp L467-L524


def _seg_min_dist_to_origin(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = b - a
    dd = d @ d
    if dd == 0:
        return math.sqrt(a @ a)
    t = min(1.0, max(0.0, -(a @ d) / dd))
    p = a + t * d
    return math.sqrt(p @ p)


def _edge_conflict(s1a, s1b, s2a, s2b, rsum, grid):                   # environment.hpp L664-L681
    if rsum < grid * 0.5:
        return s1a == s2b and s1b == s2a
    a = (s2a[0] - s1a[0], s2a[1] - s1a[1], s2a[2] - s1a[2])
    b = (s2b[0] - s1b[0], s2b[1] - s1b[1], s2b[2] - s1b[2])
    return _seg_min_dist_to_origin(a, b) * grid <= rsum


def _vertex_conflict(s1, s2, rsum, grid):                             # environment.hpp L656-L663
    if rsum < grid:
        return s1 == s2
    return math.dist(s1, s2) * grid < rsum


def plan_paths(starts, wants, free, horizon, radius, grid, rng, goal_gap=1):
    """Prioritised layered space-time search.  Agent ai moves for `horizon` steps (waits allowed) without vertex / edge
    conflicts with the agents planned before it, and parks on a goal cell wants[ai] steps (Manhattan) from its start
    when one is reachable, else on the farthest reachable one.  Returns ([N][horizon+1] cells, goals) or None."""
    N = len(starts)
    paths, goals = [], []
    dims = free.shape
    rsum = 2 * radius
    for ai in range(N):
        layers = [{starts[ai]: None}]
        for t in range(horizon):
            nxt = {}
            for s in layers[-1]:
                for mv in _MOVES:
                    n = (s[0] + mv[0], s[1] + mv[1], s[2] + mv[2])
                    if n in nxt:
                        continue
                    if not (0 <= n[0] < dims[0] and 0 <= n[1] < dims[1] and 0 <= n[2] < dims[2]) or not free[n]:
                        continue
                    bad = False
                    for pj in paths:
                        if _vertex_conflict(n, pj[t + 1], rsum, grid) or _edge_conflict(s, n, pj[t], pj[t + 1], rsum, grid):
                            bad = True
                            break
                    if not bad:
                        nxt[n] = s
            if not nxt:
                return None
            layers.append(nxt)
        s0 = starts[ai]
        cand = [c for c in layers[-1] if c[2] == s0[2]
                and all(max(abs(c[0] - g[0]), abs(c[1] - g[1])) >= goal_gap for g in goals)
                and all(c != st for k, st in enumerate(starts) if k > ai)]
        if not cand:
            return None
        dist = [abs(c[0] - s0[0]) + abs(c[1] - s0[1]) for c in cand]
        best = [c for c, d in zip(cand, dist) if d == wants[ai]]
        if not best:
            if ai == 0:
                return None                                           # agent 0 pins the makespan
            dmax = max(dist)
            best = [c for c, d in zip(cand, dist) if d == dmax]
        best.sort()
        goal = best[int(rng.integers(len(best)))]
        path = [goal]
        for t in range(horizon, 0, -1):
            path.append(layers[t][path[-1]])
        paths.append(path[::-1])
        goals.append(goal)
    return paths, goals


# ---------------------------------------------------------------------------------------------------------------------
# mission generator
# ---------------------------------------------------------------------------------------------------------------------
def synth_mission(N, M, rho, seed, param=None, max_tries=50):
    """One seeded mission of the fixed-M family. Returns a dict of numpy arrays."""
    p = dict(DEFAULT_PARAM)
    if param:
        p.update(param)
    if M < 3:
        raise ValueError("M must be >= 3 (makespan + 2 with makespan >= 1)")
    L = 10.0 * math.sqrt(max(N, 64) / 64.0)
    half = round(L / 2 / p["grid_xy_res"]) * p["grid_xy_res"]
    wxy = (-half, -half, half, half)
    for attempt in range(max_tries):
        rng = np.random.Generator(np.random.MT19937(seed * 1000 + attempt))
        m = _try_mission(N, M, rho, rng, p, wxy)
        if m is not None:
            m["seed"] = seed
            m["attempt"] = attempt
            return m
    raise RuntimeError("synth_mission: no valid mission after %d tries (N=%d M=%d rho=%g)" % (max_tries, N, M, rho))


def _try_mission(N, M, rho, rng, p, wxy):
    res = p["world_resolution"]
    world = World(wxy[0], wxy[1], p["world_z_min"], wxy[2], wxy[3], p["world_z_max"], res)
    area = (wxy[2] - wxy[0]) * (wxy[3] - wxy[1])
    for _ in range(int(round(rho * area))):
        cx, cy = rng.uniform(wxy[0], wxy[2]), rng.uniform(wxy[1], wxy[3])
        world.add_pillar(cx, cy, rng.uniform(p["obs_h_min"], p["obs_h_max"], (3, 3)))
    world.update(1.0)
    # planning grid (init_traj_planner.hpp L19-L30) and its obstacles (ecbs_planner.hpp L80-L109)
    gxy, gz = p["grid_xy_res"], p["grid_z_res"]
    gmin = [math.ceil((wxy[0] - SP_EPSILON) / gxy) * gxy, math.ceil((wxy[1] - SP_EPSILON) / gxy) * gxy,
            math.ceil((p["world_z_min"] - SP_EPSILON) / gz) * gz]
    gmax = [math.floor((wxy[2] + SP_EPSILON) / gxy) * gxy, math.floor((wxy[3] + SP_EPSILON) / gxy) * gxy,
            math.floor((p["world_z_max"] + SP_EPSILON) / gz) * gz]
    dims = (int(round((gmax[0] - gmin[0]) / gxy)) + 1, int(round((gmax[1] - gmin[1]) / gxy)) + 1,
            int(round((gmax[2] - gmin[2]) / gz)) + 1)
    r = p["radius"]
    free = np.zeros(dims, bool)
    for ix in range(dims[0]):
        for iy in range(dims[1]):
            for iz in range(dims[2]):
                d = world.get_distance(gmin[0] + ix * gxy, gmin[1] + iy * gxy, gmin[2] + iz * gz)
                free[ix, iy, iz] = not (d < r + p["grid_margin"])
    horizon = M - 2
    # starts: distinct free cells on the lowest grid layer, >= 2 cells apart while the world has room for that
    # (every other cell of the grid), otherwise any distinct free cells; goals: distinct free cells within reach
    def far(c, others, d):
        return all(max(abs(c[0] - o[0]), abs(c[1] - o[1])) >= d for o in others)

    lattice = [(ix, iy, 0) for ix in range(0, dims[0], 2) for iy in range(0, dims[1], 2) if free[ix, iy, 0]]
    cells = [(ix, iy, 0) for ix in range(dims[0]) for iy in range(dims[1]) if free[ix, iy, 0]]
    pool = lattice if len(lattice) >= N else cells
    if len(pool) < N:
        return None
    starts = [pool[k] for k in rng.permutation(len(pool))[:N]]
    wants = [horizon] + [int(rng.integers(1, horizon + 1)) for _ in range(N - 1)]   # agent 0 pins the makespan to M-2
    out = plan_paths(starts, wants, free, horizon, r, gxy, rng)
    if out is None:
        return None
    paths, goals = out
    T = np.arange(M + 1, dtype=float) * p["time_step"]

    def pos(c):
        return (c[0] * gxy + gmin[0], c[1] * gxy + gmin[1], c[2] * gz + gmin[2])

    start = np.zeros((N, 9))
    goal = np.zeros((N, 9))
    init_traj = np.zeros((N, M + 1, 3), np.float32)
    for ai in range(N):
        start[ai, :3] = pos(starts[ai])
        goal[ai, :3] = pos(goals[ai])
        pts = [start[ai, :3]] + [pos(c) for c in paths[ai]] + [goal[ai, :3]]
        init_traj[ai] = np.asarray(pts, np.float32)                    # octomap::point3d
    cor = Corridor(world, p, wxy)
    sfc = []
    for ai in range(N):
        out = cor.sfc_for_agent(init_traj[ai], T, r)
        if out is None:
            return None
        sfc.append(out)
    rsfc_n, rsfc_t, ok = rsfc_from_init_traj(init_traj, T, p["downwash"])
    if not ok:
        return None
    return dict(N=N, M=M, T=T, start=start, goal=goal, radius=np.full(N, r), max_vel=np.full((N, 3), p["max_vel"]),
                max_acc=np.full((N, 3), p["max_acc"]), sfc=sfc, rsfc_n=rsfc_n, rsfc_t=rsfc_t, init_traj=init_traj,
                world_xy=np.array(wxy), n_pillars=int(round(rho * area)), downwash=p["downwash"],
                edt=world.edt, edt_k0=world.k0.copy(), world_z=(p["world_z_min"], p["world_z_max"]), resolution=res)


def pack(missions):
    """Concatenate missions of identical (N, M) into the flat arrays of rbpe_problem (include/rbpe.h)."""
    N, M = missions[0]["N"], missions[0]["M"]
    assert all(m["N"] == N and m["M"] == M for m in missions)
    offs, base, boxes, tend = [], [0], [], []
    for m in missions:
        o = [0]
        for b, t in m["sfc"]:
            boxes.append(b)
            tend.append(t)
            o.append(o[-1] + len(t))
        offs.append(o)
        base.append(base[-1] + o[-1])
    return dict(
        N=N, M=M, count=len(missions),
        T=np.ascontiguousarray([m["T"] for m in missions], np.float64),
        start=np.ascontiguousarray([m["start"] for m in missions], np.float64),
        goal=np.ascontiguousarray([m["goal"] for m in missions], np.float64),
        radius=np.ascontiguousarray([m["radius"] for m in missions], np.float64),
        sfc_offs=np.ascontiguousarray(offs, np.int32), sfc_base=np.ascontiguousarray(base, np.int32),
        sfc_box=np.ascontiguousarray(np.concatenate(boxes), np.float64),
        sfc_t=np.ascontiguousarray(np.concatenate(tend), np.float64),
        rsfc_n=np.ascontiguousarray([m["rsfc_n"] for m in missions], np.float32),
        rsfc_t=np.ascontiguousarray([m["rsfc_t"] for m in missions], np.float64),
        init_traj=np.ascontiguousarray([m["init_traj"] for m in missions], np.float32),
    )


def dump_text(m, path):
    """Text dump of one mission for swarm_simulator_b200/host/planner_cli (format documented in planner_cli.cpp)."""
    N, M = m["N"], m["M"]
    with open(path, "w") as f:
        f.write("%d %d\n" % (N, M))
        f.write(" ".join(repr(float(t)) for t in m["T"]) + "\n")
        for qi in range(N):
            vals = list(m["start"][qi]) + list(m["goal"][qi]) + [m["radius"][qi]] + list(m["max_vel"][qi]) + list(m["max_acc"][qi])
            f.write(" ".join(repr(float(v)) for v in vals) + "\n")
        for qi in range(N):
            for j in range(M + 1):
                f.write(" ".join(repr(float(v)) for v in m["init_traj"][qi, j]) + "\n")
        for qi in range(N):
            boxes, tend = m["sfc"][qi]
            f.write("%d\n" % len(tend))
            for b, t in zip(boxes, tend):
                f.write(" ".join(repr(float(v)) for v in list(b) + [t]) + "\n")
        for p in range(N * (N - 1) // 2):
            for ri in range(M):
                f.write(" ".join(repr(float(v)) for v in list(m["rsfc_n"][p, ri]) + [m["rsfc_t"][p, ri]]) + "\n")


def dump_world_text(m, path):
    """Text dump of world + mission + initTraj for swarm_simulator_b200/host/corridor_cli (format in corridor_cli.cpp)."""
    edt = m["edt"]
    with open(path, "w") as f:
        f.write("%r %d %d %d %d %d %d\n" % (float(m["resolution"]), m["edt_k0"][0], m["edt_k0"][1], m["edt_k0"][2], *edt.shape))
        f.write(" ".join("%.9g" % v for v in edt.reshape(-1)) + "\n")
        wxy = m["world_xy"]
        f.write("%r %r %r %r %r %r\n" % (float(wxy[0]), float(wxy[1]), float(m["world_z"][0]), float(wxy[2]), float(wxy[3]), float(m["world_z"][1])))
        f.write("%d %d\n" % (m["N"], m["M"]))
        f.write(" ".join(repr(float(t)) for t in m["T"]) + "\n")
        f.write(" ".join(repr(float(r)) for r in m["radius"]) + "\n")
        for qi in range(m["N"]):
            for j in range(m["M"] + 1):
                f.write(" ".join(repr(float(v)) for v in m["init_traj"][qi, j]) + "\n")


# ---------------------------------------------------------------------------------------------------------------------
# mission packs: generated missions stored compactly (tests/golden/missions_*.npz, written by
# tests/golden/make_missions.py), so that tests and bench.py get hundreds of distinct seeded missions without paying
# ~1.5 s of Python path planning per 64-agent mission.  RSFC is not stored: it is a pure function of init_traj
# (rsfc_from_init_traj); everything else is stored in the precision the generator produced (float64 boxes carry the
# reference's `i += res` rounding noise, init_traj is float32 like octomap::point3d).
# ---------------------------------------------------------------------------------------------------------------------
PACK_FIELDS = ("N", "M", "rho", "seed", "T", "start", "goal", "radius", "init_traj", "sfc_offs", "sfc_box", "sfc_t", "downwash")


def save_pack(missions, rhos, path):
    N, M = missions[0]["N"], missions[0]["M"]
    assert all(m["N"] == N and m["M"] == M for m in missions)
    offs, boxes, tend = [], [], []
    for m in missions:
        o = [0]
        for b, t in m["sfc"]:
            boxes.append(np.asarray(b, np.float64)); tend.append(np.asarray(t, np.float64)); o.append(o[-1] + len(t))
        offs.append(o)
        assert not np.any(m["start"][:, 3:]) and not np.any(m["goal"][:, 3:])
    np.savez_compressed(
        path, N=N, M=M, rho=np.asarray(rhos, np.float64), seed=np.asarray([m["seed"] for m in missions], np.int64),
        T=np.asarray([m["T"] for m in missions], np.float64),
        start=np.asarray([m["start"][:, :3] for m in missions], np.float64),
        goal=np.asarray([m["goal"][:, :3] for m in missions], np.float64),
        radius=np.asarray([m["radius"] for m in missions], np.float64),
        init_traj=np.asarray([m["init_traj"] for m in missions], np.float32),
        sfc_offs=np.asarray(offs, np.int32), sfc_box=np.concatenate(boxes), sfc_t=np.concatenate(tend),
        downwash=np.asarray([m["downwash"] for m in missions], np.float64))


def load_pack(path, select=None):
    """Missions of a pack as the dicts synth_mission returns (without the world / distance map). `select`: indices."""
    z = np.load(path)
    N, M = int(z["N"]), int(z["M"])
    offs = z["sfc_offs"]
    base = np.concatenate([[0], np.cumsum(offs[:, -1])])
    box, tend = z["sfc_box"], z["sfc_t"]
    idx = range(len(z["seed"])) if select is None else select
    out = []
    for c in idx:
        start = np.zeros((N, 9)); goal = np.zeros((N, 9))
        start[:, :3] = z["start"][c]; goal[:, :3] = z["goal"][c]
        sfc = []
        for qi in range(N):
            lo, hi = base[c] + offs[c, qi], base[c] + offs[c, qi + 1]
            sfc.append((box[lo:hi].copy(), tend[lo:hi].copy()))
        T = z["T"][c].copy()
        dw = float(z["downwash"][c])
        rsfc_n, rsfc_t, ok = rsfc_from_init_traj(z["init_traj"][c], T, dw)
        assert ok
        out.append(dict(N=N, M=M, T=T, start=start, goal=goal, radius=z["radius"][c].copy(),
                        max_vel=np.full((N, 3), DEFAULT_PARAM["max_vel"]), max_acc=np.full((N, 3), DEFAULT_PARAM["max_acc"]),
                        sfc=sfc, rsfc_n=rsfc_n, rsfc_t=rsfc_t, init_traj=z["init_traj"][c].copy(), downwash=dw,
                        seed=int(z["seed"][c]), rho=float(z["rho"][c])))
    return out
