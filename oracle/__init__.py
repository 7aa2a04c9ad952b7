"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  See ``oracle/rbp_oracle.h`` for what it restates
(/root/reference/swarm_planner/include/rbp_planner.hpp L33-L206, L327-L700, L849-L881).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OK, INFEASIBLE, NOT_CONVERGED, BAD_ARG = 0, 1, 2, 3


def build(force=False):
    """Compile liboracle.so with the system gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "rbp_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src),
                                                   os.path.getmtime(os.path.join(_HERE, "rbp_oracle.h")))):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "CC=gcc"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return _LIB_PATH


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


class CProblem(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("sequential", C.c_int), ("batch_size", C.c_int),
                ("batch_iter", C.c_int), ("iteration", C.c_int),
                ("T", _dp), ("start", _dp), ("goal", _dp), ("radius", _dp),
                ("sfc_offs", _ip), ("sfc_box", _dp), ("sfc_t", _dp),
                ("rsfc_n", _fp), ("rsfc_t", _dp), ("init_traj", _fp)]


class CQP(C.Structure):
    _fields_ = [("nv", C.c_int), ("ne", C.c_int), ("mi", C.c_int),
                ("qnnz", C.c_int), ("qi", _ip), ("qj", _ip), ("qv", _dp),
                ("a_ptr", _ip), ("a_idx", _ip), ("a_val", _dp), ("b", _dp),
                ("g_ptr", _ip), ("g_idx", _ip), ("g_val", _dp), ("h", _dp),
                ("perm_x", _ip), ("perm_y", _ip),
                ("n_box_rows", C.c_int), ("n_rsfc_rows", C.c_int)]


class COpts(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("tol_gap", C.c_double), ("tol_res", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_build_Q_base.argtypes = [_dp, _dp]
        L.oracle_build_Aeq_base.argtypes = [_dp, C.c_int, _dp]
        L.oracle_build_deq.argtypes = [C.POINTER(CProblem), _dp]
        L.oracle_build_dlq.argtypes = [C.POINTER(CProblem), _dp, _dp, _dp]
        L.oracle_build_dlq.restype = C.c_int
        L.oracle_build_dummy.argtypes = [C.POINTER(CProblem), _dp]
        L.oracle_set_batch.argtypes = [C.POINTER(CProblem), _ip, _ip]
        L.oracle_set_batch.restype = C.c_int
        L.oracle_populate.argtypes = [C.POINTER(CProblem), _dp, C.c_int]
        L.oracle_populate.restype = C.POINTER(CQP)
        L.oracle_qp_free.argtypes = [C.POINTER(CQP)]
        L.oracle_solve_qp.argtypes = [C.POINTER(CQP), C.POINTER(COpts), _dp, _dp, _ip, _dp]
        L.oracle_solve_qp.restype = C.c_int
        L.oracle_update.argtypes = [C.POINTER(CProblem), C.POINTER(COpts), _dp, _dp, _dp, _ip, _ip, C.c_int]
        L.oracle_update.restype = C.c_int
        L.oracle_update_many.argtypes = [C.POINTER(CProblem), C.c_int, C.POINTER(COpts),
                                         C.POINTER(_dp), C.POINTER(_dp), _ip, C.c_int]
        L.oracle_update_many.restype = C.c_int
        L.oracle_rsfc.argtypes = [C.c_int, C.c_int, _fp, _dp, C.c_double, _fp, _dp]
        L.oracle_rsfc.restype = C.c_int
        L.oracle_safety_metrics.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
        L.oracle_safety_metrics.restype = C.c_int
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f(a):
    return a.ctypes.data_as(_fp)


class Problem:
    """Owns contiguous numpy copies of one mission and the matching C struct.

    Arguments mirror PlanResult / Mission / Param of the reference (sp_const.hpp L16-L28,
    mission.hpp L10-L19, param.hpp L8-L42):
      T[M+1], start[N,9], goal[N,9], radius[N], sfc = list (per agent) of (box[nb,6], t_end[nb]),
      rsfc_n[P,M,3] float32 + rsfc_t[P,M] (pairs qi<qj lexicographic), init_traj[N,M+1,3] float32.
    """

    def __init__(self, T, start, goal, radius, sfc_offs, sfc_box, sfc_t, rsfc_n, rsfc_t, init_traj,
                 sequential=True, batch_size=4, batch_iter=-1, iteration=1):
        self.T = np.ascontiguousarray(T, np.float64)
        self.start = np.ascontiguousarray(start, np.float64)
        self.goal = np.ascontiguousarray(goal, np.float64)
        self.radius = np.ascontiguousarray(radius, np.float64)
        self.sfc_offs = np.ascontiguousarray(sfc_offs, np.int32)
        self.sfc_box = np.ascontiguousarray(sfc_box, np.float64)
        self.sfc_t = np.ascontiguousarray(sfc_t, np.float64)
        self.rsfc_n = np.ascontiguousarray(rsfc_n, np.float32)
        self.rsfc_t = np.ascontiguousarray(rsfc_t, np.float64)
        self.init_traj = np.ascontiguousarray(init_traj, np.float32)
        self.N = int(self.start.shape[0])
        self.M = int(self.T.shape[0] - 1)
        P = self.N * (self.N - 1) // 2
        assert self.start.shape == (self.N, 9) and self.goal.shape == (self.N, 9)
        assert self.sfc_offs.shape == (self.N + 1,)
        assert self.rsfc_n.shape == (P, self.M, 3), (self.rsfc_n.shape, (P, self.M, 3))
        assert self.rsfc_t.shape == (P, self.M)
        assert self.init_traj.shape == (self.N, self.M + 1, 3)
        self.c = CProblem(self.N, self.M, int(bool(sequential)), int(batch_size), int(batch_iter),
                          int(iteration), _d(self.T), _d(self.start), _d(self.goal), _d(self.radius),
                          _i(self.sfc_offs), _d(self.sfc_box), _d(self.sfc_t), _f(self.rsfc_n),
                          _d(self.rsfc_t), _f(self.init_traj))

    # -- pieces of buildConstMtx ------------------------------------------------------------------
    def deq(self):
        out = np.zeros((self.N * (3 * self.M + 3), 3))
        lib().oracle_build_deq(C.byref(self.c), _d(out))
        return out

    def dlq(self):
        P = self.N * (self.N - 1) // 2
        ub = np.zeros((self.N, 6 * self.M, 3)); lbn = np.zeros_like(ub)
        rel = np.zeros((max(P, 1), 6 * self.M, 3))
        rc = lib().oracle_build_dlq(C.byref(self.c), _d(ub), _d(lbn), _d(rel))
        return rc, ub, lbn, rel[:P]

    def dummy(self):
        out = np.zeros((self.N * 6 * self.M, 3))
        lib().oracle_build_dummy(C.byref(self.c), _d(out))
        return out

    def set_batch(self):
        bs, bi = C.c_int(), C.c_int()
        n = lib().oracle_set_batch(C.byref(self.c), C.byref(bs), C.byref(bi))
        return n, bs.value, bi.value

    def populate(self, dummy, l):
        d = np.ascontiguousarray(dummy, np.float64) if dummy is not None else None
        q = lib().oracle_populate(C.byref(self.c), _d(d) if d is not None else None, int(l))
        if not q:
            raise ValueError("empty batch %d" % l)
        return QP(q, owned=True)

    def update(self, max_iter=0, tol_gap=0.0, tol_res=0.0):
        """RBPPlanner::update() minus timeScale.  Returns dict(status, coef[N,3,6M], ctrl[N,3,6M], ...)."""
        _, bs, bit = self.set_batch()
        nrec = max(1, self.c.iteration * max(bit, 1))
        coef = np.zeros((self.N, 3, 6 * self.M)); ctrl = np.zeros_like(coef)
        obj = np.zeros(nrec); its = np.zeros(nrec, np.int32); st = np.full(nrec, -1, np.int32)
        o = COpts(int(max_iter), float(tol_gap), float(tol_res))
        rc = lib().oracle_update(C.byref(self.c), C.byref(o), _d(coef), _d(ctrl), _d(obj), _i(its), _i(st), 1)
        return dict(status=rc, coef=coef, ctrl=ctrl, batch_obj=obj, batch_iters=its, batch_status=st)


class QP:
    """View of an oracle_qp (either produced by oracle_populate or built from numpy arrays)."""

    def __init__(self, ptr=None, owned=False, **arrs):
        self._owned = owned
        self._keep = arrs
        if ptr is not None:
            self.p = ptr
        else:
            a = {k: np.ascontiguousarray(v) for k, v in arrs.items() if k not in ("nv", "ne", "mi")}
            self._keep = a
            s = CQP()
            s.nv, s.ne, s.mi = int(arrs["nv"]), int(arrs["ne"]), int(arrs["mi"])
            s.qnnz = len(a["qv"])
            s.qi, s.qj, s.qv = _i(a["qi"]), _i(a["qj"]), _d(a["qv"])
            s.a_ptr, s.a_idx, s.a_val, s.b = _i(a["a_ptr"]), _i(a["a_idx"]), _d(a["a_val"]), _d(a["b"])
            s.g_ptr, s.g_idx, s.g_val, s.h = _i(a["g_ptr"]), _i(a["g_idx"]), _d(a["g_val"]), _d(a["h"])
            s.perm_x, s.perm_y = _i(a["perm_x"]), _i(a["perm_y"])
            self._s = s
            self.p = C.pointer(s)

    def __del__(self):
        if getattr(self, "_owned", False) and self.p:
            lib().oracle_qp_free(self.p)
            self.p = None

    @property
    def nv(self): return self.p.contents.nv
    @property
    def ne(self): return self.p.contents.ne
    @property
    def mi(self): return self.p.contents.mi
    @property
    def n_box_rows(self): return self.p.contents.n_box_rows
    @property
    def n_rsfc_rows(self): return self.p.contents.n_rsfc_rows

    def _arr(self, ptr, n, dt):
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)

    def dense(self):
        """(Q dense [nv,nv], A dense, b, G dense, h) -- small cases only."""
        c = self.p.contents
        Q = np.zeros((c.nv, c.nv))
        qi, qj, qv = self._arr(c.qi, c.qnnz, np.int64), self._arr(c.qj, c.qnnz, np.int64), self._arr(c.qv, c.qnnz, float)
        np.add.at(Q, (qi, qj), qv)
        A, b = self.csr("a")
        G, h = self.csr("g")
        return Q, A, b, G, h

    def csr_raw(self, which):
        c = self.p.contents
        m = c.ne if which == "a" else c.mi
        ptr = self._arr(getattr(c, which + "_ptr"), m + 1, np.int64)
        idx = self._arr(getattr(c, which + "_idx"), int(ptr[-1]), np.int64)
        val = self._arr(getattr(c, which + "_val"), int(ptr[-1]), float)
        rhs = self._arr(c.b if which == "a" else c.h, m, float)
        return ptr, idx, val, rhs

    def csr(self, which):
        ptr, idx, val, rhs = self.csr_raw(which)
        m = len(rhs)
        D = np.zeros((m, self.nv))
        rows = np.repeat(np.arange(m), np.diff(ptr))
        np.add.at(D, (rows, idx), val)
        return D, rhs

    def solve(self, max_iter=0, tol_gap=0.0, tol_res=0.0):
        x = np.zeros(self.nv); obj = C.c_double(); it = C.c_int(); res = np.zeros(4)
        o = COpts(int(max_iter), float(tol_gap), float(tol_res))
        st = lib().oracle_solve_qp(self.p, C.byref(o), _d(x), C.byref(obj), C.byref(it), _d(res))
        return dict(status=st, x=x, obj=obj.value, iters=it.value, res=res)


def Q_base_and_basis():
    Q = np.zeros((6, 6)); B = np.zeros((6, 6))
    lib().oracle_build_Q_base(_d(Q), _d(B))
    return Q, B


def Aeq_base(T):
    T = np.ascontiguousarray(T, np.float64)
    M = len(T) - 1
    A = np.zeros((3 * M + 3, 6 * M))
    lib().oracle_build_Aeq_base(_d(T), M, _d(A))
    return A


def update_many(problems, nthreads=0, max_iter=0, tol_gap=0.0, tol_res=0.0):
    """OpenMP over independent missions (CPU-baseline helper). Returns (coef list, ctrl list, status)."""
    n = len(problems)
    arr = (CProblem * n)(*[p.c for p in problems])
    coef = [np.zeros((p.N, 3, 6 * p.M)) for p in problems]
    ctrl = [np.zeros((p.N, 3, 6 * p.M)) for p in problems]
    cp = (_dp * n)(*[_d(a) for a in coef]); tp = (_dp * n)(*[_d(a) for a in ctrl])
    st = np.zeros(n, np.int32)
    o = COpts(int(max_iter), float(tol_gap), float(tol_res))
    lib().oracle_update_many(arr, n, C.byref(o), cp, tp, _i(st), int(nthreads))
    return coef, ctrl, st


def rsfc(init_traj, T, downwash):
    """Corridor::updateRelBox restatement (float32). Returns (rsfc_n[P,M,3] f32, rsfc_t[P,M], collided)."""
    tr = np.ascontiguousarray(init_traj, np.float32)
    T = np.ascontiguousarray(T, np.float64)
    N, M = tr.shape[0], tr.shape[1] - 1
    P = N * (N - 1) // 2
    n = np.zeros((max(P, 1), M, 3), np.float32); t = np.zeros((max(P, 1), M))
    rc = lib().oracle_rsfc(N, M, _f(tr), _d(T), float(downwash), _f(n), _d(t))
    return n[:P], t[:P], bool(rc)


def safety_metrics(coef, T, radius, downwash=2.0, dt=0.1):
    """(safety_margin_ratio, time of the minimum, total flight length, samples) for coef[N,3,6M] (highest power first)."""
    c = np.ascontiguousarray(coef, np.float64); T = np.ascontiguousarray(T, np.float64)
    r = np.ascontiguousarray(radius, np.float64)
    N, M = c.shape[0], len(T) - 1
    a, b, l = C.c_double(), C.c_double(), C.c_double()
    nt = lib().oracle_safety_metrics(N, M, _d(c), _d(T), _d(r), float(downwash), float(dt), C.byref(a), C.byref(b), C.byref(l))
    return a.value, b.value, l.value, nt
