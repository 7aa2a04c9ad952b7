"""ctypes binding of the C ABI in include/rbpe.h (librbpe.so: hand-written sm_100a kernels + host API).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbpe.so")

OK, INFEASIBLE, NOT_CONVERGED, BAD_ARG, CUDA_ERROR = 0, 1, 2, 3, 4
MODE_GAUSS_SEIDEL, MODE_JACOBI = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


class RbpeConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("max_iter", C.c_int), ("tol_gap", C.c_double), ("tol_res", C.c_double),
                ("smem_budget", C.c_size_t), ("reserved", C.c_int * 6)]


class RbpeProblem(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("sequential", C.c_int), ("batch_size", C.c_int),
                ("batch_iter", C.c_int), ("iteration", C.c_int),
                ("T", _dp), ("start", _dp), ("goal", _dp), ("radius", _dp),
                ("sfc_offs", _ip), ("sfc_base", _ip), ("sfc_box", _dp), ("sfc_t", _dp),
                ("rsfc_n", _fp), ("rsfc_t", _dp), ("init_traj", _fp)]


class RbpeResult(C.Structure):
    _fields_ = [("coef", _dp), ("ctrl", _dp), ("qp_obj", _dp), ("qp_iters", _ip), ("qp_status", _ip),
                ("qp_res", _dp), ("status", _ip)]


class RbpeTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("assemble_ms", C.c_float), ("solve_ms", C.c_float), ("d2h_ms", C.c_float),
                ("total_ms", C.c_float), ("kernel_launches", C.c_int)]


EXPORTS = ["rbpe_create", "rbpe_destroy", "rbpe_last_error", "rbpe_set_batch", "rbpe_solve_many", "rbpe_solve",
           "rbpe_upload", "rbpe_assemble", "rbpe_run", "rbpe_run_jacobi_range", "rbpe_set_ctrl", "rbpe_download", "rbpe_device_ctrl",
           "rbpe_device_coef", "rbpe_stream", "rbpe_sync", "rbpe_last_timing", "rbpe_timer_start", "rbpe_timer_stop",
           "rbpe_corridor_rsfc", "rbpe_safety_metrics", "rbpe_peer_export", "rbpe_peer_attach", "rbpe_peer_attach_local",
           "rbpe_run_jacobi_fused", "rbpe_peer_status", "rbpe_host_alloc", "rbpe_host_free", "rbpe_convert", "rbpe_device_status", "rbpe_last_solver"]
IPC_HANDLE_BYTES = 64

_lib = None


def load_library(path=None):
    """dlopen librbpe.so and declare prototypes. Raises OSError when the extension has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise OSError("librbpe.so not found at %s -- build it with `python __graft_entry__.py` "
                      "(there is no CPU fallback)" % p)
    L = C.CDLL(p)
    L.rbpe_create.argtypes = [C.POINTER(RbpeConfig), C.POINTER(C.c_void_p)]
    L.rbpe_create.restype = C.c_int
    L.rbpe_destroy.argtypes = [C.c_void_p]
    L.rbpe_destroy.restype = None
    L.rbpe_last_error.argtypes = [C.c_void_p]
    L.rbpe_last_error.restype = C.c_char_p
    L.rbpe_set_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip]
    L.rbpe_set_batch.restype = C.c_int
    L.rbpe_solve_many.argtypes = [C.c_void_p, C.POINTER(RbpeProblem), C.c_int, C.c_int, C.POINTER(RbpeResult)]
    L.rbpe_solve_many.restype = C.c_int
    L.rbpe_solve.argtypes = [C.c_void_p, C.POINTER(RbpeProblem), C.POINTER(RbpeResult)]
    L.rbpe_solve.restype = C.c_int
    L.rbpe_upload.argtypes = [C.c_void_p, C.POINTER(RbpeProblem), C.c_int]
    L.rbpe_upload.restype = C.c_int
    L.rbpe_assemble.argtypes = [C.c_void_p]
    L.rbpe_assemble.restype = C.c_int
    L.rbpe_timer_start.argtypes = [C.c_void_p]
    L.rbpe_timer_start.restype = C.c_int
    L.rbpe_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.rbpe_timer_stop.restype = C.c_int
    L.rbpe_run.argtypes = [C.c_void_p, C.c_int]
    L.rbpe_run.restype = C.c_int
    L.rbpe_run_jacobi_range.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.rbpe_run_jacobi_range.restype = C.c_int
    if path is None or hasattr(L, "rbpe_host_alloc"):
        L.rbpe_host_alloc.argtypes = [C.c_size_t]
        L.rbpe_host_alloc.restype = C.c_void_p
        L.rbpe_host_free.argtypes = [C.c_void_p]
        L.rbpe_host_free.restype = None
    if path is None or hasattr(L, "rbpe_run_jacobi_fused"):   # (tools/gpu_ab.py may load older builds of the library)
        L.rbpe_run_jacobi_fused.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.rbpe_run_jacobi_fused.restype = C.c_int
        L.rbpe_peer_export.argtypes = [C.c_void_p, C.c_char_p]
        L.rbpe_peer_export.restype = C.c_int
        L.rbpe_peer_attach.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        L.rbpe_peer_attach.restype = C.c_int
        L.rbpe_peer_attach_local.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.rbpe_peer_attach_local.restype = C.c_int
        L.rbpe_peer_status.argtypes = [C.c_void_p]
        L.rbpe_peer_status.restype = C.c_int
    L.rbpe_set_ctrl.argtypes = [C.c_void_p, _dp]
    L.rbpe_set_ctrl.restype = C.c_int
    L.rbpe_download.argtypes = [C.c_void_p, C.POINTER(RbpeResult)]
    L.rbpe_download.restype = C.c_int
    L.rbpe_device_ctrl.argtypes = [C.c_void_p]
    L.rbpe_device_ctrl.restype = C.c_void_p
    L.rbpe_device_coef.argtypes = [C.c_void_p]
    L.rbpe_device_coef.restype = C.c_void_p
    if path is None or hasattr(L, "rbpe_convert"):
        L.rbpe_convert.argtypes = [C.c_void_p]
        L.rbpe_convert.restype = C.c_int
        L.rbpe_device_status.argtypes = [C.c_void_p]
        L.rbpe_device_status.restype = C.c_void_p
    if path is None or hasattr(L, "rbpe_last_solver"):
        L.rbpe_last_solver.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.rbpe_last_solver.restype = C.c_int
    L.rbpe_stream.argtypes = [C.c_void_p]
    L.rbpe_stream.restype = C.c_void_p
    L.rbpe_sync.argtypes = [C.c_void_p]
    L.rbpe_sync.restype = C.c_int
    L.rbpe_last_timing.argtypes = [C.c_void_p, C.POINTER(RbpeTiming)]
    L.rbpe_last_timing.restype = C.c_int
    L.rbpe_corridor_rsfc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _fp, _dp, C.c_double, _fp, _dp, _ip]
    L.rbpe_corridor_rsfc.restype = C.c_int
    L.rbpe_safety_metrics.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp, _dp, _dp]
    L.rbpe_safety_metrics.restype = C.c_int
    if path is None:
        _lib = L
    return L


def set_batch(N, sequential, batch_size, batch_iter):
    """RBPPlanner::setBatch (rbp_planner.hpp L849-L872): (effective batch_size, effective batch_iter, max batches)."""
    bs, bi = C.c_int(), C.c_int()
    bmax = load_library().rbpe_set_batch(N, int(bool(sequential)), batch_size, batch_iter, C.byref(bs), C.byref(bi))
    return bs.value, bi.value, bmax


class PackedProblem:
    """Keeps the numpy arrays of `synth.pack()` (or the caller's own) alive next to the C struct."""

    def __init__(self, packed, sequential=True, batch_size=1, batch_iter=-1, iteration=1):
        self.a = {k: v for k, v in packed.items()}
        a = self.a
        self.count, self.N, self.M = int(a["count"]), int(a["N"]), int(a["M"])
        for k, dt in (("T", np.float64), ("start", np.float64), ("goal", np.float64), ("radius", np.float64),
                      ("sfc_offs", np.int32), ("sfc_base", np.int32), ("sfc_box", np.float64),
                      ("sfc_t", np.float64), ("rsfc_n", np.float32), ("rsfc_t", np.float64),
                      ("init_traj", np.float32)):
            a[k] = np.ascontiguousarray(a[k], dt)
        self.sequential, self.batch_size, self.batch_iter, self.iteration = (bool(sequential), int(batch_size),
                                                                             int(batch_iter), int(iteration))
        self.c = RbpeProblem(self.N, self.M, int(self.sequential), self.batch_size, self.batch_iter, self.iteration,
                             a["T"].ctypes.data_as(_dp), a["start"].ctypes.data_as(_dp),
                             a["goal"].ctypes.data_as(_dp), a["radius"].ctypes.data_as(_dp),
                             a["sfc_offs"].ctypes.data_as(_ip), a["sfc_base"].ctypes.data_as(_ip),
                             a["sfc_box"].ctypes.data_as(_dp), a["sfc_t"].ctypes.data_as(_dp),
                             a["rsfc_n"].ctypes.data_as(_fp), a["rsfc_t"].ctypes.data_as(_dp),
                             a["init_traj"].ctypes.data_as(_fp))

    def effective_batching(self):
        bs = self.batch_size if self.batch_size > 0 else 1
        bmax = -(-self.N // bs)
        if self.sequential:
            bi = self.batch_iter
            if bi < 0 or bi > bmax:
                bi = bmax
            return bs, bi
        return self.N, 1

    def h2d_bytes(self):
        return int(sum(self.a[k].nbytes for k in ("T", "start", "goal", "radius", "sfc_offs", "sfc_base", "sfc_box",
                                                   "sfc_t", "rsfc_n", "rsfc_t", "init_traj")))


class _PinnedPool:
    """numpy arrays over page-locked memory from rbpe_host_alloc (freed when the pool dies)."""

    def __init__(self):
        self.lib = load_library()
        self.ptrs = []

    def zeros(self, shape, dtype=np.float64):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.lib.rbpe_host_alloc(max(n, 1))
        if not p:
            raise MemoryError("rbpe_host_alloc(%d) failed" % n)
        self.ptrs.append(p)
        a = np.frombuffer((C.c_char * max(n, 1)).from_address(p), dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        a[...] = 0
        return a

    def __del__(self):
        for p in self.ptrs:
            self.lib.rbpe_host_free(p)
        self.ptrs = []


class Result:
    def __init__(self, prob, want_ctrl=True, pinned=False):
        """pinned=True: result arrays in page-locked memory (rbpe_host_alloc), so that the D2H copies of the pipelined
        rbpe_solve_many are truly asynchronous."""
        c, N, M = prob.count, prob.N, prob.M
        _, bi = prob.effective_batching()
        nrec = max(1, prob.iteration * bi)
        self.nrec = prob.iteration * bi
        self._pool = _PinnedPool() if pinned else None
        zeros = self._pool.zeros if pinned else (lambda shape, dtype=np.float64: np.zeros(shape, dtype))
        self.coef = zeros((c, N, 3, 6 * M))
        self.ctrl = zeros((c, N, 3, 6 * M)) if want_ctrl else None
        self.qp_obj = zeros((c, nrec))
        self.qp_iters = zeros((c, nrec), np.int32)
        self.qp_status = zeros((c, nrec), np.int32)
        self.qp_res = zeros((c, nrec, 4))
        self.status = zeros(c, np.int32)
        self.c = RbpeResult(self.coef.ctypes.data_as(_dp), self.ctrl.ctypes.data_as(_dp) if want_ctrl else None,
                            self.qp_obj.ctypes.data_as(_dp), self.qp_iters.ctypes.data_as(_ip),
                            self.qp_status.ctypes.data_as(_ip), self.qp_res.ctypes.data_as(_dp),
                            self.status.ctypes.data_as(_ip))

    def d2h_bytes(self):
        return int(self.coef.nbytes + (self.ctrl.nbytes if self.ctrl is not None else 0) + self.qp_obj.nbytes + self.qp_iters.nbytes
                   + self.qp_status.nbytes + self.qp_res.nbytes + self.status.nbytes)


class Engine:
    """One engine handle = one CUDA device + one stream (include/rbpe.h). Not thread-safe."""

    def __init__(self, device=0, max_iter=0, tol_gap=0.0, tol_res=0.0, smem_budget=0, threads=0):
        self.lib = load_library()
        cfg = RbpeConfig(device, max_iter, tol_gap, tol_res, smem_budget)
        cfg.reserved[0] = threads
        h = C.c_void_p()
        rc = self.lib.rbpe_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise RuntimeError("rbpe_create failed (%d): %s" % (rc, self.lib.rbpe_last_error(None).decode()))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.rbpe_destroy(self.h)
            self.h = None

    __del__ = close

    def last_error(self):
        return self.lib.rbpe_last_error(self.h).decode()

    def solve_many(self, prob, mode=MODE_GAUSS_SEIDEL, result=None):
        r = result or Result(prob)
        rc = self.lib.rbpe_solve_many(self.h, C.byref(prob.c), prob.count, mode, C.byref(r.c))
        if rc == CUDA_ERROR or rc == BAD_ARG and not r.status.any():
            raise RuntimeError("rbpe_solve_many failed (%d): %s" % (rc, self.last_error()))
        r.rc = rc
        return r

    def upload(self, prob):
        rc = self.lib.rbpe_upload(self.h, C.byref(prob.c), prob.count)
        if rc != OK:
            raise RuntimeError("rbpe_upload failed (%d): %s" % (rc, self.last_error()))

    def assemble(self):
        rc = self.lib.rbpe_assemble(self.h)
        if rc != OK:
            raise RuntimeError("rbpe_assemble failed (%d): %s" % (rc, self.last_error()))

    def timer_start(self):
        if self.lib.rbpe_timer_start(self.h) != OK:
            raise RuntimeError(self.last_error())

    def timer_stop(self):
        ms = C.c_float()
        if self.lib.rbpe_timer_stop(self.h, C.byref(ms)) != OK:
            raise RuntimeError(self.last_error())
        return ms.value

    def run(self, mode=MODE_GAUSS_SEIDEL):
        rc = self.lib.rbpe_run(self.h, mode)
        if rc == CUDA_ERROR or rc == BAD_ARG:
            raise RuntimeError("rbpe_run failed (%d): %s" % (rc, self.last_error()))
        return rc

    def run_jacobi_range(self, b0, b1):
        rc = self.lib.rbpe_run_jacobi_range(self.h, b0, b1)
        if rc == CUDA_ERROR or rc == BAD_ARG:
            raise RuntimeError("rbpe_run_jacobi_range failed (%d): %s" % (rc, self.last_error()))
        return rc

    # ---- Jacobi exchange over peer memory (include/rbpe.h, rbpe_peer_*) ----
    def peer_export(self):
        """Pins the table buffers and returns their IPC handles (bytes, 3 x 64)."""
        buf = C.create_string_buffer(3 * IPC_HANDLE_BYTES)
        rc = self.lib.rbpe_peer_export(self.h, buf)
        if rc != OK:
            raise RuntimeError("rbpe_peer_export failed (%d): %s" % (rc, self.last_error()))
        return buf.raw

    def peer_attach(self, rank, world, all_handles):
        """all_handles: list of `world` byte strings from peer_export, rank order (other processes)."""
        blob = b"".join(all_handles)
        assert len(blob) == world * 3 * IPC_HANDLE_BYTES
        rc = self.lib.rbpe_peer_attach(self.h, rank, world, blob)
        if rc != OK:
            raise RuntimeError("rbpe_peer_attach failed (%d): %s" % (rc, self.last_error()))

    def peer_attach_local(self, rank, engines):
        """engines: the Engine objects of all ranks living in THIS process, rank order."""
        arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
        rc = self.lib.rbpe_peer_attach_local(self.h, rank, len(engines), arr)
        if rc != OK:
            raise RuntimeError("rbpe_peer_attach_local failed (%d): %s" % (rc, self.last_error()))

    def run_jacobi_fused(self, b0, b1):
        rc = self.lib.rbpe_run_jacobi_fused(self.h, b0, b1)
        if rc != OK:
            raise RuntimeError("rbpe_run_jacobi_fused failed (%d): %s" % (rc, self.last_error()))

    def peer_status(self):
        rc = self.lib.rbpe_peer_status(self.h)
        if rc != OK:
            raise RuntimeError("peer exchange failed (%d): %s" % (rc, self.last_error()))

    def set_ctrl(self, ctrl):
        c = np.ascontiguousarray(ctrl, np.float64)
        rc = self.lib.rbpe_set_ctrl(self.h, c.ctypes.data_as(_dp))
        if rc != OK:
            raise RuntimeError("rbpe_set_ctrl failed (%d): %s" % (rc, self.last_error()))
        self.sync()

    def download(self, prob, result=None):
        r = result or Result(prob)
        rc = self.lib.rbpe_download(self.h, C.byref(r.c))
        if rc == CUDA_ERROR:
            raise RuntimeError("rbpe_download failed: %s" % self.last_error())
        r.rc = rc
        return r

    def corridor_rsfc(self, init_traj, T, downwash=2.0):
        """Corridor::updateRelBox on the device. init_traj [count,N,M+1,3] f32, T [count,M+1] -> (rsfc_n, rsfc_t, collided)."""
        tr = np.ascontiguousarray(init_traj, np.float32)
        T = np.ascontiguousarray(T, np.float64)
        count, N, M = tr.shape[0], tr.shape[1], tr.shape[2] - 1
        P = N * (N - 1) // 2
        n = np.zeros((count, max(P, 1), M, 3), np.float32)
        t = np.zeros((count, max(P, 1), M))
        col = np.zeros(count, np.int32)
        rc = self.lib.rbpe_corridor_rsfc(self.h, N, M, count, tr.ctypes.data_as(_fp), T.ctypes.data_as(_dp), float(downwash),
                                         n.ctypes.data_as(_fp), t.ctypes.data_as(_dp), col.ctypes.data_as(_ip))
        if rc != OK:
            raise RuntimeError("rbpe_corridor_rsfc failed (%d): %s" % (rc, self.last_error()))
        return n[:, :P], t[:, :P], col

    def safety_metrics(self, coef, T, radius, downwash=2.0, dt=0.1):
        """RBPPublisher checks on the device. coef [count,N,3,6M] -> (min_ratio, t_at_min, length), each [count]."""
        c = np.ascontiguousarray(coef, np.float64)
        T = np.ascontiguousarray(T, np.float64)
        r = np.ascontiguousarray(radius, np.float64)
        count, N, M = c.shape[0], c.shape[1], T.shape[1] - 1
        a, b, l = np.zeros(count), np.zeros(count), np.zeros(count)
        rc = self.lib.rbpe_safety_metrics(self.h, N, M, count, c.ctypes.data_as(_dp), T.ctypes.data_as(_dp), r.ctypes.data_as(_dp),
                                          float(downwash), float(dt), a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), l.ctypes.data_as(_dp))
        if rc != OK:
            raise RuntimeError("rbpe_safety_metrics failed (%d): %s" % (rc, self.last_error()))
        return a, b, l

    def sync(self):
        return self.lib.rbpe_sync(self.h)

    def timing(self):
        t = RbpeTiming()
        self.lib.rbpe_last_timing(self.h, C.byref(t))
        return {k: getattr(t, k) for k, _ in RbpeTiming._fields_}

    def device_ctrl_ptr(self):
        return self.lib.rbpe_device_ctrl(self.h)

    def device_coef_ptr(self):
        return self.lib.rbpe_device_coef(self.h)

    def device_status_ptr(self):
        return self.lib.rbpe_device_status(self.h)

    def last_solver(self):
        """(kernel, threads per QP) of the last solver launch: 1 = pdip1_kernel, 2 = pdip1x_kernel, 3 = pdip_kernel."""
        t = C.c_int(0)
        k = self.lib.rbpe_last_solver(self.h, C.byref(t))
        return k, t.value

    def convert(self):
        rc = self.lib.rbpe_convert(self.h)
        if rc != OK:
            raise RuntimeError("rbpe_convert failed (%d): %s" % (rc, self.last_error()))

    def stream_ptr(self):
        return self.lib.rbpe_stream(self.h)
