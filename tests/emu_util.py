"""TEST INFRASTRUCTURE: runs the product kernels under tests/cpu_emu (fiber emulator, no GPU). Never used by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

from swarm_simulator_b200 import engine as E

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMU_DIR = os.path.join(_HERE, "cpu_emu")
_EMU_LIB = os.path.join(_EMU_DIR, "librbpe_emu.so")
_SRCS = [os.path.join(_EMU_DIR, "emu_driver.cpp"), os.path.join(_EMU_DIR, "cuda_emu.h"),
         os.path.join(_HERE, "..", "swarm_simulator_b200", "csrc", "rbpe_kernels.cuh"),
         os.path.join(_HERE, "..", "swarm_simulator_b200", "csrc", "rbpe_blockla.cuh"),
         os.path.join(_HERE, "..", "swarm_simulator_b200", "csrc", "rbpe_pdip1.cuh"),
         os.path.join(_HERE, "..", "swarm_simulator_b200", "csrc", "rbpe_pdip1x.cuh"),
         os.path.join(_HERE, "..", "swarm_simulator_b200", "csrc", "rbpe_types.h")]


def build():
    if (not os.path.exists(_EMU_LIB)) or os.path.getmtime(_EMU_LIB) < max(os.path.getmtime(s) for s in _SRCS):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unused", "-Wno-unknown-pragmas", "-o", _EMU_LIB,
                               _SRCS[0]])
    return _EMU_LIB


_lib = None
_variants = {}


def emu_variant(tag, flags):
    """The same kernels compiled with extra preprocessor flags (e.g. the textbook four-pass loop, -DRBPE_FUSE_COR=0 ...):
    returns a callable with emu_solve_many's signature bound to that library."""
    if tag not in _variants:
        path = os.path.join(_EMU_DIR, "librbpe_emu_%s.so" % tag)
        if (not os.path.exists(path)) or os.path.getmtime(path) < max(os.path.getmtime(s) for s in _SRCS):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wno-unused", "-Wno-unknown-pragmas"] + list(flags) +
                                  ["-o", path, _SRCS[0]])
        lib = C.CDLL(path)
        lib.emu_solve_many.argtypes = [C.POINTER(E.RbpeProblem), C.c_int, C.c_int, C.POINTER(E.RbpeResult),
                                       C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int]
        lib.emu_solve_many.restype = C.c_int
        _variants[tag] = lib

    def solve(prob, mode=0, smem_bytes=48 * 1024, max_iter=0, tol_gap=0.0, tol_res=0.0, threads=256):
        r = E.Result(prob)
        r.rc = _variants[tag].emu_solve_many(C.byref(prob.c), prob.count, mode, C.byref(r.c), smem_bytes, max_iter, tol_gap, tol_res, threads)
        return r
    return solve


def emu_solve_many(prob, mode=0, smem_bytes=48 * 1024, max_iter=0, tol_gap=0.0, tol_res=0.0, threads=256):
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_solve_many.argtypes = [C.POINTER(E.RbpeProblem), C.c_int, C.c_int, C.POINTER(E.RbpeResult),
                                        C.c_size_t, C.c_int, C.c_double, C.c_double, C.c_int]
        _lib.emu_solve_many.restype = C.c_int
    r = E.Result(prob)
    r.rc = _lib.emu_solve_many(C.byref(prob.c), prob.count, mode, C.byref(r.c), smem_bytes, max_iter, tol_gap,
                               tol_res, threads)
    return r


def emu_block_tridiag(D, O, g, threads=64):
    """D [nblk, kp, kp], O [nblk-1, kp, kp] (kp = kb rounded up to a multiple of 8, identity padded), g [nblk, kb]:
    factor + solve with the product routines of rbpe_blockla.cuh under the emulator; returns (ok, x, L_diag_blocks, L_off_blocks)."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    _lib.emu_block_tridiag.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    _lib.emu_block_tridiag.restype = C.c_int
    nblk, kb = g.shape
    Dc, Oc, gc = np.ascontiguousarray(D, np.float64).copy(), np.ascontiguousarray(O, np.float64).copy(), np.ascontiguousarray(g, np.float64).copy()
    if Oc.size == 0:
        Oc = np.zeros((1,) + Dc.shape[1:])
    ok = _lib.emu_block_tridiag(nblk, kb, Dc.ctypes.data, Oc.ctypes.data, gc.ctypes.data, threads)
    return bool(ok), gc, Dc, Oc
