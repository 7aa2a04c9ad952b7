"""The C-ABI library loads and exports every symbol include/rbpe.h declares; host-only entry points behave; and
without a CUDA device the engine fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as G
from swarm_simulator_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    G.build()
    return E.load_library()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "rbpe.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(rbpe_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no prototypes found"
    for name in declared:
        assert hasattr(lib, name), "librbpe.so does not export %s" % name
    assert sorted(E.EXPORTS) == declared


def test_set_batch_matches_reference_rules(lib):
    """RBPPlanner::setBatch (rbp_planner.hpp L849-L872)."""
    assert E.set_batch(64, True, 4, -1) == (4, 16, 16)
    assert E.set_batch(64, True, 4, 3) == (4, 3, 16)
    assert E.set_batch(64, True, 4, 99) == (4, 16, 16)
    assert E.set_batch(64, True, 4, 0) == (4, 0, 16)
    assert E.set_batch(6, True, 4, -1) == (4, 2, 2)
    assert E.set_batch(64, False, 4, 0) == (64, 1, 16)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    cfg = E.RbpeConfig(0, 0, 0.0, 0.0, 0)
    rc = lib.rbpe_create(C.byref(cfg), C.byref(h))
    assert rc == E.CUDA_ERROR and not h.value
    assert b"no CPU fallback" in lib.rbpe_last_error(None)
    with pytest.raises(RuntimeError):
        E.Engine()


def test_null_handle_is_rejected(lib):
    assert lib.rbpe_run(None, 0) == E.BAD_ARG
    assert lib.rbpe_sync(None) == E.BAD_ARG
