"""Kernel LOGIC check without a GPU: the product kernels (swarm_simulator_b200/csrc/rbpe_kernels.cuh) compiled for the
host under the fiber emulator of tests/cpu_emu and compared with the oracle.  This is test infrastructure (it guards
the indexing and control flow of the CUDA source in CI that has no device); the parity tests proper are -m gpu."""
import numpy as np
import pytest

import emu_util
import oracle_util
from swarm_simulator_b200 import engine as E, synth


@pytest.mark.parametrize("N,M,rho,seq,bs,smem", [
    (4, 3, 0.0, False, 4, 48 * 1024),   # BASELINE configs[0]: 4 agents, empty map, 3 segments, one joint batch
    (5, 4, 0.2, True, 2, 0),            # sequential, ragged last batch, everything in "global" scratch
])
def test_emulated_kernels_match_oracle(N, M, rho, seq, bs, smem):
    m = synth.synth_mission(N, M, rho, 77)
    prob = E.PackedProblem(synth.pack([m]), sequential=seq, batch_size=bs)
    r = emu_util.emu_solve_many(prob, smem_bytes=smem, threads=64)
    ro = oracle_util.oracle_problem(m, sequential=seq, batch_size=bs).update()
    assert r.rc == 0 and ro["status"] == 0
    assert np.array_equal(r.qp_iters[0], ro["batch_iters"][:r.qp_iters.shape[1]])
    assert np.abs(r.ctrl[0] - ro["ctrl"]).max() < 1e-9
    assert np.abs(r.coef[0] - ro["coef"]).max() < 1e-9
