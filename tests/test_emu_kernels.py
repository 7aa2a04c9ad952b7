"""Kernel LOGIC check without a GPU: the product kernels (swarm_simulator_b200/csrc/rbpe_kernels.cuh) compiled for the
host under the fiber emulator of tests/cpu_emu and compared with the oracle.  This is test infrastructure (it guards
the indexing and control flow of the CUDA source in CI that has no device); the parity tests proper are -m gpu."""
import numpy as np
import pytest

import emu_util
import oracle_util
from swarm_simulator_b200 import engine as E, synth


@pytest.mark.parametrize("N,M,rho,seq,bs,smem,threads", [
    (4, 3, 0.0, False, 4, 48 * 1024, 64),   # BASELINE configs[0]: 4 agents, empty map, 3 segments, one joint batch
    (5, 4, 0.2, True, 2, 0, 64),            # sequential, ragged last batch, everything in "global" scratch
    (38, 3, 0.1, True, 4, 48 * 1024, 64),   # 34 frozen agents per batch: the compaction of kept rows crosses a 32-row group
    (8, 4, 0.2, True, 4, 48 * 1024, 256),   # the launch default b = 4 with the full CTA (8 warps: register-resident diagonal blocks)
    (8, 3, 0.1, True, 4, 96 * 1024, 512),   # latency regime: 16 warps per CTA
])
def test_emulated_kernels_match_oracle(N, M, rho, seq, bs, smem, threads):
    m = synth.synth_mission(N, M, rho, 77)
    prob = E.PackedProblem(synth.pack([m]), sequential=seq, batch_size=bs)
    r = emu_util.emu_solve_many(prob, smem_bytes=smem, threads=threads)
    ro = oracle_util.oracle_problem(m, sequential=seq, batch_size=bs).update()
    assert r.rc == 0 and ro["status"] == 0
    assert np.array_equal(r.qp_iters[0], ro["batch_iters"][:r.qp_iters.shape[1]])
    assert np.abs(r.ctrl[0] - ro["ctrl"]).max() < 1e-9
    assert np.abs(r.coef[0] - ro["coef"]).max() < 1e-9


@pytest.mark.parametrize("N,M,rho,threads,mode", [
    (5, 4, 0.2, 256, 0),      # pdip1_kernel: one warp per QP, Gauss-Seidel chain
    (5, 4, 0.2, -128, 0),     # pdip1x_kernel (latency kernel): 4 warps per QP
    (12, 5, 0.1, -256, 0),    # 8 warps per QP: 17 rows per control point dealt to 8 warps
    (7, 6, 0.1, -96, 0),      # 36 control points: two lane slots; 3 warps per QP
    (6, 3, 0.0, -256, 1),     # Jacobi sweeps: one CTA per (mission, agent)
])
def test_emulated_one_agent_kernels_match_oracle(N, M, rho, threads, mode):
    """One-agent batches (plan/batch_size = 1): the warp-per-QP kernel and the several-warps-per-QP latency kernel
    (rbpe_pdip1.cuh, rbpe_pdip1x.cuh) against the oracle -- same iteration counts, control points to rounding."""
    m = synth.synth_mission(N, M, rho, 78)
    prob = E.PackedProblem(synth.pack([m]), sequential=True, batch_size=1)
    r = emu_util.emu_solve_many(prob, mode=mode, threads=threads)
    op = oracle_util.oracle_problem(m, sequential=True, batch_size=1)
    if mode == 1:   # every QP of the sweep against the table frozen before the sweep
        dummy = op.dummy()
        for l in range(N):
            x = op.populate(dummy, l).solve()
            assert x["status"] == 0 and x["iters"] == r.qp_iters[0][l]
            assert np.abs(r.ctrl[0][l] - x["x"].reshape(3, 6 * M)).max() < 1e-9
        return
    ro = op.update()
    assert r.rc == 0 and ro["status"] == 0
    assert np.array_equal(r.qp_iters[0], ro["batch_iters"][:r.qp_iters.shape[1]])
    assert np.abs(r.ctrl[0] - ro["ctrl"]).max() < 1e-9
    assert np.abs(r.coef[0] - ro["coef"]).max() < 1e-9


def test_fused_corrector_equals_the_separate_corrector_pass():
    """The kernels form the corrector's G' product from two sums of the affine pass (RBPE_FUSE_COR, RBPE_W1_FUSE_COR,
    RBPE_X1_FUSE_COR: three passes over the rows per iteration).  Against the same kernels compiled with the textbook
    separate corrector pass: same statuses and iteration counts, control points to rounding -- for the warp-per-QP kernel,
    the latency kernel and a joint batch."""
    four = emu_util.emu_variant("fourpass", ["-DRBPE_FUSE_COR=0", "-DRBPE_W1_FUSE_COR=0", "-DRBPE_X1_FUSE_COR=0"])
    for (N, M, rho, seq, bs, threads, smem) in ((9, 5, 0.2, True, 1, 256, 48 * 1024), (9, 5, 0.2, True, 1, -128, 48 * 1024),
                                                (8, 4, 0.2, True, 4, 256, 48 * 1024), (6, 3, 0.1, False, 6, 64, 48 * 1024)):
        m = synth.synth_mission(N, M, rho, 79)
        prob = E.PackedProblem(synth.pack([m]), sequential=seq, batch_size=bs)
        a = emu_util.emu_solve_many(prob, smem_bytes=smem, threads=threads)
        b = four(prob, smem_bytes=smem, threads=threads)
        assert a.rc == b.rc == 0
        assert np.array_equal(a.qp_status, b.qp_status) and np.array_equal(a.qp_iters, b.qp_iters)
        assert np.abs(a.ctrl - b.ctrl).max() < 1e-10
        assert np.allclose(a.qp_obj, b.qp_obj, rtol=1e-8, atol=1e-9)   # (the bar of the GPU-vs-oracle tests)


@pytest.mark.parametrize("kb,nblk,threads", [(18, 3, 64), (36, 2, 64), (45, 3, 96), (72, 2, 128), (27, 1, 32),
                                             (36, 4, 256), (18, 3, 512), (63, 2, 384)])   # >= 8 warps: register-resident diagonal blocks
def test_emulated_block_tridiagonal_factor_and_solve(kb, nblk, threads):
    """rbpe_blockla.cuh (DMMA tile updates, warp-level 32 x 32 diagonal blocks + inverses, blocked substitution) against
    numpy on a random SPD block tridiagonal system, for block orders that are / are not multiples of 8 and of 32."""
    rng = np.random.default_rng(kb * 10 + nblk)
    n = kb * nblk
    G = rng.standard_normal((n, n))
    A = G @ G.T + n * np.eye(n)
    for i in range(nblk):            # keep the block tridiagonal part only (diagonally dominant enough to stay SPD)
        for j in range(nblk):
            if abs(i - j) > 1:
                A[i * kb:(i + 1) * kb, j * kb:(j + 1) * kb] = 0
    A += np.eye(n) * np.abs(A).sum(1).max()
    kp = (kb + 7) // 8 * 8
    D = np.zeros((nblk, kp, kp)); O = np.zeros((max(nblk - 1, 0), kp, kp))
    for t in range(nblk):
        D[t] = np.eye(kp)
        D[t, :kb, :kb] = np.tril(A[t * kb:(t + 1) * kb, t * kb:(t + 1) * kb])
        if t + 1 < nblk:
            O[t, :kb, :kb] = A[(t + 1) * kb:(t + 2) * kb, t * kb:(t + 1) * kb]
    g = rng.standard_normal((nblk, kb))
    ok, x, Lf, Lo = emu_util.emu_block_tridiag(D, O, g, threads=threads)
    assert ok
    x_ref = np.linalg.solve(A, g.reshape(-1))
    assert np.abs(x.reshape(-1) - x_ref).max() < 1e-11 * max(1.0, np.abs(x_ref).max())
    L = np.linalg.cholesky(A)
    for t in range(nblk):
        assert np.abs(np.tril(Lf[t, :kb, :kb]) - L[t * kb:(t + 1) * kb, t * kb:(t + 1) * kb]).max() < 1e-10 * np.abs(L).max()
        if t + 1 < nblk:
            assert np.abs(Lo[t, :kb, :kb] - L[(t + 1) * kb:(t + 2) * kb, t * kb:(t + 1) * kb]).max() < 1e-10 * np.abs(L).max()
