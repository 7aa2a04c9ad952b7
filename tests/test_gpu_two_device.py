"""Two-DEVICE test of the Jacobi exchange (-m gpu; skipped on a box with one GPU): one process per GPU over NCCL, as bench.py
launches them.  Checks, on every rank:
  * the fused exchange (peer stores over NVLink + flag words, rbpe_run_jacobi_fused) and the NCCL all-gather variant build
    bit-identical control-point tables, and both equal a single-device Jacobi solve of the same missions;
  * coefficients of the agents the OTHER rank solved are converted from the exchanged control points (round-1 bug: the
    conversion ran before the exchange);
  * a batch that fails on one rank fails the mission on both (status merge), and the flag protocol reports no timeout.
Reference behaviour being preserved: per-QP results (rbp_planner.hpp L551-L688) and `update()` returning false when any
batch fails (L158-L161)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    import feas_util as fu
    from swarm_simulator_b200 import engine as E, synth, dist as D
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    ms = fu.missions("cfg3", 6)
    ms[2] = dict(ms[2], radius=np.full(64, 0.25))          # inflated radii: some agents' QPs become infeasible
    sweeps = 2
    prob = E.PackedProblem(synth.pack(ms), sequential=True, batch_size=1, iteration=sweeps)
    out = {}
    for name, fused in (("allgather", False), ("fused", True)):
        eng = E.Engine(device=rank)
        if fused:
            D.jacobi_attach_peers(eng, prob)
        D.jacobi_solve(eng, prob, sweeps, device=dev, fused=fused)
        r = eng.download(prob)
        if fused:
            eng.peer_status()
        out[name] = (r.ctrl.copy(), r.coef.copy(), r.status.copy())
        eng.close()
    single = None
    if rank == 0:                                           # the same two sweeps on one device
        eng = E.Engine(device=0)
        eng.upload(prob); eng.assemble()
        for _ in range(sweeps):
            eng.run_jacobi_range(0, 64)
        r = eng.download(prob)
        single = (r.ctrl.copy(), r.coef.copy(), r.status.copy())
        eng.close()
    dist.barrier()
    q.put((rank, out, single))
    dist.destroy_process_group()


def test_two_device_jacobi_exchange_tables_coefficients_and_status():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    recs = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    by_rank = {r[0]: r for r in recs}
    ctrl1, coef1, st1 = by_rank[0][2]
    assert st1[2] != 0 and not st1[[0, 1, 3, 4, 5]].any()        # the inflated mission fails, the others plan
    ok = st1 == 0
    for rank in (0, 1):
        for name in ("allgather", "fused"):
            ctrl, coef, st = by_rank[rank][1][name]
            assert np.array_equal(st, st1), (rank, name, st, st1)                      # merged over ranks
            assert np.array_equal(ctrl[ok], ctrl1[ok]), (rank, name)                   # bit-identical tables
            assert np.array_equal(coef[ok], coef1[ok]), (rank, name)                   # ... and coefficients, for ALL agents
