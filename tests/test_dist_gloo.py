"""N>1 host logic on CPU: world_size 2 over gloo.  Mission sharding covers every mission exactly once, and the Jacobi
exchange (one all-gather per sweep) rebuilds the same control-point table on every rank as a single process computes."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_util
from swarm_simulator_b200 import dist as D, synth


def test_shard_missions_partitions():
    for count in (1, 7, 64, 1184):
        for world in (1, 2, 3, 8):
            parts = [D.shard_missions(count, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == count
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert D.agent_range(7, 3, 3, 2, 0) == (0, 6) and D.agent_range(7, 3, 3, 2, 1) == (6, 7)


def _sweep_oracle(m, table, bs, batches):
    """Solve the given batches of one Jacobi sweep with the oracle against the frozen `table` [N,3,6M]."""
    op = oracle_util.oracle_problem(m, sequential=True, batch_size=bs)
    dummy = np.ascontiguousarray(table.transpose(0, 2, 1).reshape(-1, 3))
    out = {}
    for l in batches:
        x = op.populate(dummy, l).solve()
        assert x["status"] == 0
        nb = min(bs, m["N"] - l * bs)
        out[l] = x["x"].reshape(3, nb, 6 * m["M"]).transpose(1, 0, 2)
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = synth.synth_mission(5, 4, 0.2, 21)
    bs, nbatch = 2, 3
    op = oracle_util.oracle_problem(m, sequential=True, batch_size=bs)
    table = torch.from_numpy(op.dummy().reshape(5, 24, 3).transpose(0, 2, 1).copy())[None]   # [1, N, 3, 6M]
    for sweep in range(2):
        frozen = table[0].numpy().copy()
        b0, b1 = D.batch_range(nbatch, world, rank)
        for l, ctrl in _sweep_oracle(m, frozen, bs, range(b0, b1)).items():
            table[0, l * bs:l * bs + ctrl.shape[0]] = torch.from_numpy(ctrl.copy())
        D.exchange_ctrl(table, 5, bs, nbatch)
    # mission status words: a batch that failed on one rank fails the mission everywhere
    status = torch.tensor([0, 2 if rank == 1 else 0, 1 if rank == 0 else 0, 0], dtype=torch.int32)
    D.merge_status(status)
    q.put((rank, table.numpy().copy(), status.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_jacobi_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    recs = [q.get(timeout=120) for _ in range(2)]
    got = {r[0]: r[1] for r in recs}
    for r in recs:
        assert r[2].tolist() == [0, 2, 1, 0]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference of the same two sweeps
    m = synth.synth_mission(5, 4, 0.2, 21)
    op = oracle_util.oracle_problem(m, sequential=True, batch_size=2)
    table = op.dummy().reshape(5, 24, 3).transpose(0, 2, 1).copy()
    for sweep in range(2):
        frozen = table.copy()
        for l, ctrl in _sweep_oracle(m, frozen, 2, range(3)).items():
            table[l * 2:l * 2 + ctrl.shape[0]] = ctrl
    assert np.array_equal(got[0], got[1])
    assert np.array_equal(got[0][0], table)
