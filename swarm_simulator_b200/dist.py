"""Multi-GPU plumbing (one process per GPU, torch.distributed): how the path shards, and the one exchange step it has.

* Missions are independent: `shard_missions` splits them over ranks with no data-path collective (weak scaling).
* Inside a mission the reference's batches form a Gauss-Seidel chain through `dummy` (rbp_planner.hpp L140-L201), which
  does not shard.  The Jacobi relaxation does: every batch of a sweep is solved against the control-point table frozen
  before the sweep, ranks take contiguous ranges of batches, and ONE all-gather of the solved control points per sweep
  rebuilds the table everywhere (`exchange_ctrl`).  The payload is tiny (18 M doubles per agent), so the collective is
  latency bound.

The functions work on any torch tensor / process group (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_missions(count, world, rank):
    """Contiguous slice of `count` missions owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def batch_range(nbatch, world, rank):
    """Contiguous range of batches [b0, b1) of every mission solved by `rank` during a Jacobi sweep."""
    return shard_missions(nbatch, world, rank)


def agent_range(N, bs, nbatch, world, rank):
    b0, b1 = batch_range(nbatch, world, rank)
    return min(b0 * bs, N), min(b1 * bs, N)


def exchange_ctrl(table, N, bs, nbatch, group=None):
    """All-gather of the control points each rank has just solved.

    table: [count, N, 3, 6M] float64 (the engine's resident `dummy`, or a CPU tensor in tests); on return every rank
    holds the union of all ranks' solved agents.  One collective call."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return table
    rank = dist.get_rank(group)
    ranges = [agent_range(N, bs, nbatch, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in ranges)
    send = table.new_zeros((table.shape[0], width) + tuple(table.shape[2:]))
    lo, hi = ranges[rank]
    send[:, :hi - lo] = table[:, lo:hi]
    recv = table.new_empty((world * send.shape[0],) + tuple(send.shape[1:]))
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view((world,) + tuple(send.shape))
    for r, (lo, hi) in enumerate(ranges):
        if r != rank and hi > lo:
            table[:, lo:hi] = recv[r, :, :hi - lo]
    return table


class _DevicePtr:
    """__cuda_array_interface__ view of a raw device pointer (the engine's resident table)."""

    def __init__(self, ptr, shape, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def engine_ctrl_tensor(eng, count, N, M, device):
    """torch view (no copy) of the engine's control-point table [count, N, 3, 6M]."""
    return torch.as_tensor(_DevicePtr(eng.device_ctrl_ptr(), (count, N, 3, 6 * M)), device=device)


def engine_status_tensor(eng, count, device):
    """torch view (no copy) of the engine's per-mission status words [count] (int32)."""
    return torch.as_tensor(_DevicePtr(eng.device_status_ptr(), (count,), "<i4"), device=device)


def merge_status(status, group=None):
    """A batch that failed on one rank fails the mission on every rank: element-wise MAX of the status words
    (0 OK < 1 infeasible < 2 not converged < 3 bad input), as RBPPlanner::update() returns false when ANY batch fails
    (rbp_planner.hpp L158-L161).  One small all-reduce per solve."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(status, op=dist.ReduceOp.MAX, group=group)
    return status


def jacobi_attach_peers(eng, prob, group=None):
    """One-time set-up of the fused exchange: upload + assemble (allocates the tables), export the IPC handles of the
    two table buffers and the flag words, gather them over the process group (host side, any backend) and open the
    peers' buffers.  Afterwards `jacobi_solve(..., fused=True)` needs no collective at all."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    eng.upload(prob)
    eng.assemble()
    err = None
    try:
        mine = eng.peer_export()
    except RuntimeError as ex:
        mine, err = None, ex
    handles = [None] * world
    if world > 1:
        dist.all_gather_object(handles, mine, group=group)
    else:
        handles[0] = mine
    if err is None and all(h is not None for h in handles):
        try:
            eng.peer_attach(rank, world, handles)
        except RuntimeError as ex:
            err = ex
    elif err is None:
        err = RuntimeError("a peer could not export its tables")
    if world > 1:   # every rank learns whether every rank attached (also: nobody stores remotely before all have opened)
        oks = [None] * world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks) and err is None:
            err = RuntimeError("peer attach failed on rank(s) %s" % [r for r, o in enumerate(oks) if not o])
    if err is not None:
        raise err


def jacobi_solve(eng, prob, sweeps, group=None, device=None, fused=False, upload=True):
    """Jacobi mode on the ranks of `group`: inputs replicated, batches sharded.
    upload=False: the inputs of `prob` are already resident on the device (a previous call uploaded them); only the
    assembly kernel runs again (it resets `dummy` to the initial trajectory).
    fused=False: one NCCL all-gather of the solved control points per sweep (`exchange_ctrl`).
    fused=True (after `jacobi_attach_peers`): the sweep kernel's epilogue stores the solved control points into every
    rank's next table over NVLink peer memory and raises a flag there; no collective call, no host synchronisation."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bs, nbatch = prob.effective_batching()
    b0, b1 = batch_range(nbatch, world, rank)
    if upload:
        eng.upload(prob)
    eng.assemble()
    if fused:
        for _ in range(sweeps):
            eng.run_jacobi_fused(b0, b1)
    else:
        table = engine_ctrl_tensor(eng, prob.count, prob.N, prob.M, device) if world > 1 else None
        for _ in range(sweeps):
            eng.run_jacobi_range(b0, b1)
            if world > 1:
                eng.sync()                      # the engine has its own stream; the collective runs on torch's
                exchange_ctrl(table, prob.N, bs, nbatch, group)
                torch.cuda.current_stream().synchronize()
        if world > 1:
            eng.convert()                       # the sweep converted before the exchange: redo it on the complete table
    if world > 1:                               # every rank reports the mission status every rank would
        eng.sync()
        merge_status(engine_status_tensor(eng, prob.count, device), group)
        torch.cuda.current_stream().synchronize()
    return eng
