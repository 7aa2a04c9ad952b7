"""Multi-GPU check of the Jacobi exchange (run under torchrun, one rank per GPU):
fused peer-store exchange (rbpe_run_jacobi_fused) vs sweep + NCCL all-gather: identical tables, and device time of both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from swarm_simulator_b200 import engine as E, synth, dist as D


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    missions, sweeps = int(os.environ.get("J_MISSIONS", 64)), int(os.environ.get("J_SWEEPS", 2))
    ms = [synth.synth_mission(64, 5, 0.2, 3000 + i) for i in range(4)]
    prob = E.PackedProblem(synth.pack([ms[i % 4] for i in range(missions)]), sequential=True, batch_size=1, iteration=sweeps)
    # A: sweep + all-gather
    ea = E.Engine(device=local)
    D.jacobi_solve(ea, prob, sweeps, device=dev); ea.sync()
    dist.barrier(); torch.cuda.synchronize()
    ea.timer_start()
    for _ in range(5):
        D.jacobi_solve(ea, prob, sweeps, device=dev)
    ms_a = ea.timer_stop() / 5
    ra = ea.download(prob)
    # B: fused peer-store exchange
    eb = E.Engine(device=local)
    D.jacobi_attach_peers(eb, prob)
    D.jacobi_solve(eb, prob, sweeps, fused=True); eb.sync()
    dist.barrier(); torch.cuda.synchronize()
    eb.timer_start()
    for _ in range(5):
        D.jacobi_solve(eb, prob, sweeps, fused=True)
    ms_b = eb.timer_stop() / 5
    rb = eb.download(prob)
    eb.peer_status()
    same = bool(np.array_equal(ra.ctrl, rb.ctrl))
    t = torch.tensor([ms_a, ms_b], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    print("rank %d/%d: fused == all-gather tables: %s | all-gather %.3f ms, fused %.3f ms per %d-sweep solve of %d missions (max over ranks)" % (
        rank, world, same, t[0].item(), t[1].item(), sweeps, missions), flush=True)
    assert same
    ea.close(); eb.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
