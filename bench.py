#!/usr/bin/env python
"""bench.py -- agent-QPs/sec of the RBP trajectory-QP hot path (BASELINE.json metric) on N B200s of one node.

A "step" = one pass of the hot path (assembly kernel + PDIP kernel + conversion kernel) over one batch of
synthetic 64-agent, 5-segment random-forest missions (BASELINE.json configs[2]), every agent's QP solved
in the reference's sequential order (plan/sequential=true, plan/batch_size=1: RBPPlanner semantics, Gauss-Seidel through
`dummy`).  Missions are independent, so ranks shard missions (weak scaling, no data-path collective).

  value  : whole-job agent-QPs/s with the inputs resident in HBM (CUDA events on the engine's stream, max over ranks)
  e2e    : the same through the C ABI call rbpe_solve_many() with pinned HOST buffers (H2D + kernels + D2H inside)
  --impl reference : the CPU oracle (oracle/, the restatement of the reference's CPLEX path) on all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "agent-QPs/sec (whole box) for 64-agent RBP random-forest"
UNIT = "agent-QPs/s"
N_AGENTS, M_SEG, RHO, CONFIG_ID = 64, 5, 0.2, 3


def dense_flops_per_iter(b, M):
    """dense reduced-KKT flops of one interior-point iteration of a batch QP (nv=18bM, ne=9b(M+1))."""
    nv, ne = 18.0 * b * M, 9.0 * b * (M + 1)
    return nv ** 3 / 3 + nv * nv * ne + nv * ne * ne + ne ** 3 / 3


def structured_flops_per_iter(b, M, N, rows=None):
    """What the kernel executes: block tridiagonal Cholesky over M-1 knot blocks of 9b + 3 row passes of ~44 flops per
    KEPT row (residual, affine + corrector right-hand side, final ratio test; rows=None counts every row of
    populatebyrow, an upper bound)."""
    kb = 9.0 * b
    if rows is None:
        rows = b * (6 * M - 6) * (6 + (N - b)) + b * (b - 1) / 2 * (6 * M - 6)
    return (M - 1) * (kb ** 3 / 3) + (M - 2) * 2 * kb ** 3 + 4 * (M - 1) * 2 * kb * kb * 2 + 3 * rows * 44


WORKLOAD = ("64 agents, random forest rho=0.2, 5-segment degree-5, sequential batch_size=1 "
            "(BASELINE configs[2]; reference Gauss-Seidel order)")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def make_pool(n, rank, pack="cfg3"):
    """`n` distinct seeded missions for `rank` from the committed pack (tests/golden/missions_*.npz, written by
    tests/golden/make_missions.py from swarm_simulator_b200/synth.py; seeds 1000 * config + trial)."""
    from swarm_simulator_b200 import synth
    z = np.load(os.path.join(GOLDEN, "missions_%s.npz" % pack))
    total = len(z["seed"])
    sel = [(rank * n + i) % total for i in range(n)]
    return synth.load_pack(os.path.join(GOLDEN, "missions_%s.npz" % pack), select=sel)


def kept_rows_mean(missions, ctrls):
    """Mean number of inequality rows per agent-QP that survive the presolve (rows on fixed control points and rows the
    bound-based redundancy test drops are never touched by the kernel): numpy restatement on the final control points."""
    tot, nqp = 0, 0
    for m, ctrl in zip(missions, ctrls):
        N, M = m["N"], m["M"]
        x = np.transpose(ctrl, (0, 2, 1))                      # [N, 6M, 3]
        seg_of = np.repeat(np.arange(M), 6)
        ub = np.zeros((N, 6 * M, 3)); lb = np.zeros((N, 6 * M, 3))
        for qi, (boxes, tend) in enumerate(m["sfc"]):
            bi = 0
            for mm in range(M):
                while bi < len(tend) and tend[bi] < m["T"][mm + 1]:
                    bi += 1
                b = boxes[min(bi, len(tend) - 1)]
                ub[qi, mm * 6:(mm + 1) * 6] = b[3:]; lb[qi, mm * 6:(mm + 1) * 6] = b[:3]
        live = np.ones(6 * M, bool); live[:3] = False; live[-3:] = False
        qi_, qj_ = np.triu_indices(N, 1)
        nrm = m["rsfc_n"][:, seg_of, :].astype(np.float64)    # [P, 6M, 3] (ri == m for these missions: one RSFC entry per segment)
        for qa in range(N):
            sel_lo = qi_ == qa; sel_hi = qj_ == qa
            n = np.concatenate([nrm[sel_lo], -nrm[sel_hi]])   # qa < qo: +n, qa > qo: -n
            other = np.concatenate([qj_[sel_lo], qi_[sel_hi]])
            h = -(m["radius"][qa] + m["radius"][other])[:, None] + (n * x[other]).sum(-1)
            amax = np.maximum(n * ub[qa][None], n * lb[qa][None]).sum(-1)
            keep = ~(amax < h - 1e-9 * np.maximum(1.0, np.abs(h)))
            tot += int(keep[:, live].sum()) + 6 * int(live.sum())
            nqp += 1
    return tot / max(nqp, 1)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def pin(packed):
    """Move the mission arrays into pinned host memory (torch is plumbing here: pinned allocator only)."""
    import torch
    out = dict(packed)
    keep = []
    for k, v in packed.items():
        if isinstance(v, np.ndarray) and v.size:
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            keep.append(t)
            out[k] = t.numpy()
    out["_pinned"] = keep
    return out


def run_reference(args, rank, world):
    """Reference arm: the CPU oracle (float64 restatement of the reference's CPLEX path) on all host threads."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    import oracle_util
    cores = os.cpu_count() or 1
    pool = make_pool(args.pool, 0)
    nmis = args.ref_missions or 16 * cores            # >= 16 missions per host thread per step (dynamic OpenMP schedule)
    probs = [oracle_util.oracle_problem(pool[i % len(pool)], sequential=True, batch_size=1) for i in range(nmis)]
    for _ in range(min(args.warmup, 2)):
        oracle.update_many(probs[:max(len(pool), cores)], nthreads=cores)
    t0 = time.perf_counter()
    bad = 0
    for _ in range(args.steps):
        _, _, st = oracle.update_many(probs, nthreads=cores)
        bad += int((st != 0).sum())
    dt = (time.perf_counter() - t0) / args.steps
    bad //= max(args.steps, 1)
    nqp = (len(probs) - bad) * N_AGENTS          # only the QPs of missions whose update() succeeds count
    val = nqp / dt
    sample = "%d missions x %d agents per step (%d distinct, seeds %d..), oracle.update_many, OpenMP over missions, %d threads" % (
        len(probs), N_AGENTS, len(pool), 1000 * CONFIG_ID, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "missions_per_step": len(probs), "agent_qps_per_step": nqp,
                   "failed_missions_per_step": bad, "distinct_missions": len(pool),
                   "sample_note": "bounded sample of the engine arm's workload (same missions, same order, same tolerances)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU oracle (not CPLEX: proprietary, absent)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--missions", type=int, default=0, help="missions per rank per step (0 = 48 per SM: 7104 = three pipeline chunks of 16 warp-resident QP chains per SM)")
    ap.add_argument("--pool", type=int, default=64, help="distinct synthetic missions per rank (from the committed pack, tiled to --missions)")
    ap.add_argument("--ref-missions", type=int, default=0, help="missions per step of the CPU arm (0 = 16 per host thread)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--joint-missions", type=int, default=592, help="missions of the joint-batch leg (configs[1]); 0 = skip")
    ap.add_argument("--jacobi-missions", type=int, default=64, help="missions (replicated on every rank) of the Jacobi leg; 0 = skip")
    ap.add_argument("--jacobi-sweeps", type=int, default=2)
    ap.add_argument("--jacobi-missions-large", type=int, default=1184, help="missions of the throughput-bound Jacobi leg; 0 = skip")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the BASELINE configs[3] / configs[4] legs")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-mission latency leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # rank 0 prints exactly ONE JSON line on stdout: native libraries (NCCL's version banner) write to fd 1 directly, so
    # fd 1 is pointed at stderr for the whole run and the result goes to the saved descriptor
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import __graft_entry__ as G
    if rank == 0:
        G.build()
    barrier()
    from swarm_simulator_b200 import engine as E, synth

    count = args.missions or 48 * 148
    pool = make_pool(args.pool, rank)
    packed = pin(synth.pack([pool[i % len(pool)] for i in range(count)]))
    prob = E.PackedProblem(packed, sequential=True, batch_size=1)
    res = E.Result(prob, want_ctrl=True, pinned=True)    # page-locked like the inputs; coefficients (msgs_traj_coef) AND the updated `dummy` come back
    eng = E.Engine(device=local)
    h2d, d2h = prob.h2d_bytes(), res.d2h_bytes()
    nqp = count * N_AGENTS
    l2_note = "inputs_larger_than_L2" if h2d > 126e6 else "l2_flushed_between_steps"
    flush_buf = None if h2d > 126e6 else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush():
        if flush_buf is not None:
            flush_buf.zero_()
            torch.cuda.synchronize()

    # ---- warm-up (e2e path: exercises H2D, all three kernels, D2H) ----
    for _ in range(args.warmup):
        r = eng.solve_many(prob, result=res)
    # a mission whose update() fails is reported (config.failed_missions_per_step) and its QPs do not count; the committed
    # packs contain none (tests/test_feasibility_classifier.py)
    assert r.rc in (E.OK, E.INFEASIBLE, E.NOT_CONVERGED), (r.rc, eng.last_error())
    failed_missions = int((res.status != 0).sum())
    nqp_ok = (count - failed_missions) * N_AGENTS
    iters_total = int(res.qp_iters.sum())
    iters_mean = float(res.qp_iters.mean())
    launches0 = eng.timing()["kernel_launches"]

    clocks = ClockSampler(local)
    clocks.start()
    # ---- region A: inputs resident in HBM, K steps back to back on the engine's stream ----
    eng.upload(prob)
    eng.run()            # untimed: sizes the scratch arena of the resident path (the warm-up above went through the pipelined call)
    eng.sync()
    barrier()
    t_dev = 0.0
    for _ in range(args.steps):
        flush()
        eng.timer_start()
        eng.run()
        dt_step = eng.timer_stop()
        t_dev += dt_step
        if rank == 0:
            print("resident step: %.2f ms  %s" % (dt_step, eng.timing()), file=sys.stderr)
    barrier()
    eng.download(prob, res)
    kernel_ms = eng.timing()["solve_ms"]          # PDIP + conversion kernels of the last step (events on the stream)
    ms_step = max_over_ranks(t_dev / args.steps)
    # ---- region B: end to end through rbpe_solve_many() with pinned host buffers ----
    barrier()
    t_e2e = 0.0
    for _ in range(args.steps):
        flush()
        eng.timer_start()
        r = eng.solve_many(prob, result=res)
        t_e2e += eng.timer_stop()
    barrier()
    assert r.rc in (E.OK, E.INFEASIBLE, E.NOT_CONVERGED)
    ms_e2e = max_over_ranks(t_e2e / args.steps)
    clk = clocks.stop()
    launches = eng.timing()["kernel_launches"] - launches0
    kernel_ms = max_over_ranks(kernel_ms)

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nqp_all = sum_over_ranks(float(nqp_ok))          # converged agent-QPs per step over all ranks
    failed_all = int(sum_over_ranks(float(failed_missions)))
    value = nqp_all / (ms_step * 1e-3)
    e2e = nqp_all / (ms_e2e * 1e-3)

    # ---- secondary leg: BASELINE configs[1] (16 agents, forest 0.2, 5 segments, ONE joint QP per mission: nv = 1440, K = 2304);
    # CTA-per-QP kernel, block tridiagonal factorisation of 144 x 144 blocks on the FP64 tensor pipe (DMMA) ----
    joint = None
    if args.joint_missions > 0:
        jm = make_pool(min(64, args.joint_missions), rank, "cfg2")
        jp = E.PackedProblem(pin(synth.pack([jm[i % len(jm)] for i in range(args.joint_missions)])), sequential=False, batch_size=16)
        je = E.Engine(device=local)
        je.upload(jp); je.run(); je.sync()
        barrier()
        je.timer_start()
        for _ in range(2):
            je.run()
        ms_joint = max_over_ranks(je.timer_stop() / 2)
        jr = je.download(jp)
        it_j = int(jr.qp_iters.sum())
        joint = {"workload": "16 agents, random forest rho=0.2, 5-segment, one joint batch of 16 (BASELINE configs[1])",
                 "value": world * args.joint_missions * 16 / (ms_joint * 1e-3), "unit": UNIT, "ms_per_step": ms_joint,
                 "missions_per_gpu": args.joint_missions, "ipm_iterations_mean": float(jr.qp_iters.mean()), "failed": int((jr.status != 0).sum()),
                 "dense_tflops": dense_flops_per_iter(16, M_SEG) * it_j / (ms_joint * 1e-3) / 1e12,
                 "distinct_missions_per_gpu": len(jm),
                 "kernel": "pdip_kernel (256 threads per QP; DMMA m8n8k4 block Cholesky, rbpe_blockla.cuh)"}
        je.close()

    # ---- the other BASELINE configurations as single-GPU legs (rank 0 only; resident inputs, one warm + one timed pass):
    # configs[3] = 256 agents, rho 0.4, sequential batches of 32; configs[4] = 1024 agents, rho 0.1 .. 0.5 (batch size
    # unspecified in BASELINE.json: the per-agent b = 1 of the headline and the b = 32 of configs[3] are both run) ----
    others = None
    if not args.no_other_configs and rank == 0:
        others = []
        # (name, pack, distinct missions, missions in flight, sequential, batch size): missions are tiled so that the legs
        # are not bound by the latency of a single mission chain (one CTA / one warp per mission)
        for name, pack, nm, tile, seq, bsz in (("configs[3]: 256 agents, rho=0.4, sequential batch_size=32", "cfg4", 32, 148, True, 32),
                                               ("configs[4]: 1024 agents, rho=0.1..0.5, sequential batch_size=1", "cfg5", 5, 35, True, 1),
                                               ("configs[4]: 1024 agents, rho=0.1..0.5, sequential batch_size=32", "cfg5", 5, 35, True, 32)):
            try:
                od = make_pool(nm, 0, pack)
                om = [od[i % nm] for i in range(tile)]
                op_ = E.PackedProblem(pin(synth.pack(om)), sequential=seq, batch_size=bsz)
                oe = E.Engine(device=local)
                oe.upload(op_); oe.run(); oe.sync()
                oe.timer_start(); oe.run(); ms_o = oe.timer_stop()
                orr = oe.download(op_)
                nq = sum(m["N"] for m, st in zip(om, orr.status) if st == 0)
                others.append({"workload": name, "missions": len(om), "distinct_missions": nm, "value": nq / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o,
                               "failed_missions": int((orr.status != 0).sum()), "ipm_iterations_mean": float(orr.qp_iters.mean()),
                               "dense_tflops": dense_flops_per_iter(bsz, M_SEG) * float(orr.qp_iters.sum()) / (ms_o * 1e-3) / 1e12})
                oe.close()
            except Exception as ex:   # a leg must never take the headline down with it
                others.append({"workload": name, "error": repr(ex)})

    # ---- latency leg (rank 0): the reference's own call pattern -- ONE mission per update() (swarm_traj_planner_rbp.cpp
    # L110-L111) -- at the launch default batch_size = 4 (plan_rbp_random_forest.launch L63-L65) and at batch_size = 1, plus one
    # Jacobi sweep of one mission (64 independent QPs: the latency of a single QP).  Kernel time (assemble + solve + convert) with
    # inputs resident, and end to end through rbpe_solve_many with host buffers.  One-agent batches of a handful of missions run
    # on the several-warps-per-QP latency kernel (rbpe_pdip1x.cuh), joint batches on 16-warp CTAs (rbpe_api.cu).
    latency = None
    if not args.no_latency and rank == 0:
        latency = []
        lm = make_pool(1, 0)
        for name, bsz, mode in (("one 64-agent mission, sequential batch_size=4 (launch default)", 4, E.MODE_GAUSS_SEIDEL),
                                ("one 64-agent mission, sequential batch_size=1", 1, E.MODE_GAUSS_SEIDEL),
                                ("one Jacobi sweep of one 64-agent mission (64 independent QPs)", 1, E.MODE_JACOBI)):
            try:
                lp = E.PackedProblem(pin(synth.pack(lm)), sequential=True, batch_size=bsz)
                le = E.Engine(device=local)
                le.upload(lp); le.run(mode); le.sync()
                best = 1e30
                for _ in range(5):
                    le.timer_start(); le.run(mode); best = min(best, le.timer_stop())
                lr = le.download(lp)
                solver = le.last_solver()
                t0 = time.perf_counter()
                for _ in range(3):
                    lr2 = le.solve_many(lp, mode=mode)
                ms_call = (time.perf_counter() - t0) / 3 * 1e3
                latency.append({"workload": name, "kernel_ms": best, "call_ms_host_buffers": ms_call, "qps": int((lr.qp_status == 0).sum()),
                                "solver": {1: "pdip1_kernel", 2: "pdip1x_kernel", 3: "pdip_kernel"}.get(solver[0], "?") + " (%d threads per QP)" % solver[1],
                                "ipm_iterations_mean": float(lr.qp_iters.mean()), "failed_missions": int((lr.status != 0).sum())})
                le.close()
            except Exception as ex:
                latency.append({"workload": name, "error": repr(ex)})

    # ---- secondary leg: Jacobi mode (north-star's agent sharding): the SAME missions on every rank, each rank solves its
    # range of agents of every mission against the frozen table; the exchange of the solved control points is fused into the
    # sweep kernel (peer stores over NVLink), with the NCCL all-gather variant timed beside it ----
    jac = None
    if args.jacobi_missions > 0:
        from swarm_simulator_b200 import dist as D
        jpool = make_pool(min(args.pool, 64), 0)
        jprob = E.PackedProblem(pin(synth.pack([jpool[i % len(jpool)] for i in range(args.jacobi_missions)])),
                                sequential=True, batch_size=1, iteration=args.jacobi_sweeps)
        dev = torch.device("cuda", local)

        def jacobi_leg(fused, jprob_=None, resident=False):
            jp_ = jprob_ or jprob
            jeng = E.Engine(device=local)
            if fused:
                D.jacobi_attach_peers(jeng, jp_)                            # one-time exchange of IPC handles (host side)
            D.jacobi_solve(jeng, jp_, args.jacobi_sweeps, device=dev, fused=fused)   # warm-up (also allocates, uploads)
            jeng.sync()
            l0 = jeng.timing()["kernel_launches"]
            barrier()
            jeng.timer_start()                               # CUDA events on the engine's stream; every phase in between
            for _ in range(args.steps):                      # (H2D, sweeps, exchange) is stream- or host-synchronised
                D.jacobi_solve(jeng, jp_, args.jacobi_sweeps, device=dev, fused=fused, upload=not resident)
            ms_j = max_over_ranks(jeng.timer_stop() / args.steps)
            barrier()
            jr = jeng.download(jp_)
            if fused:
                jeng.peer_status()
            out = (ms_j, int(jeng.timing()["kernel_launches"] - l0), int((jr.status != 0).sum()))
            jeng.close()
            return out

        fused_ok = world > 1
        try:
            ms_j, jl, jfail = jacobi_leg(fused=fused_ok)
        except RuntimeError as ex:      # e.g. no peer access between two devices: report the collective variant instead
            print("fused Jacobi exchange unavailable (%s); timing the all-gather variant only" % ex, file=sys.stderr)
            fused_ok = False
            ms_j, jl, jfail = jacobi_leg(fused=False)
        jac = {"value": args.jacobi_missions * N_AGENTS * args.jacobi_sweeps / (ms_j * 1e-3), "unit": UNIT,
               "scaling": "strong", "missions": args.jacobi_missions, "sweeps": args.jacobi_sweeps, "ms_per_step": ms_j,
               "exchange": ("fused into the sweep kernel: stores of the solved control points (%d B per agent) into every rank's "
                            "next table over NVLink peer memory + flag words; no collective call" % (18 * M_SEG * 8)) if fused_ok
                           else ("single GPU: none" if world == 1 else "1 NCCL all-gather of control points per sweep"),
               "includes": "H2D of inputs, assembly, %d sweeps, exchange" % args.jacobi_sweeps,
               "gpu_launches": jl, "failed": jfail}
        if fused_ok:   # the baseline it replaces: sweep kernel, then one NCCL all-gather of control points per sweep
            ms_ag, _, _ = jacobi_leg(fused=False)
            jac["allgather_baseline"] = {"value": args.jacobi_missions * N_AGENTS * args.jacobi_sweeps / (ms_ag * 1e-3),
                                         "ms_per_step": ms_ag, "collective": "1 NCCL all-gather of control points per sweep"}
        # the same leg with the inputs resident (no H2D inside): what remains is assembly + sweeps + exchange
        try:
            ms_r, _, _ = jacobi_leg(fused=fused_ok, resident=True)
            jac["resident"] = {"value": args.jacobi_missions * N_AGENTS * args.jacobi_sweeps / (ms_r * 1e-3), "ms_per_step": ms_r}
        except RuntimeError as ex:
            print("resident Jacobi leg skipped: %s" % ex, file=sys.stderr)
        # ... and a batch large enough to be throughput bound on 8 GPUs (the 64-mission leg is bound by the latency of one
        # QP: 4096 QPs are less than two waves of one GPU's 2368 resident warps)
        if args.jacobi_missions_large > 0:
            try:
                lpool = make_pool(min(args.pool, 64), 0)
                lprob = E.PackedProblem(pin(synth.pack([lpool[i % len(lpool)] for i in range(args.jacobi_missions_large)])),
                                        sequential=True, batch_size=1, iteration=args.jacobi_sweeps)
                ms_l, _, lfail = jacobi_leg(fused=fused_ok, jprob_=lprob, resident=True)
                jac["large_batch"] = {"missions": args.jacobi_missions_large, "ms_per_step": ms_l, "failed": lfail,
                                      "value": args.jacobi_missions_large * N_AGENTS * args.jacobi_sweeps / (ms_l * 1e-3),
                                      "note": "inputs resident; same missions on every rank, agents sharded"}
            except RuntimeError as ex:
                print("large Jacobi leg skipped: %s" % ex, file=sys.stderr)

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "missions_per_step_per_gpu": count, "agent_qps_per_step": int(nqp_all), "failed_missions_per_step": failed_all, "distinct_missions_per_gpu": len(pool),
                   "parallelism": "missions sharded over %d GPU(s), no collective" % world, "cache": l2_note,
                   "ipm_iterations_mean": iters_mean, "tol_gap": 1e-10, "tol_res": 1e-9},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(launches),
        "clocks": clk,
    }
    if joint:
        out["joint_batch"] = joint
    if others:
        out["other_configs"] = others
    if latency is not None:
        out["latency"] = latency
    if jac:
        out["jacobi_mode"] = jac
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
        tf32_peak, tf32_src = bf16 / 2.0, ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 dense proxy)" if peaks
                                           else "fallback 1.4 PFLOP/s bf16 / 2")
        fp64_peak = None
        try:   # the two peaks this path is judged against, measured here (after the timed regions): cuBLAS TF32 and FP64 GEMMs
            def gemm_tflops(dtype, n, reps):
                a_ = torch.randn(n, n, device="cuda", dtype=dtype); b_ = torch.randn(n, n, device="cuda", dtype=dtype)
                torch.matmul(a_, b_); torch.cuda.synchronize()
                best = 1e9
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); torch.matmul(a_, b_); e1.record(); torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                return 2.0 * n ** 3 / (best * 1e-3) / 1e12
            torch.backends.cuda.matmul.allow_tf32 = True
            tf32_peak, tf32_src = gemm_tflops(torch.float32, 8192, 5), "measured in this run: torch.matmul TF32 8192^3, best of 5 (CUDA events)"
            fp64_peak = gemm_tflops(torch.float64, 4096, 3)
        except Exception as ex:   # keep the fallback
            print("peak measurement skipped: %r" % (ex,), file=sys.stderr)
        f_dense = dense_flops_per_iter(1, M_SEG) * iters_total          # per rank per step
        kept = kept_rows_mean(pool[:4], [res.ctrl[i] for i in range(4)])   # rows the kernel really walks (numpy, 4 missions)
        f_struct = structured_flops_per_iter(1, M_SEG, N_AGENTS, rows=kept) * iters_total
        ach = f_dense / (kernel_ms * 1e-3) / 1e12
        traffic, ncu, prof_name = None, {}, None
        for prof_name in ("r2_pdip1_ncu.json", "r1_pdip1_ncu.json"):
            try:   # DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture, scaled to this
                # launch's mission count (missions are independent work items of identical shape)
                prof = json.load(open(os.path.join(ROOT, "profiles", prof_name)))
                traffic = prof["dram_bytes_per_launch"] * count / float(prof.get("missions", 2368))
                ncu = {k: float(v["value"].replace(",", "")) for k, v in prof["metrics"].items() if k.endswith(".pct") or "pct_of_peak" in k}
                break
            except Exception:
                continue
        out["roofline"] = {
            "bound": "tensor", "kernel": "pdip1_kernel", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": ach / tf32_peak, "traffic": traffic, "peak_source": tf32_src,
            "algorithmic": "dense reduced-KKT flops (SURVEY 8d: nv^3/3 + nv^2 ne + nv ne^2 + ne^3/3 = 0.995 MFLOP per iteration at "
                           "b=1, M=5) x iterations executed: %.3e per launch" % f_dense,
            "algorithmic_bytes": h2d + d2h,
            "executed_structured_tflops": f_struct / (kernel_ms * 1e-3) / 1e12,
            "executed_structured_note": "block tridiagonal factor / solves + 44 flops per KEPT inequality row and pass (3 passes per iteration) (%.0f of the "
                                        "%d rows of populatebyrow survive the presolve on average)" % (kept, (6 * M_SEG - 6) * (N_AGENTS + 5)),
            "fp64_peak_tflops": fp64_peak,
            "executed_fraction_of_fp64_peak": (f_struct / (kernel_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
            "kernel_ms_per_launch": kernel_ms,
            "traffic_note": "ncu dram__bytes_read+write of one launch (profiles/%s) scaled to %d missions; inputs + outputs are "
                            "%.2f GB. The excess is working-set overflow of the 126 MB L2 (2368 mission chains in flight x (46 KB "
                            "control-point table + ~25 KB of row state)), not re-reads by the kernel's loops; DRAM is at ~5 %% of its "
                            "bandwidth, the kernel is instruction-issue / fetch bound (DESIGN.md section 6)" % (prof_name, count, (h2d + d2h) / 1e9),
            "ncu": {"issue_active_pct": ncu.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "fp64_pipe_pct": ncu.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                    "l2_hit_pct": ncu.get("lts__t_sector_hit_rate.pct")},
            "note": "one-agent QPs (K = 144) reduce to 36x36 block tridiagonal systems: FP64 SIMT, issue/latency bound, no tensor "
                    "work by design (SURVEY section 7 'tiny matrices'); the joint-batch kernel (b > 1) runs its factorisation on the "
                    "FP64 tensor pipe (DMMA), see the joint_batch leg and DESIGN.md section 4",
        }
        if world == 1 and not args.no_cpu_baseline:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle
            import oracle_util
            cores = os.cpu_count() or 1
            nmis = 16 * cores
            probs = [oracle_util.oracle_problem(pool[i % len(pool)], sequential=True, batch_size=1) for i in range(nmis)]
            oracle.update_many(probs[:max(cores, 8)], nthreads=cores)
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < args.cpu_seconds:
                oracle.update_many(probs, nthreads=cores)
                n += len(probs)
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": n * N_AGENTS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "%d missions (%d distinct, %d per call = 16 per thread) in %.1f s, OpenMP over missions"
                                             % (n, len(pool), len(probs), dt),
                                   "note": "CPU oracle (not CPLEX: proprietary, absent)"}
        sys.stdout.flush()
        os.write(result_fd, (json.dumps(out) + "\n").encode())
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
