// emu_driver.cpp -- TEST INFRASTRUCTURE ONLY.  Runs the product kernels (rbpe_kernels.cuh) under the fiber emulator
// of cuda_emu.h on host memory, with the same launch sequence as rbpe_api.cu, so that kernel logic can be debugged
// against the oracle without a GPU.  Not linked into, loaded by, or a fallback of the product library.
#include <stdio.h>
#include <stdint.h>

#include "cuda_emu.h"
#define RBPE_EMU 1
#include "../../swarm_simulator_b200/csrc/rbpe_kernels.cuh"
#include "../../include/rbpe.h"

using namespace rbpe;

extern "C" int emu_solve_many(const rbpe_problem *p, int count, int mode, rbpe_result *r, size_t smem_bytes,
                              int max_iter, double tol_gap, double tol_res, int threads) {
    const int N = p->N, M = p->M;
    int bs, nbatch;
    rbpe_set_batch(N, p->sequential, p->batch_size, p->batch_iter, &bs, &nbatch);
    const long P = (long)N * (N - 1) / 2;
    const size_t per = (size_t)N * 18 * M;
    std::vector<double> segbox((size_t)count * N * M * 6), segmat((size_t)count * M * SEGMAT), ctrl(count * per),
        coef(count * per), frozen(count * per);
    std::vector<float> reln((size_t)count * (P > 0 ? P : 1) * M * 3);
    std::vector<int> status(count, 0);
    AssembleArgs A;
    A.count = count; A.N = N; A.M = M; A.sequential = p->sequential;
    A.T = p->T; A.sfc_offs = p->sfc_offs; A.sfc_base = p->sfc_base; A.sfc_box = p->sfc_box; A.sfc_t = p->sfc_t;
    A.rsfc_n = p->rsfc_n; A.rsfc_t = p->rsfc_t; A.init_traj = p->init_traj;
    A.segbox = segbox.data(); A.reln = reln.data(); A.ctrl = ctrl.data(); A.segmat = segmat.data();
    A.status = status.data();
    emu::launch([&] { assemble_kernel(A); }, 4, 64, 0);

    int nrec = p->iteration * nbatch;
    if (nrec < 1) nrec = 1;
    std::vector<double> obj((size_t)count * nrec, 0.0), res((size_t)count * nrec * 4, 0.0);
    std::vector<int> its((size_t)count * nrec, 0), qst((size_t)count * nrec, 0);
    SolveArgs S;
    S.npeer = 0; S.peer_rank = 0; S.sweep_id = 0; S.done_counter = nullptr; S.work_items = 0;
    S.count = count; S.N = N; S.M = M; S.bs = bs; S.nbatch = nbatch; S.iteration = p->iteration;
    S.sequential = p->sequential; S.mode = mode; S.batch_begin = 0; S.batch_end = nbatch; S.rec_offset = 0;
    S.max_iter = max_iter > 0 ? max_iter : 100;
    S.tol_gap = tol_gap > 0 ? tol_gap : 1e-10;
    S.tol_res = tol_res > 0 ? tol_res : 1e-9;
    S.start = p->start; S.goal = p->goal; S.radius = p->radius;
    S.segbox = segbox.data(); S.reln = reln.data(); S.segmat = segmat.data();
    S.ctrl = ctrl.data(); S.ctrl_frozen = frozen.data();
    S.qp_obj = obj.data(); S.qp_iters = its.data(); S.qp_status = qst.data(); S.qp_res = res.data();
    S.nrec = nrec; S.status = status.data();
    S.scratch_stride = scratch_doubles(N, M, bs);
    S.smem_bytes = (unsigned)smem_bytes;
    S.panel_bytes = 0;
    // same kernel selection as rbpe_api.cu: one-agent batches -> warp-per-QP kernel, else CTA-per-QP kernel
    // threads < 0: the several-warps-per-QP latency kernel with -threads threads per QP
    const bool lat_kernel = (bs == 1) && threads < 0;
    const bool warp_kernel = (bs == 1) && threads != 64;   // threads == 64 forces the CTA kernel (A/B in tests)
    const int wpc = 2;
    auto launch = [&](long units) {
        if (lat_kernel) {
            const int nw = -threads / 32;
            S.scratch_stride = 0;
            S.scratch = nullptr;
            S.smem_bytes = (unsigned)(x1_smem_doubles(N, M, nw) * 8);
            emu::launch([&] { pdip1x_kernel(S); }, (unsigned)units, nw * 32, S.smem_bytes);
        } else if (warp_kernel) {
            long grid = (units + wpc - 1) / wpc;
            S.scratch_stride = w1_scratch_doubles(N, M);
            S.smem_bytes = (unsigned)(wpc * w1_smem_doubles(M) * 8);
            std::vector<double> scratch(S.scratch_stride * grid * wpc);
            S.scratch = scratch.data();
            emu::launch([&] { pdip1_kernel(S); }, (unsigned)grid, wpc * 32, S.smem_bytes);
        } else {
            S.scratch_stride = scratch_doubles(N, M, bs);
            S.panel_bytes = bs > 1 ? (unsigned)(bla_panel_doubles((int)kp_of(bs)) * 8) : 0;   // same carve-out as rbpe_api.cu
            S.smem_bytes = (unsigned)smem_bytes + S.panel_bytes;
            std::vector<double> scratch(S.scratch_stride * units);
            S.scratch = scratch.data();
            emu::launch([&] { pdip_kernel(S); }, (unsigned)units, threads, S.smem_bytes);
        }
    };
    if (nbatch > 0) {
        if (mode == 0) {
            launch(count);
        } else {
            for (int it = 0; it < p->iteration; it++) {
                frozen = ctrl;
                S.rec_offset = it * nbatch;
                launch((long)count * nbatch);
            }
        }
    }
    ConvertArgs C;
    C.count = count; C.N = N; C.M = M; C.ctrl = ctrl.data(); C.segmat = segmat.data(); C.coef = coef.data();
    emu::launch([&] { convert_kernel(C); }, 4, 64, 0);

    memcpy(r->coef, coef.data(), coef.size() * 8);
    if (r->ctrl) memcpy(r->ctrl, ctrl.data(), ctrl.size() * 8);
    if (r->qp_obj) memcpy(r->qp_obj, obj.data(), (size_t)count * p->iteration * nbatch * 8);
    if (r->qp_iters) memcpy(r->qp_iters, its.data(), (size_t)count * p->iteration * nbatch * 4);
    if (r->qp_status) memcpy(r->qp_status, qst.data(), (size_t)count * p->iteration * nbatch * 4);
    if (r->qp_res) memcpy(r->qp_res, res.data(), (size_t)count * p->iteration * nbatch * 32);
    int rc = 0;
    for (int c = 0; c < count; c++) {
        if (r->status) r->status[c] = status[c];
        if (!rc && status[c]) rc = status[c];
    }
    return rc;
}

// same definition as the product's (rbpe_api.cu); restated here because the emulator links nothing of it
extern "C" int rbpe_set_batch(int N, int sequential, int batch_size, int batch_iter, int *ebs, int *ebi) {
    if (batch_size <= 0) batch_size = 1;
    int bmax = (N + batch_size - 1) / batch_size;
    if (sequential) {
        if (batch_iter < 0 || batch_iter > bmax) batch_iter = bmax;
    } else {
        batch_size = N;
        batch_iter = 1;
    }
    *ebs = batch_size;
    *ebi = batch_iter;
    return bmax;
}

// Block tridiagonal factorisation / solve of rbpe_blockla.cuh on caller-provided blocks (test_emu_kernels.py compares with
// numpy): Dall [nblk][kp*kp] (lower triangles used), Oall [nblk-1][kp*kp], g [nblk*kb] -> in place: factor, solution.
extern "C" int emu_block_tridiag(int nblk, int kb, double *Dall, double *Oall, double *g, int threads) {
    const int kp = bla_kp(kb);
    std::vector<double> Linv((size_t)nblk * bla_ninv(kp) * BLA_W * BLA_W), w((size_t)nblk * kp), y((size_t)nblk * kp + 32), flag(1);
    int ok = 1;
    emu::launch([&] {
        bool f = factor_bt_blk(nblk, kp, Dall, Oall, Linv.data(), flag.data());
        if (threadIdx.x == 0) ok = f ? 1 : 0;
        __syncthreads();
        if (kp <= 2 * BLA_W) solve_bt_small(nblk, kb, kp, Dall, Oall, Linv.data(), g, w.data(), y.data());   // same rule as kkt_solve
        else solve_bt_blk(nblk, kb, kp, Dall, Oall, Linv.data(), g, w.data(), y.data());
    }, 1, threads, 0);
    return ok;
}
