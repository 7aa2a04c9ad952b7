// rbpe_api.cu -- host side of the C ABI in include/rbpe.h: device memory, H2D/D2H, kernel launches, timing.
// One handle = one device + one stream.  No CPU fallback: without a usable CUDA device every call fails.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rbpe.h"
#include "rbpe_kernels.cuh"
#include "rbpe_post.cuh"

using namespace rbpe;

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool pinned = false;   // exported to peers (IPC): must not be reallocated
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (pinned) return cudaErrorInvalidValue;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return (T *)p; }
};

char g_create_error[512] = "";

}  // namespace

struct rbpe_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of the pipelined rbpe_solve_many
    cudaStream_t s_k2 = nullptr;                     // second compute stream: kernels of consecutive chunks overlap (no tail gaps)
    std::vector<cudaEvent_t> pev;                    // per-chunk events of the pipeline
    int chunk = 2368;                                // missions per pipeline stage (RBPE_CHUNK)
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int max_iter = 100;
    double tol_gap = 1e-10, tol_res = 1e-9;
    size_t smem_budget = 0, smem_optin = 0;
    size_t pdip_dyn_cap = 0;   // largest dynamic shared memory pdip_kernel can be launched with (opt-in limit minus its static part)
    int threads = 128;
    int threads_forced = 0;   // RBPE_THREADS / rbpe_config.reserved[0] given: applies to joint batches too
    int force_cta = 0;   // RBPE_KERNEL=cta: never use the warp-per-QP kernel (A/B testing)
    size_t x1_dyn_cap = 0;   // same for the several-warps-per-QP latency kernel (pdip1x_kernel)
    int lat_mode = -1;       // RBPE_LAT=0 / 1: never / always use the latency kernel where it fits (default: by work-item count)
#ifdef RBPE_W1_V1
    int lat_warps = 0;       // (the round-1 layout has no latency kernel)
#else
    int lat_warps = X1_MAXW; // RBPE_LAT_WARPS: warps per QP of the latency kernel
#endif
    int lat_warps_forced = 0;
    int last_solver = 0, last_threads = 0;   // rbpe_last_solver
    int sm_count = 0;
    char err[512] = "";
    // resident problem
    bool resident = false, assembled = false;
    long launches = 0;
    cudaEvent_t tev[2] = {nullptr, nullptr};
    int count = 0, N = 0, M = 0, sequential = 0, bs = 1, nbatch = 0, iteration = 1, nrec = 1, sweep = 0;
    DevBuf T, start, goal, radius, sfc_offs, sfc_base, sfc_box, sfc_t, rsfc_n, rsfc_t, init_traj;
    DevBuf post_a, post_b, post_c, post_d, post_e;   // scratch of rbpe_corridor_rsfc / rbpe_safety_metrics
    DevBuf segbox, reln, segmat, ctrl, frozen, coef, qp_obj, qp_iters, qp_status, qp_res, status, scratch;
    rbpe_timing timing;
    struct PinnedInts {   // page-locked (a pageable destination would make the per-chunk status copies blocking)
        int *p = nullptr;
        size_t cap = 0;
        bool resize(size_t n) {
            if (n <= cap) return true;
            if (p) cudaFreeHost(p);
            p = nullptr; cap = 0;
            if (cudaHostAlloc((void **)&p, (n + 64) * sizeof(int), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return false; }
            cap = n + 64;
            return true;
        }
        int *data() { return p; }
        int &operator[](size_t i) { return p[i]; }
        ~PinnedInts() { if (p) cudaFreeHost(p); }
    } host_status;
    // Jacobi exchange over peer memory (rbpe_peer_*): tables[0] / tables[1] are the two physical control-point buffers of
    // every rank as exported (ctrl, frozen); `phys_cur` tells which one currently is h->ctrl
    int peer_rank = -1, peer_world = 0, phys_cur = 0;
    double *peer_tables[2][RBPE_MAX_PEERS] = {};
    unsigned long long *peer_flags[RBPE_MAX_PEERS] = {};
    std::vector<void *> ipc_opened;
    DevBuf flags;   // [RBPE_MAX_PEERS] u64 flag words, then u32 done counter, then i32 error
    unsigned long long sweep_id = 0;
};

static int fail(rbpe_handle *h, int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h ? h->err : g_create_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(h, RBPE_CUDA_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" int rbpe_set_batch(int N, int sequential, int batch_size, int batch_iter, int *ebs, int *ebi) {
    // RBPPlanner::setBatch, rbp_planner.hpp L849-L872
    if (batch_size <= 0) batch_size = 1;
    int bmax = (N + batch_size - 1) / batch_size;
    if (sequential) {
        if (batch_iter < 0 || batch_iter > bmax) batch_iter = bmax;
    } else {
        batch_size = N;
        batch_iter = 1;
    }
    if (ebs) *ebs = batch_size;
    if (ebi) *ebi = batch_iter;
    return bmax;
}

extern "C" int rbpe_create(const rbpe_config *cfg, rbpe_handle **out) {
    if (!out) return RBPE_BAD_ARG;
    *out = nullptr;
    rbpe_handle *h = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RBPE_CUDA_ERROR, "no CUDA device (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    int dev = cfg ? cfg->device : 0;
    if (dev < 0 || dev >= ndev) return fail(nullptr, RBPE_BAD_ARG, "device %d out of range (%d devices)", dev, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, dev)) != cudaSuccess)
        return fail(nullptr, RBPE_CUDA_ERROR, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, RBPE_CUDA_ERROR, "device %d is sm_%d%d; this library holds sm_100a code only", dev, prop.major,
                    prop.minor);
    h = new rbpe_handle();
    h->device = dev;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (cfg) {
        if (cfg->max_iter > 0) h->max_iter = cfg->max_iter;
        if (cfg->tol_gap > 0) h->tol_gap = cfg->tol_gap;
        if (cfg->tol_res > 0) h->tol_res = cfg->tol_res;
        h->smem_budget = cfg->smem_budget;
        int th = cfg->reserved[0];   // CTA size of the PDIP kernel (tuning knob): 32..256, multiple of 32
        if (th >= 32 && th <= CTA_THREADS_MAX && th % 32 == 0) { h->threads = th; h->threads_forced = 1; }
    }
    // tuning overrides (documented in DESIGN.md): RBPE_THREADS, RBPE_SMEM_KB
    if (const char *e = getenv("RBPE_THREADS")) { int th = atoi(e); if (th >= 32 && th <= CTA_THREADS_MAX && th % 32 == 0) { h->threads = th; h->threads_forced = 1; } }
    if (const char *e = getenv("RBPE_KERNEL")) h->force_cta = (strcmp(e, "cta") == 0);
    if (const char *e = getenv("RBPE_CHUNK")) { int c = atoi(e); if (c > 0) h->chunk = c; }
    if (const char *e = getenv("RBPE_SMEM_KB")) { long kb = atol(e); if (kb > 0) h->smem_budget = (size_t)kb * 1024; }
    if (h->smem_budget == 0) h->smem_budget = 48 * 1024;
    // the joint-batch factorisation holds 17 KB of static shared memory next to the dynamic budget
    if (h->smem_budget + 20 * 1024 > h->smem_optin) h->smem_budget = h->smem_optin - 20 * 1024;
    memset(&h->timing, 0, sizeof(h->timing));
    if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        fail(nullptr, RBPE_CUDA_ERROR, "stream creation: %s", cudaGetErrorString(e));
        delete h;
        return RBPE_CUDA_ERROR;
    }
    for (int i = 0; i < 7; i++) cudaEventCreate(&h->ev[i]);
    cudaEventCreate(&h->tev[0]);
    cudaEventCreate(&h->tev[1]);
    {
        cudaFuncAttributes fa;
        size_t st = 36 * 1024;
        if (cudaFuncGetAttributes(&fa, pdip_kernel) == cudaSuccess) st = fa.sharedSizeBytes;
        h->pdip_dyn_cap = (h->smem_optin - st - 1024) & ~(size_t)15;
        cudaFuncSetAttribute(pdip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pdip_dyn_cap);
    }
    cudaFuncSetAttribute(pdip1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin);
#ifndef RBPE_W1_V1
    {
        cudaFuncAttributes fa;
        size_t st = 1024;
        if (cudaFuncGetAttributes(&fa, pdip1x_kernel) == cudaSuccess) st = fa.sharedSizeBytes;
        h->x1_dyn_cap = (h->smem_optin - st - 1024) & ~(size_t)15;
        cudaFuncSetAttribute(pdip1x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->x1_dyn_cap);
        if (const char *e = getenv("RBPE_LAT")) h->lat_mode = atoi(e) != 0;
        if (const char *e = getenv("RBPE_LAT_WARPS")) { int w = atoi(e); if (w >= 1 && w <= X1_MAXW) { h->lat_warps = w; h->lat_warps_forced = 1; } }
    }
#endif
    *out = h;
    return RBPE_OK;
}

extern "C" void rbpe_destroy(rbpe_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf *all[] = {&h->T, &h->start, &h->goal, &h->radius, &h->sfc_offs, &h->sfc_base, &h->sfc_box, &h->sfc_t,
                     &h->rsfc_n, &h->rsfc_t, &h->init_traj, &h->segbox, &h->reln, &h->segmat, &h->ctrl, &h->frozen,
                     &h->coef, &h->qp_obj, &h->qp_iters, &h->qp_status, &h->qp_res, &h->status, &h->scratch, &h->post_a, &h->post_b, &h->post_c,
                     &h->post_d, &h->post_e, &h->flags};
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    for (DevBuf *b : all) b->release();
    for (int i = 0; i < 7; i++)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 2; i++)
        if (h->tev[i]) cudaEventDestroy(h->tev[i]);
    for (cudaEvent_t e : h->pev) cudaEventDestroy(e);
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->s_k2) cudaStreamDestroy(h->s_k2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *rbpe_last_error(const rbpe_handle *h) { return h ? h->err : g_create_error; }

// page-locked host memory for the caller's input / result buffers: with pageable buffers every cudaMemcpyAsync of the
// pipelined rbpe_solve_many blocks the host, and the chunks' copies and kernels no longer overlap
extern "C" void *rbpe_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void rbpe_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

static int up(rbpe_handle *h, DevBuf &b, const void *src, size_t bytes) {
    if (bytes == 0) return RBPE_OK;
    CU(b.reserve(bytes));
    CU(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return RBPE_OK;
}

extern "C" int rbpe_upload(rbpe_handle *h, const rbpe_problem *p, int count) {
    if (!h) return RBPE_BAD_ARG;
    if (!p || count <= 0 || p->N <= 0 || p->M <= 0 || p->M > MAX_M || p->batch_size <= 0 || p->iteration < 0)
        return fail(h, RBPE_BAD_ARG, "bad problem shape (N=%d M=%d batch_size=%d count=%d)", p ? p->N : -1, p ? p->M : -1,
                    p ? p->batch_size : -1, count);
    if (!p->T || !p->start || !p->goal || !p->radius || !p->sfc_offs || !p->sfc_base || !p->sfc_box || !p->sfc_t ||
        (p->N > 1 && (!p->rsfc_n || !p->rsfc_t)) || (p->sequential && !p->init_traj))
        return fail(h, RBPE_BAD_ARG, "null input array");
    CU(cudaSetDevice(h->device));
    h->resident = false;
    const int N = p->N, M = p->M;
    const size_t P = (size_t)N * (N - 1) / 2;
    h->count = count; h->N = N; h->M = M; h->sequential = p->sequential ? 1 : 0; h->iteration = p->iteration;
    rbpe_set_batch(N, h->sequential, p->batch_size, p->batch_iter, &h->bs, &h->nbatch);
    h->nrec = h->iteration * h->nbatch;
    if (h->nrec < 1) h->nrec = 1;
    const size_t nbox = (size_t)p->sfc_base[count];

    CU(cudaEventRecord(h->ev[0], h->stream));
    int rc;
    if ((rc = up(h, h->T, p->T, (size_t)count * (M + 1) * 8))) return rc;
    if ((rc = up(h, h->start, p->start, (size_t)count * N * 9 * 8))) return rc;
    if ((rc = up(h, h->goal, p->goal, (size_t)count * N * 9 * 8))) return rc;
    if ((rc = up(h, h->radius, p->radius, (size_t)count * N * 8))) return rc;
    if ((rc = up(h, h->sfc_offs, p->sfc_offs, (size_t)count * (N + 1) * 4))) return rc;
    if ((rc = up(h, h->sfc_base, p->sfc_base, (size_t)(count + 1) * 4))) return rc;
    if ((rc = up(h, h->sfc_box, p->sfc_box, nbox * 6 * 8))) return rc;
    if ((rc = up(h, h->sfc_t, p->sfc_t, nbox * 8))) return rc;
    if ((rc = up(h, h->rsfc_n, p->rsfc_n, (size_t)count * P * M * 3 * 4))) return rc;
    if ((rc = up(h, h->rsfc_t, p->rsfc_t, (size_t)count * P * M * 8))) return rc;
    if (p->sequential && (rc = up(h, h->init_traj, p->init_traj, (size_t)count * N * (M + 1) * 3 * 4))) return rc;
    CU(cudaEventRecord(h->ev[1], h->stream));

    h->resident = true;
    h->assembled = false;
    return RBPE_OK;
}

// k1 on the resident raw inputs: (re)builds segbox / reln / segmat and resets the control-point table to `dummy`
extern "C" int rbpe_assemble(rbpe_handle *h) {
    if (!h) return RBPE_BAD_ARG;
    if (!h->resident) return fail(h, RBPE_BAD_ARG, "rbpe_assemble: nothing uploaded");
    CU(cudaSetDevice(h->device));
    const int N = h->N, M = h->M, count = h->count;
    const size_t P = (size_t)N * (N - 1) / 2, per = (size_t)N * 18 * M;
    CU(cudaEventRecord(h->ev[1], h->stream));
    CU(h->segbox.reserve((size_t)count * N * M * 6 * 8));
    CU(h->reln.reserve((size_t)count * (P ? P : 1) * M * 3 * 4));
    CU(h->segmat.reserve((size_t)count * M * SEGMAT * 8));
    CU(h->ctrl.reserve(count * per * 8));
    CU(h->frozen.reserve(count * per * 8));
    CU(h->coef.reserve(count * per * 8));
    CU(h->qp_obj.reserve((size_t)count * h->nrec * 8));
    CU(h->qp_iters.reserve((size_t)count * h->nrec * 4));
    CU(h->qp_status.reserve((size_t)count * h->nrec * 4));
    CU(h->qp_res.reserve((size_t)count * h->nrec * 32));
    CU(h->status.reserve((size_t)count * 4));
    CU(cudaMemsetAsync(h->status.p, 0, (size_t)count * 4, h->stream));
    CU(cudaMemsetAsync(h->qp_obj.p, 0, (size_t)count * h->nrec * 8, h->stream));
    CU(cudaMemsetAsync(h->qp_iters.p, 0, (size_t)count * h->nrec * 4, h->stream));
    CU(cudaMemsetAsync(h->qp_status.p, 0, (size_t)count * h->nrec * 4, h->stream));
    CU(cudaMemsetAsync(h->qp_res.p, 0, (size_t)count * h->nrec * 32, h->stream));

    AssembleArgs A;
    A.count = count; A.N = N; A.M = M; A.sequential = h->sequential;
    A.T = h->T.as<double>(); A.sfc_offs = h->sfc_offs.as<int>(); A.sfc_base = h->sfc_base.as<int>();
    A.sfc_box = h->sfc_box.as<double>(); A.sfc_t = h->sfc_t.as<double>();
    A.rsfc_n = h->rsfc_n.as<float>(); A.rsfc_t = h->rsfc_t.as<double>(); A.init_traj = h->init_traj.as<float>();
    A.segbox = h->segbox.as<double>(); A.reln = h->reln.as<float>(); A.ctrl = h->ctrl.as<double>();
    A.segmat = h->segmat.as<double>(); A.status = h->status.as<int>();
    const long per_mission = (long)N + (long)P * M + (long)per + M;
    long total = per_mission * count;
    int blocks = (int)((total + 255) / 256);
    int cap = h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    assemble_kernel<<<blocks, 256, 0, h->stream>>>(A);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev[2], h->stream));
    h->launches++;
    h->sweep = 0;
    h->assembled = true;
    return RBPE_OK;
}

// Dynamic shared memory of the CTA-per-QP kernel: the first `budget` bytes of a CTA's scratch live in shared memory, the
// rest in the L2-resident arena (layout() in rbpe_kernels.cuh takes the small hot arrays first).  Throughput regime (more
// CTAs than SMs): the configured budget (default 48 KB -> two CTAs per SM).  Latency regime (a handful of missions: at most
// one CTA per SM anyway): everything the SM has, so that vectors, reduced Hessian and factor stay on chip.
static size_t smem_for(const rbpe_handle *h, size_t scratch_d, long units = -1) {
    size_t need = scratch_d * 8 + 1024;
    size_t budget = h->smem_budget;
    if (units >= 0 && units <= h->sm_count && !getenv("RBPE_SMEM_KB")) budget = h->pdip_dyn_cap;
    size_t s = need < budget ? need : budget;
    return (s + 15) & ~(size_t)15;
}

static int launch_convert(rbpe_handle *h) {
    ConvertArgs C;
    C.count = h->count; C.N = h->N; C.M = h->M;
    C.ctrl = h->ctrl.as<double>(); C.segmat = h->segmat.as<double>(); C.coef = h->coef.as<double>();
    long total = (long)h->count * h->N * 18 * h->M;
    int blocks = (int)((total + 255) / 256), cap = h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    convert_kernel<<<blocks, 256, 0, h->stream>>>(C);
    CU(cudaGetLastError());
    h->launches++;
    return RBPE_OK;
}

static int fill_solve_args(rbpe_handle *h, SolveArgs &S, int mode, int grid) {
    S.count = h->count; S.N = h->N; S.M = h->M; S.bs = h->bs; S.nbatch = h->nbatch; S.iteration = h->iteration;
    S.sequential = h->sequential; S.mode = mode; S.batch_begin = 0; S.batch_end = h->nbatch; S.rec_offset = 0;
    S.max_iter = h->max_iter; S.tol_gap = h->tol_gap; S.tol_res = h->tol_res;
    S.start = h->start.as<double>(); S.goal = h->goal.as<double>(); S.radius = h->radius.as<double>();
    S.segbox = h->segbox.as<double>(); S.reln = h->reln.as<float>(); S.segmat = h->segmat.as<double>();
    S.ctrl = h->ctrl.as<double>(); S.ctrl_frozen = h->frozen.as<double>();
    S.qp_obj = h->qp_obj.as<double>(); S.qp_iters = h->qp_iters.as<int>(); S.qp_status = h->qp_status.as<int>();
    S.qp_res = h->qp_res.as<double>(); S.nrec = h->nrec; S.status = h->status.as<int>();
    S.npeer = 0; S.peer_rank = 0; S.sweep_id = 0; S.done_counter = nullptr; S.work_items = 0;
    for (int p = 0; p < RBPE_MAX_PEERS; p++) { S.peer_ctrl[p] = nullptr; S.peer_flags[p] = nullptr; }
    S.scratch_stride = scratch_doubles(h->N, h->M, h->bs);
    S.smem_bytes = (unsigned)smem_for(h, S.scratch_stride);
    S.panel_bytes = 0;
    CU(h->scratch.reserve(S.scratch_stride * 8 * (size_t)grid));
    S.scratch = h->scratch.as<double>();
    return RBPE_OK;
}

// QPs per CTA of the warp-per-QP kernel (one-agent batches); 0 = use the CTA-per-QP kernel
static int warps_per_cta(const rbpe_handle *h) {
    if (h->bs != 1 || h->force_cta) return 0;
    size_t per = w1_smem_doubles(h->M) * 8;
    size_t budget = h->smem_optin < 200 * 1024 ? h->smem_optin : 200 * 1024;
    int w = (int)(budget / per);
    if (w > W1_WARPS) w = W1_WARPS;
    return w;
}

// Warps per QP of the latency kernel (pdip1x_kernel), 0 = not used.  It is the kernel for one-agent batches when the
// work items are too few to fill the SMs with one warp each (a handful of missions: the reference's own use), and only
// where the row state of a QP fits in shared memory.  Up to two CTAs per SM: beyond that pdip1_kernel's one warp per QP
// has the higher throughput.
static int warps_per_qp(const rbpe_handle *h, long units) {
#ifdef RBPE_W1_V1
    return 0;
#else
    if (h->bs != 1 || h->force_cta || h->lat_mode == 0) return 0;
    int nw = h->lat_warps;
    const size_t bytes = x1_smem_doubles(h->N, h->M, nw) * 8;
    if (bytes > h->x1_dyn_cap) return 0;
    if (h->lat_mode == 1) return nw;
    // measured (tools/gpu_latency.py, 64 agents, one Jacobi sweep): 8 warps per QP win up to two CTAs per SM (0.59 / 0.61 /
    // 0.85 ms at 64 / 128 / 256 QPs against 1.17 / 1.18 / 1.19 ms of pdip1_kernel), 4 warps per QP up to four (512 QPs:
    // 0.91 ms against 1.48 ms), beyond that one warp per QP (1 024 QPs: 2.0 ms against 1.5 ms)
    const long per_sm = (2 * (bytes + 1024) <= 227 * 1024) ? 2 : 1;
    if (units <= (long)h->sm_count * per_sm) return nw;
    if (!h->lat_warps_forced && nw > 4 && units <= 2 * (long)h->sm_count * per_sm && x1_smem_doubles(h->N, h->M, 4) * 8 <= h->x1_dyn_cap) return 4;
    return 0;
#endif
}

// launches k2 over `units` independent work items (missions in mode 0, (mission, batch) pairs in mode 1)
// S must have been filled by fill_solve_args (mode, ranges, record offsets already set by the caller)
// `st`: stream to launch on (default: the engine's); `slot0` / `slots`: this launch uses work-item slots [slot0, slot0 + units)
// of a scratch arena sized for `slots` items (pipelined call: chunks in flight on two streams must not share scratch)
static int launch_pdip_prepared(rbpe_handle *h, SolveArgs &S, long units, cudaStream_t st = nullptr, long slot0 = 0, long slots = 0) {
    int wpc = warps_per_cta(h);
    S.work_items = (unsigned)units;
    if (!st) st = h->stream;
    if (slots < units) slots = units;
    const int xw = warps_per_qp(h, units);
    if (xw > 0) {
#ifndef RBPE_W1_V1
        S.scratch_stride = 0;
        S.smem_bytes = (unsigned)(x1_smem_doubles(h->N, h->M, xw) * 8);
        S.scratch = h->scratch.as<double>();
        pdip1x_kernel<<<(unsigned)units, xw * 32, S.smem_bytes, st>>>(S);
        h->last_solver = 2; h->last_threads = xw * 32;
#endif
    } else if (wpc > 0) {
        long grid = (units + wpc - 1) / wpc;
        S.scratch_stride = w1_scratch_doubles(h->N, h->M);
        S.smem_bytes = (unsigned)(wpc * w1_smem_doubles(h->M) * 8);
        CU(h->scratch.reserve(S.scratch_stride * 8 * (size_t)(slots + wpc)));
        S.scratch = h->scratch.as<double>() + (size_t)slot0 * S.scratch_stride;
        pdip1_kernel<<<(unsigned)grid, wpc * 32, S.smem_bytes, st>>>(S);
        h->last_solver = 1; h->last_threads = 32;
    } else {
        S.scratch_stride = scratch_doubles(h->N, h->M, h->bs);
        S.smem_bytes = (unsigned)smem_for(h, S.scratch_stride, units);
        S.panel_bytes = 0;
        // TMA-staged panel of the block factorisation, in front of the arena: on by default where it measured faster (blocks of
        // order > 144, i.e. b >= 17: +6.5 % at b = 32; -4 % at b <= 16 where it costs the second CTA per SM its shared memory);
        // RBPE_TMA=1 / 0 forces it on / off
        const char *tma_env = getenv("RBPE_TMA");
        const bool use_tma = tma_env ? atoi(tma_env) != 0 : kp_of(h->bs) > 144;
        if (h->bs > 1 && use_tma) {
            size_t pb = bla_panel_doubles((int)kp_of(h->bs)) * 8;
            const size_t cap = h->pdip_dyn_cap;
            if (pb + 16 * 1024 <= cap) {   // the arena keeps at least 16 KB; in the latency regime it shrinks to make room
                if (S.smem_bytes + pb > cap) S.smem_bytes = (unsigned)((cap - pb) & ~(size_t)15);
                S.panel_bytes = (unsigned)pb;
                S.smem_bytes += (unsigned)pb;
            }
        }
        CU(h->scratch.reserve(S.scratch_stride * 8 * (size_t)slots));
        S.scratch = h->scratch.as<double>() + (size_t)slot0 * S.scratch_stride;
        // joint batches use the full CTA (CTA-wide DMMA factorisation); one-agent batches through this kernel keep the knob
        // latency regime (at most one CTA per SM): 16 warps per CTA -- a lone CTA is bound by the dependent-issue latency of each
        // warp (~6 cycles per instruction), so halving the per-warp share of the row passes and the CTA-wide loops pays (one
        // 64-agent mission at b = 4: 66.9 -> 53.1 ms); with more CTAs than SMs, 8 warps and two CTAs per SM
        const int joint_threads = (units <= h->sm_count) ? CTA_THREADS_MAX : CTA_THREADS;
        const int threads = (h->bs > 1 && !h->threads_forced) ? joint_threads : h->threads;
        pdip_kernel<<<(unsigned)units, threads, S.smem_bytes, st>>>(S);
        h->last_solver = 3; h->last_threads = threads;
    }
    CU(cudaGetLastError());
    h->launches++;
    return RBPE_OK;
}
static int launch_pdip(rbpe_handle *h, SolveArgs &S, int mode, long units) {
    int rc;
    if ((rc = fill_solve_args(h, S, mode, 1))) return rc;
    return launch_pdip_prepared(h, S, units);
}

extern "C" int rbpe_run(rbpe_handle *h, int mode) {
    if (!h) return RBPE_BAD_ARG;
    if (!h->resident) return fail(h, RBPE_BAD_ARG, "rbpe_run: nothing uploaded");
    int rc;
    if ((rc = rbpe_assemble(h))) return rc;   // k1 is part of the hot path; it also resets `dummy`, so runs are repeatable
    CU(cudaEventRecord(h->ev[3], h->stream));
    if (h->nbatch > 0 && h->iteration > 0) {
        SolveArgs S;
        if (mode == RBPE_MODE_GAUSS_SEIDEL) {
            if ((rc = launch_pdip(h, S, 0, h->count))) return rc;
        } else {
            for (int it = 0; it < h->iteration; it++) {
                CU(cudaMemcpyAsync(h->frozen.p, h->ctrl.p, (size_t)h->count * h->N * 18 * h->M * 8, cudaMemcpyDeviceToDevice,
                                   h->stream));
                if ((rc = fill_solve_args(h, S, 1, 1))) return rc;
                S.rec_offset = it * h->nbatch;
                if ((rc = launch_pdip_prepared(h, S, (long)h->count * h->nbatch))) return rc;
            }
        }
    }
    if ((rc = launch_convert(h))) return rc;
    CU(cudaEventRecord(h->ev[4], h->stream));
    return RBPE_OK;
}

extern "C" int rbpe_run_jacobi_range(rbpe_handle *h, int b0, int b1) {
    if (!h) return RBPE_BAD_ARG;
    if (!h->resident || !h->assembled) return fail(h, RBPE_BAD_ARG, "rbpe_run_jacobi_range: call rbpe_upload and rbpe_assemble first");
    if (b0 < 0 || b1 > h->nbatch || b0 > b1) return fail(h, RBPE_BAD_ARG, "batch range [%d,%d) outside [0,%d)", b0, b1, h->nbatch);
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev[3], h->stream));
    int rc;
    CU(cudaMemcpyAsync(h->frozen.p, h->ctrl.p, (size_t)h->count * h->N * 18 * h->M * 8, cudaMemcpyDeviceToDevice, h->stream));
    if (b1 > b0) {
        SolveArgs S;
        if ((rc = fill_solve_args(h, S, 1, 1))) return rc;
        S.batch_begin = b0; S.batch_end = b1;
        S.rec_offset = (h->iteration > 0 ? h->sweep % h->iteration : 0) * h->nbatch;
        if ((rc = launch_pdip_prepared(h, S, (long)h->count * (b1 - b0)))) return rc;
    }
    h->sweep++;
    if ((rc = launch_convert(h))) return rc;
    CU(cudaEventRecord(h->ev[4], h->stream));
    return RBPE_OK;
}


// ---- Jacobi exchange over NVLink peer memory ---------------------------------------------------------------------
static int peer_prepare(rbpe_handle *h) {
    if (!h->resident || !h->assembled) return fail(h, RBPE_BAD_ARG, "rbpe_peer_*: call rbpe_upload and rbpe_assemble first");
    CU(cudaSetDevice(h->device));
    CU(h->flags.reserve(RBPE_MAX_PEERS * 8 + 64));
    CU(cudaMemsetAsync(h->flags.p, 0, RBPE_MAX_PEERS * 8 + 64, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    {   // no allocation may happen once peers spin on our flags (cudaMalloc can wait for running kernels): size the
        // scratch arena for a sweep over every batch now
        const long units = (long)h->count * (h->nbatch > 0 ? h->nbatch : 1);
        const int wpc = warps_per_cta(h);
        size_t bytes = wpc > 0 ? w1_scratch_doubles(h->N, h->M) * 8 * (size_t)((units + wpc - 1) / wpc) * wpc
                               : scratch_doubles(h->N, h->M, h->bs) * 8 * (size_t)units;
        CU(h->scratch.reserve(bytes));
    }
    h->ctrl.pinned = h->frozen.pinned = h->flags.pinned = true;
    h->sweep_id = 0;
    h->phys_cur = 0;
    return RBPE_OK;
}

extern "C" int rbpe_peer_export(rbpe_handle *h, unsigned char *handles) {
    if (!h || !handles) return RBPE_BAD_ARG;
    int rc = peer_prepare(h);
    if (rc) return rc;
    cudaIpcMemHandle_t m[3];
    CU(cudaIpcGetMemHandle(&m[0], h->ctrl.p));
    CU(cudaIpcGetMemHandle(&m[1], h->frozen.p));
    CU(cudaIpcGetMemHandle(&m[2], h->flags.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == RBPE_IPC_HANDLE_BYTES, "IPC handle size");
    memcpy(handles, m, sizeof(m));
    return RBPE_OK;
}

extern "C" int rbpe_peer_attach(rbpe_handle *h, int rank, int world, const unsigned char *all_handles) {
    if (!h || !all_handles || world < 1 || world > RBPE_MAX_PEERS || rank < 0 || rank >= world)
        return fail(h, RBPE_BAD_ARG, "rbpe_peer_attach: bad rank / world (%d / %d, at most %d ranks)", rank, world, RBPE_MAX_PEERS);
    if (!h->ctrl.pinned) return fail(h, RBPE_BAD_ARG, "rbpe_peer_attach: call rbpe_peer_export first");
    CU(cudaSetDevice(h->device));
    for (int p = 0; p < world; p++) {
        if (p == rank) {
            h->peer_tables[0][p] = h->ctrl.as<double>(); h->peer_tables[1][p] = h->frozen.as<double>();
            h->peer_flags[p] = h->flags.as<unsigned long long>();
            continue;
        }
        cudaIpcMemHandle_t m[3];
        memcpy(m, all_handles + (size_t)p * 3 * RBPE_IPC_HANDLE_BYTES, sizeof(m));
        void *ptr[3];
        for (int i = 0; i < 3; i++) {
            CU(cudaIpcOpenMemHandle(&ptr[i], m[i], cudaIpcMemLazyEnablePeerAccess));
            h->ipc_opened.push_back(ptr[i]);
        }
        h->peer_tables[0][p] = (double *)ptr[0]; h->peer_tables[1][p] = (double *)ptr[1];
        h->peer_flags[p] = (unsigned long long *)ptr[2];
    }
    h->peer_rank = rank; h->peer_world = world;
    return RBPE_OK;
}

// same-process variant (several handles of one process, e.g. the single-GPU tests): raw device pointers, no IPC
extern "C" int rbpe_peer_attach_local(rbpe_handle *h, int rank, int world, rbpe_handle *const *peers) {
    if (!h || !peers || world < 1 || world > RBPE_MAX_PEERS || rank < 0 || rank >= world || peers[rank] != h)
        return fail(h, RBPE_BAD_ARG, "rbpe_peer_attach_local: bad rank / world / peer list");
    for (int p = 0; p < world; p++)
        if (!peers[p] || !peers[p]->ctrl.pinned) return fail(h, RBPE_BAD_ARG, "rbpe_peer_attach_local: peer %d not exported", p);
    CU(cudaSetDevice(h->device));
    for (int p = 0; p < world; p++) {
        if (peers[p]->device != h->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, h->device, peers[p]->device));
            if (!can) return fail(h, RBPE_CUDA_ERROR, "device %d cannot access device %d", h->device, peers[p]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(peers[p]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
            cudaGetLastError();
        }
        h->peer_tables[0][p] = peers[p]->ctrl.as<double>(); h->peer_tables[1][p] = peers[p]->frozen.as<double>();
        h->peer_flags[p] = peers[p]->flags.as<unsigned long long>();
    }
    h->peer_rank = rank; h->peer_world = world;
    return RBPE_OK;
}

// One Jacobi sweep over batches [b0, b1) with the exchange fused into the sweep kernel: the kernel reads the current
// table, stores every solved batch into the NEXT table of every rank over peer memory, and its last work item raises
// this rank's flag on every peer; a one-CTA kernel then waits for all ranks' flags (no host synchronisation, no
// collective call), the tables swap roles and the conversion kernel runs on the completed table.
extern "C" int rbpe_run_jacobi_fused(rbpe_handle *h, int b0, int b1) {
    if (!h) return RBPE_BAD_ARG;
    if (!h->resident || !h->assembled) return fail(h, RBPE_BAD_ARG, "rbpe_run_jacobi_fused: call rbpe_upload and rbpe_assemble first");
    if (h->peer_world < 1) return fail(h, RBPE_BAD_ARG, "rbpe_run_jacobi_fused: no peers attached");
    if ((long)h->nbatch * h->bs < h->N)
        return fail(h, RBPE_BAD_ARG, "rbpe_run_jacobi_fused: truncated schedules (batch_iter < ceil(N/batch_size)) leave the next table incomplete");
    if (b0 < 0 || b1 > h->nbatch || b0 > b1) return fail(h, RBPE_BAD_ARG, "batch range [%d,%d) outside [0,%d)", b0, b1, h->nbatch);
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->ev[3], h->stream));
    int rc;
    h->sweep_id++;
    unsigned long long *fl = h->flags.as<unsigned long long>();
    int *err = (int *)(fl + RBPE_MAX_PEERS) + 1;
    if (b1 > b0) {
        SolveArgs S;
        if ((rc = fill_solve_args(h, S, 1, 1))) return rc;
        S.batch_begin = b0; S.batch_end = b1;
        S.rec_offset = (h->iteration > 0 ? h->sweep % h->iteration : 0) * h->nbatch;
        S.ctrl_frozen = h->ctrl.as<double>();   // the complete current table; nothing is written to it during the sweep
        S.npeer = h->peer_world; S.peer_rank = h->peer_rank; S.sweep_id = h->sweep_id;
        for (int p = 0; p < h->peer_world; p++) { S.peer_ctrl[p] = h->peer_tables[h->phys_cur ^ 1][p]; S.peer_flags[p] = h->peer_flags[p]; }
        S.done_counter = (unsigned int *)(fl + RBPE_MAX_PEERS);
        if ((rc = launch_pdip_prepared(h, S, (long)h->count * (b1 - b0)))) return rc;
    } else {   // nothing to solve on this rank: only raise the flags (host-issued peer copies on the stream)
        for (int p = 0; p < h->peer_world; p++)
            CU(cudaMemcpyAsync(h->peer_flags[p] + h->peer_rank, &h->sweep_id, 8, cudaMemcpyHostToDevice, h->stream));
    }
    peer_wait_kernel<<<1, 32, 0, h->stream>>>(fl, h->peer_world, h->sweep_id, err);
    CU(cudaGetLastError());
    h->launches++;
    std::swap(h->ctrl, h->frozen);
    h->phys_cur ^= 1;
    h->sweep++;
    if ((rc = launch_convert(h))) return rc;
    CU(cudaEventRecord(h->ev[4], h->stream));
    return RBPE_OK;
}

// 0 when every flag wait so far has completed in time; RBPE_CUDA_ERROR after a timeout (a peer died or fell behind by > 20 s)
extern "C" int rbpe_peer_status(rbpe_handle *h) {
    if (!h) return RBPE_BAD_ARG;
    if (!h->flags.p) return RBPE_OK;
    CU(cudaSetDevice(h->device));
    int e = 0;
    CU(cudaMemcpyAsync(&e, (int *)(h->flags.as<unsigned long long>() + RBPE_MAX_PEERS) + 1, 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (e) return fail(h, RBPE_CUDA_ERROR, "peer flag wait timed out");
    return RBPE_OK;
}

extern "C" int rbpe_download(rbpe_handle *h, rbpe_result *r) {
    if (!h || !r) return RBPE_BAD_ARG;
    if (!h->resident) return fail(h, RBPE_BAD_ARG, "rbpe_download: nothing uploaded");
    CU(cudaSetDevice(h->device));
    const size_t per = (size_t)h->N * 18 * h->M * 8, c = h->count;
    const size_t nr = (size_t)h->iteration * h->nbatch;
    CU(cudaEventRecord(h->ev[6], h->stream));
    if (r->coef) CU(cudaMemcpyAsync(r->coef, h->coef.p, c * per, cudaMemcpyDeviceToHost, h->stream));
    if (r->ctrl) CU(cudaMemcpyAsync(r->ctrl, h->ctrl.p, c * per, cudaMemcpyDeviceToHost, h->stream));
    if (nr) {
        if (r->qp_obj) CU(cudaMemcpyAsync(r->qp_obj, h->qp_obj.p, c * nr * 8, cudaMemcpyDeviceToHost, h->stream));
        if (r->qp_iters) CU(cudaMemcpyAsync(r->qp_iters, h->qp_iters.p, c * nr * 4, cudaMemcpyDeviceToHost, h->stream));
        if (r->qp_status) CU(cudaMemcpyAsync(r->qp_status, h->qp_status.p, c * nr * 4, cudaMemcpyDeviceToHost, h->stream));
        if (r->qp_res) CU(cudaMemcpyAsync(r->qp_res, h->qp_res.p, c * nr * 32, cudaMemcpyDeviceToHost, h->stream));
    }
    if (!h->host_status.resize(c)) return fail(h, RBPE_CUDA_ERROR, "cudaHostAlloc failed");
    CU(cudaMemcpyAsync(h->host_status.data(), h->status.p, c * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev[5], h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int rc = RBPE_OK;
    for (size_t i = 0; i < c; i++) {
        if (r->status) r->status[i] = h->host_status[i];
        if (rc == RBPE_OK && h->host_status[i] != RBPE_OK) rc = h->host_status[i];
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->timing.h2d_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->timing.assemble_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]) == cudaSuccess) h->timing.solve_ms = ms;
    else h->timing.solve_ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[6], h->ev[5]) == cudaSuccess) h->timing.d2h_ms = ms;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]) == cudaSuccess) h->timing.total_ms = ms;
    cudaGetLastError();
    if (rc != RBPE_OK) snprintf(h->err, sizeof(h->err), "a mission ended with status %d (1 infeasible, 2 not converged, 3 bad input)", rc);
    return rc;
}

// ---- pipelined whole-path call: missions are cut into chunks; chunk k+1 is uploaded (copy stream) while chunk k is being
// assembled / solved / converted (engine stream) and chunk k-1 is downloaded (second copy stream).  Reference semantics
// only (Gauss-Seidel inside every mission); missions are independent, so the chunking changes nothing in the results.
static int solve_many_pipelined(rbpe_handle *h, const rbpe_problem *p, int count, rbpe_result *r) {
    CU(cudaSetDevice(h->device));
    const int N = p->N, M = p->M;
    const size_t P = (size_t)N * (N - 1) / 2, per = (size_t)N * 18 * M;
    h->resident = false;
    h->count = count; h->N = N; h->M = M; h->sequential = p->sequential ? 1 : 0; h->iteration = p->iteration;
    rbpe_set_batch(N, h->sequential, p->batch_size, p->batch_iter, &h->bs, &h->nbatch);
    h->nrec = h->iteration * h->nbatch;
    if (h->nrec < 1) h->nrec = 1;
    const size_t nrec = h->nrec, nr_out = (size_t)h->iteration * h->nbatch;
    const size_t nbox = (size_t)p->sfc_base[count];
    if (!h->s_h2d) CU(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    if (!h->s_d2h) CU(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    if (!h->s_k2) CU(cudaStreamCreateWithFlags(&h->s_k2, cudaStreamNonBlocking));
    const int nchunk = (count + h->chunk - 1) / h->chunk;
    while ((int)h->pev.size() < 2 * nchunk) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pev.push_back(e);
    }
    // device buffers for the whole call
    CU(h->T.reserve((size_t)count * (M + 1) * 8)); CU(h->start.reserve((size_t)count * N * 9 * 8));
    CU(h->goal.reserve((size_t)count * N * 9 * 8)); CU(h->radius.reserve((size_t)count * N * 8));
    CU(h->sfc_offs.reserve((size_t)count * (N + 1) * 4)); CU(h->sfc_base.reserve((size_t)(count + 1) * 4));
    CU(h->sfc_box.reserve(nbox * 6 * 8)); CU(h->sfc_t.reserve(nbox * 8));
    CU(h->rsfc_n.reserve((size_t)count * (P ? P : 1) * M * 3 * 4)); CU(h->rsfc_t.reserve((size_t)count * (P ? P : 1) * M * 8));
    CU(h->init_traj.reserve((size_t)count * N * (M + 1) * 3 * 4));
    CU(h->segbox.reserve((size_t)count * N * M * 6 * 8)); CU(h->reln.reserve((size_t)count * (P ? P : 1) * M * 3 * 4));
    CU(h->segmat.reserve((size_t)count * M * SEGMAT * 8));
    CU(h->ctrl.reserve(count * per * 8)); CU(h->frozen.reserve(count * per * 8)); CU(h->coef.reserve(count * per * 8));
    CU(h->qp_obj.reserve((size_t)count * nrec * 8)); CU(h->qp_iters.reserve((size_t)count * nrec * 4));
    CU(h->qp_status.reserve((size_t)count * nrec * 4)); CU(h->qp_res.reserve((size_t)count * nrec * 32));
    CU(h->status.reserve((size_t)count * 4));
    if (!h->host_status.resize(count)) return fail(h, RBPE_CUDA_ERROR, "cudaHostAlloc failed");
    CU(cudaEventRecord(h->ev[0], h->stream));
    CU(cudaStreamWaitEvent(h->s_h2d, h->ev[0], 0));   // the copy stream starts after whatever the engine stream was doing
    CU(cudaStreamWaitEvent(h->s_k2, h->ev[0], 0));
    CU(cudaMemcpyAsync(h->sfc_base.p, p->sfc_base, (size_t)(count + 1) * 4, cudaMemcpyHostToDevice, h->s_h2d));
    for (int k = 0; k < nchunk; k++) {
        const int c0 = k * h->chunk, c1 = (c0 + h->chunk < count) ? c0 + h->chunk : count, n = c1 - c0;
        const size_t b0 = (size_t)p->sfc_base[c0], b1 = (size_t)p->sfc_base[c1];
        // ---- H2D of the chunk (copy stream) ----
#define H2D(buf, src, elem_bytes, per_mission)                                                                            \
    CU(cudaMemcpyAsync((char *)h->buf.p + (size_t)c0 * (per_mission) * (elem_bytes), (const char *)(src) + (size_t)c0 * (per_mission) * (elem_bytes), \
                       (size_t)n * (per_mission) * (elem_bytes), cudaMemcpyHostToDevice, h->s_h2d))
        H2D(T, p->T, 8, (size_t)(M + 1));
        H2D(start, p->start, 8, (size_t)N * 9);
        H2D(goal, p->goal, 8, (size_t)N * 9);
        H2D(radius, p->radius, 8, (size_t)N);
        H2D(sfc_offs, p->sfc_offs, 4, (size_t)(N + 1));
        if (b1 > b0) {
            CU(cudaMemcpyAsync((char *)h->sfc_box.p + b0 * 48, (const char *)p->sfc_box + b0 * 48, (b1 - b0) * 48, cudaMemcpyHostToDevice, h->s_h2d));
            CU(cudaMemcpyAsync((char *)h->sfc_t.p + b0 * 8, (const char *)p->sfc_t + b0 * 8, (b1 - b0) * 8, cudaMemcpyHostToDevice, h->s_h2d));
        }
        if (P) {
            H2D(rsfc_n, p->rsfc_n, 4, P * M * 3);
            H2D(rsfc_t, p->rsfc_t, 8, P * M);
        }
        if (p->sequential) H2D(init_traj, p->init_traj, 4, (size_t)N * (M + 1) * 3);
#undef H2D
        CU(cudaEventRecord(h->pev[2 * k], h->s_h2d));
        // ---- kernels of the chunk (alternating compute streams: the next chunk's CTAs fill the SMs as this one's drain) ----
        cudaStream_t ks = (k & 1) ? h->s_k2 : h->stream;
        CU(cudaStreamWaitEvent(ks, h->pev[2 * k], 0));
        CU(cudaMemsetAsync((char *)h->status.p + (size_t)c0 * 4, 0, (size_t)n * 4, ks));
        CU(cudaMemsetAsync((char *)h->qp_obj.p + (size_t)c0 * nrec * 8, 0, (size_t)n * nrec * 8, ks));
        CU(cudaMemsetAsync((char *)h->qp_iters.p + (size_t)c0 * nrec * 4, 0, (size_t)n * nrec * 4, ks));
        CU(cudaMemsetAsync((char *)h->qp_status.p + (size_t)c0 * nrec * 4, 0, (size_t)n * nrec * 4, ks));
        CU(cudaMemsetAsync((char *)h->qp_res.p + (size_t)c0 * nrec * 32, 0, (size_t)n * nrec * 32, ks));
        AssembleArgs A;
        A.count = n; A.N = N; A.M = M; A.sequential = h->sequential;
        A.T = h->T.as<double>() + (size_t)c0 * (M + 1); A.sfc_offs = h->sfc_offs.as<int>() + (size_t)c0 * (N + 1);
        A.sfc_base = h->sfc_base.as<int>() + c0; A.sfc_box = h->sfc_box.as<double>(); A.sfc_t = h->sfc_t.as<double>();
        A.rsfc_n = h->rsfc_n.as<float>() + (size_t)c0 * P * M * 3; A.rsfc_t = h->rsfc_t.as<double>() + (size_t)c0 * P * M;
        A.init_traj = h->init_traj.as<float>() + (size_t)c0 * N * (M + 1) * 3;
        A.segbox = h->segbox.as<double>() + (size_t)c0 * N * M * 6; A.reln = h->reln.as<float>() + (size_t)c0 * P * M * 3;
        A.ctrl = h->ctrl.as<double>() + c0 * per; A.segmat = h->segmat.as<double>() + (size_t)c0 * M * SEGMAT;
        A.status = h->status.as<int>() + c0;
        {
            const long per_mission = (long)N + (long)P * M + (long)per + M;
            long total = per_mission * n;
            int blocks = (int)((total + 255) / 256), cap = h->sm_count * 16;
            if (blocks > cap) blocks = cap;
            if (blocks < 1) blocks = 1;
            assemble_kernel<<<blocks, 256, 0, ks>>>(A);
            CU(cudaGetLastError());
            h->launches++;
        }
        if (h->nbatch > 0 && h->iteration > 0) {
            SolveArgs S;
            int rc;
            if ((rc = fill_solve_args(h, S, 0, 1))) return rc;
            S.count = n;
            S.start += (size_t)c0 * N * 9; S.goal += (size_t)c0 * N * 9; S.radius += (size_t)c0 * N;
            S.segbox += (size_t)c0 * N * M * 6; S.reln += (size_t)c0 * P * M * 3; S.segmat += (size_t)c0 * M * SEGMAT;
            S.ctrl += c0 * per; S.ctrl_frozen += c0 * per;
            S.qp_obj += (size_t)c0 * nrec; S.qp_iters += (size_t)c0 * nrec; S.qp_status += (size_t)c0 * nrec;
            S.qp_res += (size_t)c0 * nrec * 4; S.status += c0;
            if ((rc = launch_pdip_prepared(h, S, n, ks, c0, count))) return rc;
        }
        {
            ConvertArgs C;
            C.count = n; C.N = N; C.M = M;
            C.ctrl = h->ctrl.as<double>() + c0 * per; C.segmat = h->segmat.as<double>() + (size_t)c0 * M * SEGMAT;
            C.coef = h->coef.as<double>() + c0 * per;
            long total = (long)n * per;
            int blocks = (int)((total + 255) / 256), cap = h->sm_count * 16;
            if (blocks > cap) blocks = cap;
            convert_kernel<<<blocks, 256, 0, ks>>>(C);
            CU(cudaGetLastError());
            h->launches++;
        }
        CU(cudaEventRecord(h->pev[2 * k + 1], ks));
        // ---- D2H of the chunk (second copy stream) ----
        CU(cudaStreamWaitEvent(h->s_d2h, h->pev[2 * k + 1], 0));
        if (r->coef) CU(cudaMemcpyAsync(r->coef + c0 * per, (char *)h->coef.p + c0 * per * 8, n * per * 8, cudaMemcpyDeviceToHost, h->s_d2h));
        if (r->ctrl) CU(cudaMemcpyAsync(r->ctrl + c0 * per, (char *)h->ctrl.p + c0 * per * 8, n * per * 8, cudaMemcpyDeviceToHost, h->s_d2h));
        if (nr_out) {   // host records are [count][iteration*batch_iter] == the device layout (nrec == nr_out when nr_out > 0)
            if (r->qp_obj) CU(cudaMemcpyAsync(r->qp_obj + (size_t)c0 * nr_out, (char *)h->qp_obj.p + (size_t)c0 * nrec * 8, (size_t)n * nr_out * 8, cudaMemcpyDeviceToHost, h->s_d2h));
            if (r->qp_iters) CU(cudaMemcpyAsync(r->qp_iters + (size_t)c0 * nr_out, (char *)h->qp_iters.p + (size_t)c0 * nrec * 4, (size_t)n * nr_out * 4, cudaMemcpyDeviceToHost, h->s_d2h));
            if (r->qp_status) CU(cudaMemcpyAsync(r->qp_status + (size_t)c0 * nr_out, (char *)h->qp_status.p + (size_t)c0 * nrec * 4, (size_t)n * nr_out * 4, cudaMemcpyDeviceToHost, h->s_d2h));
            if (r->qp_res) CU(cudaMemcpyAsync(r->qp_res + (size_t)c0 * nr_out * 4, (char *)h->qp_res.p + (size_t)c0 * nrec * 32, (size_t)n * nr_out * 32, cudaMemcpyDeviceToHost, h->s_d2h));
        }
        CU(cudaMemcpyAsync(h->host_status.data() + c0, (char *)h->status.p + (size_t)c0 * 4, (size_t)n * 4, cudaMemcpyDeviceToHost, h->s_d2h));
    }
    CU(cudaStreamSynchronize(h->s_d2h));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->s_k2));
    CU(cudaStreamSynchronize(h->s_h2d));
    h->resident = true;
    h->assembled = true;
    h->sweep = 0;
    int rc = RBPE_OK;
    for (int i = 0; i < count; i++) {
        if (r->status) r->status[i] = h->host_status[i];
        if (rc == RBPE_OK && h->host_status[i] != RBPE_OK) rc = h->host_status[i];
    }
    if (rc != RBPE_OK) snprintf(h->err, sizeof(h->err), "a mission ended with status %d (1 infeasible, 2 not converged, 3 bad input)", rc);
    return rc;
}

extern "C" int rbpe_solve_many(rbpe_handle *h, const rbpe_problem *p, int count, int mode, rbpe_result *r) {
    if (h && p && r && mode == RBPE_MODE_GAUSS_SEIDEL && count >= 2 * h->chunk && p->N > 0 && p->M > 0 && p->M <= MAX_M &&
        p->batch_size > 0 && p->iteration >= 0 && p->T && p->start && p->goal && p->radius && p->sfc_offs && p->sfc_base &&
        p->sfc_box && p->sfc_t && (p->N == 1 || (p->rsfc_n && p->rsfc_t)) && (!p->sequential || p->init_traj))
        return solve_many_pipelined(h, p, count, r);
    int rc = rbpe_upload(h, p, count);
    if (rc) return rc;
    if ((rc = rbpe_run(h, mode))) return rc;
    return rbpe_download(h, r);
}

extern "C" int rbpe_solve(rbpe_handle *h, const rbpe_problem *p, rbpe_result *r) {
    int base[2] = {0, 0};
    if (!h) return RBPE_BAD_ARG;
    if (!p || p->N <= 0 || !p->sfc_offs) return fail(h, RBPE_BAD_ARG, "bad problem");
    rbpe_problem q = *p;
    if (!q.sfc_base) {  // single mission: CSR offsets already absolute
        base[1] = p->sfc_offs[p->N];
        q.sfc_base = base;
    }
    return rbpe_solve_many(h, &q, 1, RBPE_MODE_GAUSS_SEIDEL, r);
}

// overwrite the resident control-point table (`dummy`) from host memory: [count][N][3][6M]
extern "C" int rbpe_set_ctrl(rbpe_handle *h, const double *ctrl) {
    if (!h || !ctrl) return RBPE_BAD_ARG;
    if (!h->resident || !h->assembled) return fail(h, RBPE_BAD_ARG, "rbpe_set_ctrl: call rbpe_upload and rbpe_assemble first");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->ctrl.p, ctrl, (size_t)h->count * h->N * 18 * h->M * 8, cudaMemcpyHostToDevice, h->stream));
    return RBPE_OK;
}

extern "C" double *rbpe_device_ctrl(rbpe_handle *h) { return h ? h->ctrl.as<double>() : nullptr; }
extern "C" double *rbpe_device_coef(rbpe_handle *h) { return h ? h->coef.as<double>() : nullptr; }
// Bernstein -> monomial conversion of the CURRENT control-point table (k3 alone).  Multi-rank Jacobi with the NCCL
// exchange (dist.jacobi_solve, fused = false) calls it after the last all-gather: rbpe_run_jacobi_range converts before the
// exchange, i.e. with the other ranks' agents still at their pre-sweep control points.
extern "C" int rbpe_convert(rbpe_handle *h) {
    if (!h) return RBPE_BAD_ARG;
    if (!h->resident || !h->assembled) return fail(h, RBPE_BAD_ARG, "rbpe_convert: call rbpe_upload and rbpe_assemble first");
    CU(cudaSetDevice(h->device));
    return launch_convert(h);
}
// device address of the per-mission status words [count] (int32): ranks of a Jacobi solve all-reduce (MAX) them
extern "C" int *rbpe_device_status(rbpe_handle *h) { return h ? h->status.as<int>() : nullptr; }
extern "C" int rbpe_last_solver(rbpe_handle *h, int *threads_per_qp) {
    if (!h) return 0;
    if (threads_per_qp) *threads_per_qp = h->last_threads;
    return h->last_solver;
}
extern "C" void *rbpe_stream(rbpe_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int rbpe_sync(rbpe_handle *h) {
    if (!h) return RBPE_BAD_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return RBPE_OK;
}
extern "C" int rbpe_last_timing(const rbpe_handle *h, rbpe_timing *t) {
    if (!h || !t) return RBPE_BAD_ARG;
    *t = h->timing;
    t->kernel_launches = (int)h->launches;
    return RBPE_OK;
}
// CUDA-event stopwatch on the engine's stream (the stream every kernel and copy of this handle is issued on)
extern "C" int rbpe_timer_start(rbpe_handle *h) {
    if (!h) return RBPE_BAD_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventRecord(h->tev[0], h->stream));
    return RBPE_OK;
}
extern "C" int rbpe_timer_stop(rbpe_handle *h, float *ms) {
    if (!h || !ms) return RBPE_BAD_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaEventRecord(h->tev[1], h->stream));
    CU(cudaEventSynchronize(h->tev[1]));
    CU(cudaEventElapsedTime(ms, h->tev[0], h->tev[1]));
    return RBPE_OK;
}

// Corridor::updateRelBox (rbp_corridor.hpp L338-L398) for `count` missions: host in, host out
extern "C" int rbpe_corridor_rsfc(rbpe_handle *h, int N, int M, int count, const float *init_traj, const double *T,
                                  double downwash, float *rsfc_n, double *rsfc_t, int *collided) {
    if (!h) return RBPE_BAD_ARG;
    if (N < 1 || M < 1 || count < 1 || !init_traj || !T || !collided || (N > 1 && (!rsfc_n || !rsfc_t)) || !(downwash > 0))
        return fail(h, RBPE_BAD_ARG, "rbpe_corridor_rsfc: bad argument");
    CU(cudaSetDevice(h->device));
    const size_t P = (size_t)N * (N - 1) / 2;
    const size_t b_traj = (size_t)count * N * (M + 1) * 3 * 4, b_T = (size_t)count * (M + 1) * 8;
    const size_t b_n = (size_t)count * (P ? P : 1) * M * 3 * 4, b_t = (size_t)count * (P ? P : 1) * M * 8;
    CU(h->post_a.reserve(b_traj)); CU(h->post_b.reserve(b_T)); CU(h->post_c.reserve(b_n)); CU(h->post_d.reserve(b_t));
    CU(h->post_e.reserve((size_t)count * 4));
    CU(cudaMemcpyAsync(h->post_a.p, init_traj, b_traj, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->post_b.p, T, b_T, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemsetAsync(h->post_e.p, 0, (size_t)count * 4, h->stream));
    RsfcArgs A;
    A.count = count; A.N = N; A.M = M; A.downwash = downwash;
    A.init_traj = h->post_a.as<float>(); A.T = h->post_b.as<double>();
    A.rsfc_n = h->post_c.as<float>(); A.rsfc_t = h->post_d.as<double>(); A.collided = h->post_e.as<int>();
    long total = (long)count * N * N * M;
    int blocks = (int)((total + 255) / 256), cap = h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    rsfc_kernel<<<blocks, 256, 0, h->stream>>>(A);
    CU(cudaGetLastError());
    h->launches++;
    if (P) {
        CU(cudaMemcpyAsync(rsfc_n, h->post_c.p, (size_t)count * P * M * 3 * 4, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(rsfc_t, h->post_d.p, (size_t)count * P * M * 8, cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaMemcpyAsync(collided, h->post_e.p, (size_t)count * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return RBPE_OK;
}

// RBPPublisher post-hoc checks (rbp_publisher.hpp L117-L127): safety_margin_ratio (collision-free iff >= 1), the time
// of its minimum, and the total flight length, per mission.  coef [count][N][3][6M] as rbpe_result.coef / msgs_traj_coef.
extern "C" int rbpe_safety_metrics(rbpe_handle *h, int N, int M, int count, const double *coef, const double *T,
                                   const double *radius, double downwash, double dt, double *min_ratio, double *t_at_min,
                                   double *length) {
    if (!h) return RBPE_BAD_ARG;
    if (N < 1 || M < 1 || count < 1 || !coef || !T || !radius || !(downwash > 0) || !(dt > 0) || !min_ratio || !t_at_min || !length)
        return fail(h, RBPE_BAD_ARG, "rbpe_safety_metrics: bad argument");
    CU(cudaSetDevice(h->device));
    int nt_max = 1;
    for (int c = 0; c < count; c++) {
        int nt = (int)floor(T[(size_t)c * (M + 1) + M] / dt);
        if (nt > nt_max) nt_max = nt;
    }
    const size_t b_coef = (size_t)count * N * 18 * M * 8, b_T = (size_t)count * (M + 1) * 8, b_r = (size_t)count * N * 8;
    const size_t b_ratio = (size_t)count * nt_max * 8, b_len = (size_t)count * N * nt_max * 8;
    CU(h->post_a.reserve(b_coef)); CU(h->post_b.reserve(b_T)); CU(h->post_c.reserve(b_r)); CU(h->post_d.reserve(b_ratio));
    CU(h->post_e.reserve(b_len));
    CU(cudaMemcpyAsync(h->post_a.p, coef, b_coef, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->post_b.p, T, b_T, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->post_c.p, radius, b_r, cudaMemcpyHostToDevice, h->stream));
    MetricsArgs A;
    A.count = count; A.N = N; A.M = M; A.nt_max = nt_max; A.downwash = downwash; A.dt = dt;
    A.coef = h->post_a.as<double>(); A.T = h->post_b.as<double>(); A.radius = h->post_c.as<double>();
    A.ratio_t = h->post_d.as<double>(); A.seglen = h->post_e.as<double>();
    size_t smem = ((size_t)N * 6 + 8) * 8;
    if (smem > h->smem_optin) return fail(h, RBPE_BAD_ARG, "rbpe_safety_metrics: N=%d needs %zu B of shared memory", N, smem);
    cudaFuncSetAttribute(metrics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin);
    metrics_kernel<<<dim3(nt_max, count), 256, smem, h->stream>>>(A);
    CU(cudaGetLastError());
    h->launches++;
    std::vector<double> ratio((size_t)count * nt_max), seglen((size_t)count * N * nt_max);
    CU(cudaMemcpyAsync(ratio.data(), h->post_d.p, b_ratio, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(seglen.data(), h->post_e.p, b_len, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < count; c++) {   // final reductions in the reference's loop order (L685-L695, L776-L797)
        int nt = (int)floor(T[(size_t)c * (M + 1) + M] / dt);
        double best = 1e9, tb = 0, len = 0;
        for (int i = 0; i < nt; i++)
            if (ratio[(size_t)c * nt_max + i] < best) { best = ratio[(size_t)c * nt_max + i]; tb = i * dt; }
        for (int q = 0; q < N; q++)
            for (int i = 0; i + 1 < nt; i++) len += seglen[((size_t)c * N + q) * nt_max + i];
        min_ratio[c] = best; t_at_min[c] = tb; length[c] = len;
    }
    return RBPE_OK;
}

#ifdef RBPE_PROFILE
// profile builds only (tools/): cumulative clock64 ticks of thread 0 per phase {setup, row passes, factor, solves, other}
extern "C" int rbpe_prof_read(unsigned long long *out, int reset) {
    unsigned long long z[16] = {0}; z[13] = ~0ull;
    if (cudaMemcpyFromSymbol(out, rbpe::g_prof, sizeof(z)) != cudaSuccess) return RBPE_CUDA_ERROR;
    if (reset && cudaMemcpyToSymbol(rbpe::g_prof, z, sizeof(z)) != cudaSuccess) return RBPE_CUDA_ERROR;
    return RBPE_OK;
}
#endif
