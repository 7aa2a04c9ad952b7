/*
 * rbpe.h -- C ABI of the B200 batched RBP trajectory-QP engine ("rbpe").
 *
 * Drop-in boundary for the inside of SwarmPlanning::RBPPlanner::update()
 *   /root/reference/swarm_planner/include/rbp_planner.hpp L33-L84
 * i.e. buildConstMtx (L100-L109), solveQP (L111-L206: populatebyrow L551-L688 + the CPLEX call L158 +
 * Bernstein->monomial conversion and `dummy` propagation L167-L196), with the batch partition of
 * setBatch (L849-L872).  The host-side mirror of the reference class
 * (swarm_simulator_b200/host/rbp_planner.hpp) and the Python binding are thin wrappers over this file.
 *
 * Plain pointers and sizes only.  The caller owns every host buffer; the engine owns all device memory.
 * No exceptions, no callbacks, no C++ or torch types cross this boundary.  There is no CPU fallback:
 * every entry point that computes fails with RBPE_CUDA_ERROR when no sm_100 device is usable.
 */
#ifndef RBPE_H
#define RBPE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* return / status codes (RBPPlanner::update() returns `false` for anything but RBPE_OK, cf. L62-L69, L158-L161) */
enum {
    RBPE_OK = 0,
    RBPE_INFEASIBLE = 1,     /* `!cplex.solve()` -> throw(-1), rbp_planner.hpp L158-L161 */
    RBPE_NOT_CONVERGED = 2,
    RBPE_BAD_ARG = 3,
    RBPE_CUDA_ERROR = 4
};

enum { RBPE_MODE_GAUSS_SEIDEL = 0, /* reference semantics: batches solved in order through `dummy` (L140-L201) */
       RBPE_MODE_JACOBI = 1        /* every batch of a sweep solved against the frozen table, exchanged after */ };

typedef struct rbpe_handle rbpe_handle;

typedef struct rbpe_config {
    int device;             /* CUDA ordinal */
    int max_iter;           /* interior-point iteration cap (<=0 -> 100) */
    double tol_gap;         /* complementarity gap (<=0 -> 1e-10), relative to max(1,|obj|) */
    double tol_res;         /* relative primal/dual residual (<=0 -> 1e-9) */
    size_t smem_budget;     /* bytes of dynamic shared memory a QP may use (0 -> engine default) */
    int reserved[6];        /* reserved[0]: CTA size of the PDIP kernel (0 -> default 128); rest must be 0 */
} rbpe_config;

/*
 * One mission = the inputs RBPPlanner reads: Mission (mission.hpp L10-L19), Param (param.hpp L8-L42) and
 * PlanResult{T, initTraj, SFC, RSFC} (sp_const.hpp L16-L28).
 * rbpe_solve_many() takes `count` missions of IDENTICAL shape (N, M, batching) packed back to back:
 * every pointer below addresses [count] consecutive per-mission blocks.
 */
typedef struct rbpe_problem {
    int N;                  /* mission.qn */
    int M;                  /* planResult->T.size()-1 (rbp_planner.hpp L35) */
    int sequential;         /* param.sequential */
    int batch_size;         /* param.batch_size */
    int batch_iter;         /* param.batch_iter (<0 or too large -> ceil(N/batch_size), L856-L859) */
    int iteration;          /* param.iteration */
    const double *T;        /* [count][M+1]  planResult->T */
    const double *start;    /* [count][N][9] mission.startState (pos, vel, acc) */
    const double *goal;     /* [count][N][9] mission.goalState */
    const double *radius;   /* [count][N]    mission.quad_size */
    const int *sfc_offs;    /* [count][N+1]  CSR over planResult->SFC[qi], offsets relative to the mission's first box */
    const int *sfc_base;    /* [count+1]     first box of every mission inside sfc_box / sfc_t */
    const double *sfc_box;  /* [nbox_total][6] xmin,ymin,zmin,xmax,ymax,zmax  (SFC[qi][bi].first) */
    const double *sfc_t;    /* [nbox_total]    SFC[qi][bi].second */
    const float *rsfc_n;    /* [count][P][M][3] RSFC[qi][qj][ri].first (float32), P=N(N-1)/2 pairs qi<qj lexicographic */
    const double *rsfc_t;   /* [count][P][M]    RSFC[qi][qj][ri].second */
    const float *init_traj; /* [count][N][M+1][3] planResult->initTraj (float32); may be NULL iff !sequential */
} rbpe_problem;

typedef struct rbpe_result {
    double *coef;        /* [count][N][3][6M]  per agent the msgs_traj_coef layout (column-major M(n+1) x 3, L286-L289):
                            per segment 6 monomial coefficients, highest power first, local unnormalised time */
    double *ctrl;        /* [count][N][3][6M]  Bernstein control points (final `dummy`); may be NULL */
    double *qp_obj;      /* [count][iteration*batch_iter]  cplex.getObjValue() convention x'Qx; may be NULL */
    int *qp_iters;       /* [count][iteration*batch_iter]  interior-point iterations; may be NULL */
    int *qp_status;      /* [count][iteration*batch_iter]  RBPE_* per batch QP; may be NULL */
    double *qp_res;      /* [count][iteration*batch_iter][4] gap, |rp|, |rd|, |rg| at exit; may be NULL */
    int *status;         /* [count] RBPE_* per mission (first failing batch, as update() would abort) */
} rbpe_result;

/* timing of the last rbpe_solve_many() on the engine's stream (CUDA events), milliseconds */
typedef struct rbpe_timing {
    float h2d_ms, assemble_ms, solve_ms, d2h_ms, total_ms;
    int kernel_launches;
} rbpe_timing;

int rbpe_create(const rbpe_config *cfg, rbpe_handle **out);
void rbpe_destroy(rbpe_handle *h);
const char *rbpe_last_error(const rbpe_handle *h); /* never NULL; h may be NULL (creation errors) */

/* Page-locked host memory for input / result buffers (optional; any host memory works).  With pageable buffers the
 * asynchronous copies of the pipelined rbpe_solve_many degrade to blocking ones and copies stop overlapping kernels. */
void *rbpe_host_alloc(size_t bytes);
void rbpe_host_free(void *p);

/* effective batch partition of RBPPlanner::setBatch (L849-L872): returns ceil(N/batch_size) */
int rbpe_set_batch(int N, int sequential, int batch_size, int batch_iter, int *eff_batch_size, int *eff_batch_iter);

/* Whole hot path for `count` independent missions with HOST buffers: H2D, assembly kernel, PDIP kernel,
 * conversion, D2H.  Returns RBPE_OK when every mission succeeded, else the first non-OK code
 * (per-mission detail in result->status). */
int rbpe_solve_many(rbpe_handle *h, const rbpe_problem *p, int count, int mode, rbpe_result *r);

/* Single mission == what RBPPlanner::update() does once. */
int rbpe_solve(rbpe_handle *h, const rbpe_problem *p, rbpe_result *r);

/* ---- resident (device-side) interface: inputs stay in HBM across calls; used for kernel-only timing and
 * for the multi-GPU Jacobi mode where the caller exchanges control points between sweeps. ---- */
int rbpe_upload(rbpe_handle *h, const rbpe_problem *p, int count);          /* H2D of the raw inputs */
int rbpe_assemble(rbpe_handle *h);                                           /* assembly kernel; resets `dummy` */
int rbpe_run(rbpe_handle *h, int mode);                                      /* assembly + PDIP + conversion kernels on resident inputs */
/* Jacobi sweep restricted to batches [batch_begin, batch_end) of every resident mission (agent sharding);
 * needs rbpe_upload + rbpe_assemble first; does not reset the control-point table */
int rbpe_run_jacobi_range(rbpe_handle *h, int batch_begin, int batch_end);
/* ---- Jacobi exchange over NVLink peer memory (one process per GPU, or several handles in one process) ----
 * Replaces "sweep, then all-gather of control points" (the exchange step the north-star names; the reference itself
 * has no such step: rbp_planner.hpp L140-L201 is a Gauss-Seidel chain) by ONE sweep kernel whose epilogue stores every
 * solved batch straight into the next table of every rank and raises a flag there.
 *   rbpe_peer_export   after rbpe_upload + rbpe_assemble: pins the two table buffers and the flag words and returns
 *                      their 3 IPC handles (3 x RBPE_IPC_HANDLE_BYTES bytes); exchange them between ranks with any
 *                      host-side transport (bench.py / dist.py use torch.distributed.all_gather_object)
 *   rbpe_peer_attach   all_handles = [world][3][RBPE_IPC_HANDLE_BYTES], rank order
 *   rbpe_peer_attach_local  same, for handles living in the calling process (raw device pointers)
 *   rbpe_run_jacobi_fused   one sweep over batches [batch_begin, batch_end) of every resident mission; every rank
 *                      must call it the same number of times (ranks with an empty range included)
 *   rbpe_peer_status   RBPE_CUDA_ERROR after a flag wait timed out (about 20 s: a peer died) */
#define RBPE_IPC_HANDLE_BYTES 64
int rbpe_peer_export(rbpe_handle *h, unsigned char *handles);
int rbpe_peer_attach(rbpe_handle *h, int rank, int world, const unsigned char *all_handles);
int rbpe_peer_attach_local(rbpe_handle *h, int rank, int world, rbpe_handle *const *peers);
int rbpe_run_jacobi_fused(rbpe_handle *h, int batch_begin, int batch_end);
int rbpe_peer_status(rbpe_handle *h);
int rbpe_set_ctrl(rbpe_handle *h, const double *ctrl);                       /* H2D: overwrite `dummy` [count][N][3][6M] */
int rbpe_download(rbpe_handle *h, rbpe_result *r);                           /* D2H of results */
/* device pointers of the resident control-point table [count][N][3][6M] (f64) and coefficient table */
double *rbpe_device_ctrl(rbpe_handle *h);
double *rbpe_device_coef(rbpe_handle *h);
int *rbpe_device_status(rbpe_handle *h);                                     /* [count] mission status words (int32) on the device */
/* k3 alone on the current control-point table (after a collective exchange of the Jacobi mode); replaces the conversion
 * loop of rbp_planner.hpp L167-L196 for the agents other ranks solved */
int rbpe_convert(rbpe_handle *h);
/* which solver kernel the last k2 launch used (diagnostics, tests): 0 none yet, 1 = one warp per QP (pdip1_kernel),
 * 2 = several warps per QP (pdip1x_kernel, latency regime), 3 = one CTA per QP (pdip_kernel); *threads_per_qp = 32, 32 * warps, CTA size */
int rbpe_last_solver(rbpe_handle *h, int *threads_per_qp);
void *rbpe_stream(rbpe_handle *h);                                           /* cudaStream_t */
int rbpe_sync(rbpe_handle *h);
int rbpe_last_timing(const rbpe_handle *h, rbpe_timing *t);                  /* kernel_launches = total since creation */
/* CUDA-event stopwatch on the engine's stream: start synchronises first; stop returns the elapsed milliseconds */
int rbpe_timer_start(rbpe_handle *h);
int rbpe_timer_stop(rbpe_handle *h, float *ms);

/* ---- the two pure-arithmetic neighbours of the path (host buffers in, host buffers out) ---- */
/* Corridor::updateRelBox (rbp_corridor.hpp L338-L398): RSFC normals (float32, octomath::Vector3 semantics) and end
 * times from initTraj [count][N][M+1][3]; collided[c] = 1 where the reference reports "initial trajectories are
 * collided" (L385-L388).  Output layout = rbpe_problem.rsfc_n / rsfc_t. */
int rbpe_corridor_rsfc(rbpe_handle *h, int N, int M, int count, const float *init_traj, const double *T, double downwash,
                       float *rsfc_n, double *rsfc_t, int *collided);
/* RBPPublisher::plot (rbp_publisher.hpp L117-L127; update_quad_state L670-L683, update_safety_margin_ratio L769-L798,
 * trajectory_length_sum L685-L695): per mission the safety margin ratio (collision-free iff >= 1), the sample time of
 * its minimum and the total flight length, sampled every dt (the reference uses 0.1). */
int rbpe_safety_metrics(rbpe_handle *h, int N, int M, int count, const double *coef, const double *T, const double *radius,
                        double downwash, double dt, double *min_ratio, double *t_at_min, double *length);

#ifdef __cplusplus
}
#endif
#endif
