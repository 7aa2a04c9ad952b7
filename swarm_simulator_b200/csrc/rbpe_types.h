// rbpe_types.h -- launch descriptors shared by the host API (rbpe_api.cu) and the kernels.
#pragma once
#include <stddef.h>

namespace rbpe {

constexpr int NCP = 6;            // n+1 control points per segment (n = 5 only: rbp_planner.hpp L328, L361)
constexpr int CTA_THREADS = 256;  // PDIP kernel block size (throughput regime: two CTAs per SM)
constexpr int CTA_THREADS_MAX = 512;  // ... in the latency regime (at most one CTA per SM): same 128-register budget per thread
constexpr int MAX_M = 64;
constexpr int RBPE_MAX_PEERS = 8;   // one node: up to 8 GPUs behind one NVSwitch

// status codes == include/rbpe.h
enum { ST_OK = 0, ST_INFEASIBLE = 1, ST_NOT_CONVERGED = 2, ST_BAD_ARG = 3 };

// Raw mission inputs on the device (what rbpe_problem points to, after H2D) + what the assembly kernel makes of them.
struct AssembleArgs {
    int count, N, M, sequential;
    const double *T;         // [count][M+1]
    const int *sfc_offs;     // [count][N+1]
    const int *sfc_base;     // [count+1]
    const double *sfc_box;   // [nbox][6]
    const double *sfc_t;     // [nbox]
    const float *rsfc_n;     // [count][P][M][3]
    const double *rsfc_t;    // [count][P][M]
    const float *init_traj;  // [count][N][M+1][3]
    // outputs
    double *segbox;          // [count][N][M][6]   box chosen for every segment (build_dlq box part, L443-L474)
    float *reln;             // [count][P][M][3]   normal chosen for every pair/segment (build_dlq RSFC part, L476-L504)
    double *ctrl;            // [count][N][3][6M]  `dummy` (build_dummy L513-L549), zeros when !sequential
    double *segmat;          // [count][M][SEGMAT] per-segment constants, see SEGMAT_* below
    int *status;             // [count] ST_BAD_ARG when an SFC / RSFC look-up runs off the end (UB in the reference)
};

// per-segment constant block: AL[3][6], AR[3][6], qscale, tpow[6], CL[3][3], CR[3][3], RQ[6][6]
constexpr int SEGMAT_AL = 0;      // left-knot equality coefficients of this segment  (build_Aeq_base L353-L405)
constexpr int SEGMAT_AR = 18;     // right-knot equality coefficients
constexpr int SEGMAT_QS = 36;     // dt^(-2 phi + 1) (build_Q_p L349-L351)
constexpr int SEGMAT_TP = 37;     // (1/dt)^(5-j), j = 0..5 (timeMatrix L695-L700)
constexpr int SEGMAT_CL = 44;     // control points 0..2 = CL[i][d] * (pos, vel, acc)[d] at the left knot
constexpr int SEGMAT_CR = 53;     // control points 3..5 = CR[i-3][d] * state[d] at the right knot
constexpr int SEGMAT_RQ = 62;     // 6x6 cost Hessian over (left state, right state): C'(2 dt^-5 Q_base)C
constexpr int SEGMAT = 98;

struct SolveArgs {
    int count, N, M;
    int bs;             // effective batch size (agents per QP)
    int nbatch;         // effective batch_iter (QPs per outer iteration)
    int iteration;
    int sequential;
    int mode;           // 0 Gauss-Seidel chain inside one CTA per mission, 1 Jacobi (one CTA per (mission, batch))
    int batch_begin, batch_end;  // Jacobi: range of batches solved by this launch
    int rec_offset;     // Jacobi: record slot offset (outer iteration * nbatch)
    int max_iter;
    double tol_gap, tol_res;
    const double *start, *goal, *radius;  // [count][N][9], [count][N][9], [count][N]
    const double *segbox;   // [count][N][M][6]
    const float *reln;      // [count][P][M][3]
    const double *segmat;   // [count][M][SEGMAT]
    double *ctrl;           // [count][N][3][6M] control-point table, updated in place by solved batches
    const double *ctrl_frozen;  // Jacobi: table the RSFC rows are built from (copy of ctrl before the sweep)
    double *qp_obj;         // [count][nrec]
    int *qp_iters;          // [count][nrec]
    int *qp_status;         // [count][nrec]
    double *qp_res;         // [count][nrec][4]
    int nrec;
    int *status;            // [count] mission status (assembly may already have set BAD_ARG)
    // Jacobi exchange fused into the sweep (multi-GPU, NVLink peer memory): every solved batch is stored into the next
    // table of every rank (own one included), and the last work item of the launch raises this rank's flag on every peer
    int npeer;              // 0: single-GPU behaviour (results go to `ctrl` only)
    int peer_rank;
    double *peer_ctrl[RBPE_MAX_PEERS];               // next control-point table [count][N][3][6M] of every rank
    unsigned long long *peer_flags[RBPE_MAX_PEERS];  // flag words [RBPE_MAX_PEERS] of every rank; slot [peer_rank] is ours there
    unsigned long long sweep_id;                     // value to raise (monotone)
    unsigned int *done_counter;                      // work items finished in this launch (reset by the last one)
    unsigned int work_items;
    double *scratch;        // per-CTA global scratch
    size_t scratch_stride;  // doubles per CTA
    unsigned smem_bytes;    // dynamic shared memory given to the kernel
    unsigned panel_bytes;   // leading part of it reserved for the TMA-staged factor panel (joint batches; 0 = no staging)
};

struct ConvertArgs {
    int count, N, M;
    const double *ctrl;     // [count][N][3][6M]
    const double *segmat;
    double *coef;           // [count][N][3][6M]
};

}  // namespace rbpe
