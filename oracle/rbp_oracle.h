/*
 * rbp_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C float64 restatement of the reference hot path
 *   SwarmPlanning::RBPPlanner::update()   /root/reference/swarm_planner/include/rbp_planner.hpp L33-L84
 * i.e. constraint assembly (L100-L109, L327-L549), the per-batch QP build (populatebyrow L551-L688),
 * the sequential (Gauss-Seidel) batch loop with Bernstein->monomial conversion and `dummy` update
 * (solveQP L111-L206), timeMatrix (L695-L700) and the batch partition (L849-L881).
 *
 * The arithmetic of the solve itself lives in a closed-source dependency that is NOT under
 * /root/reference: IBM ILOG CPLEX 12.10 (CMakeLists.txt L41-L47; README L23 says 12.9), called through
 * the Concert API at rbp_planner.hpp L115, L144-L165. Its published algorithm for continuous convex QPs
 * (barrier = primal-dual predictor-corrector interior point on the normal equations, complementarity
 * tolerance 1e-8) is restated here as a Mehrotra predictor-corrector method; parity is anchored on the
 * reference's only frozen CPLEX artefacts, log/QPmodel.lp + log/coef*.csv (see tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#ifndef RBP_ORACLE_H
#define RBP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_problem {
    int N;              /* mission.qn */
    int M;              /* T.size()-1, rbp_planner.hpp L35 */
    int sequential;     /* param.sequential */
    int batch_size;     /* param.batch_size */
    int batch_iter;     /* param.batch_iter (-1 => all) */
    int iteration;      /* param.iteration */
    const double *T;         /* [M+1] segment times */
    const double *start;     /* [N][9] mission.startState (pos, vel, acc) */
    const double *goal;      /* [N][9] mission.goalState */
    const double *radius;    /* [N] mission.quad_size */
    const int    *sfc_offs;  /* [N+1] CSR offsets into sfc_box / sfc_t : planResult->SFC[qi] */
    const double *sfc_box;   /* [nbox][6] xmin,ymin,zmin,xmax,ymax,zmax */
    const double *sfc_t;     /* [nbox] box end time (.second) */
    const float  *rsfc_n;    /* [P][M][3] RSFC[qi][qj][ri].first (float32), pairs qi<qj lexicographic */
    const double *rsfc_t;    /* [P][M]    RSFC[qi][qj][ri].second */
    const float  *init_traj; /* [N][M+1][3] planResult->initTraj (float32) */
} oracle_problem;

/* One batch QP in general sparse form:  min x'Qx  s.t.  A x = b,  G x <= h   (objective WITHOUT 1/2,
 * as the reference's IloMinimize(cost), rbp_planner.hpp L581-L605). Variable order = reference order
 * row = k*offset_dim + bi*offset_quad + m*6 + i (L561); rows in the order populatebyrow adds them. */
typedef struct oracle_qp {
    int nv, ne, mi;
    int qnnz; int *qi, *qj; double *qv;          /* Q triplets (both triangles, as L597-L599) */
    int *a_ptr, *a_idx; double *a_val, *b;       /* CSR equalities */
    int *g_ptr, *g_idx; double *g_val, *h;       /* CSR inequalities in <= form */
    int *perm_x;  /* [nv] position of variable in segment-major order (solver ordering only) */
    int *perm_y;  /* [ne] position of equality row in knot-major order */
    int n_box_rows, n_rsfc_rows;
} oracle_qp;

typedef struct oracle_solver_opts {
    int max_iter;      /* default 100 */
    double tol_gap;    /* default 1e-10 : mu = s'z/mi relative to max(1,|obj|) */
    double tol_res;    /* default 1e-9  : relative primal/dual residual */
} oracle_solver_opts;

/* status codes shared with include/rbpe.h */
enum { ORACLE_OK = 0, ORACLE_INFEASIBLE = 1, ORACLE_NOT_CONVERGED = 2, ORACLE_BAD_ARG = 3 };

/* constants: Q_base[36], basis[36] row-major (L327-L347) */
void oracle_build_Q_base(double *Q_base, double *basis);
/* Aeq_base ((3M+3) x 6M, row-major) from T (L353-L405) */
void oracle_build_Aeq_base(const double *T, int M, double *Aeq);
/* deq [N*(3M+3)][3] row-major (L408-L432) */
void oracle_build_deq(const oracle_problem *p, double *deq);
/* compact dlq (L435-L511): box_ub[N][6M][3], box_lbneg[N][6M][3] (= -lower), rel_n[P][6M][3] (double)
 * returns 0, or ORACLE_BAD_ARG if a box / rsfc index runs off the end (UB in the reference). */
int oracle_build_dlq(const oracle_problem *p, double *box_ub, double *box_lbneg, double *rel_n);
/* dummy [N*6M][3] (L513-L549) */
void oracle_build_dummy(const oracle_problem *p, double *dummy);
/* batch partition (L849-L872): writes effective batch_size / batch_iter, returns batch count */
int oracle_set_batch(const oracle_problem *p, int *eff_batch_size, int *eff_batch_iter);

/* populatebyrow for batch l (L551-L688). dummy may be NULL iff !sequential. */
oracle_qp *oracle_populate(const oracle_problem *p, const double *dummy, int l);
void oracle_qp_free(oracle_qp *qp);

/* Mehrotra predictor-corrector; x[nv] out. obj = x'Qx (cplex.getObjValue convention). */
int oracle_solve_qp(const oracle_qp *qp, const oracle_solver_opts *opts, double *x, double *obj,
                    int *iters, double *res_out /*[4]: gap, rp, rd, rg or NULL*/);

/* Whole update() minus timeScale/ROS: coef[N][3][6M] (column-major per agent = msgs_traj_coef layout,
 * rows highest power first), ctrl[N][3][6M] (final dummy / Bernstein control points; for !sequential the
 * solved control points), per-batch obj/iters/status arrays sized iteration*batch_iter (may be NULL). */
int oracle_update(const oracle_problem *p, const oracle_solver_opts *opts, double *coef, double *ctrl,
                  double *batch_obj, int *batch_iters, int *batch_status, int nthreads_unused);

/* Solve `count` independent missions with OpenMP across missions (CPU baseline timing helper). */
int oracle_update_many(const oracle_problem *ps, int count, const oracle_solver_opts *opts,
                       double **coef, double **ctrl, int *status, int nthreads);

/* Corridor::updateRelBox (rbp_corridor.hpp L338-L398) in float32 (octomath::Vector3 semantics). Returns 1 if a normal is zero. */
int oracle_rsfc(int N, int M, const float *init_traj, const double *T, double downwash, float *rsfc_n, double *rsfc_t);

/* RBPPublisher post-hoc checks (rbp_publisher.hpp L47-L51, L169-L183, L670-L695, L769-L798): returns the sample count. */
int oracle_safety_metrics(int N, int M, const double *coef, const double *T, const double *radius, double downwash,
                          double dt, double *min_ratio, double *t_at_min, double *length);

#ifdef __cplusplus
}
#endif
#endif
