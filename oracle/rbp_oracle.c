/*
 * rbp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See rbp_oracle.h.
 *
 * Every function names the lines of /root/reference/swarm_planner/include/rbp_planner.hpp ("RP")
 * it restates. No reference source text is copied; hard-coded matrices are re-derived in comments
 * and checked by tests/test_oracle_matrices.py against their closed forms.
 */
#include "rbp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NCP 6 /* n+1 control points per segment (n = 5 only, RP L328, L361) */
#define PHI 3

/* ------------------------------------------------------------------------------------------------
 * Constants (RP L327-L347).
 * Q_base = 60^2 * D3' * G2 * D3 : D3 = third forward difference (3x6), G2[i][j] = C(2,i)C(2,j)/(5 C(4,i+j))
 * = Gram matrix of the degree-2 Bernstein basis; i.e. the integral of the squared third derivative of a
 * quintic Bezier curve on [0,1]. basis[i][j] = coefficient of t^(5-j) in B_i^5(t).
 * ---------------------------------------------------------------------------------------------- */
static double binom(int n, int k) {
    double r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}

void oracle_build_Q_base(double *Q, double *basis) {
    double D3[3][6] = {{0}}, G2[3][3];
    for (int r = 0; r < 3; r++) { /* third difference: -1, 3, -3, 1 */
        D3[r][r] = -1; D3[r][r + 1] = 3; D3[r][r + 2] = -3; D3[r][r + 3] = 1;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) G2[i][j] = binom(2, i) * binom(2, j) / (5.0 * binom(4, i + j));
    for (int a = 0; a < 6; a++)
        for (int b = 0; b < 6; b++) {
            double s = 0;
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) s += D3[i][a] * G2[i][j] * D3[j][b];
            Q[a * 6 + b] = rint(3600.0 * s); /* entries are exact integers (multiples of 120) */
        }
    /* B_i^5(t) = C(5,i) t^i (1-t)^(5-i) = C(5,i) sum_l C(5-i,l) (-1)^l t^(i+l); column j <-> power 5-j */
    for (int i = 0; i < 36; i++) basis[i] = 0;
    for (int i = 0; i < 6; i++)
        for (int l = 0; l <= 5 - i; l++) {
            int pw = i + l;
            basis[i * 6 + (5 - pw)] = binom(5, i) * binom(5 - i, l) * ((l & 1) ? -1.0 : 1.0);
        }
}

/* i-th forward difference stencil at tau=0 (A_0.row(i)) / backward at tau=1 (A_T.row(i)), RP L362-L374 */
static void diff_rows(int i, double *a0, double *aT) {
    for (int c = 0; c < 6; c++) a0[c] = aT[c] = 0;
    for (int c = 0; c <= i; c++) {
        double v = binom(i, c) * (((i - c) & 1) ? -1.0 : 1.0); /* (-1)^(i-c) C(i,c) */
        a0[c] = v;
        aT[5 - i + c] = v;
    }
}

void oracle_build_Aeq_base(const double *T, int M, double *Aeq) {
    int rows = 2 * PHI + (M - 1) * PHI, cols = M * NCP;
    memset(Aeq, 0, sizeof(double) * rows * cols);
    double a0[6], aT[6];
    int nn = 1;
    for (int i = 0; i < PHI; i++) { /* waypoints, RP L380-L387 */
        diff_rows(i, a0, aT);
        double s0 = pow(T[1] - T[0], -i) * nn, sT = pow(T[M] - T[M - 1], -i) * nn;
        for (int c = 0; c < 6; c++) {
            Aeq[i * cols + c] = s0 * a0[c];
            Aeq[(PHI + i) * cols + NCP * (M - 1) + c] = sT * aT[c];
        }
        nn = nn * (5 - i);
    }
    for (int m = 1; m < M; m++) { /* continuity, RP L390-L399 */
        nn = 1;
        for (int j = 0; j < PHI; j++) {
            diff_rows(j, a0, aT);
            double sp = pow(T[m] - T[m - 1], -j) * nn, sn = -pow(T[m + 1] - T[m], -j) * nn;
            int r = 2 * PHI + PHI * (m - 1) + j;
            for (int c = 0; c < 6; c++) {
                Aeq[r * cols + NCP * (m - 1) + c] = sp * aT[c];
                Aeq[r * cols + NCP * m + c] = sn * a0[c];
            }
            nn = nn * (5 - j);
        }
    }
}

void oracle_build_deq(const oracle_problem *p, double *deq) { /* RP L408-L432 */
    int rows = 2 * PHI + (p->M - 1) * PHI;
    memset(deq, 0, sizeof(double) * p->N * rows * 3);
    for (int qi = 0; qi < p->N; qi++)
        for (int k = 0; k < 3; k++)
            for (int d = 0; d < 3; d++) {
                deq[(qi * rows + d) * 3 + k] = p->start[qi * 9 + k + 3 * d];
                deq[(qi * rows + PHI + d) * 3 + k] = p->goal[qi * 9 + k + 3 * d];
            }
}

static long pair_index(int N, int qi, int qj) { /* lexicographic (qi<qj), = `iter` of RP L477-L503 */
    return (long)qi * N - (long)qi * (qi + 1) / 2 + (qj - qi - 1);
}

int oracle_build_dlq(const oracle_problem *p, double *box_ub, double *box_lbneg, double *rel_n) {
    int N = p->N, M = p->M, rc = 0;
    for (int qi = 0; qi < N; qi++) { /* RP L443-L474: bi is monotone over m (not reset) */
        int nb = p->sfc_offs[qi + 1] - p->sfc_offs[qi], bi = 0;
        for (int m = 0; m < M; m++) {
            while (bi < nb && p->sfc_t[p->sfc_offs[qi] + bi] < p->T[m + 1]) bi++;
            int b = bi;
            if (b >= nb) { rc = ORACLE_BAD_ARG; b = nb - 1; } /* reference reads past the end (UB) */
            const double *box = p->sfc_box + 6 * (size_t)(p->sfc_offs[qi] + b);
            for (int i = 0; i < NCP; i++)
                for (int k = 0; k < 3; k++) {
                    box_ub[((size_t)qi * NCP * M + NCP * m + i) * 3 + k] = box[3 + k];
                    box_lbneg[((size_t)qi * NCP * M + NCP * m + i) * 3 + k] = -box[k];
                }
        }
    }
    for (int qi = 0; qi < N; qi++) /* RP L476-L504: ri reset for every m */
        for (int qj = qi + 1; qj < N; qj++) {
            long it = pair_index(N, qi, qj);
            for (int m = 0; m < M; m++) {
                int ri = 0;
                while (ri < M && p->rsfc_t[it * M + ri] < p->T[m + 1]) ri++;
                if (ri >= M) { rc = ORACLE_BAD_ARG; ri = M - 1; }
                const float *nv = p->rsfc_n + (it * M + ri) * 3;
                for (int i = 0; i < NCP; i++)
                    for (int k = 0; k < 3; k++)
                        rel_n[((size_t)it * NCP * M + NCP * m + i) * 3 + k] = (double)nv[k];
            }
        }
    return rc;
}

void oracle_build_dummy(const oracle_problem *p, double *dummy) { /* RP L513-L549 */
    int N = p->N, M = p->M, oq = M * NCP, path = M + 1;
    for (int qi = 0; qi < N; qi++) {
        int m = 0, idx = 0;
        while (m < M) {
            if (idx >= path - 1) {
                idx = path - 1;
                for (int j = 0; j < NCP; j++)
                    for (int k = 0; k < 3; k++)
                        dummy[((size_t)qi * oq + m * NCP + j) * 3 + k] =
                            (double)p->init_traj[((size_t)qi * path + idx) * 3 + k];
            } else {
                for (int j = 0; j < NCP; j++) {
                    int a = (j < NCP / 2) ? 0 : 1;
                    for (int k = 0; k < 3; k++) /* (1-a)*pi_idx + a*pi_{idx+1}: float promoted to double */
                        dummy[((size_t)qi * oq + m * NCP + j) * 3 + k] =
                            (1 - a) * (double)p->init_traj[((size_t)qi * path + idx) * 3 + k] +
                            a * (double)p->init_traj[((size_t)qi * path + idx + 1) * 3 + k];
                }
            }
            m++;
            idx++;
        }
    }
}

int oracle_set_batch(const oracle_problem *p, int *ebs, int *ebi) { /* RP L849-L872 */
    int bs = p->batch_size, bit = p->batch_iter, N = p->N;
    int batch_max_iter = (int)ceil((double)N / (double)bs);
    if (p->sequential) {
        if (bit < 0 || bit > batch_max_iter) bit = batch_max_iter;
    } else {
        bs = N;
        bit = 1;
    }
    *ebs = bs;
    *ebi = bit;
    return batch_max_iter; /* batches.resize(batch_max_iter) uses the OLD batch_size (harmless) */
}

static int quad_in_batch(int qi, int l, int bs, int N) { /* RP L874-L881 on batches[qi / batch_size] */
    (void)N;
    return (qi / bs == l) ? qi - l * bs : -1;
}

/* ------------------------------------------------------------------------------------------------
 * populatebyrow (RP L551-L688)
 * ---------------------------------------------------------------------------------------------- */
void oracle_qp_free(oracle_qp *q) {
    if (!q) return;
    free(q->qi); free(q->qj); free(q->qv);
    free(q->a_ptr); free(q->a_idx); free(q->a_val); free(q->b);
    free(q->g_ptr); free(q->g_idx); free(q->g_val); free(q->h);
    free(q->perm_x); free(q->perm_y);
    free(q);
}

oracle_qp *oracle_populate(const oracle_problem *p, const double *dummy, int l) {
    int N = p->N, M = p->M, bs, bit;
    oracle_set_batch(p, &bs, &bit);
    int q0 = l * bs, nb = (q0 + bs <= N) ? bs : N - q0;
    if (nb <= 0) return NULL;
    int oq = M * NCP, od = nb * oq, nv = 3 * od, neq1 = 2 * PHI + (M - 1) * PHI;
    long P = (long)N * (N - 1) / 2;
    oracle_qp *q = (oracle_qp *)calloc(1, sizeof(oracle_qp));
    q->nv = nv;
    q->ne = 3 * nb * neq1;

    double Qb[36], basis[36];
    oracle_build_Q_base(Qb, basis);
    double *Aeq = (double *)malloc(sizeof(double) * neq1 * oq);
    oracle_build_Aeq_base(p->T, M, Aeq);
    double *deq = (double *)malloc(sizeof(double) * N * neq1 * 3);
    oracle_build_deq(p, deq);
    double *ub = (double *)malloc(sizeof(double) * (size_t)N * oq * 3);
    double *lbn = (double *)malloc(sizeof(double) * (size_t)N * oq * 3);
    double *rel = (double *)malloc(sizeof(double) * (size_t)(P > 0 ? P : 1) * oq * 3);
    oracle_build_dlq(p, ub, lbn, rel);

    /* cost, RP L581-L605: Q_p = Q_base * dt^(-2 phi + 1), skip exact zeros */
    q->qi = (int *)malloc(sizeof(int) * 36 * 3 * nb * M);
    q->qj = (int *)malloc(sizeof(int) * 36 * 3 * nb * M);
    q->qv = (double *)malloc(sizeof(double) * 36 * 3 * nb * M);
    int c = 0;
    for (int k = 0; k < 3; k++)
        for (int bi = 0; bi < nb; bi++)
            for (int m = 0; m < M; m++) {
                double sc = pow(p->T[m + 1] - p->T[m], -2 * PHI + 1);
                for (int i = 0; i < NCP; i++)
                    for (int j = 0; j < NCP; j++) {
                        double v = Qb[i * 6 + j] * sc;
                        if (v != 0) {
                            q->qi[c] = k * od + bi * oq + m * NCP + i;
                            q->qj[c] = k * od + bi * oq + m * NCP + j;
                            q->qv[c++] = v;
                        }
                    }
            }
    q->qnnz = c;

    /* equalities, RP L608-L622 */
    q->a_ptr = (int *)malloc(sizeof(int) * (q->ne + 1));
    q->a_idx = (int *)malloc(sizeof(int) * q->ne * 12);
    q->a_val = (double *)malloc(sizeof(double) * q->ne * 12);
    q->b = (double *)malloc(sizeof(double) * q->ne);
    q->perm_y = (int *)malloc(sizeof(int) * q->ne);
    int r = 0, nz = 0;
    for (int k = 0; k < 3; k++)
        for (int bi = 0; bi < nb; bi++)
            for (int i = 0; i < neq1; i++) {
                q->a_ptr[r] = nz;
                for (int j = 0; j < oq; j++)
                    if (Aeq[i * oq + j] != 0) {
                        q->a_idx[nz] = k * od + bi * oq + j;
                        q->a_val[nz++] = Aeq[i * oq + j];
                    }
                q->b[r] = deq[((q0 + bi) * neq1 + i) * 3 + k];
                int knot = (i < 3) ? 0 : (i < 6) ? M : (i - 6) / 3 + 1, d = (i < 6) ? i % 3 : (i - 6) % 3;
                q->perm_y[r] = knot * 9 * nb + (bi * 3 + k) * 3 + d;
                r++;
            }
    q->a_ptr[r] = nz;

    /* inequalities, RP L626-L684, all converted to <= form */
    long nrel = 0;
    for (int qi = 0; qi < N; qi++)
        for (int qj = qi + 1; qj < N; qj++)
            if (quad_in_batch(qi, l, bs, N) >= 0 || quad_in_batch(qj, l, bs, N) >= 0) nrel += oq;
    q->n_box_rows = 2 * nv;
    q->n_rsfc_rows = (int)nrel;
    q->mi = q->n_box_rows + q->n_rsfc_rows;
    q->g_ptr = (int *)malloc(sizeof(int) * (q->mi + 1));
    q->g_idx = (int *)malloc(sizeof(int) * ((size_t)q->n_box_rows + 6 * (size_t)nrel));
    q->g_val = (double *)malloc(sizeof(double) * ((size_t)q->n_box_rows + 6 * (size_t)nrel));
    q->h = (double *)malloc(sizeof(double) * q->mi);
    r = 0; nz = 0;
    for (int k = 0; k < 3; k++)
        for (int bi = 0; bi < nb; bi++) {
            int qi = q0 + bi;
            for (int j = 0; j < oq; j++) {
                int idx = k * od + bi * oq + j;
                q->g_ptr[r] = nz; q->g_idx[nz] = idx; q->g_val[nz++] = 1.0;    /*  x <= ub  */
                q->h[r++] = ub[((size_t)qi * oq + j) * 3 + k];
                q->g_ptr[r] = nz; q->g_idx[nz] = idx; q->g_val[nz++] = -1.0;   /* -x <= -lb */
                q->h[r++] = lbn[((size_t)qi * oq + j) * 3 + k];
            }
        }
    for (int qi = 0; qi < N; qi++)
        for (int qj = qi + 1; qj < N; qj++) {
            int bi = quad_in_batch(qi, l, bs, N), bj = quad_in_batch(qj, l, bs, N);
            if (bi < 0 && bj < 0) continue;
            long it = pair_index(N, qi, qj);
            double rr = p->radius[qi] + p->radius[qj];
            for (int j = 0; j < oq; j++) {
                const double *nv3 = rel + ((size_t)it * oq + j) * 3;
                q->g_ptr[r] = nz;
                double hc = -rr;
                if (bi >= 0 && bj < 0) { /* n.(dummy_qj - x_qi) >= r  <=>  n.x_qi <= n.dummy_qj - r */
                    for (int k = 0; k < 3; k++) {
                        hc += nv3[k] * dummy[((size_t)qj * oq + j) * 3 + k];
                        if (nv3[k] != 0) { q->g_idx[nz] = k * od + bi * oq + j; q->g_val[nz++] = nv3[k]; }
                    }
                } else if (bi < 0 && bj >= 0) { /* n.(x_qj - dummy_qi) >= r <=> -n.x_qj <= -n.dummy_qi - r */
                    for (int k = 0; k < 3; k++) {
                        hc -= nv3[k] * dummy[((size_t)qi * oq + j) * 3 + k];
                        if (nv3[k] != 0) { q->g_idx[nz] = k * od + bj * oq + j; q->g_val[nz++] = -nv3[k]; }
                    }
                } else { /* n.(x_qj - x_qi) >= r <=> n.x_qi - n.x_qj <= -r */
                    for (int k = 0; k < 3; k++)
                        if (nv3[k] != 0) {
                            q->g_idx[nz] = k * od + bi * oq + j; q->g_val[nz++] = nv3[k];
                            q->g_idx[nz] = k * od + bj * oq + j; q->g_val[nz++] = -nv3[k];
                        }
                }
                q->h[r++] = hc;
            }
        }
    q->g_ptr[r] = nz;

    q->perm_x = (int *)malloc(sizeof(int) * nv);
    for (int k = 0; k < 3; k++)
        for (int bi = 0; bi < nb; bi++)
            for (int m = 0; m < M; m++)
                for (int i = 0; i < NCP; i++)
                    q->perm_x[k * od + bi * oq + m * NCP + i] = ((m * nb + bi) * 3 + k) * NCP + i;
    free(Aeq); free(deq); free(ub); free(lbn); free(rel);
    return q;
}

/* ------------------------------------------------------------------------------------------------
 * Newton systems by the null-space method over the C2 knot states.
 *
 * The equality rows (build_Aeq_base, RP L353-L405) say: the (pos, vel, acc) of segment m-1 at tau=1 equals
 * that of segment m at tau=0, and the first / last knot equal the start / goal state.  Per segment and
 * (agent, axis) the 6x6 matrix E_m = [rows of the left knot ; rows of the right knot] restricted to the
 * segment's 6 control points is invertible (quintic Hermite interpolation), so with the knot states
 * sigma = (s_1 .. s_{M-1}) as free variables  x = x_p + Z sigma  parametrises {x : A x = b} exactly, and
 *     H dx + A' dy = r1, A dx = 0   <=>   (Z' H Z) dsigma = Z' r1,  dx = Z dsigma.
 * Z'HZ is symmetric positive definite (Q is positive definite on null(A)) and block tridiagonal over knots
 * (blocks of 9b).  Same Newton direction as the KKT system, without ever inverting H = 2Q + G'WG, which is
 * numerically singular near convergence (Q has rank 3 per 6x6 block and inactive rows' weights vanish).
 * This plays the role of CPLEX's own presolve / free-variable handling; E_m^-1 is obtained numerically from A.
 * ---------------------------------------------------------------------------------------------- */
static int chol_dense(int n, double *A, int ld) { /* lower triangle, in place */
    for (int j = 0; j < n; j++) {
        double d = A[j * ld + j];
        for (int t = 0; t < j; t++) d -= A[j * ld + t] * A[j * ld + t];
        if (!(d > 0)) return -1;
        d = sqrt(d);
        A[j * ld + j] = d;
        for (int i = j + 1; i < n; i++) {
            double v = A[i * ld + j];
            for (int t = 0; t < j; t++) v -= A[i * ld + t] * A[j * ld + t];
            A[i * ld + j] = v / d;
        }
    }
    return 0;
}
static int inv6(const double *E, double *Ei) { /* Gauss-Jordan with partial pivoting */
    double a[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { a[i][j] = E[i * 6 + j]; a[i][6 + j] = (i == j); }
    for (int c = 0; c < 6; c++) {
        int pv = c;
        for (int r = c + 1; r < 6; r++) if (fabs(a[r][c]) > fabs(a[pv][c])) pv = r;
        if (a[pv][c] == 0) return -1;
        if (pv != c) for (int j = 0; j < 12; j++) { double t = a[c][j]; a[c][j] = a[pv][j]; a[pv][j] = t; }
        double d = a[c][c];
        for (int j = 0; j < 12; j++) a[c][j] /= d;
        for (int r = 0; r < 6; r++) if (r != c) {
            double f = a[r][c];
            if (f != 0) for (int j = 0; j < 12; j++) a[r][j] -= f * a[c][j];
        }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ei[i * 6 + j] = a[i][6 + j];
    return 0;
}

typedef struct {
    const oracle_qp *q;
    int nv, ne, mi, M, nb, n, kb, nr;
    int *seg, *loc;       /* per variable: segment, local index (agent*3+axis)*6 + i */
    int *rseg;            /* per inequality row: segment of its variables */
    double *Zc;           /* [nv][6]: x_v = xp_v + sum_d Zc[v][d] s_m[ak][d] + Zc[v][3+d] s_{m+1}[ak][d] */
    double *xp;           /* [nv] particular solution (sigma = 0) */
    double *Hs;           /* [M][n*n] H per segment, full symmetric */
    double *Wd, *Wo;      /* reduced Hessian: [(M-1)][kb*kb] diagonal, [(M-2)][kb*kb] blocks (t+1,t) */
    double *T, *Wm;       /* scratch n*2kb, 2kb*2kb */
    double *w;
} nsp_t;

static void csr_mulv(int m, const int *ptr, const int *idx, const double *val, const double *x, double *y) {
    for (int r = 0; r < m; r++) {
        double s = 0;
        for (int t = ptr[r]; t < ptr[r + 1]; t++) s += val[t] * x[idx[t]];
        y[r] = s;
    }
}
static void csr_mulTv_add(int m, const int *ptr, const int *idx, const double *val, const double *y, double *x) {
    for (int r = 0; r < m; r++) {
        double yr = y[r];
        if (yr == 0) continue;
        for (int t = ptr[r]; t < ptr[r + 1]; t++) x[idx[t]] += val[t] * yr;
    }
}
static void p_mulv(const oracle_qp *q, const double *x, double *y) { /* y = (Q+Q')x */
    for (int i = 0; i < q->nv; i++) y[i] = 0;
    for (int t = 0; t < q->qnnz; t++) {
        y[q->qi[t]] += q->qv[t] * x[q->qj[t]];
        y[q->qj[t]] += q->qv[t] * x[q->qi[t]];
    }
}

static void nsp_free(nsp_t *K) {
    free(K->seg); free(K->loc); free(K->rseg); free(K->Zc); free(K->xp); free(K->Hs); free(K->Wd); free(K->Wo);
    free(K->T); free(K->Wm); free(K->w);
}

/* returns 0, or ORACLE_BAD_ARG when the equality structure is not the knot structure described above */
static int nsp_init(nsp_t *K, const oracle_qp *q) {
    memset(K, 0, sizeof(*K));
    K->q = q; K->nv = q->nv; K->ne = q->ne; K->mi = q->mi;
    int nv = q->nv, ne = q->ne;
    /* nv = 18 nb M, ne = 9 nb (M+1) */
    if (nv <= 0 || nv % 18 || ne % 9) return ORACLE_BAD_ARG;
    int a18 = nv / 18, a9 = ne / 9; /* nb*M, nb*(M+1) */
    int nb = a9 - a18;
    if (nb <= 0 || a18 % nb) return ORACLE_BAD_ARG;
    int M = a18 / nb, n = 18 * nb, kb = 9 * nb;
    K->M = M; K->nb = nb; K->n = n; K->kb = kb; K->nr = kb * (M - 1);
    K->seg = (int *)malloc(sizeof(int) * nv); K->loc = (int *)malloc(sizeof(int) * nv);
    int *varof = (int *)malloc(sizeof(int) * nv), *rowof = (int *)malloc(sizeof(int) * ne);
    for (int v = 0; v < nv; v++) { int p = q->perm_x[v]; K->seg[v] = p / n; K->loc[v] = p % n; varof[p] = v; }
    for (int r = 0; r < ne; r++) rowof[q->perm_y[r]] = r;
    K->Zc = (double *)calloc((size_t)nv * 6, sizeof(double));
    K->xp = (double *)calloc(nv, sizeof(double));
    int rc = 0;
    for (int m = 0; m < M && !rc; m++)
        for (int ak = 0; ak < 3 * nb && !rc; ak++) {
            double E[36] = {0}, Ei[36], rhs[6];
            for (int side = 0; side < 2; side++)
                for (int d = 0; d < 3; d++) {
                    int r = rowof[(m + side) * kb + ak * 3 + d];
                    for (int t = q->a_ptr[r]; t < q->a_ptr[r + 1]; t++) {
                        int c = q->a_idx[t], p = q->perm_x[c];
                        if (p / n == m) {
                            if ((p % n) / 6 != ak) rc = ORACLE_BAD_ARG;
                            else E[(side * 3 + d) * 6 + p % 6] = q->a_val[t];
                        } else if (p / n != m + (side ? 1 : -1)) rc = ORACLE_BAD_ARG;
                    }
                    /* the knot's right-hand side belongs to the side that closes it: the first knot to
                     * segment 0, every other knot to the segment on its left */
                    rhs[side * 3 + d] = (side == 1 || m == 0) ? q->b[r] : 0.0;
                }
            if (rc || inv6(E, Ei)) { rc = ORACLE_BAD_ARG; break; }
            for (int i = 0; i < 6; i++) {
                int v = varof[m * n + ak * 6 + i];
                double xpv = 0;
                for (int j = 0; j < 6; j++) xpv += Ei[i * 6 + j] * rhs[j];
                K->xp[v] = xpv;
                for (int d = 0; d < 3; d++) {
                    /* interior knot t: s_t := (rows of knot t on segment t-1) x_{t-1} - b_t  = -(rows on segment t) x_t */
                    K->Zc[(size_t)v * 6 + d] = (m > 0) ? -Ei[i * 6 + d] : 0.0;
                    K->Zc[(size_t)v * 6 + 3 + d] = (m < M - 1) ? Ei[i * 6 + 3 + d] : 0.0;
                }
            }
        }
    free(varof); free(rowof);
    if (rc) return rc;
    K->rseg = (int *)malloc(sizeof(int) * (q->mi + 1));
    for (int r = 0; r < q->mi; r++) {
        int sg = -1;
        for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1]; t++) {
            int s2 = K->seg[q->g_idx[t]];
            if (sg < 0) sg = s2; else if (sg != s2) return ORACLE_BAD_ARG;
        }
        K->rseg[r] = sg;
    }
    K->Hs = (double *)malloc(sizeof(double) * (size_t)M * n * n);
    K->Wd = (double *)malloc(sizeof(double) * (size_t)(M > 1 ? M - 1 : 1) * kb * kb);
    K->Wo = (double *)malloc(sizeof(double) * (size_t)(M > 2 ? M - 2 : 1) * kb * kb);
    K->T = (double *)malloc(sizeof(double) * (size_t)n * 2 * kb);
    K->Wm = (double *)malloc(sizeof(double) * (size_t)4 * kb * kb);
    K->w = (double *)malloc(sizeof(double) * (q->mi + 1));
    return 0;
}

/* out[nr] = Z' g */
static void nsp_Zt(const nsp_t *K, const double *g, double *out) {
    for (int i = 0; i < K->nr; i++) out[i] = 0;
    for (int v = 0; v < K->nv; v++) {
        int m = K->seg[v], ak = K->loc[v] / 6;
        const double *z = K->Zc + (size_t)v * 6;
        if (m > 0) for (int d = 0; d < 3; d++) out[(m - 1) * K->kb + ak * 3 + d] += z[d] * g[v];
        if (m < K->M - 1) for (int d = 0; d < 3; d++) out[m * K->kb + ak * 3 + d] += z[3 + d] * g[v];
    }
}
/* out[nv] = Z sigma */
static void nsp_Z(const nsp_t *K, const double *sg, double *out) {
    for (int v = 0; v < K->nv; v++) {
        int m = K->seg[v], ak = K->loc[v] / 6;
        const double *z = K->Zc + (size_t)v * 6;
        double s = 0;
        if (m > 0) for (int d = 0; d < 3; d++) s += z[d] * sg[(m - 1) * K->kb + ak * 3 + d];
        if (m < K->M - 1) for (int d = 0; d < 3; d++) s += z[3 + d] * sg[m * K->kb + ak * 3 + d];
        out[v] = s;
    }
}

/* reduced Hessian Z'(P + G'WG)Z and its block tridiagonal Cholesky */
static int nsp_factor(nsp_t *K, const double *w) {
    const oracle_qp *q = K->q;
    int M = K->M, n = K->n, kb = K->kb, k2 = 2 * kb;
    if (M < 2) return 0;
    memset(K->Hs, 0, sizeof(double) * (size_t)M * n * n);
    for (int t = 0; t < q->qnnz; t++) { /* P = Q + Q' */
        int a = q->qi[t], b = q->qj[t];
        double *Hm = K->Hs + (size_t)K->seg[a] * n * n;
        Hm[K->loc[a] * n + K->loc[b]] += q->qv[t];
        Hm[K->loc[b] * n + K->loc[a]] += q->qv[t];
    }
    for (int r = 0; r < q->mi; r++) {
        double wr = w[r];
        if (wr == 0 || K->rseg[r] < 0) continue;
        double *Hm = K->Hs + (size_t)K->rseg[r] * n * n;
        for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1]; t++) {
            double va = wr * q->g_val[t];
            int la = K->loc[q->g_idx[t]];
            for (int u = q->g_ptr[r]; u < q->g_ptr[r + 1]; u++) Hm[la * n + K->loc[q->g_idx[u]]] += va * q->g_val[u];
        }
    }
    memset(K->Wd, 0, sizeof(double) * (size_t)(M - 1) * kb * kb);
    if (M > 2) memset(K->Wo, 0, sizeof(double) * (size_t)(M - 2) * kb * kb);
    /* Zc of the variable at local index l of segment m: need a map (m, l) -> v; loc/seg give the inverse, so
     * gather the per-segment Z coefficients into a dense [n][6] table first */
    double *zt = (double *)malloc(sizeof(double) * (size_t)M * n * 6);
    for (int v = 0; v < K->nv; v++) memcpy(zt + ((size_t)K->seg[v] * n + K->loc[v]) * 6, K->Zc + (size_t)v * 6, 48);
    for (int m = 0; m < M; m++) {
        const double *Hm = K->Hs + (size_t)m * n * n, *zm = zt + (size_t)m * n * 6;
        /* T[r][(side,ak,d)] = sum_i Hm[r][(ak,i)] z[(ak,i)][side*3+d] */
        for (int r = 0; r < n; r++)
            for (int side = 0; side < 2; side++)
                for (int ak = 0; ak < 3 * K->nb; ak++)
                    for (int d = 0; d < 3; d++) {
                        double s = 0;
                        for (int i = 0; i < 6; i++) s += Hm[r * n + ak * 6 + i] * zm[(ak * 6 + i) * 6 + side * 3 + d];
                        K->T[(size_t)r * k2 + side * kb + ak * 3 + d] = s;
                    }
        for (int side = 0; side < 2; side++)
            for (int ak = 0; ak < 3 * K->nb; ak++)
                for (int d = 0; d < 3; d++) {
                    int row = side * kb + ak * 3 + d;
                    for (int c = 0; c < k2; c++) {
                        double s = 0;
                        for (int i = 0; i < 6; i++) s += zm[(ak * 6 + i) * 6 + side * 3 + d] * K->T[(size_t)(ak * 6 + i) * k2 + c];
                        K->Wm[(size_t)row * k2 + c] = s;
                    }
                }
        for (int r = 0; r < kb; r++)
            for (int c = 0; c < kb; c++) {
                if (m > 0) K->Wd[(size_t)(m - 1) * kb * kb + r * kb + c] += K->Wm[(size_t)r * k2 + c];                  /* LL -> knot m */
                if (m < M - 1) K->Wd[(size_t)m * kb * kb + r * kb + c] += K->Wm[(size_t)(kb + r) * k2 + kb + c];         /* RR -> knot m+1 */
                if (m > 0 && m < M - 1) K->Wo[(size_t)(m - 1) * kb * kb + r * kb + c] = K->Wm[(size_t)(kb + r) * k2 + c]; /* RL */
            }
    }
    free(zt);
    /* block tridiagonal Cholesky: D_t <- chol(D_t - O_{t-1} O_{t-1}'), O_t <- O_t D_t^-T */
    for (int t = 0; t < M - 1; t++) {
        double *D = K->Wd + (size_t)t * kb * kb;
        if (t > 0) {
            const double *O = K->Wo + (size_t)(t - 1) * kb * kb;
            for (int r = 0; r < kb; r++)
                for (int c = 0; c <= r; c++) {
                    double s = 0;
                    for (int k = 0; k < kb; k++) s += O[r * kb + k] * O[c * kb + k];
                    D[r * kb + c] -= s;
                }
        }
        if (chol_dense(kb, D, kb)) return -1;
        if (t < M - 2) {
            double *O = K->Wo + (size_t)t * kb * kb;
            for (int r = 0; r < kb; r++)
                for (int c = 0; c < kb; c++) {
                    double v = O[r * kb + c];
                    for (int k = 0; k < c; k++) v -= O[r * kb + k] * D[c * kb + k];
                    O[r * kb + c] = v / D[c * kb + c];
                }
        }
    }
    return 0;
}
/* g (nr) <- (Z'HZ)^-1 g */
static void nsp_solve(const nsp_t *K, double *g) {
    int M = K->M, kb = K->kb;
    for (int t = 0; t < M - 1; t++) {
        const double *D = K->Wd + (size_t)t * kb * kb;
        double *gt = g + t * kb;
        if (t > 0) {
            const double *O = K->Wo + (size_t)(t - 1) * kb * kb;
            for (int r = 0; r < kb; r++) {
                double s = 0;
                for (int k = 0; k < kb; k++) s += O[r * kb + k] * g[(t - 1) * kb + k];
                gt[r] -= s;
            }
        }
        for (int r = 0; r < kb; r++) {
            double v = gt[r];
            for (int k = 0; k < r; k++) v -= D[r * kb + k] * gt[k];
            gt[r] = v / D[r * kb + r];
        }
    }
    for (int t = M - 2; t >= 0; t--) {
        const double *D = K->Wd + (size_t)t * kb * kb;
        double *gt = g + t * kb;
        if (t < M - 2) {
            const double *O = K->Wo + (size_t)t * kb * kb;
            for (int c = 0; c < kb; c++) {
                double s = 0;
                for (int k = 0; k < kb; k++) s += O[k * kb + c] * g[(t + 1) * kb + k];
                gt[c] -= s;
            }
        }
        for (int r = kb - 1; r >= 0; r--) {
            double v = gt[r];
            for (int k = r + 1; k < kb; k++) v -= D[k * kb + r] * gt[k];
            gt[r] = v / D[r * kb + r];
        }
    }
}

static double inf_norm(const double *v, int n) {
    double m = 0;
    for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > m) m = a; }
    return m;
}

/* Rows all of whose variables are fixed by the start / goal equalities (control points 0..2 of the first and
 * 3..5 of the last segment: their Z rows vanish) are constant: checked against their right-hand side with
 * CPLEX's default feasibility tolerance 1e-6 and dropped, as a presolve would.  A start or goal lying on a face
 * of its SFC box, which the corridor generator produces routinely, would otherwise leave no strict interior. */
#define PRESOLVE_FEAS_TOL 1e-6
static int presolve_dead_rows(const nsp_t *K, double *h, unsigned char *dead, int *n_live) {
    const oracle_qp *q = K->q;
    int rc = 0, live = 0;
    for (int r = 0; r < q->mi; r++) {
        int all = q->g_ptr[r + 1] > q->g_ptr[r];
        double gx = 0;
        for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1] && all; t++) {
            int c = q->g_idx[t];
            for (int d = 0; d < 6; d++) if (K->Zc[(size_t)c * 6 + d] != 0) { all = 0; break; }
            gx += q->g_val[t] * K->xp[c];
        }
        dead[r] = (unsigned char)all;
        if (all) { if (gx - h[r] > PRESOLVE_FEAS_TOL) rc = ORACLE_INFEASIBLE; }
    }
    /* Bound-based row redundancy (also a standard presolve reduction): the box rows are singleton rows, i.e. simple
     * bounds lb <= x <= ub on every variable.  A row with several columns whose maximal activity over those bounds,
     * sum_t max(g_t ub_t, g_t lb_t), stays below its right-hand side can never be active and is dropped: the feasible
     * set, hence the minimiser, is unchanged.  In a 64-agent mission 60-95 % of the RSFC rows go this way (far-apart
     * agents).  Margin 1e-9 relative so that borderline rows are kept on every platform. */
    {
        int nv = q->nv;
        double *ub = (double *)malloc(sizeof(double) * (nv + 1)), *lb = (double *)malloc(sizeof(double) * (nv + 1));
        for (int i = 0; i < nv; i++) { ub[i] = 1e300; lb[i] = -1e300; }
        for (int r = 0; r < q->n_box_rows && r < q->mi; r++)   /* the box rows of populatebyrow (RP L626-L635) */
            if (q->g_ptr[r + 1] - q->g_ptr[r] == 1) {
                int c = q->g_idx[q->g_ptr[r]];
                double g = q->g_val[q->g_ptr[r]], b = h[r] / g;
                if (g > 0) { if (b < ub[c]) ub[c] = b; } else if (g < 0) { if (b > lb[c]) lb[c] = b; }
            }
        /* Bounds that leave no interior at the feasibility tolerance (|ub - lb| < 2e-6: a corridor box of zero width,
         * e.g. an agent flying along the world boundary next to a pillar) would be a fixed variable to CPLEX's presolve,
         * and any point within 1e-6 of the bounds is feasible to it.  The interior-point method needs an interior: such
         * a pair of bounds is opened to [mid - 1e-6, mid + 1e-6]. */
        for (int r = 0; r < q->n_box_rows && r < q->mi; r++)
            if (!dead[r] && q->g_ptr[r + 1] - q->g_ptr[r] == 1) {
                int c = q->g_idx[q->g_ptr[r]];
                double g = q->g_val[q->g_ptr[r]], wdt = ub[c] - lb[c];
                if (ub[c] < 1e300 && lb[c] > -1e300 && wdt < 2 * PRESOLVE_FEAS_TOL && wdt > -2 * PRESOLVE_FEAS_TOL) {
                    double mid = 0.5 * (ub[c] + lb[c]);
                    if (g > 0 && h[r] / g == ub[c]) h[r] = g * (mid + PRESOLVE_FEAS_TOL);
                    if (g < 0 && h[r] / g == lb[c]) h[r] = g * (mid - PRESOLVE_FEAS_TOL);
                }
            }
        for (int i = 0; i < nv; i++) {
            double wdt = ub[i] - lb[i];
            if (ub[i] < 1e300 && lb[i] > -1e300 && wdt < 2 * PRESOLVE_FEAS_TOL && wdt > -2 * PRESOLVE_FEAS_TOL) {
                double mid = 0.5 * (ub[i] + lb[i]);
                ub[i] = mid + PRESOLVE_FEAS_TOL; lb[i] = mid - PRESOLVE_FEAS_TOL;
            }
        }
        /* The same rule for the two control points that C0 continuity identifies (control point 5 of segment m = control
         * point 0 of segment m + 1, RP L390-L399): consecutive corridor boxes that only share a face pin the knot to that
         * face -- a fixed variable to CPLEX's presolve.  The binding faces of the INTERSECTION of the two boxes are moved
         * 1e-6 apart (worlds/map32.bt of the reference's smoke loop produces such a pair). */
        {
            int *varof = (int *)malloc(sizeof(int) * nv);
            for (int v = 0; v < nv; v++) varof[K->seg[v] * K->n + K->loc[v]] = v;
            double *ub0 = (double *)malloc(sizeof(double) * nv), *lb0 = (double *)malloc(sizeof(double) * nv);
            memcpy(ub0, ub, sizeof(double) * nv); memcpy(lb0, lb, sizeof(double) * nv);
            for (int m = 0; m + 1 < K->M; m++)
                for (int ak = 0; ak < 3 * K->nb; ak++) {
                    int va = varof[m * K->n + ak * 6 + 5], vb = varof[(m + 1) * K->n + ak * 6];
                    if (ub0[va] >= 1e300 || ub0[vb] >= 1e300 || lb0[va] <= -1e300 || lb0[vb] <= -1e300) continue;
                    double lo = lb0[va] > lb0[vb] ? lb0[va] : lb0[vb], hi = ub0[va] < ub0[vb] ? ub0[va] : ub0[vb], wdt = hi - lo;
                    if (wdt < 2 * PRESOLVE_FEAS_TOL && wdt > -2 * PRESOLVE_FEAS_TOL) {
                        double mid = 0.5 * (hi + lo);
                        int vv[2] = {va, vb};
                        for (int t = 0; t < 2; t++) {
                            if (ub[vv[t]] < mid + PRESOLVE_FEAS_TOL) ub[vv[t]] = mid + PRESOLVE_FEAS_TOL;
                            if (lb[vv[t]] > mid - PRESOLVE_FEAS_TOL) lb[vv[t]] = mid - PRESOLVE_FEAS_TOL;
                        }
                    }
                }
            free(varof); free(ub0); free(lb0);
            for (int r = 0; r < q->n_box_rows && r < q->mi; r++)   /* loosened bounds back into the rows */
                if (!dead[r] && q->g_ptr[r + 1] - q->g_ptr[r] == 1) {
                    int c = q->g_idx[q->g_ptr[r]];
                    double g = q->g_val[q->g_ptr[r]];
                    if (g > 0 && ub[c] < 1e300 && h[r] / g < ub[c]) h[r] = g * ub[c];
                    if (g < 0 && lb[c] > -1e300 && h[r] / g > lb[c]) h[r] = g * lb[c];
                }
        }
        for (int r = q->n_box_rows; r < q->mi; r++) {          /* the RSFC rows (RP L636-L684) */
            if (dead[r] || q->g_ptr[r + 1] == q->g_ptr[r]) continue;
            double amax = 0;
            int bounded = 1;
            for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1]; t++) {
                int c = q->g_idx[t];
                double g = q->g_val[t];
                if (ub[c] >= 1e300 || lb[c] <= -1e300) { bounded = 0; break; }
                double a = g * ub[c], b = g * lb[c];
                amax += (a > b) ? a : b;
            }
            if (bounded && amax < h[r] - 1e-9 * fmax(1.0, fabs(h[r]))) dead[r] = 2;
        }
        free(ub); free(lb);
    }
    for (int r = 0; r < q->mi; r++) if (!dead[r]) live++;
    *n_live = live;
    return rc;
}

#define TOL_DUAL_FLOOR 1e-6 /* CPLEX EpOpt default; see the acceptance rule in oracle_solve_qp */
/* Farkas certificate z >= 0, h'z < 0: |(GZ)'z| / (-h'z) < ratio proves that no sigma with |sigma|_1 < 1/ratio satisfies the
 * rows (knot states -- metres, m/s, m/s^2 -- beyond 1e6 are outside any mission). */
#define CERT_RATIO 1e-6
#define CERT_RATIO_BREAKDOWN 1e-4
/* Mehrotra predictor-corrector on  min x'Qx  s.t. Ax=b, Gx+s=h, s>=0  (P = Q+Q'). */
int oracle_solve_qp(const oracle_qp *q, const oracle_solver_opts *opts, double *x, double *obj_out,
                    int *iters_out, double *res_out) {
    int nv = q->nv, ne = q->ne, mi = q->mi;
    int max_iter = (opts && opts->max_iter > 0) ? opts->max_iter : 100;
    double tol_gap = (opts && opts->tol_gap > 0) ? opts->tol_gap : 1e-10;
    double tol_res = (opts && opts->tol_res > 0) ? opts->tol_res : 1e-9;
    nsp_t K;
    int status = ORACLE_NOT_CONVERGED, it = 0, n_live = mi;
    double gap = 0, obj = 0, nrd = 0, nrg = 0, hn = 0;
    if (nsp_init(&K, q)) {
        nsp_free(&K);
        if (obj_out) *obj_out = 0;
        if (iters_out) *iters_out = 0;
        return ORACLE_BAD_ARG;
    }
    int nr = K.nr;
    double *s = (double *)malloc(sizeof(double) * (mi + 1)), *z = (double *)malloc(sizeof(double) * (mi + 1));
    double *rd = (double *)malloc(sizeof(double) * nv), *rg = (double *)malloc(sizeof(double) * (mi + 1));
    double *r1 = (double *)malloc(sizeof(double) * nv), *tt = (double *)malloc(sizeof(double) * (mi + 1));
    double *dx = (double *)malloc(sizeof(double) * nv), *dxa = (double *)malloc(sizeof(double) * nv);
    double *gx = (double *)malloc(sizeof(double) * (mi + 1)), *px = (double *)malloc(sizeof(double) * nv);
    double *gxa = (double *)malloc(sizeof(double) * (mi + 1)), *hs = (double *)malloc(sizeof(double) * (mi + 1));
    double *sg = (double *)malloc(sizeof(double) * (nr + 1)), *ax = (double *)malloc(sizeof(double) * (ne + 1));
    unsigned char *dead = (unsigned char *)calloc(mi + 1, 1);
    double *h = (double *)malloc(sizeof(double) * (mi + 1));   /* right-hand sides after presolve */
    memcpy(h, q->h, sizeof(double) * mi);
    if (presolve_dead_rows(&K, h, dead, &n_live)) { status = ORACLE_INFEASIBLE; goto done; }
    for (int r = 0; r < mi; r++) if (!dead[r] && fabs(h[r]) > hn) hn = fabs(h[r]);
    for (int i = 0; i < nv; i++) x[i] = K.xp[i];
    if (nr == 0) { /* a single segment: start and goal fix everything */
        p_mulv(q, x, px);
        for (int i = 0; i < nv; i++) obj += 0.5 * x[i] * px[i];
        status = ORACLE_OK;
        goto done;
    }
    /* reduced right-hand side h - G x_p (used by the infeasibility certificate) */
    csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, K.xp, hs);
    for (int r = 0; r < mi; r++) hs[r] = h[r] - hs[r];

    /* initial point (least-squares start, W = I): min 1/2 x'(P+G'G)x - (G'h)'x  s.t. Ax = b */
    for (int r = 0; r < mi; r++) K.w[r] = dead[r] ? 0.0 : 1.0;
    if (nsp_factor(&K, K.w)) goto done;
    p_mulv(q, x, px);
    for (int i = 0; i < nv; i++) r1[i] = -px[i];
    for (int r = 0; r < mi; r++) tt[r] = dead[r] ? 0.0 : hs[r];
    csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, tt, r1);
    nsp_Zt(&K, r1, sg);
    nsp_solve(&K, sg);
    nsp_Z(&K, sg, dx);
    for (int i = 0; i < nv; i++) x[i] += dx[i];
    csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, x, gx);
    {
        double ap = -1e300, ad = -1e300;
        for (int r = 0; r < mi; r++) { z[r] = gx[r] - h[r]; s[r] = -z[r]; }
        for (int r = 0; r < mi; r++) { if (dead[r]) continue; if (-s[r] > ap) ap = -s[r]; if (-z[r] > ad) ad = -z[r]; }
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { s[r] = 1.0; z[r] = 0.0; continue; }
            if (ap >= 0) s[r] += 1.0 + ap;
            if (ad >= 0) z[r] += 1.0 + ad;
        }
    }

    for (it = 0; it < max_iter; it++) {
        p_mulv(q, x, px);
        for (int i = 0; i < nv; i++) rd[i] = px[i];
        csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, z, rd);   /* rd = Px + G'z; its part in range(A') is the multiplier */
        csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, x, gx);
        double mu = 0, hz = 0, zmax = 0;
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { rg[r] = 0; continue; }
            rg[r] = gx[r] + s[r] - h[r]; mu += s[r] * z[r]; hz += hs[r] * z[r];
            if (z[r] > zmax) zmax = z[r];
        }
        mu /= (n_live > 0 ? n_live : 1);
        obj = 0;
        for (int i = 0; i < nv; i++) obj += 0.5 * x[i] * px[i];
        gap = mu;
        nsp_Zt(&K, rd, sg);
        nrd = inf_norm(sg, nr);
        nrg = inf_norm(rg, mi);
        double dscale = 1.0 + inf_norm(px, nv);
        if (!(mu == mu) || !(nrd == nrd)) { status = ORACLE_NOT_CONVERGED; break; }
        {
            int gap_ok = gap <= tol_gap * fmax(1.0, fabs(obj)) && nrg <= tol_res * (1 + hn);
            if (gap_ok && nrd <= tol_res * dscale) { status = ORACLE_OK; break; }
            /* Round-off floor of the dual residual.  On (nearly) degenerate QPs x converges like sqrt(mu) while the
             * weights z/s of the active rows grow like 1/mu, so the multiplier step dz = .. - w (G dx) carries noise
             * eps w |dx| that GROWS as mu falls: the strict dual test can become unreachable although complementarity
             * and primal feasibility are converged (seeds 3029, 3194, 3220 of the 64-agent workload: strictly feasible
             * by an independent LP, CPLEX solves them).  Once complementarity and primal feasibility meet their strict
             * tolerances the iterate is therefore accepted at CPLEX's own optimality tolerance (EpOpt, default 1e-6,
             * relative to the gradient scale) instead of iterating into a numerically singular factorisation. */
            if (gap_ok && nrd <= TOL_DUAL_FLOOR * dscale) { status = ORACLE_OK; break; }
        }
        /* infeasibility certificate of the reduced problem {G Z sigma <= h - G x_p}: z >= 0, (GZ)'z ~ 0, (h - G x_p)'z < 0.
         * By LP duality the largest uniform slack of the rows is min (h - G x_p)'z / sum(z) over such z, so the row set is
         * infeasible beyond the feasibility tolerance only if (h - G x_p)'z < -1e-6 sum(z); sum(z) >= max(z) is used.  A
         * QP whose rows have no interior but are consistent (slack exactly 0) must not be certified infeasible. */
        double cert = 1e300;
        if (hz < -PRESOLVE_FEAS_TOL * zmax) {
            for (int i = 0; i < nv; i++) r1[i] = 0;
            csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, z, r1);
            nsp_Zt(&K, r1, sg);
            cert = inf_norm(sg, nr) / (-hz);
            if (cert < CERT_RATIO) { status = ORACLE_INFEASIBLE; break; }
        }
        for (int r = 0; r < mi; r++) K.w[r] = dead[r] ? 0.0 : z[r] / s[r];
        if (nsp_factor(&K, K.w)) {
            /* the weights of an infeasible QP diverge by orders of magnitude per iteration; if the factorisation gives up
             * before the certificate is sharp, a certificate that already excludes every |sigma| < 1e4 decides */
            status = cert < CERT_RATIO_BREAKDOWN ? ORACLE_INFEASIBLE : ORACLE_NOT_CONVERGED;
            break;
        }
        /* affine direction: rc = s.z  =>  G' coefficient -(w rg - z) */
        for (int r = 0; r < mi; r++) tt[r] = dead[r] ? 0.0 : -(K.w[r] * rg[r] - z[r]);
        for (int i = 0; i < nv; i++) r1[i] = -rd[i];
        csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, tt, r1);
        nsp_Zt(&K, r1, sg);
        nsp_solve(&K, sg);
        nsp_Z(&K, sg, dxa);
        csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, dxa, gxa);
        double aa = 1.0;
        for (int r = 0; r < mi; r++) {
            if (dead[r]) continue;
            double dsa = -rg[r] - gxa[r], dza = -z[r] - K.w[r] * dsa;
            if (dsa < 0) { double a = -s[r] / dsa; if (a < aa) aa = a; }
            if (dza < 0) { double a = -z[r] / dza; if (a < aa) aa = a; }
        }
        double mua = 0;
        for (int r = 0; r < mi; r++) {
            if (dead[r]) continue;
            double dsa = -rg[r] - gxa[r], dza = -z[r] - K.w[r] * dsa;
            mua += (s[r] + aa * dsa) * (z[r] + aa * dza);
        }
        mua /= (n_live > 0 ? n_live : 1);
        double sigma = (mu > 0) ? (mua / mu) * (mua / mu) * (mua / mu) : 0;
        /* corrector: rc = s.z + dsa.dza - sigma mu.  (The CUDA kernels form G'tt from two sums of their affine pass,
           G'(-(z rg - s z - dsa dza)/s) - sigma mu G'(1/s), instead of a pass of their own: same iterates up to rounding,
           tests/test_emu_kernels.py::test_fused_corrector_equals_the_separate_corrector_pass.) */
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { tt[r] = 0; continue; }
            double dsa = -rg[r] - gxa[r], dza = -z[r] - K.w[r] * dsa;
            double rc = s[r] * z[r] + dsa * dza - sigma * mu;
            tt[r] = -(z[r] * rg[r] - rc) / s[r];
        }
        for (int i = 0; i < nv; i++) r1[i] = -rd[i];
        csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, tt, r1);
        nsp_Zt(&K, r1, sg);
        nsp_solve(&K, sg);
        nsp_Z(&K, sg, dx);
        csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, dx, gx);
        double am = 1e300;
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { tt[r] = 0; rg[r] = 0; continue; }
            double dsa = -rg[r] - gxa[r], dza = -z[r] - K.w[r] * dsa;
            double rc = s[r] * z[r] + dsa * dza - sigma * mu;
            double ds = -rg[r] - gx[r], dz = (-rc - z[r] * ds) / s[r];
            tt[r] = ds; rg[r] = dz; /* reuse as ds, dz */
            if (ds < 0) { double a = -s[r] / ds; if (a < am) am = a; }
            if (dz < 0) { double a = -z[r] / dz; if (a < am) am = a; }
        }
        double al = fmin(1.0, 0.99 * am);
        for (int i = 0; i < nv; i++) x[i] += al * dx[i];
        for (int r = 0; r < mi; r++) { if (dead[r]) continue; s[r] += al * tt[r]; z[r] += al * rg[r]; }
    }
done:
    if (res_out) {
        double nrp = 0;
        if (ne > 0 && status != ORACLE_BAD_ARG) {
            csr_mulv(ne, q->a_ptr, q->a_idx, q->a_val, x, ax);
            for (int r = 0; r < ne; r++) { double a = fabs(ax[r] - q->b[r]); if (a > nrp) nrp = a; }
        }
        res_out[0] = gap; res_out[1] = nrp; res_out[2] = nrd; res_out[3] = nrg;
    }
    if (obj_out) *obj_out = obj;
    if (iters_out) *iters_out = it;
    nsp_free(&K);
    free(s); free(z); free(rd); free(rg); free(r1); free(tt); free(dx); free(dxa); free(gx); free(px);
    free(gxa); free(hs); free(sg); free(ax); free(dead); free(h);
    return status;
}

/* ------------------------------------------------------------------------------------------------
 * update(): solveQP loop (RP L111-L206) with conversion (L167-L196) and timeMatrix (L695-L700)
 * ---------------------------------------------------------------------------------------------- */
static void ctrl_to_coef(const double *basis, double dt, const double *c6 /*stride 3*/, int stride, double *out6) {
    /* c = sum_i vals_i * (basis * diag((1/dt)^(5-j))).row(i) */
    for (int j = 0; j < 6; j++) {
        double s = 0, tm = pow(1.0 / dt, 5 - j);
        for (int i = 0; i < 6; i++) s += c6[i * stride] * (basis[i * 6 + j] * tm);
        out6[j] = s;
    }
}

int oracle_update(const oracle_problem *p, const oracle_solver_opts *opts, double *coef, double *ctrl,
                  double *batch_obj, int *batch_iters, int *batch_status, int nthreads_unused) {
    (void)nthreads_unused;
    int N = p->N, M = p->M, oq = M * NCP, bs, bit;
    if (N <= 0 || M <= 0 || p->batch_size <= 0) return ORACLE_BAD_ARG;
    int nbatches = oracle_set_batch(p, &bs, &bit);
    (void)nbatches;
    double Qb[36], basis[36];
    oracle_build_Q_base(Qb, basis);
    double *dummy = (double *)calloc((size_t)N * oq * 3, sizeof(double));
    if (p->sequential) oracle_build_dummy(p, dummy);
    memset(coef, 0, sizeof(double) * (size_t)N * 3 * oq);
    int rc = ORACLE_OK;

    if (p->sequential && bit == 0) { /* RP L119-L138: publish the initial trajectory */
        for (int k = 0; k < 3; k++)
            for (int qi = 0; qi < N; qi++)
                for (int m = 0; m < M; m++)
                    ctrl_to_coef(basis, p->T[m + 1] - p->T[m], dummy + ((size_t)qi * oq + m * NCP) * 3 + k, 3,
                                 coef + ((size_t)qi * 3 + k) * oq + m * NCP);
        goto finish;
    }
    {
        int batch_max_iter = (int)ceil((double)N / (double)bs);
        int rec = 0;
        for (int iter = 0; iter < p->iteration; iter++)
            for (int l = 0; l < bit; l++) {
                oracle_qp *q = oracle_populate(p, dummy, l);
                if (!q) { rc = ORACLE_BAD_ARG; goto finish; }
                double *vals = (double *)malloc(sizeof(double) * q->nv), obj = 0;
                int its = 0;
                int st = oracle_solve_qp(q, opts, vals, &obj, &its, NULL);
                if (batch_obj) batch_obj[rec] = obj;
                if (batch_iters) batch_iters[rec] = its;
                if (batch_status) batch_status[rec] = st;
                rec++;
                if (st != ORACLE_OK) { /* RP L158-L161: throw(-1) -> update returns false */
                    free(vals); oracle_qp_free(q); rc = st; goto finish;
                }
                int q0 = l * bs, nb = (q0 + bs <= N) ? bs : N - q0, od = nb * oq;
                for (int k = 0; k < 3; k++)
                    for (int qi = 0; qi < N; qi++)
                        for (int m = 0; m < M; m++) {
                            int bi = quad_in_batch(qi, l, bs, N);
                            double dt = p->T[m + 1] - p->T[m];
                            double *out = coef + ((size_t)qi * 3 + k) * oq + m * NCP;
                            if (bi >= 0) {
                                ctrl_to_coef(basis, dt, vals + k * od + bi * oq + m * NCP, 1, out);
                                if (p->sequential)
                                    for (int i = 0; i < NCP; i++)
                                        dummy[((size_t)qi * oq + m * NCP + i) * 3 + k] = vals[k * od + bi * oq + m * NCP + i];
                                else /* not in the reference (dummy unused there); kept so ctrl[] is defined */
                                    for (int i = 0; i < NCP; i++)
                                        dummy[((size_t)qi * oq + m * NCP + i) * 3 + k] = vals[k * od + bi * oq + m * NCP + i];
                            } else if (p->sequential && bit < batch_max_iter) {
                                ctrl_to_coef(basis, dt, dummy + ((size_t)qi * oq + m * NCP) * 3 + k, 3, out);
                            }
                        }
                free(vals);
                oracle_qp_free(q);
            }
    }
finish:
    if (ctrl)
        for (int qi = 0; qi < N; qi++)
            for (int k = 0; k < 3; k++)
                for (int j = 0; j < oq; j++) ctrl[((size_t)qi * 3 + k) * oq + j] = dummy[((size_t)qi * oq + j) * 3 + k];
    free(dummy);
    return rc;
}

int oracle_update_many(const oracle_problem *ps, int count, const oracle_solver_opts *opts, double **coef,
                       double **ctrl, int *status, int nthreads) {
    int bad = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : bad)
#endif
    for (int i = 0; i < count; i++) {
        int st = oracle_update(&ps[i], opts, coef[i], ctrl ? ctrl[i] : NULL, NULL, NULL, NULL, 1);
        if (status) status[i] = st;
        if (st != ORACLE_OK) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------------------------------------
 * Corridor::updateRelBox (rbp_corridor.hpp L338-L398): relative safe flight corridor normals.
 * All arithmetic is octomap::point3d = octomath::Vector3 arithmetic, i.e. FLOAT32 with these semantics
 * (octomap is not vendored under /root/reference; restated from its public header math/Vector3.h):
 *   operator-, operator*(float), operator/=(float): component-wise float;
 *   dot(), norm_sq(): float products and sums, returned as double; norm() = sqrt of that double;
 *   normalize(): len = norm(); if (len > 0) *this /= (float)len.
 * For pair (qi<qj) and segment iter=1..M: a, b = relative positions at iter-1, iter with z divided by the
 * downwash coefficient; m = closest point of segment a-b to the origin (candidates a, b, and the foot of the
 * perpendicular if it falls inside); normalise; divide z by downwash again.  Returns 0, or 1 if some normal
 * is zero ("initial trajectories are collided", L385-L388).
 * ---------------------------------------------------------------------------------------------- */
static double v3_norm(const float *v) { float s = v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; return sqrt((double)s); }
static double v3_dot(const float *a, const float *b) { float s = a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; return (double)s; }
static void v3_normalize(float *v) {
    double len = v3_norm(v);
    if (len > 0) { float f = (float)len; v[0] /= f; v[1] /= f; v[2] /= f; }
}
int oracle_rsfc(int N, int M, const float *init_traj /*[N][M+1][3]*/, const double *T, double downwash,
                float *rsfc_n /*[P][M][3]*/, double *rsfc_t /*[P][M]*/) {
    int rc = 0;
    long it = 0;
    for (int qi = 0; qi < N; qi++)
        for (int qj = qi + 1; qj < N; qj++, it++)
            for (int iter = 1; iter <= M; iter++) {
                const float *pi0 = init_traj + ((size_t)qi * (M + 1) + iter - 1) * 3, *pj0 = init_traj + ((size_t)qj * (M + 1) + iter - 1) * 3;
                const float *pi1 = pi0 + 3, *pj1 = pj0 + 3;
                float a[3], b[3], m[3], n[3], c[3], ca[3], cb[3];
                for (int k = 0; k < 3; k++) { a[k] = pj0[k] - pi0[k]; b[k] = pj1[k] - pi1[k]; }
                a[2] = (float)((double)a[2] / downwash);      /* a.z() = a.z() / param.downwash (float / double -> float) */
                b[2] = (float)((double)b[2] / downwash);
                if (a[0] == b[0] && a[1] == b[1] && a[2] == b[2]) {
                    for (int k = 0; k < 3; k++) m[k] = a[k];
                } else {
                    for (int k = 0; k < 3; k++) m[k] = a[k];
                    double dist_min = v3_norm(a), dist = v3_norm(b);
                    if (dist_min > dist) { for (int k = 0; k < 3; k++) m[k] = b[k]; dist_min = dist; }
                    for (int k = 0; k < 3; k++) n[k] = b[k] - a[k];
                    v3_normalize(n);
                    float f = (float)v3_dot(a, n);            /* n * a.dot(n): the double narrows to operator*(float) */
                    for (int k = 0; k < 3; k++) c[k] = a[k] - n[k] * f;
                    dist = v3_norm(c);
                    for (int k = 0; k < 3; k++) { ca[k] = c[k] - a[k]; cb[k] = c[k] - b[k]; }
                    if (v3_dot(ca, cb) < 0 && dist_min > dist) for (int k = 0; k < 3; k++) m[k] = c[k];
                }
                v3_normalize(m);
                m[2] = (float)((double)m[2] / downwash);
                if (v3_norm(m) == 0) rc = 1;
                float *o = rsfc_n + ((size_t)it * M + iter - 1) * 3;
                o[0] = m[0]; o[1] = m[1]; o[2] = m[2];
                rsfc_t[(size_t)it * M + iter - 1] = T[iter];
            }
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * RBPPublisher post-hoc checks (rbp_publisher.hpp): sampling t_i = i*dt, i < floor(T_M/dt) (L47-L51); segment
 * of a sample = last m with T[m] < t (strict), local time t - T[m] (timeMatrix L169-L183); position = sum_j
 * coef_j tseg^j (update_quad_state L670-L683); safety_margin_ratio = min over samples and pairs of
 * sqrt(dx^2 + dy^2 + (dz/downwash)^2) / (r_i + r_j) (L769-L798) -- collision-free iff >= 1; flight length =
 * sum over agents of the polyline through the samples (L685-L695).
 * Powers of tseg are formed by repeated multiplication and sums run lowest power first (the reference calls pow()
 * and Eigen's product; bit-level agreement with that binary is unattainable, the boolean outcome is what matters).
 * coef: [N][3][6M], highest power first per segment (the msgs_traj_coef layout).
 * ---------------------------------------------------------------------------------------------- */
static void eval_pos(const double *coef, int M, const double *T, double t, double *p /*[3]*/) {
    int index = 0;
    double tseg = 0;
    for (int m = 0; m < M; m++) {
        if (T[m] < t) { tseg = T[m]; index = m; } else break;
    }
    tseg = t - tseg;
    for (int k = 0; k < 3; k++) {
        const double *c = coef + (size_t)k * 6 * M + index * 6;
        double s = 0, pw = 1;
        for (int j = 0; j < 6; j++) { s = s + c[5 - j] * pw; pw = pw * tseg; }
        p[k] = s;
    }
}
int oracle_safety_metrics(int N, int M, const double *coef, const double *T, const double *radius, double downwash,
                          double dt, double *min_ratio, double *t_at_min, double *length) {
    int nt = (int)floor(T[M] / dt);
    double best = 1e9, tbest = 0, len = 0;
    double *pos = (double *)malloc(sizeof(double) * (size_t)N * 3 * (nt > 0 ? nt : 1));
    for (int i = 0; i < nt; i++)
        for (int q = 0; q < N; q++) eval_pos(coef + (size_t)q * 18 * M, M, T, i * dt, pos + ((size_t)i * N + q) * 3);
    for (int i = 0; i < nt; i++)
        for (int qi = 0; qi < N; qi++)
            for (int qj = qi + 1; qj < N; qj++) {
                const double *a = pos + ((size_t)i * N + qi) * 3, *b = pos + ((size_t)i * N + qj) * 3;
                double dx = a[0] - b[0], dy = a[1] - b[1], dz = (a[2] - b[2]) / downwash;
                double ratio = sqrt(dx * dx + dy * dy + dz * dz) / (radius[qi] + radius[qj]);
                if (ratio < best) { best = ratio; tbest = i * dt; }
            }
    for (int q = 0; q < N; q++)
        for (int i = 0; i + 1 < nt; i++) {
            const double *a = pos + ((size_t)i * N + q) * 3, *b = pos + ((size_t)(i + 1) * N + q) * 3;
            double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
            len += sqrt(dx * dx + dy * dy + dz * dz);
        }
    free(pos);
    *min_ratio = best; *t_at_min = tbest; *length = len;
    return nt;
}
