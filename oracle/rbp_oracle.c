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
 * Envelope (skyline) Cholesky: row r stores columns first[r]..r contiguously.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int n; int *first; size_t *off; double *v; } env_t;

static void env_alloc(env_t *e, int n, const int *first) {
    e->n = n;
    e->first = (int *)malloc(sizeof(int) * n);
    e->off = (size_t *)malloc(sizeof(size_t) * (n + 1));
    size_t t = 0;
    for (int r = 0; r < n; r++) { e->first[r] = first[r]; e->off[r] = t; t += (size_t)(r - first[r] + 1); }
    e->off[n] = t;
    e->v = (double *)malloc(sizeof(double) * (t ? t : 1));
}
static void env_free(env_t *e) { free(e->first); free(e->off); free(e->v); }
#define ENV(e, r, c) ((e)->v[(e)->off[r] + (size_t)((c) - (e)->first[r])])

/* Pivots that elimination has driven below PIVOT_REL of their original diagonal belong to rows that are
 * numerically dependent on earlier ones (an equality among variables that active inequality rows already pin:
 * LICQ fails there and the multiplier is not unique).  They are replaced by a huge value, which zeroes that
 * component of the solve (the "Cholesky-infinity" device of barrier codes); the primal solution is unaffected. */
#define PIVOT_REL 1e-13
#define PIVOT_BIG 1e128
static int env_chol(env_t *e, int safeguard) {
    for (int r = 0; r < e->n; r++) {
        int fr = e->first[r];
        double orig = ENV(e, r, r);
        for (int c = fr; c <= r; c++) {
            int fc = e->first[c], t0 = fr > fc ? fr : fc;
            double s = ENV(e, r, c);
            const double *lr = &ENV(e, r, t0), *lc = &ENV(e, c, t0);
            for (int t = 0; t < c - t0; t++) s -= lr[t] * lc[t];
            if (c < r) ENV(e, r, c) = s / ENV(e, c, c);
            else {
                if (!(s == s)) return -1;
                if (safeguard) { if (!(s > PIVOT_REL * orig) || !(s > 0)) s = PIVOT_BIG; }
                else if (!(s > 0)) return -1;
                ENV(e, r, r) = sqrt(s);
            }
        }
    }
    return 0;
}
/* forward solve in place; v nonzero only in [*lo, *hi]; tracks the range (sparse right-hand sides) */
static void env_fwd(const env_t *e, double *v, int *lo, int *hi) {
    int l0 = *lo, last = *hi;
    for (int r = l0; r < e->n; r++) {
        int fr = e->first[r];
        if (r > *hi && fr > last) { v[r] = 0; continue; }
        int t0 = fr > l0 ? fr : l0;
        double s = v[r];
        for (int t = t0; t < r; t++) s -= ENV(e, r, t) * v[t];
        v[r] = s / ENV(e, r, r);
        if (v[r] != 0) last = r;
    }
    *hi = last;
}
static void env_bwd(const env_t *e, double *v) {
    for (int r = e->n - 1; r >= 0; r--) {
        double x = v[r] / ENV(e, r, r);
        v[r] = x;
        if (x != 0)
            for (int c = e->first[r]; c < r; c++) v[c] -= ENV(e, r, c) * x;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Mehrotra predictor-corrector on  min 1/2 x'Px  s.t. Ax=b, Gx+s=h, s>=0   with P = Q+Q'.
 * Normal equations: H = P + G'WG (block diagonal over segments in perm_x order), S = A H^-1 A'
 * (block tridiagonal over knots in perm_y order); both factored with the envelope Cholesky.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const oracle_qp *q;
    int nv, ne, mi;
    env_t H, S;
    int *px, *py;            /* perms */
    double *Y; int *ylo, *yhi; /* Y = L^-1 A' (column c = eq row in perm order), dense nv per column */
    double *w;               /* z/s */
    double *t1, *t2, *t3;    /* scratch nv, nv, ne */
} kkt_t;

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

static int kkt_init(kkt_t *K, const oracle_qp *q) {
    memset(K, 0, sizeof(*K));
    K->q = q; K->nv = q->nv; K->ne = q->ne; K->mi = q->mi;
    int nv = q->nv, ne = q->ne;
    K->px = q->perm_x; K->py = q->perm_y;
    int *first = (int *)malloc(sizeof(int) * nv);
    for (int i = 0; i < nv; i++) first[i] = i;
    for (int t = 0; t < q->qnnz; t++) {
        int a = K->px[q->qi[t]], b = K->px[q->qj[t]];
        if (a < b) { int s = a; a = b; b = s; }
        if (b < first[a]) first[a] = b;
    }
    for (int r = 0; r < q->mi; r++) {
        int mn = nv;
        for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1]; t++) { int a = K->px[q->g_idx[t]]; if (a < mn) mn = a; }
        for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1]; t++) { int a = K->px[q->g_idx[t]]; if (mn < first[a]) first[a] = mn; }
    }
    /* an envelope must be monotone enough for fill: first[r] <= min over c in [first[r], r] handled by
     * the algorithm itself (entries outside a row's envelope are structurally zero and stay zero). */
    env_alloc(&K->H, nv, first);
    free(first);
    K->Y = (double *)malloc(sizeof(double) * (size_t)nv * (ne ? ne : 1));
    K->ylo = (int *)malloc(sizeof(int) * (ne + 1));
    K->yhi = (int *)malloc(sizeof(int) * (ne + 1));
    K->w = (double *)malloc(sizeof(double) * (q->mi + 1));
    K->t1 = (double *)malloc(sizeof(double) * nv);
    K->t2 = (double *)malloc(sizeof(double) * nv);
    K->t3 = (double *)malloc(sizeof(double) * (ne + 1));
    K->S.n = 0;
    return 0;
}
static void kkt_free(kkt_t *K) {
    env_free(&K->H);
    if (K->S.n) env_free(&K->S);
    free(K->Y); free(K->ylo); free(K->yhi); free(K->w); free(K->t1); free(K->t2); free(K->t3);
}

/* factor with weights w (mi) */
static int kkt_factor(kkt_t *K, const double *w) {
    const oracle_qp *q = K->q;
    int nv = K->nv, ne = K->ne;
    env_t *H = &K->H;
    memset(H->v, 0, sizeof(double) * H->off[nv]);
    for (int t = 0; t < q->qnnz; t++) { /* P = Q + Q' */
        int a = K->px[q->qi[t]], b = K->px[q->qj[t]];
        if (a >= b) ENV(H, a, b) += q->qv[t]; else ENV(H, b, a) += q->qv[t];
        if (a == b) ENV(H, a, a) += q->qv[t];
    }
    /* note: for a != b the pair (i,j),(j,i) both appear in the triplets, each adds Q_ij to the lower entry:
     * total 2 Q_ij = P_ij. For a == b the single triplet adds Q_ii twice = P_ii. */
    for (int r = 0; r < q->mi; r++) {
        double wr = w[r];
        for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1]; t++) {
            int a = K->px[q->g_idx[t]];
            double va = wr * q->g_val[t];
            for (int u = q->g_ptr[r]; u < q->g_ptr[r + 1]; u++) {
                int b = K->px[q->g_idx[u]];
                if (b <= a) ENV(H, a, b) += va * q->g_val[u];
            }
        }
    }
    if (env_chol(H, 0)) return -1;
    if (ne == 0) return 0;
    /* Y columns */
    for (int c = 0; c < ne; c++) { K->ylo[c] = nv; K->yhi[c] = -1; }
    memset(K->Y, 0, sizeof(double) * (size_t)nv * ne);
    for (int r = 0; r < ne; r++) {
        int c = K->py[r];
        double *y = K->Y + (size_t)c * nv;
        for (int t = q->a_ptr[r]; t < q->a_ptr[r + 1]; t++) {
            int a = K->px[q->a_idx[t]];
            y[a] = q->a_val[t];
            if (a < K->ylo[c]) K->ylo[c] = a;
            if (a > K->yhi[c]) K->yhi[c] = a;
        }
    }
    for (int c = 0; c < ne; c++)
        if (K->yhi[c] >= 0) env_fwd(H, K->Y + (size_t)c * nv, &K->ylo[c], &K->yhi[c]);
    /* S = Y'Y with envelope from range intersections */
    int *first = (int *)malloc(sizeof(int) * ne);
    for (int r = 0; r < ne; r++) {
        first[r] = r;
        for (int c = 0; c < r; c++) {
            int lo = K->ylo[r] > K->ylo[c] ? K->ylo[r] : K->ylo[c];
            int hi = K->yhi[r] < K->yhi[c] ? K->yhi[r] : K->yhi[c];
            if (lo <= hi) { first[r] = c; break; }
        }
    }
    if (K->S.n) env_free(&K->S);
    env_alloc(&K->S, ne, first);
    free(first);
    for (int r = 0; r < ne; r++)
        for (int c = K->S.first[r]; c <= r; c++) {
            int lo = K->ylo[r] > K->ylo[c] ? K->ylo[r] : K->ylo[c];
            int hi = K->yhi[r] < K->yhi[c] ? K->yhi[r] : K->yhi[c];
            double s = 0;
            const double *yr = K->Y + (size_t)r * nv, *yc = K->Y + (size_t)c * nv;
            for (int t = lo; t <= hi; t++) s += yr[t] * yc[t];
            ENV(&K->S, r, c) = s;
        }
    if (env_chol(&K->S, 1)) return -2;
    return 0;
}

/* solve  H dx + A' dy = r1 ; A dx = r2   (r1 in original variable order, r2 in original row order) */
static void kkt_solve(kkt_t *K, const double *r1, const double *r2, double *dx, double *dy) {
    int nv = K->nv, ne = K->ne;
    double *u = K->t1, *g = K->t3;
    for (int i = 0; i < nv; i++) u[K->px[i]] = r1[i];
    int lo = 0, hi = nv - 1;
    env_fwd(&K->H, u, &lo, &hi);
    if (ne) {
        for (int r = 0; r < ne; r++) {
            int c = K->py[r];
            const double *y = K->Y + (size_t)c * nv;
            double s = 0;
            for (int t = K->ylo[c]; t <= K->yhi[c]; t++) s += y[t] * u[t];
            g[c] = s - r2[r];
        }
        lo = 0; hi = ne - 1;
        env_fwd(&K->S, g, &lo, &hi);
        env_bwd(&K->S, g);
        for (int r = 0; r < ne; r++) dy[r] = g[K->py[r]];
        for (int c = 0; c < ne; c++) {
            const double *y = K->Y + (size_t)c * nv;
            double gc = g[c];
            for (int t = K->ylo[c]; t <= K->yhi[c]; t++) u[t] -= y[t] * gc;
        }
    }
    env_bwd(&K->H, u);
    for (int i = 0; i < nv; i++) dx[i] = u[K->px[i]];
}

static double inf_norm(const double *v, int n) {
    double m = 0;
    for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > m) m = a; }
    return m;
}


/* ------------------------------------------------------------------------------------------------
 * Presolve (what CPLEX's presolve does to this model before the barrier sees it): singleton equality
 * rows fix their column -- here the start rows fix control points 0..2 of the first segment and the
 * goal rows fix 3..5 of the last one (rows 0-5 of Aeq_base are triangular in them, RP L380-L387) --
 * and inequality rows all of whose columns are fixed are checked against their right-hand side
 * (feasibility tolerance 1e-6, CPLEX's default) and dropped.  Without this a start or goal lying on a
 * face of its SFC box, which the corridor generator produces routinely, leaves the QP without a strict
 * interior.  dead[r] = 1 for dropped rows.  Returns 0, or ORACLE_INFEASIBLE.
 * ---------------------------------------------------------------------------------------------- */
#define PRESOLVE_FEAS_TOL 1e-6
static int presolve_dead_rows(const oracle_qp *q, unsigned char *dead, int *n_live) {
    int nv = q->nv, ne = q->ne, mi = q->mi, rc = 0;
    unsigned char *fixed = (unsigned char *)calloc(nv + 1, 1), *used = (unsigned char *)calloc(ne + 1, 1);
    double *xf = (double *)calloc(nv + 1, sizeof(double));
    int progress = 1;
    while (progress) {
        progress = 0;
        for (int r = 0; r < ne; r++) {
            if (used[r]) continue;
            int nfree = 0, col = -1;
            double rhs = q->b[r], a = 0;
            for (int t = q->a_ptr[r]; t < q->a_ptr[r + 1]; t++) {
                int c = q->a_idx[t];
                if (fixed[c]) rhs -= q->a_val[t] * xf[c];
                else { nfree++; col = c; a = q->a_val[t]; }
            }
            if (nfree == 1 && a != 0) {
                fixed[col] = 1; xf[col] = rhs / a; used[r] = 1; progress = 1;
            }
        }
    }
    int live = 0;
    for (int r = 0; r < mi; r++) {
        int all = q->g_ptr[r + 1] > q->g_ptr[r];
        double gx = 0;
        for (int t = q->g_ptr[r]; t < q->g_ptr[r + 1]; t++) {
            int c = q->g_idx[t];
            if (!fixed[c]) { all = 0; break; }
            gx += q->g_val[t] * xf[c];
        }
        dead[r] = (unsigned char)all;
        if (all) { if (gx - q->h[r] > PRESOLVE_FEAS_TOL) rc = ORACLE_INFEASIBLE; }
        else live++;
    }
    *n_live = live;
    free(fixed); free(used); free(xf);
    return rc;
}

int oracle_solve_qp(const oracle_qp *q, const oracle_solver_opts *opts, double *x, double *obj_out,
                    int *iters_out, double *res_out) {
    int nv = q->nv, ne = q->ne, mi = q->mi;
    int max_iter = (opts && opts->max_iter > 0) ? opts->max_iter : 100;
    double tol_gap = (opts && opts->tol_gap > 0) ? opts->tol_gap : 1e-10;
    double tol_res = (opts && opts->tol_res > 0) ? opts->tol_res : 1e-9;
    kkt_t K;
    kkt_init(&K, q);
    double *y = (double *)calloc(ne + 1, sizeof(double)), *s = (double *)malloc(sizeof(double) * (mi + 1));
    double *z = (double *)malloc(sizeof(double) * (mi + 1));
    double *rd = (double *)malloc(sizeof(double) * nv), *rp = (double *)malloc(sizeof(double) * (ne + 1));
    double *rg = (double *)malloc(sizeof(double) * (mi + 1)), *r1 = (double *)malloc(sizeof(double) * nv);
    double *r2 = (double *)malloc(sizeof(double) * (ne + 1)), *tt = (double *)malloc(sizeof(double) * (mi + 1));
    double *dx = (double *)malloc(sizeof(double) * nv), *dy = (double *)malloc(sizeof(double) * (ne + 1));
    double *ds = (double *)malloc(sizeof(double) * (mi + 1)), *dz = (double *)malloc(sizeof(double) * (mi + 1));
    double *dsa = (double *)malloc(sizeof(double) * (mi + 1)), *dza = (double *)malloc(sizeof(double) * (mi + 1));
    double *gx = (double *)malloc(sizeof(double) * (mi + 1)), *px = (double *)malloc(sizeof(double) * nv);
    int status = ORACLE_NOT_CONVERGED, it = 0;
    unsigned char *dead = (unsigned char *)calloc(mi + 1, 1);
    int n_live = mi;
    double bn = inf_norm(q->b, ne), hn = 0, gap = 0, obj = 0;
    if (presolve_dead_rows(q, dead, &n_live)) { status = ORACLE_INFEASIBLE; goto done; }
    for (int r = 0; r < mi; r++) if (!dead[r] && fabs(q->h[r]) > hn) hn = fabs(q->h[r]);

    /* initial point (standard least-squares start): W = I */
    for (int r = 0; r < mi; r++) K.w[r] = 1.0;
    if (kkt_factor(&K, K.w)) { status = ORACLE_NOT_CONVERGED; goto done; }
    for (int i = 0; i < nv; i++) r1[i] = 0;
    for (int r = 0; r < mi; r++) tt[r] = dead[r] ? 0.0 : q->h[r];
    csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, tt, r1);
    kkt_solve(&K, r1, q->b, x, y);
    csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, x, gx);
    {
        double ap = -1e300, ad = -1e300;
        for (int r = 0; r < mi; r++) { z[r] = gx[r] - q->h[r]; s[r] = -z[r]; }
        for (int r = 0; r < mi; r++) { if (dead[r]) continue; if (-s[r] > ap) ap = -s[r]; if (-z[r] > ad) ad = -z[r]; }
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { s[r] = 1.0; z[r] = 0.0; continue; }  /* dropped row: no multiplier, unit weight in H */
            if (ap >= 0) s[r] += 1.0 + ap;
            if (ad >= 0) z[r] += 1.0 + ad;
        }
    }

    for (it = 0; it < max_iter; it++) {
        /* residuals */
        p_mulv(q, x, px);
        for (int i = 0; i < nv; i++) rd[i] = px[i];
        csr_mulTv_add(ne, q->a_ptr, q->a_idx, q->a_val, y, rd);
        csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, z, rd);
        csr_mulv(ne, q->a_ptr, q->a_idx, q->a_val, x, rp);
        for (int r = 0; r < ne; r++) rp[r] -= q->b[r];
        csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, x, gx);
        double mu = 0;
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { rg[r] = 0; continue; }
            rg[r] = gx[r] + s[r] - q->h[r]; mu += s[r] * z[r];
        }
        mu /= (n_live > 0 ? n_live : 1);
        obj = 0;
        for (int i = 0; i < nv; i++) obj += 0.5 * x[i] * px[i];
        gap = mu;
        double nrp = inf_norm(rp, ne), nrg = inf_norm(rg, mi), nrd = inf_norm(rd, nv);
        double dscale = 1.0 + inf_norm(px, nv);
        if (res_out) { res_out[0] = gap; res_out[1] = nrp; res_out[2] = nrd; res_out[3] = nrg; }
        if (!(mu == mu) || !(nrd == nrd)) { status = ORACLE_NOT_CONVERGED; break; }
        if (gap <= tol_gap * fmax(1.0, fabs(obj)) && nrp <= tol_res * (1 + bn) && nrg <= tol_res * (1 + hn) &&
            nrd <= tol_res * dscale) {
            status = ORACLE_OK;
            break;
        }
        /* primal infeasibility certificate: z>=0, G'z + A'y ~ 0, h'z + b'y < 0 */
        {
            double hz = 0;
            for (int r = 0; r < mi; r++) if (!dead[r]) hz += q->h[r] * z[r];
            for (int r = 0; r < ne; r++) hz += q->b[r] * y[r];
            if (hz < 0) {
                for (int i = 0; i < nv; i++) r1[i] = 0;
                csr_mulTv_add(ne, q->a_ptr, q->a_idx, q->a_val, y, r1);
                csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, z, r1);
                if (inf_norm(r1, nv) / (-hz) < 1e-8) { status = ORACLE_INFEASIBLE; break; }
            }
        }
        for (int r = 0; r < mi; r++) K.w[r] = dead[r] ? 1.0 : z[r] / s[r];
        if (kkt_factor(&K, K.w)) { status = ORACLE_NOT_CONVERGED; break; }
        /* affine direction: rc = s.z  =>  t = (z.rg - rc)/s = w.rg - z */
        for (int r = 0; r < mi; r++) tt[r] = dead[r] ? 0.0 : K.w[r] * rg[r] - z[r];
        for (int i = 0; i < nv; i++) r1[i] = -rd[i];
        for (int r = 0; r < mi; r++) tt[r] = -tt[r];
        csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, tt, r1);
        for (int r = 0; r < ne; r++) r2[r] = -rp[r];
        kkt_solve(&K, r1, r2, dx, dy);
        csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, dx, gx);
        double aa = 1.0;
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { dsa[r] = 0; dza[r] = 0; continue; }
            dsa[r] = -rg[r] - gx[r];
            dza[r] = -z[r] - K.w[r] * dsa[r];
            if (dsa[r] < 0) { double a = -s[r] / dsa[r]; if (a < aa) aa = a; }
            if (dza[r] < 0) { double a = -z[r] / dza[r]; if (a < aa) aa = a; }
        }
        double mua = 0;
        for (int r = 0; r < mi; r++) if (!dead[r]) mua += (s[r] + aa * dsa[r]) * (z[r] + aa * dza[r]);
        mua /= (n_live > 0 ? n_live : 1);
        double sigma = (mu > 0) ? pow(mua / mu, 3.0) : 0;
        /* corrector: rc = s.z + dsa.dza - sigma mu */
        for (int r = 0; r < mi; r++) {
            double rc = s[r] * z[r] + dsa[r] * dza[r] - sigma * mu;
            tt[r] = dead[r] ? 0.0 : -(z[r] * rg[r] - rc) / s[r];
        }
        for (int i = 0; i < nv; i++) r1[i] = -rd[i];
        csr_mulTv_add(mi, q->g_ptr, q->g_idx, q->g_val, tt, r1);
        kkt_solve(&K, r1, r2, dx, dy);
        csr_mulv(mi, q->g_ptr, q->g_idx, q->g_val, dx, gx);
        double am = 1e300;
        for (int r = 0; r < mi; r++) {
            if (dead[r]) { ds[r] = 0; dz[r] = 0; continue; }
            double rc = s[r] * z[r] + dsa[r] * dza[r] - sigma * mu;
            ds[r] = -rg[r] - gx[r];
            dz[r] = (-rc - z[r] * ds[r]) / s[r];
            if (ds[r] < 0) { double a = -s[r] / ds[r]; if (a < am) am = a; }
            if (dz[r] < 0) { double a = -z[r] / dz[r]; if (a < am) am = a; }
        }
        double al = fmin(1.0, 0.99 * am);
        for (int i = 0; i < nv; i++) x[i] += al * dx[i];
        for (int r = 0; r < ne; r++) y[r] += al * dy[r];
        for (int r = 0; r < mi; r++) { s[r] += al * ds[r]; z[r] += al * dz[r]; }
    }
done:
    if (obj_out) *obj_out = obj;
    if (iters_out) *iters_out = it;
    kkt_free(&K);
    free(dead);
    free(y); free(s); free(z); free(rd); free(rp); free(rg); free(r1); free(r2); free(tt);
    free(dx); free(dy); free(ds); free(dz); free(dsa); free(dza); free(gx); free(px);
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
