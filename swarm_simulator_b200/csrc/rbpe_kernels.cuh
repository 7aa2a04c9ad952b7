// rbpe_kernels.cuh -- sm_100a kernels of the batched RBP trajectory-QP engine.
//
//   assemble_kernel  (k1)  constraint assembly: per-segment SFC box, per pair/segment RSFC normal, initial
//                          control points, equality / cost / conversion constants
//                          (rbp_planner.hpp build_Q_base..build_dummy, L327-L549)
//   pdip_kernel      (k2)  one CTA per mission (Gauss-Seidel chain over its batches, L140-L201) or per
//                          (mission, batch) (Jacobi): builds the batch QP of populatebyrow (L551-L688)
//                          implicitly -- rows are never materialised as a matrix -- and solves it with a
//                          Mehrotra predictor-corrector interior-point method in FP64.  Newton systems use the
//                          null-space method over the C2 knot states (pos, vel, acc at every interior knot):
//                          x = x_p + Z sigma parametrises {Ax = b} exactly (quintic Hermite <-> Bernstein), and
//                          Z'(2Q + G'WG)Z is SPD and block tridiagonal over knots (blocks of 9b).
//   convert_kernel   (k3)  Bernstein control points -> monomial coefficients (L167-L196, timeMatrix L695-L700)
//
// The same source compiles under tests/cpu_emu (a fiber emulator of the CUDA execution model used to debug the
// kernel logic where no GPU exists); that harness is test infrastructure only and is never loaded by the product.
#pragma once
#include "rbpe_types.h"

#ifndef RBPE_EMU
#include <cuda_runtime.h>
#define RBPE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include <math.h>

#ifdef RBPE_EMU
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
#endif

namespace rbpe {

// ------------------------------------------------------------------------------------------------------------
// constants (build_Q_base, L327-L347): Q_base = 60^2 D3' G2 D3 (jerk Gram matrix in the quintic Bernstein basis),
// basis[i][j] = coefficient of t^(5-j) in B_i^5(t).  Both are integer matrices; derived, not copied.
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline double binom_d(int n, int k) {
    double r = 1;
    for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
    return r;
}
__host__ __device__ inline double q_base_entry(int a, int b) {
    // third forward difference rows r: (-1, 3, -3, 1) at columns r..r+3
    double s = 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            int ca = a - i, cb = b - j;
            if (ca < 0 || ca > 3 || cb < 0 || cb > 3) continue;
            const double st[4] = {-1, 3, -3, 1};
            double g2 = binom_d(2, i) * binom_d(2, j) / (5.0 * binom_d(4, i + j));
            s += st[ca] * g2 * st[cb];
        }
    return rint(3600.0 * s);
}
// The same matrix in exact integer arithmetic (3600/5 = 720 and C(4,k) divides 720), usable in constant expressions.
constexpr int binom_i(int n, int k) { return k == 0 ? 1 : binom_i(n, k - 1) * (n - k + 1) / k; }
constexpr int q_base_int(int a, int b) {
    int s = 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            int ca = a - i, cb = b - j;
            if (ca < 0 || ca > 3 || cb < 0 || cb > 3) continue;
            int sa = (ca == 0 || ca == 2) ? (ca == 0 ? -1 : -3) : (ca == 1 ? 3 : 1);
            int sb = (cb == 0 || cb == 2) ? (cb == 0 ? -1 : -3) : (cb == 1 ? 3 : 1);
            s += sa * binom_i(2, i) * binom_i(2, j) * (720 / binom_i(4, i + j)) * sb;
        }
    return s;
}
__host__ __device__ inline double basis_entry(int i, int j) {  // coefficient of t^(5-j) in B_i^5
    int pw = 5 - j, l = pw - i;
    if (l < 0 || l > 5 - i) return 0.0;
    return binom_d(5, i) * binom_d(5 - i, l) * ((l & 1) ? -1.0 : 1.0);
}
// i-th forward difference stencil at tau=0 / backward at tau=1 (A_0.row(i), A_T.row(i), L362-L374)
__host__ __device__ inline double diff0_entry(int i, int c) {
    if (c > i) return 0.0;
    return binom_d(i, c) * (((i - c) & 1) ? -1.0 : 1.0);
}
__host__ __device__ inline double diffT_entry(int i, int c) {
    int cc = c - (5 - i);
    if (cc < 0 || cc > i) return 0.0;
    return binom_d(i, cc) * (((i - cc) & 1) ? -1.0 : 1.0);
}

// 3x3 inverse by cofactors (row-major)
__host__ __device__ inline void inv3(const double *a, double *o) {
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

__host__ __device__ inline long pair_index(int N, int qi, int qj) {  // lexicographic qi<qj (`iter`, L477-L503)
    return (long)qi * N - (long)qi * (qi + 1) / 2 + (qj - qi - 1);
}

// ------------------------------------------------------------------------------------------------------------
// per-QP scratch size (doubles) for a batch of bs agents; must match Layout below
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t al2(size_t nd) { return (nd + 1) & ~(size_t)1; }
// leading dimension of the reduced-Hessian blocks: 9 for one-agent batches (register routines), else 9b rounded up to
// a multiple of 8 (DMMA tiles, rbpe_blockla.cuh)
__host__ __device__ inline size_t kp_of(int nb) { return nb <= 1 ? 9 : (size_t)((9 * nb + 7) & ~7); }
__host__ __device__ inline size_t scratch_doubles(int N, int M, int bs) {
    size_t n = 18 * (size_t)bs, kb = 9 * (size_t)bs, nv = n * M, nr = kb * (size_t)(M > 1 ? M - 1 : 0);
    size_t NE = (size_t)(N - bs > 0 ? N - bs : 0), rext = (size_t)bs * M * 6 * NE;
    if (bs > 0 && N % bs) {  // the last, smaller batch sees more frozen agents
        size_t nl = (size_t)(N % bs), rl = nl * M * 6 * (N - nl);
        if (rl > rext) rext = rl;
    }
    size_t rint_ = (size_t)bs * (bs - 1) / 2 * 6 * M;
    size_t kp = kp_of(bs), ninv = (kp + 31) / 32;
    size_t t = 104 + 36;
    t += 14 * al2(nv) + 2 * al2(nr) + al2(nr > 32 ? nr : 32);
    t += al2((size_t)M * bs * 36) + al2(rint_ * 6);                  // Dcp, Dint
    t += al2((size_t)(M > 1 ? M - 1 : 1) * kp * kp) + al2((size_t)(M > 2 ? M - 2 : 1) * kp * kp);
    if (bs > 1) t += al2((size_t)(M > 1 ? M - 1 : 1) * ninv * 1024) + al2((size_t)(M > 1 ? M - 1 : 1) * kp) + al2((size_t)(M > 1 ? M - 1 : 1) * kp + 32);   // Linv, wk, yk
    t += 4 * al2(rext) + 3 * al2((rext + 1) / 2) + al2(((size_t)M * bs * 6 + 1) / 2);
    t += 7 * al2(rint_) + 3 * al2((rint_ + 1) / 2);
    return t;
}

#ifdef __CUDACC__
#define RBPE_DEV __device__ __forceinline__
#define RBPE_NOINLINE __device__ __noinline__
#define RBPE_STATIC_SMEM(type, name, n) __shared__ __align__(16) type name[n]
#else
#define RBPE_DEV inline
#define RBPE_NOINLINE inline
#define RBPE_STATIC_SMEM(type, name, n) static type name[n]   /* emulator: blocks run one after another */
#endif


// ---- Jacobi exchange over peer memory (fused epilogue of the sweep kernels) -------------------------------------
// One work item (a warp of pdip1_kernel, a CTA of pdip_kernel) has finished: make its stores visible system-wide, count
// it, and let the last one raise this rank's flag in every peer's flag array (stores over NVLink, no collective call).
RBPE_DEV void peer_signal_done(const SolveArgs &S, bool leader) {
#if defined(__CUDACC__)
    if (S.npeer <= 0 || !leader) return;
    __threadfence_system();
    unsigned int prev = atomicAdd(S.done_counter, 1u);
    if (prev + 1 == S.work_items) {
        *S.done_counter = 0;
        __threadfence_system();
        for (int p = 0; p < S.npeer; p++) {
            volatile unsigned long long *f = S.peer_flags[p] + S.peer_rank;
            *f = S.sweep_id;
        }
        __threadfence_system();
    }
#endif
}

// waits until every rank has raised its flag to `sweep_id` in OUR flag array; err[0] = 1 on timeout (about 20 s: a peer
// that is merely late -- first-touch allocations, a slower host thread -- must not be mistaken for a dead one)
__global__ void peer_wait_kernel(const unsigned long long *flags, int world, unsigned long long sweep_id, int *err) {
#if defined(__CUDACC__)
    const int p = threadIdx.x;
    if (p >= world) return;
    const long long t0 = clock64();
    const volatile unsigned long long *f = flags + p;
    while (*f < sweep_id) {
        if (clock64() - t0 > 40000000000LL) { atomicExch(err, 1); return; }
        __nanosleep(200);
    }
    __threadfence_system();
#endif
}

// Phase timers of the CTA-per-QP kernel (tools/gpu_joint.py with a -DRBPE_PROFILE build; never in the product build)
#if defined(RBPE_PROFILE) && defined(__CUDACC__)
__device__ unsigned long long g_prof[16];
#define PROF_DECL long long prof_t0 = clock64()
#define PROF(slot) do { if (threadIdx.x == 0) { long long t1_ = clock64(); atomicAdd(&g_prof[slot], (unsigned long long)(t1_ - prof_t0)); prof_t0 = t1_; } __syncwarp(); } while (0)
#else
#define PROF_DECL
#define PROF(slot)
#endif

}  // namespace rbpe
#include "rbpe_blockla.cuh"
namespace rbpe {

#if defined(__CUDACC__) || defined(RBPE_EMU)

// ------------------------------------------------------------------------------------------------------------
// k1: assembly
// ------------------------------------------------------------------------------------------------------------
__global__ void assemble_kernel(AssembleArgs A) {
    const int N = A.N, M = A.M;
    const long P = (long)N * (N - 1) / 2;
    const long per_box = N, per_rel = P * M, per_dummy = (long)N * 3 * 6 * M, per_seg = M;
    const long per_mission = per_box + per_rel + per_dummy + per_seg;
    const long total = per_mission * A.count;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        int c = (int)(g / per_mission);
        long w = g - (long)c * per_mission;
        const double *T = A.T + (size_t)c * (M + 1);
        if (w < per_box) {
            // build_dlq box part (L443-L474): first box whose end time >= T[m+1]; bi is monotone over m
            int qi = (int)w;
            const int *offs = A.sfc_offs + (size_t)c * (N + 1);
            int base = A.sfc_base[c] + offs[qi], nb = offs[qi + 1] - offs[qi], bi = 0;
            for (int m = 0; m < M; m++) {
                while (bi < nb && A.sfc_t[base + bi] < T[m + 1]) bi++;
                int b = bi;
                if (b >= nb) { A.status[c] = ST_BAD_ARG; b = nb - 1; }
                double *o = A.segbox + (((size_t)c * N + qi) * M + m) * 6;
                for (int k = 0; k < 6; k++) o[k] = (b >= 0) ? A.sfc_box[(size_t)(base + b) * 6 + k] : 0.0;
            }
            continue;
        }
        w -= per_box;
        if (w < per_rel) {
            // build_dlq RSFC part (L476-L504): ri restarts from 0 for every segment
            long it = w / M;
            int m = (int)(w - it * M), ri = 0;
            const double *rt = A.rsfc_t + ((size_t)c * P + it) * M;
            while (ri < M && rt[ri] < T[m + 1]) ri++;
            if (ri >= M) { A.status[c] = ST_BAD_ARG; ri = M - 1; }
            const float *src = A.rsfc_n + (((size_t)c * P + it) * M + ri) * 3;
            float *dst = A.reln + (((size_t)c * P + it) * M + m) * 3;
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
            continue;
        }
        w -= per_rel;
        if (w < per_dummy) {
            // build_dummy (L513-L549): control points 0..2 = pi_m, 3..5 = pi_{m+1}; float32 promoted to double
            int j = (int)(w % (6 * M));
            long r = w / (6 * M);
            int k = (int)(r % 3), qi = (int)(r / 3);
            double v = 0.0;
            if (A.sequential) {
                int m = j / 6, i = j % 6, idx = m + ((i < 3) ? 0 : 1);
                v = (double)A.init_traj[(((size_t)c * N + qi) * (M + 1) + idx) * 3 + k];
            }
            A.ctrl[(size_t)c * per_dummy + w] = v;
            continue;
        }
        w -= per_dummy;
        {
            // per-segment constants: build_Aeq_base (L353-L405), build_Q_p (L349-L351), timeMatrix (L695-L700)
            int m = (int)w;
            double dt = T[m + 1] - T[m];
            double *o = A.segmat + ((size_t)c * M + m) * SEGMAT;
            double nn = 1;
            for (int d = 0; d < 3; d++) {
                double sc = pow(dt, (double)-d) * nn;
                for (int i = 0; i < 6; i++) {
                    o[SEGMAT_AL + d * 6 + i] = (m == 0 ? sc : -sc) * diff0_entry(d, i);
                    o[SEGMAT_AR + d * 6 + i] = sc * diffT_entry(d, i);
                }
                nn = nn * (5 - d);
            }
            o[SEGMAT_QS] = pow(dt, -5.0);
            for (int j = 0; j < 6; j++) o[SEGMAT_TP + j] = pow(1.0 / dt, (double)(5 - j));
            // Hermite <-> Bernstein: control points 0..2 = CL * (pos, vel, acc at the left knot), 3..5 = CR * (right knot)
            double EL[9], ER[9];
            for (int d = 0; d < 3; d++)
                for (int i = 0; i < 3; i++) {
                    double v = o[SEGMAT_AL + d * 6 + i];
                    EL[d * 3 + i] = (m == 0) ? v : -v;     // interior knots carry the minus sign of the continuity row
                    ER[d * 3 + i] = o[SEGMAT_AR + d * 6 + 3 + i];
                }
            inv3(EL, o + SEGMAT_CL);
            inv3(ER, o + SEGMAT_CR);
            // reduced cost Hessian of the segment over (left state, right state): C' (2 dt^-5 Q) C
            for (int r = 0; r < 6; r++)
                for (int cc = 0; cc < 6; cc++) {
                    const double *Cr = o + (r < 3 ? SEGMAT_CL : SEGMAT_CR), *Cc = o + (cc < 3 ? SEGMAT_CL : SEGMAT_CR);
                    double sum = 0;
                    for (int i = 0; i < 3; i++)
                        for (int ii = 0; ii < 3; ii++)
                            sum += Cr[i * 3 + r % 3] * q_base_entry((r < 3 ? 0 : 3) + i, (cc < 3 ? 0 : 3) + ii) * Cc[ii * 3 + cc % 3];
                    o[SEGMAT_RQ + r * 6 + cc] = 2.0 * o[SEGMAT_QS] * sum;
                }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// k3: conversion (L167-L196): c_j = sum_i ctrl_i * (basis[i][j] * (1/dt)^(5-j)), highest power first
// ------------------------------------------------------------------------------------------------------------
__global__ void convert_kernel(ConvertArgs A) {
    const long per = (long)A.N * 3 * 6 * A.M, total = per * A.count;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        int c = (int)(g / per);
        long w = g - (long)c * per;
        int j = (int)(w % 6), m = (int)((w / 6) % A.M);
        const double *src = A.ctrl + (size_t)c * per + (w - j);
        const double *tp = A.segmat + ((size_t)c * A.M + m) * SEGMAT + SEGMAT_TP;
        double s = 0;
        for (int i = 0; i < 6; i++) s += src[i] * (basis_entry(i, j) * tp[j]);
        A.coef[(size_t)c * per + w] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------
// k2: PDIP
// ------------------------------------------------------------------------------------------------------------
struct QP {
    int N, M, nb, q0, NE, n, kb, kp, nv, nr, nrext, nrint, mi;   // kp: leading dimension of the Wd / Wo blocks (kp_of)
    int c;  // mission
    const double *start, *goal, *radius, *segbox, *segmat;
    const float *reln;
    const double *ctrl_src;
    // x-space vectors (nv), segment-major: v = m*18nb + (a*3+k)*6 + i
    double *x, *dxa, *dx, *rdx, *ub, *lbn, *sub, *zub, *slb, *zlb, *vA, *vB, *tub, *tlb;
    // knot-space vectors (nr): r = (t-1)*9nb + (a*3+k)*3 + d, t = 1..M-1
    double *sg, *sg2;
    double *dinv;        // [max(nr, 32)] column exchange buffer of the 9 x 9 routines (one-agent batches)
    double *Wd, *Wo;     // reduced Hessian Z'HZ: (M-1) diagonal blocks, (M-2) blocks (t+1,t), each kp x kp (9nb used)
    double *panel;       // joint batches: shared-memory panel the factorisation stages with TMA (rbpe_blockla.cuh), or null
    double *Linv, *wk, *yk;   // joint batches: inverted 32 x 32 diagonal blocks of the factor, solve work vectors (rbpe_blockla.cuh)
    double *Dcp;         // [M*nb*6][6]  sum_rows w g g' restricted to one control point (3x3 symmetric over axes)
    double *Dint;        // [nrint][6]   -w n n' of a row between two batch agents
    double *he, *se, *ze, *te;   // per row: right-hand side, slack, multiplier, t = 1/(s z)  (1/s = t z, 1/z = t s)
    int *cnt_ext;                // [M*nb*6] rows against frozen agents KEPT by the presolve for control point (a, m, i); they are
                                 // stored compacted at the front of the control point's NE slots (pruned rows are never read again)
    float *nex, *ney, *nez;
    double *hi, *si, *zi, *ti;
    double *si_w, *zi_w, *ti_w;  // write side of the (s, z) pair of rows between two batch agents during the fused residual pass
    float *nix, *niy, *niz;
    double *red;  // 104 doubles: 6 per warp (up to 16 warps), verdict of the factorisation at [100]
    double *QB;   // 36 doubles: Q_base
};

struct Arena {
    unsigned char *sm;
    size_t sm_left;
    double *gl;
    RBPE_DEV double *take(size_t nd) {
        nd = al2(nd);
        size_t bytes = nd * 8;
        if (bytes <= sm_left) {
            double *p = (double *)sm;
            sm += bytes;
            sm_left -= bytes;
            return p;
        }
        double *p = gl;
        gl += nd;
        return p;
    }
};

RBPE_DEV void layout(QP &q, unsigned char *smem, size_t smem_bytes, double *gscratch) {
    Arena a;
    a.sm = smem;
    a.sm_left = smem_bytes;
    a.gl = gscratch;
    q.red = a.take(104);
    q.QB = a.take(36);
    // Throughput regime (a 48 KB arena, two CTAs per SM): the vectors of the iteration and the state of the box rows first.
    // Latency regime (the whole SM's shared memory for one CTA, rbpe_api.cu smem_for): the reduced Hessian, its factor and the
    // inverted diagonal blocks come before the box-row state -- for b = 4, N = 64 they then all fit (the substitutions and the
    // L21 products read them many times per iteration; from L2 every read is a ~250-cycle round trip for a lone CTA).
    const bool hot_first = smem_bytes > 128 * 1024;
    double **vv[14] = {&q.x, &q.dxa, &q.dx, &q.rdx, &q.ub, &q.lbn, &q.sub, &q.zub, &q.slb, &q.zlb, &q.vA, &q.vB, &q.tub, &q.tlb};
    const bool box_state[14] = {false, false, false, false, true, true, true, true, true, true, false, false, true, true};
    for (int i = 0; i < 14; i++)
        if (!hot_first || !box_state[i]) *vv[i] = a.take(q.nv);
    q.sg = a.take(q.nr);
    q.sg2 = a.take(q.nr);
    q.dinv = a.take(q.nr > 32 ? q.nr : 32);
    q.Dcp = a.take((size_t)q.M * q.nb * 36);
    q.Dint = a.take((size_t)q.nrint * 6);
    q.Wd = a.take((size_t)(q.M > 1 ? q.M - 1 : 1) * q.kp * q.kp);
    q.Wo = a.take((size_t)(q.M > 2 ? q.M - 2 : 1) * q.kp * q.kp);
    q.Linv = q.wk = q.yk = nullptr;
    if (q.nb > 1) {
        q.Linv = a.take((size_t)(q.M > 1 ? q.M - 1 : 1) * bla_ninv(q.kp) * BLA_W * BLA_W);
        q.wk = a.take((size_t)(q.M > 1 ? q.M - 1 : 1) * q.kp);
        q.yk = a.take((size_t)(q.M > 1 ? q.M - 1 : 1) * q.kp + 32);
    }
    if (hot_first)
        for (int i = 0; i < 14; i++)
            if (box_state[i]) *vv[i] = a.take(q.nv);
    q.cnt_ext = (int *)a.take(((size_t)q.M * q.nb * 6 + 1) / 2);
    q.he = a.take(q.nrext); q.se = a.take(q.nrext); q.ze = a.take(q.nrext); q.te = a.take(q.nrext);
    q.nex = (float *)a.take(((size_t)q.nrext + 1) / 2);
    q.ney = (float *)a.take(((size_t)q.nrext + 1) / 2);
    q.nez = (float *)a.take(((size_t)q.nrext + 1) / 2);
    q.hi = a.take(q.nrint); q.si = a.take(q.nrint); q.zi = a.take(q.nrint); q.ti = a.take(q.nrint);
    q.si_w = a.take(q.nrint); q.zi_w = a.take(q.nrint); q.ti_w = a.take(q.nrint);
    q.nix = (float *)a.take(((size_t)q.nrint + 1) / 2);
    q.niy = (float *)a.take(((size_t)q.nrint + 1) / 2);
    q.niz = (float *)a.take(((size_t)q.nrint + 1) / 2);
}

template <class T>
RBPE_DEV T shfl_down_t(T v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }

// CTA-wide reduction of the slots of v = (sum, sum, max, max, max, min) selected by MASK; every thread returns the same
// values (fixed order).  Unselected slots are left untouched.
template <int MASK>
RBPE_DEV void block_reduce6(double *v, double *red) {
    for (int o = 16; o > 0; o >>= 1) {
        if (MASK & 1) v[0] += shfl_down_t(v[0], o);
        if (MASK & 2) v[1] += shfl_down_t(v[1], o);
        if (MASK & 4) v[2] = fmax(v[2], shfl_down_t(v[2], o));
        if (MASK & 8) v[3] = fmax(v[3], shfl_down_t(v[3], o));
        if (MASK & 16) v[4] = fmax(v[4], shfl_down_t(v[4], o));
        if (MASK & 32) v[5] = fmin(v[5], shfl_down_t(v[5], o));
    }
    int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0)
        for (int i = 0; i < 6; i++)
            if (MASK & (1 << i)) red[w * 6 + i] = v[i];
    __syncthreads();
    if (MASK & 1) v[0] = 0;
    if (MASK & 2) v[1] = 0;
    if (MASK & 4) v[2] = -1e300;
    if (MASK & 8) v[3] = -1e300;
    if (MASK & 16) v[4] = -1e300;
    if (MASK & 32) v[5] = 1e300;
    for (int i = 0; i < nw; i++) {
        if (MASK & 1) v[0] += red[i * 6 + 0];
        if (MASK & 2) v[1] += red[i * 6 + 1];
        if (MASK & 4) v[2] = fmax(v[2], red[i * 6 + 2]);
        if (MASK & 8) v[3] = fmax(v[3], red[i * 6 + 3]);
        if (MASK & 16) v[4] = fmax(v[4], red[i * 6 + 4]);
        if (MASK & 32) v[5] = fmin(v[5], red[i * 6 + 5]);
    }
}

enum { P_INIT = 0, P_START, P_SHIFT, P_RES, P_AFF, P_COR, P_STEP, P_DEAD };
// 1: no corrector pass over the rows.  The corrector's coefficient -(z rg - rc) / s with rc = s z + dsa dza - sigma mu is
// affine in sigma mu, which is only known after the affine pass's reductions: that pass accumulates G'(part without
// sigma mu) in vA and G'(1 / s) in vB, and the solver forms vA - sigma mu vB -- three passes per iteration instead of four.
// 0: the separate pass (tools only, A/B).
#ifndef RBPE_FUSE_COR
#define RBPE_FUSE_COR 1
#endif

constexpr double PRESOLVE_FEAS_TOL = 1e-6;  // CPLEX's default feasibility tolerance, for rows made constant by the endpoints
// Dual-residual floor of the acceptance rule (CPLEX's optimality tolerance EpOpt, default 1e-6); see pdip_solve and
// oracle/rbp_oracle.c (TOL_DUAL_FLOOR) for the reasoning.
constexpr double TOL_DUAL_FLOOR = 1e-6;
// Farkas certificate ratios |(GZ)'z| / (-h'z): regular test, and the looser one that decides when the factorisation of a
// diverging (infeasible) QP breaks down first.  Reasoning in oracle/rbp_oracle.c (CERT_RATIO).
constexpr double CERT_RATIO = 1e-6, CERT_RATIO_BREAKDOWN = 1e-4;
// right-hand side marker of a row removed by the bound-based redundancy presolve
constexpr double ROW_PRUNED = 1e300;

struct Acc {  // lane-local reductions of a row pass
    double s1, s2, mx, mx2, mn;
};

// One inequality row.  in: h, s, z, t = 1/(s z), gx = g.x, ga = g.dx_aff, gd = g.dx.  sa/sb: pass scalars.
// out: cA, cB (coefficients of g in the two G' products), w (weight of g g' in H); s, z, t may be rewritten.
// One division per row and iteration (t, after the step); 1/s = t z and 1/z = t s everywhere else.  The ratio tests
// run in max form: step = 1 / max_r(-ds_r / s_r, -dz_r / z_r).
// max that keeps the accumulator when v is NaN (what fmax does here) without fmax's NaN-handling sequence
RBPE_DEV double rmax2(double acc, double v) { return (v > acc) ? v : acc; }
template <int MODE>
RBPE_DEV void row_eval(double h, double &s, double &z, double &t, double gx, double ga, double gd, double sa, double sb,
                       bool owner, double &cA, double &cB, double &w, Acc &acc) {
    cA = 0; cB = 0; w = 0;
    if (MODE == P_DEAD) { acc.mx = rmax2(acc.mx, gx - h); return; }   // constant row: violation of its right-hand side
    if (MODE == P_INIT) { w = 1.0; cA = h - gx; return; }
    if (MODE == P_START) {  // z = Gx - h, s = -z (least-squares start)
        z = gx - h; s = -z;
        if (owner) { acc.mx = rmax2(acc.mx, -s); acc.mx2 = rmax2(acc.mx2, -z); }
        return;
    }
    if (MODE == P_SHIFT) { s += sa; z += sb; t = bla_rcp(s * z); return; }
    double rs = t * z;
    if (MODE == P_RES && sb != 0.0) {
        // pending step of the previous iteration (sa = its sigma*mu, sb = its step length), fused into this pass:
        // s += al ds, z += al dz with ds, dz recomputed from the old point; then the residual at the new point
        double rgo = gx + s - h, wo = z * rs;
        double dsa = -rgo - ga, dza = -z - wo * dsa;
        double rc = s * z + dsa * dza - sa;
        double ds = -rgo - gd, dz = (-rc - z * ds) * rs;
        s += sb * ds; z += sb * dz;
        gx += sb * gd;
        t = bla_rcp(s * z);
        rs = t * z;
    }
    double rg = gx + s - h;
    w = z * rs;
    if (MODE == P_RES) {
        cA = z;
        cB = -(w * rg - z);
        if (owner) { acc.s1 += s * z; acc.s2 += h * z; acc.mx = rmax2(acc.mx, fabs(rg)); acc.mx2 = rmax2(acc.mx2, z); }
        return;
    }
    double rz = t * s;
    double dsa = -rg - ga, dza = -z - w * dsa;
    if (MODE == P_AFF) {
        acc.mx = rmax2(acc.mx, rmax2(-dsa * rs, -dza * rz));
        // sum (s + a dsa)(z + a dza) = s'z + a * s1 + a^2 * s2 for whatever step a comes out of the ratio test
        if (owner) { acc.s1 += s * dza + z * dsa; acc.s2 += dsa * dza; }
#if RBPE_FUSE_COR
        cA = -(z * rg - (s * z + dsa * dza)) * rs; cB = rs;
#endif
        return;
    }
    double rc = s * z + dsa * dza - sa;  // sa = sigma * mu
    if (MODE == P_COR) { cA = -(z * rg - rc) * rs; return; }
    double ds = -rg - gd, dz = (-rc - z * ds) * rs;
    if (MODE == P_STEP) acc.mx = rmax2(acc.mx, rmax2(-ds * rs, -dz * rz));
}

// Control points fixed by the start / goal equalities: 0..2 of the first segment, 3..5 of the last one.
RBPE_DEV bool cp_dead(const QP &q, int m, int i) { return (m == 0 && i < 3) || (m == q.M - 1 && i >= 3); }

// All inequality rows touching control point (m, a, i) of the batch, executed by a GROUP OF FOUR LANES (eight control
// points per warp): the kept rows against agents outside the batch (L643-L668), the rows against the other agents of
// the batch (L669-L680; evaluated from both ends, owned by the lower index) and the six box rows (L626-L635: x <= ub,
// -x <= -lb), each kind in its own loop so that the lanes of a warp stay on the same path; lane g of the group takes the
// rows g, g + 4, ...; G'(.) and sum w g g' of the group are added with two shuffle rounds.
// Round-2 history of this function (one 64-agent b = 4 mission, where the row passes were 40 % of the time, ncu r2f): one
// WARP per control point used ~21 of its 32 lanes, paid a 5-round butterfly over 12 values and its prologue per control
// point -- 6 control points x ~500 instructions per warp and pass at 16 warps; four lanes per control point run the same rows
// in ~800 instructions per warp and pass.  (A lone CTA pays ~6 cycles per instruction and warp.)
constexpr int CPG = 4;   // lanes per control point
template <int MODE>
RBPE_DEV void cp_group(const QP &q, const bool active, int m, int a, int i, double sa, double sb, Acc &acc) {
    constexpr bool WR = (MODE == P_START || MODE == P_SHIFT || MODE == P_RES);
    constexpr bool VEC = (MODE == P_INIT || MODE == P_RES || MODE == P_COR || (RBPE_FUSE_COR && MODE == P_AFF));
    constexpr bool VECB = (MODE == P_RES || (RBPE_FUSE_COR && MODE == P_AFF));   // the pass also produces vB
    constexpr bool MAT = (MODE == P_INIT || MODE == P_RES);
    const int g = threadIdx.x & (CPG - 1);
    double vA0 = 0, vA1 = 0, vA2 = 0, vB0 = 0, vB1 = 0, vB2 = 0;
    double Dxx = 0, Dxy = 0, Dxz = 0, Dyy = 0, Dyz = 0, Dzz = 0;
    int v0 = 0;
    if (active) {
        const int base = m * q.n;
        v0 = base + a * 18 + i;
        const double x0 = q.x[v0], x1 = q.x[v0 + 6], x2 = q.x[v0 + 12];
        const double a0 = q.dxa[v0], a1 = q.dxa[v0 + 6], a2 = q.dxa[v0 + 12];
        const double d0 = q.dx[v0], d1 = q.dx[v0 + 6], d2 = q.dx[v0 + 12];
        {   // rows against agents outside the batch
            const int task = (a * q.M + m) * 6 + i;
            const size_t rb = (size_t)task * q.NE;
            const int cnt = q.cnt_ext[task];     // kept rows only (compacted by setup_rows)
            for (int e = g; e < cnt; e += CPG) {
                const size_t r = rb + e;
                const double n0 = q.nex[r], n1 = q.ney[r], n2 = q.nez[r];
                double h = q.he[r], s = q.se[r], z = q.ze[r], t = q.te[r], cA, cB, w;
                row_eval<MODE>(h, s, z, t, n0 * x0 + n1 * x1 + n2 * x2, n0 * a0 + n1 * a1 + n2 * a2,
                               n0 * d0 + n1 * d1 + n2 * d2, sa, sb, true, cA, cB, w, acc);
                if (WR) { q.se[r] = s; q.ze[r] = z; q.te[r] = t; }
                if (VEC) { vA0 += cA * n0; vA1 += cA * n1; vA2 += cA * n2; vB0 += cB * n0; vB1 += cB * n1; vB2 += cB * n2; }
                if (MAT) {
                    const double w0 = w * n0, w1 = w * n1, w2 = w * n2;
                    Dxx += w0 * n0; Dxy += w0 * n1; Dxz += w0 * n2; Dyy += w1 * n1; Dyz += w1 * n2; Dzz += w2 * n2;
                }
            }
        }
        // rows between two agents of the batch
        for (int oo = g; oo < q.nb - 1; oo += CPG) {
            const int o = oo < a ? oo : oo + 1;
            const int lo = a < o ? a : o, hi = a < o ? o : a;
            const size_t r = ((size_t)lo * q.nb - (size_t)lo * (lo + 1) / 2 + (hi - lo - 1)) * 6 * q.M + m * 6 + i;
            const double h = q.hi[r];
            if (h >= ROW_PRUNED) continue;
            const bool own = (a == lo);
            // the start-point passes only touch a row through its owner (the partner would read (s, z, t) while the owner
            // rewrites them and use nothing of what it read)
            if (!own && (MODE == P_START || MODE == P_SHIFT)) continue;
            const double sg = own ? 1.0 : -1.0;
            const double n0 = sg * q.nix[r], n1 = sg * q.niy[r], n2 = sg * q.niz[r];
            const int vo = base + o * 18 + i;
            const double gx = n0 * (x0 - q.x[vo]) + n1 * (x1 - q.x[vo + 6]) + n2 * (x2 - q.x[vo + 12]);
            const double ga = n0 * (a0 - q.dxa[vo]) + n1 * (a1 - q.dxa[vo + 6]) + n2 * (a2 - q.dxa[vo + 12]);
            const double gd = n0 * (d0 - q.dx[vo]) + n1 * (d1 - q.dx[vo + 6]) + n2 * (d2 - q.dx[vo + 12]);
            double s = q.si[r], z = q.zi[r], t = q.ti[r], cA, cB, w;
            row_eval<MODE>(h, s, z, t, gx, ga, gd, sa, sb, own, cA, cB, w, acc);
            if (WR && own) {
                // such a row is evaluated from both of its control points; in the residual pass, which also advances
                // (s, z), the owner writes to the other half of a double buffer so that the partner still reads the old pair
                if (MODE == P_RES) { q.si_w[r] = s; q.zi_w[r] = z; q.ti_w[r] = t; }
                else { q.si[r] = s; q.zi[r] = z; q.ti[r] = t; }
            }
            if (VEC) { vA0 += cA * n0; vA1 += cA * n1; vA2 += cA * n2; vB0 += cB * n0; vB1 += cB * n1; vB2 += cB * n2; }
            if (MAT) {
                const double w0 = w * n0, w1 = w * n1, w2 = w * n2;
                Dxx += w0 * n0; Dxy += w0 * n1; Dxz += w0 * n2; Dyy += w1 * n1; Dyz += w1 * n2; Dzz += w2 * n2;
                if (own) {  // coupling block between the two agents at this control point: -w n n'
                    double *D = q.Dint + r * 6;
                    D[0] = -w0 * n0; D[1] = -w0 * n1; D[2] = -w0 * n2; D[3] = -w1 * n1; D[4] = -w1 * n2; D[5] = -w2 * n2;
                }
            }
        }
        // box rows: x_k <= ub (side 0), -x_k <= -lb (side 1)
        for (int bx = g; bx < 6; bx += CPG) {
            const int k = bx >> 1, v = v0 + 6 * k;
            const bool lower = bx & 1;
            const double sg = lower ? -1.0 : 1.0;
            const double xk = k == 0 ? x0 : (k == 1 ? x1 : x2), ak = k == 0 ? a0 : (k == 1 ? a1 : a2), dk = k == 0 ? d0 : (k == 1 ? d1 : d2);
            double *ps = (lower ? q.slb : q.sub) + v, *pz = (lower ? q.zlb : q.zub) + v, *pt = (lower ? q.tlb : q.tub) + v;
            double s = *ps, z = *pz, t = *pt, cA, cB, w;
            row_eval<MODE>(lower ? q.lbn[v] : q.ub[v], s, z, t, sg * xk, sg * ak, sg * dk, sa, sb, true, cA, cB, w, acc);
            if (WR) { *ps = s; *pz = z; *pt = t; }
            const double cAs = cA * sg, cBs = cB * sg;
            if (k == 0) { vA0 += cAs; vB0 += cBs; Dxx += w; }
            else if (k == 1) { vA1 += cAs; vB1 += cBs; Dyy += w; }
            else { vA2 += cAs; vB2 += cBs; Dzz += w; }
        }
    }
    __syncwarp();
    if (VEC) {
#pragma unroll
        for (int o = 1; o < CPG; o <<= 1) {
            vA0 += __shfl_xor_sync(0xffffffffu, vA0, o); vA1 += __shfl_xor_sync(0xffffffffu, vA1, o); vA2 += __shfl_xor_sync(0xffffffffu, vA2, o);
            if (VECB) { vB0 += __shfl_xor_sync(0xffffffffu, vB0, o); vB1 += __shfl_xor_sync(0xffffffffu, vB1, o); vB2 += __shfl_xor_sync(0xffffffffu, vB2, o); }
        }
        if (active && g == 0) {
            q.vA[v0] = vA0; q.vA[v0 + 6] = vA1; q.vA[v0 + 12] = vA2;
            if (VECB) { q.vB[v0] = vB0; q.vB[v0 + 6] = vB1; q.vB[v0 + 12] = vB2; }
        }
    }
    if (MAT) {
#pragma unroll
        for (int o = 1; o < CPG; o <<= 1) {
            Dxx += __shfl_xor_sync(0xffffffffu, Dxx, o); Dxy += __shfl_xor_sync(0xffffffffu, Dxy, o); Dxz += __shfl_xor_sync(0xffffffffu, Dxz, o);
            Dyy += __shfl_xor_sync(0xffffffffu, Dyy, o); Dyz += __shfl_xor_sync(0xffffffffu, Dyz, o); Dzz += __shfl_xor_sync(0xffffffffu, Dzz, o);
        }
        if (active && g == 0) {
            double *D = q.Dcp + ((size_t)(m * q.nb + a) * 6 + i) * 6;
            D[0] = Dxx; D[1] = Dxy; D[2] = Dxz; D[3] = Dyy; D[4] = Dyz; D[5] = Dzz;
        }
    }
}

template <int MODE>
RBPE_DEV void row_pass(const QP &q, double sa, double sb, Acc &out) {
    Acc acc;
    acc.s1 = 0; acc.s2 = 0; acc.mx = (MODE == P_AFF || MODE == P_STEP) ? 0.0 : -1e300; acc.mx2 = -1e300; acc.mn = 1e300;
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, ntask = q.M * q.nb * 6;
    const int slot = (threadIdx.x & 31) / CPG, per = 32 / CPG;
    for (int t0 = warp * per; t0 < ntask; t0 += nw * per) {   // eight consecutive control points per warp and round
        const int t = t0 + slot;
        bool active = t < ntask;
        int i = 0, a = 0, m = 0;
        if (active) {
            i = t % 6;
            const int ma = t / 6;
            if (q.nb == 1) { a = 0; m = ma; } else { a = ma % q.nb; m = ma / q.nb; }
            active = cp_dead(q, m, i) == (MODE == P_DEAD);   // live passes skip fixed control points and vice versa
        }
        cp_group<MODE>(q, active, m, a, i, sa, sb, acc);
    }
    if (MODE == P_SHIFT || MODE == P_COR || MODE == P_INIT) { __syncthreads(); return; }
    double v[6] = {acc.s1, acc.s2, acc.mx, acc.mx2, -1e300, acc.mn};
    constexpr int MASK = (MODE == P_RES) ? (1 + 2 + 4 + 8) : (MODE == P_AFF) ? (1 + 2 + 4) : (MODE == P_START) ? (4 + 8) : 4;
    block_reduce6<MASK>(v, q.red);
    out.s1 = v[0]; out.s2 = v[1]; out.mx = v[2]; out.mx2 = v[3]; out.mn = v[5];
}

// ---- knot space ------------------------------------------------------------------------------------------------
RBPE_DEV int sym6(int k, int kk) {  // index of (k,kk) in (xx, xy, xz, yy, yz, zz)
    int lo = k < kk ? k : kk, hi = k < kk ? kk : k;
    return lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
}
// out (nr) = Z' vec (nv)
RBPE_DEV void Zt_apply(const QP &q, const double *vec, double *out) {
    for (int r = threadIdx.x; r < q.nr; r += blockDim.x) {
        int t = r / q.kb + 1, cc = r % q.kb, ak = cc / 3, d = cc % 3;
        const double *CR = q.segmat + (t - 1) * SEGMAT + SEGMAT_CR, *CL = q.segmat + t * SEGMAT + SEGMAT_CL;
        const double *vl = vec + (t - 1) * q.n + ak * 6 + 3, *vr = vec + t * q.n + ak * 6;
        double s = 0;
        for (int j = 0; j < 3; j++) s += CR[j * 3 + d] * vl[j] + CL[j * 3 + d] * vr[j];
        out[r] = s;
    }
}
// out (nv) = Z sg (nr)
RBPE_DEV void Z_apply(const QP &q, const double *sg, double *out) {
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        int m = v / q.n, r = v % q.n, ak = r / 6, i = r % 6;
        double s = 0;
        if (i < 3) {
            if (m > 0) {
                const double *C = q.segmat + m * SEGMAT + SEGMAT_CL + i * 3, *g = sg + (m - 1) * q.kb + ak * 3;
                s = C[0] * g[0] + C[1] * g[1] + C[2] * g[2];
            }
        } else if (m < q.M - 1) {
            const double *C = q.segmat + m * SEGMAT + SEGMAT_CR + (i - 3) * 3, *g = sg + m * q.kb + ak * 3;
            s = C[0] * g[0] + C[1] * g[1] + C[2] * g[2];
        }
        out[v] = s;
    }
}

// reduced Hessian Z'(2Q + G'WG)Z from the per-control-point blocks left by the last P_INIT / P_RES pass
RBPE_DEV void build_W(const QP &q) {
    const int kb = q.kb, kp = q.kp, kk = kp * kp, M = q.M;
    // a thread owns entry (r, c) of EVERY knot's block: the index arithmetic (agent, axis, derivative of row and column, pair
    // number) is done once per entry instead of once per entry and knot (this phase was 11 % of the joint kernel at b = 16)
    for (int idx = threadIdx.x; idx < kk; idx += blockDim.x) {
        const int r = idx / kp, c = idx - r * kp;
        if (r >= kb || c >= kb || c > r) {
            const double v = (r >= kb || c >= kb) ? ((r == c) ? 1.0 : 0.0) : 0.0;   // identity padding up to a multiple of 8; zero above the diagonal
            for (int t = 1; t < M; t++) q.Wd[(size_t)(t - 1) * kk + idx] = v;
        } else {
            const int a = r / 9, k = (r % 9) / 3, d = r % 3, a2 = c / 9, k2 = (c % 9) / 3, d2 = c % 3;
            const int e = sym6(k, k2);
            const bool same_agent = a == a2, same_axis = same_agent && k == k2;
            // per-control-point blocks: rows of one agent, or the rows between two batch agents (a > a2)
            const size_t pp = same_agent ? 0 : (size_t)a2 * q.nb - (size_t)a2 * (a2 + 1) / 2 + (a - a2 - 1);
            for (int t = 1; t < M; t++) {
                const double *sl = q.segmat + (t - 1) * SEGMAT, *sr = q.segmat + t * SEGMAT;
                const double *CR = sl + SEGMAT_CR, *CL = sr + SEGMAT_CL;
                const double *Dl = same_agent ? q.Dcp + ((size_t)((t - 1) * q.nb + a) * 6 + 3) * 6 + e : q.Dint + (pp * 6 * M + (t - 1) * 6 + 3) * 6 + e;
                const double *Dr = same_agent ? q.Dcp + ((size_t)(t * q.nb + a) * 6) * 6 + e : q.Dint + (pp * 6 * M + t * 6) * 6 + e;
                double s = 0;
                for (int j = 0; j < 3; j++) s += CR[j * 3 + d] * CR[j * 3 + d2] * Dl[j * 6] + CL[j * 3 + d] * CL[j * 3 + d2] * Dr[j * 6];
                if (same_axis) s += sl[SEGMAT_RQ + (3 + d) * 6 + 3 + d2] + sr[SEGMAT_RQ + d * 6 + d2];
                q.Wd[(size_t)(t - 1) * kk + idx] = s;
            }
        }
        // blocks (t+1, t), t = 1..M-2: only the cost couples neighbouring knots, per (agent, axis)
        const bool coupled = r < kb && c < kb && r / 3 == c / 3;
        for (int t = 1; t < M - 1; t++)
            q.Wo[(size_t)(t - 1) * kk + idx] = coupled ? q.segmat[t * SEGMAT + SEGMAT_RQ + (3 + r % 3) * 6 + c % 3] : 0.0;
    }
}

// one-agent batches: warp 0 factors out of registers, verdict through q.red[100]; joint batches: CTA-wide DMMA routines
RBPE_NOINLINE bool kkt_factor(const QP &q) {
    PROF_DECL;
    __syncthreads();
    build_W(q);
    __syncthreads();
    PROF(5);
    if (q.kb == 9) {
        if ((threadIdx.x >> 5) == 0) {
            bool ok = factor_bt9v<0>(q.M - 1, q.Wd, q.Wo, q.dinv);
            if (threadIdx.x == 0) q.red[100] = ok ? 0.0 : 1.0;
        }
        __syncthreads();
        return q.red[100] == 0.0;
    }
    return factor_bt_blk(q.M - 1, q.kp, q.Wd, q.Wo, q.Linv, q.red + 100, q.panel);
}

// dxout (nv) = Z (Z'HZ)^-1 Z' r   with r (nv) in x-space; uses q.sg
RBPE_NOINLINE void kkt_solve(const QP &q, const double *r, double *dxout) {
    PROF_DECL;
    __syncthreads();
    Zt_apply(q, r, q.sg);
    __syncthreads();
    PROF(10);
    if (q.kb == 9) {
        if ((threadIdx.x >> 5) == 0) solve_bt9v<0>(q.M - 1, q.Wd, q.Wo, q.sg, q.dinv);
    } else {
        if (q.kp <= 2 * BLA_W) solve_bt_small(q.M - 1, q.kb, q.kp, q.Wd, q.Wo, q.Linv, q.sg, q.wk, q.yk);
        else solve_bt_blk(q.M - 1, q.kb, q.kp, q.Wd, q.Wo, q.Linv, q.sg, q.wk, q.yk);
    }
    __syncthreads();
    PROF(11);
    Z_apply(q, q.sg, dxout);
    __syncthreads();
    PROF(12);
}

// Bounds of one axis of a corridor box as the solver sees them.  A box of (numerically) zero width -- an agent flying
// along the world boundary next to a pillar gets one -- is a fixed variable to CPLEX's presolve, and any point within its
// feasibility tolerance (1e-6) of the bounds is feasible to it; an interior-point method needs an interior, so such a pair
// of bounds is opened to [mid - 1e-6, mid + 1e-6].  Same rule in oracle/rbp_oracle.c (presolve_dead_rows).
RBPE_DEV void box_bounds(const double *box, int k, double &lb, double &ub) {
    lb = box[k]; ub = box[3 + k];
    const double wdt = ub - lb;
    if (wdt < 2 * PRESOLVE_FEAS_TOL && wdt > -2 * PRESOLVE_FEAS_TOL) {
        const double mid = 0.5 * (ub + lb);
        ub = mid + PRESOLVE_FEAS_TOL; lb = mid - PRESOLVE_FEAS_TOL;
    }
}

// Bounds of control point i of segment m (box = the segment's box, its neighbours' boxes at box -+ 6): the segment's own
// box, and for the two control points that C0 continuity identifies with a control point of the neighbouring segment
// (i = 5 with the next segment's i = 0) the same tolerance rule applied to the INTERSECTION of the two boxes: consecutive
// corridor boxes that only share a face (worlds/map32.bt of the reference's smoke loop has one) pin the knot to that face,
// which CPLEX's presolve turns into a fixed variable; here the two binding faces are moved 1e-6 apart.
RBPE_DEV void cp_bounds(const double *box, int M, int m, int i, int k, double &lb, double &ub) {
    box_bounds(box, k, lb, ub);
    const double *ob = (i == 5 && m + 1 < M) ? box + 6 : ((i == 0 && m > 0) ? box - 6 : nullptr);
    if (!ob) return;
    double olb, oub;
    box_bounds(ob, k, olb, oub);
    const double lo = lb > olb ? lb : olb, hi = ub < oub ? ub : oub, wdt = hi - lo;
    if (wdt < 2 * PRESOLVE_FEAS_TOL && wdt > -2 * PRESOLVE_FEAS_TOL) {
        const double mid = 0.5 * (hi + lo);
        if (ub < mid + PRESOLVE_FEAS_TOL) ub = mid + PRESOLVE_FEAS_TOL;
        if (lb > mid - PRESOLVE_FEAS_TOL) lb = mid - PRESOLVE_FEAS_TOL;
    }
}

// rows of populatebyrow for the batch starting at agent q0 (h, normals; s = z = 1 until the start point is known);
// x <- particular solution x_p (fixed control points from the start / goal states, zero elsewhere)
RBPE_DEV void setup_rows(const QP &q) {
    const int M = q.M, N = q.N;
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        int m = v / q.n, r = v % q.n, a = r / 18, k = (r % 18) / 6, i = r % 6;
        const double *box = q.segbox + ((size_t)(q.q0 + a) * M + m) * 6;
        {
            double lb, ub;
            cp_bounds(box, M, m, i, k, lb, ub);
            q.ub[v] = ub;
            q.lbn[v] = -lb;
        }
        q.sub[v] = 1; q.zub[v] = 1; q.slb[v] = 1; q.zlb[v] = 1; q.tub[v] = 1; q.tlb[v] = 1;
        double xp = 0;
        if (m == 0 && i < 3) {          // build_deq rows 0..2 (L408-L432): start pos / vel / acc
            const double *C = q.segmat + SEGMAT_CL + i * 3, *st = q.start + (size_t)(q.q0 + a) * 9 + k;
            xp = C[0] * st[0] + C[1] * st[3] + C[2] * st[6];
        }
        if (m == M - 1 && i >= 3) {     // rows 3..5: goal
            const double *C = q.segmat + (M - 1) * SEGMAT + SEGMAT_CR + (i - 3) * 3, *gl = q.goal + (size_t)(q.q0 + a) * 9 + k;
            xp = C[0] * gl[0] + C[1] * gl[3] + C[2] * gl[6];
        }
        q.x[v] = xp; q.dxa[v] = 0; q.dx[v] = 0; q.vA[v] = 0; q.vB[v] = 0;
    }
    // RSFC rows against frozen agents: g = sg*n on x_a,  h = sg*n.dummy_other - (r_a + r_other).  One warp per control point
    // (a, m, i); the rows that survive the presolve are written COMPACTED, in their original order, to the front of the
    // control point's NE slots (ballot + prefix count), so that no later pass touches a pruned row: 60-95 % of the rows go,
    // and a pass over a control point of a 256-agent mission reads 1-2 lines instead of 7 (ncu r2: the row passes of a
    // single b = 4 mission spent more than half their samples waiting for these loads).
    {
        const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31, ntask = q.nb * M * 6;
        for (int task = warp; task < ntask; task += nw) {
            const int i = task % 6, m = (task / 6) % M, a = task / (6 * M);
            const int qa = q.q0 + a;
            const bool dead = cp_dead(q, m, i);
            const double *box = q.segbox + ((size_t)qa * M + m) * 6;
            double blo[3], bhi[3];
            for (int k = 0; k < 3; k++) cp_bounds(box, M, m, i, k, blo[k], bhi[k]);
            const size_t rb = (size_t)task * q.NE;
            int kept = 0;
            for (int e0 = 0; e0 < q.NE; e0 += 32) {
                const int e = e0 + lane;
                bool keep = false;
                double h = 0;
                float f0 = 0, f1 = 0, f2 = 0;
                if (e < q.NE) {
                    const int qo = (e < q.q0) ? e : e + q.nb;
                    const double sg = (qa < qo) ? 1.0 : -1.0;
                    const long it = (qa < qo) ? pair_index(N, qa, qo) : pair_index(N, qo, qa);
                    const float *nf = q.reln + ((size_t)it * M + m) * 3;
                    f0 = nf[0]; f1 = nf[1]; f2 = nf[2];
                    const double *co = q.ctrl_src + (size_t)qo * 18 * M + m * 6 + i;
                    h = -(q.radius[qa] + q.radius[qo]);
                    h += sg * ((double)f0 * co[0]);
                    h += sg * ((double)f1 * co[6 * M]);
                    h += sg * ((double)f2 * co[12 * M]);
                    if (sg < 0) { f0 = -f0; f1 = -f1; f2 = -f2; }
                    keep = true;
                    if (!dead) {
                        // Bound-based row redundancy (standard presolve): if the largest value of g.x over the bounds of the
                        // control point stays below h the row can never be active and is dropped; the feasible set is unchanged.
                        const double gg[3] = {(double)f0, (double)f1, (double)f2};
                        double amax = 0;
                        for (int k = 0; k < 3; k++) {
                            double a1 = gg[k] * bhi[k], b1 = gg[k] * blo[k];
                            amax += (a1 > b1) ? a1 : b1;
                        }
                        if (amax < h - 1e-9 * fmax(1.0, fabs(h))) keep = false;
                    }
                }
                const unsigned mask = __ballot_sync(0xffffffffu, keep);
                if (keep) {
                    const size_t r = rb + kept + __popc(mask & ((1u << lane) - 1u));
                    q.nex[r] = f0; q.ney[r] = f1; q.nez[r] = f2;
                    q.he[r] = h; q.se[r] = 1; q.ze[r] = 1; q.te[r] = 1;
                }
                kept += __popc(mask);
            }
            if (lane == 0) q.cnt_ext[task] = kept;
        }
    }
    // RSFC rows between two batch agents lo<hi: n.x_lo - n.x_hi <= -(r_lo + r_hi)
    for (int r = threadIdx.x; r < q.nrint; r += blockDim.x) {
        int j = r % (6 * M), pp = r / (6 * M), m = j / 6;
        int lo = 0, rem = pp;
        while (rem >= q.nb - 1 - lo) { rem -= q.nb - 1 - lo; lo++; }
        int hi = lo + 1 + rem;
        long it = pair_index(N, q.q0 + lo, q.q0 + hi);
        const float *nf = q.reln + ((size_t)it * M + m) * 3;
        q.nix[r] = nf[0]; q.niy[r] = nf[1]; q.niz[r] = nf[2];
        double hh = -(q.radius[q.q0 + lo] + q.radius[q.q0 + hi]);
        if (!cp_dead(q, m, j % 6)) {   // same presolve over the boxes of both control points (g = +n on lo, -n on hi)
            const double *bl = q.segbox + ((size_t)(q.q0 + lo) * M + m) * 6, *bh = q.segbox + ((size_t)(q.q0 + hi) * M + m) * 6;
            double amax = 0;
            for (int k = 0; k < 3; k++) {
                double g = (double)nf[k], lb, ub;
                if (g == 0) continue;
                cp_bounds(bl, M, m, j % 6, k, lb, ub);
                double a1 = g * ub, b1 = g * lb;
                amax += (a1 > b1) ? a1 : b1;
                cp_bounds(bh, M, m, j % 6, k, lb, ub);
                a1 = -g * ub; b1 = -g * lb;
                amax += (a1 > b1) ? a1 : b1;
            }
            if (amax < hh - 1e-9 * fmax(1.0, fabs(hh))) hh = ROW_PRUNED;
        }
        q.hi[r] = hh;
        for (int e6 = 0; e6 < 6; e6++) q.Dint[(size_t)r * 6 + e6] = 0.0;   // a dropped row contributes no coupling block
        q.si[r] = 1; q.zi[r] = 1; q.ti[r] = 1;
    }
}

// rdx = 2Q x + vA (x-space); returns lane-local partial sums through the arguments
RBPE_DEV void dual_residual(const QP &q, double &obj_part, double &mpx) {
    for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
        int m = v / q.n, i = v % 6, b6 = v - i;
        double s = 0;
        for (int j = 0; j < 6; j++) s += q.QB[i * 6 + j] * q.x[b6 + j];
        double pxv = 2.0 * q.segmat[m * SEGMAT + SEGMAT_QS] * s;
        q.rdx[v] = pxv + q.vA[v];
        obj_part += 0.5 * q.x[v] * pxv;
        mpx = fmax(mpx, fabs(pxv));
    }
}
RBPE_DEV double block_max(double v, double *red) {
    double r[6] = {0, 0, v, -1e300, -1e300, 1e300};
    block_reduce6<4>(r, red);
    return r[2];
}


// Solves the QP described by q. Returns status; x holds the solution.
RBPE_DEV int pdip_solve(const QP &q_in, int max_iter, double tol_gap, double tol_res, double *obj_out, int *it_out,
                        double *res_out) {
    QP q = q_in;
    const int tid = threadIdx.x, nt = blockDim.x;
    Acc acc;
    PROF_DECL;
    setup_rows(q);
    __syncthreads();
    PROF(0);
    int status = ST_NOT_CONVERGED, it = 0;
    double obj = 0, gap = 0, nrd = 0, nrg = 0, hn = 0;
    bool go = true;
    // presolve: rows on fixed control points are constants; check them and leave them out
    row_pass<P_DEAD>(q, 0, 0, acc);
    if (acc.mx > PRESOLVE_FEAS_TOL) { status = ST_INFEASIBLE; go = false; }
    if (go && q.nr == 0) {  // single segment: the endpoints fix everything
        double o = 0, mpx = 0;
        dual_residual(q, o, mpx);
        double r[6] = {o, 0, mpx, -1e300, -1e300, 1e300};
        block_reduce6<1>(r, q.red);
        obj = r[0];
        status = ST_OK;
        go = false;
    }
    if (go) {
        double mh = 0;
        for (int v = tid; v < q.nv; v += nt) {
            int m = v / q.n, i = v % 6;
            if (!cp_dead(q, m, i)) mh = fmax(mh, fmax(fabs(q.ub[v]), fabs(q.lbn[v])));
        }
        double live = 0;
        for (int task = tid >> 5; task < q.nb * q.M * 6; task += nt >> 5) {   // kept rows of the live control points
            const int i = task % 6, m = (task / 6) % q.M;
            if (cp_dead(q, m, i)) continue;
            const int cnt = q.cnt_ext[task];
            for (int e = tid & 31; e < cnt; e += 32) { mh = fmax(mh, fabs(q.he[(size_t)task * q.NE + e])); live += 1; }
        }
        for (int r = tid; r < q.nrint; r += nt) {
            int j = r % (6 * q.M);
            if (!cp_dead(q, j / 6, j % 6) && q.hi[r] < ROW_PRUNED) { mh = fmax(mh, fabs(q.hi[r])); live += 1; }
        }
        {
            double rr[6] = {live, 0, mh, -1e300, -1e300, 1e300};
            block_reduce6<1 + 4>(rr, q.red);
            hn = rr[2];
            q.mi = (int)(rr[0] + 0.5) + q.nb * (6 * q.M - 6) * 6;   // kept RSFC rows + the box rows of the live control points
        }
        // ---- initial point: W = I,  min 1/2 x'(P + G'G)x - (G'h)'x  over x = x_p + Z sigma ----
        row_pass<P_INIT>(q, 0, 0, acc);   // vA = G'(h - G x_p), Dcp/Dint with unit weights
        if (!kkt_factor(q)) go = false;
    }
    if (go) {
        double o = 0, mpx = 0;
        dual_residual(q, o, mpx);          // rdx = P x_p + vA
        __syncthreads();
        for (int v = tid; v < q.nv; v += nt) q.rdx[v] = 2.0 * q.vA[v] - q.rdx[v];   // G'(h - G x_p) - P x_p
        kkt_solve(q, q.rdx, q.dx);
        for (int v = tid; v < q.nv; v += nt) { q.x[v] += q.dx[v]; q.dx[v] = 0; }
        __syncthreads();
        row_pass<P_START>(q, 0, 0, acc);  // acc.mx = max(-s), acc.mx2 = max(-z)
        double ap = acc.mx, ad = acc.mx2;
        row_pass<P_SHIFT>(q, ap >= 0 ? 1.0 + ap : 0.0, ad >= 0 ? 1.0 + ad : 0.0, acc);
    }

    double sigmu = 0, al = 0;   // pending step of the previous iteration, applied inside the next residual pass
    for (it = 0; go && it < max_iter; it++) {
        // ---- residuals (with the previous step's s, z update fused in) ----
        PROF(4);
        row_pass<P_RES>(q, sigmu, al, acc);  // vA = G'z, vB = G't_aff, Dcp/Dint; s1 = s'z, s2 = h'z, mx = |rg|
        PROF(1);
        { double *t1 = q.si; q.si = q.si_w; q.si_w = t1; t1 = q.zi; q.zi = q.zi_w; q.zi_w = t1; t1 = q.ti; q.ti = q.ti_w; q.ti_w = t1; }
        if (al != 0.0) {
            for (int v = tid; v < q.nv; v += nt) q.x[v] += al * q.dx[v];
            __syncthreads();
        }
        double mu = acc.s1 / (q.mi > 0 ? q.mi : 1), hz = acc.s2;
        nrg = fmax(acc.mx, 0.0);
        double o = 0, mpx = 0;
        dual_residual(q, o, mpx);
        __syncthreads();
        Zt_apply(q, q.rdx, q.sg);
        Zt_apply(q, q.vA, q.sg2);
        __syncthreads();
        double mr = 0, mc = 0;
        for (int r = tid; r < q.nr; r += nt) { mr = fmax(mr, fabs(q.sg[r])); mc = fmax(mc, fabs(q.sg2[r])); }
        double rr[6] = {o, 0, mpx, mr, mc, 1e300};
        block_reduce6<1 + 4 + 8 + 16>(rr, q.red);
        obj = rr[0]; mpx = rr[2]; nrd = rr[3];
        double mcert = rr[4];
        gap = mu;
        if (!(mu == mu) || !(nrd == nrd)) { status = ST_NOT_CONVERGED; break; }
        {
            // Strict test first.  On (nearly) degenerate QPs the dual residual has a round-off floor that RISES as mu
            // falls (multiplier noise eps w |dx| with w = z/s ~ 1/mu and |dx| ~ sqrt(mu)); once complementarity and primal
            // feasibility meet their strict tolerances the iterate is accepted at CPLEX's own optimality tolerance (1e-6,
            // relative to the gradient scale) instead of iterating into a numerically singular factorisation.
            const bool gap_ok = gap <= tol_gap * fmax(1.0, fabs(obj)) && nrg <= tol_res * (1 + hn);
            if (gap_ok && nrd <= tol_res * (1.0 + mpx)) { status = ST_OK; break; }
            if (gap_ok && nrd <= TOL_DUAL_FLOOR * (1.0 + mpx)) { status = ST_OK; break; }
        }
        // Farkas certificate of the reduced problem.  The largest uniform slack of the rows equals min h'z / sum(z) over
        // z >= 0 with (GZ)'z = 0, so "infeasible beyond the feasibility tolerance" needs h'z < -1e-6 sum(z) (max(z) is
        // used); rows that are consistent but have no interior (slack exactly 0) are not certified infeasible.
        const double cert = (hz < -PRESOLVE_FEAS_TOL * acc.mx2) ? mcert / (-hz) : 1e300;
        if (cert < CERT_RATIO) { status = ST_INFEASIBLE; break; }
        // ---- factor with W = z/s ----
        PROF(4);
        if (!kkt_factor(q)) { status = cert < CERT_RATIO_BREAKDOWN ? ST_INFEASIBLE : ST_NOT_CONVERGED; break; }
        PROF(2);
        // ---- affine direction ----
        for (int v = tid; v < q.nv; v += nt) q.vB[v] = -q.rdx[v] + q.vB[v];
        kkt_solve(q, q.vB, q.dxa);
        PROF(3);
        row_pass<P_AFF>(q, 0, 0, acc);       // mx = ratio test (max form); s1, s2 = linear / quadratic coefficient of mu_aff(a)
        PROF(1);
        double aa = (acc.mx > 1.0) ? 1.0 / acc.mx : 1.0;
        double mua = (mu * (q.mi > 0 ? q.mi : 1) + aa * acc.s1 + aa * aa * acc.s2) / (q.mi > 0 ? q.mi : 1);
        double sigma = (mu > 0) ? (mua / mu) * (mua / mu) * (mua / mu) : 0.0;
        sigmu = sigma * mu;
        // ---- corrector ----
#if RBPE_FUSE_COR
        for (int v = tid; v < q.nv; v += nt) q.vA[v] = -q.rdx[v] + (q.vA[v] - sigmu * q.vB[v]);
#else
        row_pass<P_COR>(q, sigmu, 0, acc);
        PROF(1);
        for (int v = tid; v < q.nv; v += nt) q.vA[v] = -q.rdx[v] + q.vA[v];
#endif
        kkt_solve(q, q.vA, q.dx);
        PROF(3);
        row_pass<P_STEP>(q, sigmu, 0, acc);
        PROF(1);
        al = (0.99 < acc.mx) ? 0.99 / acc.mx : 1.0;   // min(1, 0.99 * max step); s, z, x advance at the top of the next iteration
    }
    __syncthreads();
    // |Ax - b| for the record (the parametrisation keeps it at rounding level)
    double mrp = 0;
    for (int e = tid; e < q.kb * (q.M + 1); e += nt) {
        int t = e / q.kb, cc = e % q.kb, ak = cc / 3, d = cc % 3, a = ak / 3, k = ak % 3;
        double sm = 0;
        if (t < q.M) {
            const double *sp = q.segmat + t * SEGMAT + SEGMAT_AL + d * 6, *xx = q.x + t * q.n + ak * 6;
            for (int i = 0; i < 6; i++) sm += sp[i] * xx[i];
        }
        if (t > 0) {
            const double *sp = q.segmat + (t - 1) * SEGMAT + SEGMAT_AR + d * 6, *xx = q.x + (t - 1) * q.n + ak * 6;
            for (int i = 0; i < 6; i++) sm += sp[i] * xx[i];
        }
        if (t == 0) sm -= q.start[(size_t)(q.q0 + a) * 9 + k + 3 * d];
        if (t == q.M) sm -= q.goal[(size_t)(q.q0 + a) * 9 + k + 3 * d];
        mrp = fmax(mrp, fabs(sm));
    }
    mrp = block_max(mrp, q.red);
    if (tid == 0) {
        *obj_out = obj;
        *it_out = it;
        res_out[0] = gap; res_out[1] = mrp; res_out[2] = nrd; res_out[3] = nrg;
    }
    return status;
}

#ifndef RBPE_PDIP_MINB
#define RBPE_PDIP_MINB 1
#endif
__global__ void __launch_bounds__(CTA_THREADS_MAX, RBPE_PDIP_MINB) pdip_kernel(SolveArgs S) {
    RBPE_DYN_SMEM(smem);
    const int N = S.N, M = S.M;
    int c, l_begin, l_end;
    if (S.mode == 0) {
        c = blockIdx.x; l_begin = 0; l_end = S.nbatch;
    } else {
        int per = S.batch_end - S.batch_begin;
        c = blockIdx.x / per;
        l_begin = S.batch_begin + blockIdx.x % per;
        l_end = l_begin + 1;
    }
    if (c >= S.count) return;
    if (S.status[c] != ST_OK && S.mode == 0) return;
    const long P = (long)N * (N - 1) / 2;
    QP q;
    q.N = N; q.M = M; q.c = c;
    q.start = S.start + (size_t)c * N * 9;
    q.goal = S.goal + (size_t)c * N * 9;
    q.radius = S.radius + (size_t)c * N;
    q.segbox = S.segbox + (size_t)c * N * M * 6;
    q.segmat = S.segmat + (size_t)c * M * SEGMAT;
    q.reln = S.reln + (size_t)c * P * M * 3;
    double *ctrl = S.ctrl + (size_t)c * N * 18 * M;
    q.ctrl_src = (S.mode == 0) ? ctrl : S.ctrl_frozen + (size_t)c * N * 18 * M;
    double *gs = S.scratch + (size_t)blockIdx.x * S.scratch_stride;
    const int iters = (S.mode == 0) ? S.iteration : 1;
    for (int iter = 0; iter < iters; iter++)
        for (int l = l_begin; l < l_end; l++) {
            q.q0 = l * S.bs;
            q.nb = (q.q0 + S.bs <= N) ? S.bs : N - q.q0;
            if (q.nb <= 0) continue;
            q.NE = S.sequential ? N - q.nb : 0;
            q.n = 18 * q.nb; q.kb = 9 * q.nb; q.kp = (int)kp_of(q.nb); q.nv = q.n * M; q.nr = q.kb * (M > 1 ? M - 1 : 0);
            q.nrext = q.nb * M * 6 * q.NE;
            q.nrint = q.nb * (q.nb - 1) / 2 * 6 * M;
            {   // live rows: every row on the 6M-6 control points per agent that the endpoints do not fix
                int live_cp = 6 * M - 6;
                q.mi = q.nb * live_cp * (6 + q.NE) + q.nb * (q.nb - 1) / 2 * live_cp;
            }
            {   // the front of the dynamic shared memory is the TMA panel of the factorisation (joint batches only)
                const size_t pb = (q.nb > 1 && S.panel_bytes >= bla_panel_doubles(q.kp) * 8) ? bla_panel_doubles(q.kp) * 8 : 0;
                q.panel = pb ? (double *)smem : nullptr;
                layout(q, smem + pb, S.smem_bytes - pb, gs);
            }
            if (threadIdx.x < 36) q.QB[threadIdx.x] = q_base_entry(threadIdx.x / 6, threadIdx.x % 6);
            __syncthreads();
            int rec = (S.mode == 0 ? iter * S.nbatch : S.rec_offset) + l;
            int st = pdip_solve(q, S.max_iter, S.tol_gap, S.tol_res, S.qp_obj + (size_t)c * S.nrec + rec,
                                S.qp_iters + (size_t)c * S.nrec + rec, S.qp_res + ((size_t)c * S.nrec + rec) * 4);
            if (threadIdx.x == 0) {
                S.qp_status[(size_t)c * S.nrec + rec] = st;
                if (st != ST_OK && S.status[c] == ST_OK) S.status[c] = st;
            }
            if (st != ST_OK && S.mode == 0) return;  // RBPPlanner::update() aborts on the first failed batch (L158-L161)
            if (st != ST_OK && S.npeer <= 0) continue;
            // dummy <- vals for the agents of the batch (L182-L184); with peers: into the next table of every rank (a
            // failed batch carries its old control points over, so that the next table is complete)
            for (int v = threadIdx.x; v < q.nv; v += blockDim.x) {
                int m = v / q.n, r = v % q.n, a = r / 18, k = (r % 18) / 6, i = r % 6;
                const size_t at = (size_t)(q.q0 + a) * 18 * M + (size_t)k * 6 * M + m * 6 + i;
                if (S.npeer <= 0) { ctrl[at] = q.x[v]; continue; }
                const double val = (st == ST_OK) ? q.x[v] : q.ctrl_src[at];
                for (int p = 0; p < S.npeer; p++) S.peer_ctrl[p][(size_t)c * N * 18 * M + at] = val;
            }
            __syncthreads();
        }
    peer_signal_done(S, threadIdx.x == 0);
}

#endif  // __CUDACC__ || RBPE_EMU
}  // namespace rbpe

#ifdef RBPE_W1_V1
#include "rbpe_pdip1_v1.cuh"
#else
#include "rbpe_pdip1.cuh"
#include "rbpe_pdip1x.cuh"
#endif
