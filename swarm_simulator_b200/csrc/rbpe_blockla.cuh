// rbpe_blockla.cuh -- CTA-wide blocked FP64 linear algebra for the joint-batch Newton systems (batch size b > 1).
//
// The reduced Hessian Z'(2Q + G'WG)Z of a batch QP is SPD and block tridiagonal over the M-1 interior knots with
// dense blocks of kb = 9b (36 at the reference's launch default b = 4, 144 at b = 16, 288 at b = 32).  This file
// factors and solves it with the FP64 tensor-core instruction of sm_100a, mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4;
// tcgen05 has no FP64 kind, and the interior-point method needs FP64 -- DESIGN.md section 4).
//
//   chol_tall   left-looking block-column Cholesky (32 columns per block column) of the "tall" matrix [D_t; O_t]:
//               rows 0..kp-1 are the diagonal block (lower triangle), rows kp..2kp-1 the block (t+1, t), which so
//               becomes L_{t+1,t} = O_t L_tt^-T without a separate triangular solve.  The update of a block column
//               (all earlier columns, plus L_{t,t-1} L_{t,t-1}' of the previous knot) is one DMMA k-loop per
//               8 x 32 strip, operands straight from L1/L2.
//   chol32_warp the 32 x 32 diagonal block of a block column is factored AND inverted by one warp out of registers;
//               the rows below then become L21 = A21 L11^-T = A21 X' by DMMA as well, and the two solves of an
//               interior-point iteration reuse X as matrix-vector products (2 barriers per 32 unknowns).
//   solve_bt_blk forward / backward block substitution.
//
// Matrices are row-major with leading dimension kp = kb rounded up to a multiple of 8 (identity padding).
#pragma once

namespace rbpe {

#if defined(__CUDACC__) || defined(RBPE_EMU)

constexpr int BLA_W = 32;   // columns per block column = order of the inverted diagonal blocks

__host__ __device__ inline int bla_kp(int kb) { return (kb + 7) & ~7; }
__host__ __device__ inline int bla_ninv(int kp) { return (kp + BLA_W - 1) / BLA_W; }

// d (8x8, lane (g,t) holds d[g][2t], d[g][2t+1]) += a (8x4: lane holds a[g][t]) * b (4x8: lane holds b[t][g])
RBPE_DEV void dmma884(double &d0, double &d1, double a, double b) {
#ifdef RBPE_EMU
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int k = 0; k < 4; k++) {
        double ak = __shfl_sync(0xffffffffu, a, g * 4 + k);
        double b0 = __shfl_sync(0xffffffffu, b, (2 * t) * 4 + k), b1 = __shfl_sync(0xffffffffu, b, (2 * t + 1) * 4 + k);
        d0 += ak * b0;
        d1 += ak * b1;
    }
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
#endif
}

// ---- TMA staging of the shared operand of a block column ------------------------------------------------------------
// Every 8-row strip of a block column multiplies its own earlier columns with the SAME 32 rows of L (the rows of the
// block column's diagonal block): that panel is copied once per block column into shared memory by the bulk-copy engine
// (cp.async.bulk global -> shared, completion on an mbarrier; SASS UBLKCP) and all warps feed their DMMAs from there.
// Row r of the panel lands at r * (klen + 4) doubles: the stride is 4 mod 16 doubles, so the 8 x 4 doubles a DMMA B operand
// load touches fall into distinct banks within each half warp.
constexpr int BLA_PANEL_PAD = 4;
__host__ __device__ inline size_t bla_panel_doubles(int kp) { return kp > 9 ? (size_t)BLA_W * (kp + BLA_PANEL_PAD) : 0; }

#if defined(__CUDACC__) && !defined(RBPE_EMU)
RBPE_DEV unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
RBPE_DEV void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
RBPE_DEV void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
RBPE_DEV void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// returns false after ~1 s without completion (a lost copy must not hang the device)
RBPE_DEV bool mbar_wait(unsigned long long *bar, unsigned parity) {
    const long long t0 = clock64();
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 2000000000LL) return false;
    }
    return true;
}
#endif

// Stage rows [r0, r0 + nrows) x columns [0, klen) of the row-major matrix Mx (leading dimension ld) into `panel`
// (row stride klen + BLA_PANEL_PAD).  All threads of the CTA call; on return the panel is readable by all of them.
// `bar` / `phase`: the CTA's mbarrier and its running parity (shared memory).  Generic-proxy stores to the source rows
// (the previous block column's step 3) are ordered before the async-proxy reads by the fence + barrier below.
RBPE_DEV bool stage_panel(double *panel, const double *Mx, int ld, int r0, int nrows, int klen, unsigned long long *bar, unsigned *phase) {
    const int tid = threadIdx.x;
    const int ps = klen + BLA_PANEL_PAD;
#if defined(__CUDACC__) && !defined(RBPE_EMU)
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    const unsigned par = *phase;
    if (tid < 32) {
        if (tid == 0) mbar_expect_tx(bar, (unsigned)(nrows * klen * 8));
        __syncwarp();
        if (tid < nrows) tma_bulk_g2s(panel + (size_t)tid * ps, Mx + (size_t)(r0 + tid) * ld, (unsigned)(klen * 8), bar);
    }
    const bool ok = mbar_wait(bar, par);
    __syncthreads();
    if (tid == 0) *phase = par ^ 1u;
    return ok;
#else
    __syncthreads();
    for (int i = tid; i < nrows * klen; i += blockDim.x) panel[(size_t)(i / klen) * ps + i % klen] = Mx[(size_t)(r0 + i / klen) * ld + i % klen];
    __syncthreads();
    (void)bar; (void)phase;
    return true;
#endif
}

// acc[jt] (8 x 8 tiles jt < ntj of the strip rows A0.., columns = rows B0 + 8 jt.. of Bm) += A0[.][0..klen) * Bm[.][0..klen)'
// tiles with jt > jt_max are skipped (upper triangle).  Warp-uniform arguments.
RBPE_DEV void strip_mma(double (&acc)[4][2], const double *A0, const double *B0, int ld, int ldb, int klen, int ntj, int jt_max) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double *ap = A0 + (size_t)g * ld + t;
    const double *bp = B0 + (size_t)g * ldb + t;
    const size_t tile = (size_t)8 * ldb;
    const bool u0 = ntj > 0 && jt_max >= 0, u1 = ntj > 1 && jt_max >= 1, u2 = ntj > 2 && jt_max >= 2, u3 = ntj > 3 && jt_max >= 3;
#pragma unroll 2
    for (int k0 = 0; k0 < klen; k0 += 4) {
        const double a = ap[k0];
        if (u0) dmma884(acc[0][0], acc[0][1], a, bp[k0]);
        if (u1) dmma884(acc[1][0], acc[1][1], a, bp[tile + k0]);
        if (u2) dmma884(acc[2][0], acc[2][1], a, bp[2 * tile + k0]);
        if (u3) dmma884(acc[3][0], acc[3][1], a, bp[3 * tile + k0]);
    }
}

// Static shared memory of the 32 x 32 diagonal-block routines (chol32_warp: Ls, chol32_cta: As | Rs | invs, chol32_cta_reg:
// colb | rowb | pivs).  One pool for all three -- a call uses exactly one of them -- instead of one array per routine:
// 17 KB instead of 34 KB of static shared memory, which the latency regime hands to the dynamic arena.
constexpr int BLA_POOL = 2 * 32 * 33 + 32;
RBPE_DEV double *bla_pool() {
    RBPE_STATIC_SMEM(double, pool, BLA_POOL);
    return pool;
}

// One warp: Cholesky of the wJ x wJ (wJ <= 32) diagonal block at Db (leading dimension ld, lower triangle) and its
// inverse X (32 x 32, row-major; identity beyond wJ).  Compact rolled code (the kernel is instruction-fetch bound):
// lane i owns row i of the trailing matrix in a register window that rotates by one column per step; the columns of L
// are exchanged through a shared-memory tile Ls (64 x 32, rows 32..63 zero so that the window can run past the block)
// as warp-wide broadcast loads.  No global memory traffic inside the two loops (a store followed by __syncwarp costs a
// full L2 round trip).  The inverse is a second pass of the same shape (lane c owns column c of L^-1).
// (Staging the window -- 32 / 24 / 16 / 8 live entries in four groups of steps -- saves 37 % of the load + FMA pairs but was
// 4 % slower in the same-box A/B: four unrolled bodies instead of one.)
// Returns false on a non-positive pivot.
RBPE_NOINLINE bool chol32_warp(double *Db, int ld, int wJ, double *X) {
    static_assert(64 * 32 + 32 <= BLA_POOL, "pool");
    double *Ls = bla_pool();
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    bool ok = true;
    double a[32];
    PROF_DECL;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        a[k] = (lane < wJ && k <= lane) ? Db[(size_t)lane * ld + k] : ((k == lane) ? 1.0 : 0.0);
        Ls[(32 + k) * 32 + lane] = 0.0;
    }
    PROF(10);
#pragma unroll 1
    for (int j = 0; j < wJ; j++) {
        double piv = __shfl_sync(FULL, a[0], j);
        if (!(piv > 0)) { ok = false; piv = 1.0; }
        const double inv = rsqrt(piv);
        const double a0 = a[0] * inv;                    // L[lane][j] for lane >= j
        Ls[lane * 32 + j] = (lane >= j) ? a0 : 0.0;
        if (lane == j) Ls[2048 + j] = inv;
        __syncwarp();
        const double *col = Ls + j * 33;                 // col[k * 32] = L[j + k][j]
        // loads staged 8 at a time ahead of their FMAs (at the 128-register cap the compiler otherwise pairs every
        // load with its FMA and exposes the shared-memory latency 31 times per column)
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 8) {
            double c[8];
#pragma unroll
            for (int k = 0; k < 8; k++) c[k] = col[(k0 + k + 1) * 32];   // row 32 + j of Ls is zero
#pragma unroll
            for (int k = 0; k < 8; k++) a[k0 + k] = ((k0 + k + 1 < 32) ? a[(k0 + k + 1) & 31] : 0.0) - a0 * c[k];
        }
    }
    __syncwarp();
#if defined(RBPE_PROFILE) && defined(__CUDACC__)
    if (threadIdx.x == 0 && wJ == 32) { unsigned long long d = (unsigned long long)(clock64() - prof_t0); atomicMin(&g_prof[13], d); atomicMax(&g_prof[14], d); atomicAdd(&g_prof[15], 1ull); }
#endif
    PROF(11);
    // inverse: a = residual of column `lane` of L X = I, window rotating with the row index
#pragma unroll
    for (int k = 0; k < 32; k++) a[k] = (k == lane) ? 1.0 : 0.0;
#pragma unroll 1
    for (int j = 0; j < 32; j++) {
        double xj = a[0];
        if (j < wJ) {   // warp-uniform
            xj *= Ls[2048 + j];
            const double *col = Ls + j * 33;
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 8) {
                double c[8];
#pragma unroll
                for (int k = 0; k < 8; k++) c[k] = col[(k0 + k + 1) * 32];
#pragma unroll
                for (int k = 0; k < 8; k++) a[k0 + k] = ((k0 + k + 1 < 32) ? a[(k0 + k + 1) & 31] : 0.0) - c[k] * xj;
            }
        } else {
#pragma unroll
            for (int k = 1; k < 32; k++) a[k - 1] = a[k];
            a[31] = 0.0;
        }
        X[j * BLA_W + lane] = xj;
    }
#pragma unroll 1
    for (int r = 0; r < wJ; r++)
        if (lane <= r) Db[(size_t)r * ld + lane] = Ls[r * 32 + lane];
    __syncwarp();
    PROF(12);
    return ok;
}

// CTA-wide version of chol32_warp: same right-looking elimination, same products, hence the same L and X bit for bit,
// but the rank-1 update of a column step is spread over ALL threads of the CTA (one element per thread and 32 x 32 / nt
// sub-steps) with one block barrier per column, instead of 32 dependent FMAs per lane of a single warp.  A lone warp runs
// its dependent chain at ~0.1 instructions per cycle: chol32_warp took ~96 k cycles per call and was 25-41 % of the whole
// joint-batch kernel (tools/gpu_joint.py with -DRBPE_PROFILE, r2); this version needs ~32 barriers.
// As (trailing matrix, then L unscaled by column) and Rs (right-hand side of L X = I) live in shared memory with a
// leading dimension of 33 (column reads are conflict free).  All threads of the CTA must call; returns false on a
// non-positive pivot.
RBPE_NOINLINE bool chol32_cta(double *Db, int ld, int wJ, double *X) {
    double *As = bla_pool(), *Rs = As + 32 * 33, *invs = Rs + 32 * 33;
    const int tid = threadIdx.x, nt = blockDim.x, c = tid & 31, w0 = tid >> 5, nw = nt >> 5;
    bool ok = true;
    for (int r = w0; r < 32; r += nw) {
        As[r * 33 + c] = (r < wJ && c <= r) ? Db[(size_t)r * ld + c] : ((r == c) ? 1.0 : 0.0);
        Rs[r * 33 + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < wJ; j++) {
        double piv = As[j * 33 + j];
        if (!(piv > 0)) { ok = false; piv = 1.0; }
        const double inv = rsqrt(piv);
        if (tid == 0) invs[j] = inv;
        // column j of L is As[.][j] * inv, row j of X is Rs[j][.] * inv; neither is written during this step
        const double lc = As[c * 33 + j] * inv, xj = Rs[j * 33 + c] * inv;
        if (c > j) {
            for (int r = w0; r < 32; r += nw)
                if (r > j) As[r * 33 + c] = As[r * 33 + c] - (As[r * 33 + j] * inv) * lc;
        } else {
            for (int r = w0; r < 32; r += nw)
                if (r > j) Rs[r * 33 + c] = Rs[r * 33 + c] - (As[r * 33 + j] * inv) * xj;
        }
        __syncthreads();
    }
    for (int r = w0; r < 32; r += nw) {
        if (r < wJ) {
            if (c <= r) Db[(size_t)r * ld + c] = As[r * 33 + c] * invs[c];
            X[r * BLA_W + c] = (c <= r) ? Rs[r * 33 + c] * invs[r] : 0.0;
        } else {
            X[r * BLA_W + c] = (r == c) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    return ok;
}

// Register-resident version of chol32_cta for CTAs of >= 8 warps (the joint-batch kernel's normal launch), organised for
// the LATENCY of the 32-step pivot chain -- a lone CTA runs every warp at ~6 cycles per instruction (dependent issue), so
// what counts is the number of instructions between two barriers:
//   * thread (warp w, lane c) keeps its <= 4 elements A[w + nw i][c] of the trailing matrix and of the right-hand side of
//     L X = I in registers; a column step publishes column j of A and row j of the right-hand side (64 doubles, double
//     buffered: ONE barrier per column), every thread reads the pivot, takes its reciprocal (MUFU.RCP64H + two Newton
//     steps) and updates its elements with two FMAs each;
//   * square-root free inside the chain (A[r][c] -= A[r][j] A[c][j] / pivot); the factors 1 / sqrt(pivot) are applied when
//     the results are written (one rsqrt per thread, off the chain).
// About 40 instructions per thread and column instead of ~250 (chol32_cta re-reads and re-writes shared memory and takes a
// square root per column): the diagonal blocks of one b = 4 factorisation went from 59 k to 29 k cycles per knot (clock64
// phase timers, -DRBPE_PROFILE), the column loop is 68 SASS instructions at 16 warps.
// Same contract as chol32_cta; results agree with it to rounding (not bit for bit).
RBPE_DEV double bla_rcp(double a) {  // 1/a to double rounding: hardware seed + two Newton steps (no FP64 division sequence)
#ifdef RBPE_EMU
    double r = (double)(1.0f / (float)a);
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
#endif
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
template <int RPT>   // rows per thread: 4 for 8..15 warps, 2 for 16 warps
RBPE_NOINLINE bool chol32_cta_reg(double *Db, int ld, int wJ, double *X) {
    double *colb = bla_pool(), *rowb = colb + 64, *pivs = rowb + 64;
    const int tid = threadIdx.x, nt = blockDim.x, c = tid & 31, w0 = tid >> 5, nw = nt >> 5;
    bool ok = true;
    double a[RPT], x[RPT];
    int rr[RPT];            // the thread's rows; -1 = none (a row index beyond 31)
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        const int r = w0 + nw * i;
        rr[i] = r < 32 ? r : -1;
        a[i] = (r < wJ && c <= r) ? Db[(size_t)r * ld + c] : ((r == c) ? 1.0 : 0.0);
        x[i] = (r == c) ? 1.0 : 0.0;
    }
    // publish column 0 / row 0
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        if (c == 0 && rr[i] >= 0) colb[rr[i]] = a[i];
        if (rr[i] == 0) rowb[c] = x[i];
    }
    __syncthreads();
    // The loop below is written for a small instruction count (every instruction of a lone CTA costs ~6 cycles): no branch per
    // row -- a row that is not updated in this step gets the multiplier 0 -- and the column / row selection folded into two
    // values per step (ac, xr), not two selects per element.
    double *cb = colb, *cn = colb + 32, *rb = rowb, *rn = rowb + 32;
#pragma unroll 1
    for (int j = 0; j < wJ; j++) {
        double piv = cb[j];
        if (!(piv > 0)) { ok = false; piv = 1.0; }
        if (tid == 0) pivs[j] = piv;
        const double ipiv = bla_rcp(piv);
        const bool right = c > j;
        const double ac = right ? cb[c] : 0.0;       // A[c][j] for the columns still being eliminated
        const double xr = right ? 0.0 : rb[c];       // row j of the right-hand side (columns <= j)
        const bool nextc = c == j + 1;
#pragma unroll
        for (int i = 0; i < RPT; i++) {
            const int r = rr[i];
            const bool below = r > j;
            const double f = below ? cb[r & 31] * ipiv : 0.0;   // A[r][j] / pivot
            a[i] = fma(-f, ac, a[i]);
            x[i] = fma(-f, xr, x[i]);
            if (nextc && below) cn[r] = a[i];        // next column, as soon as it is final
            if (r == j + 1) rn[c] = x[i];            // next row of the right-hand side
        }
        { double *t = cb; cb = cn; cn = t; t = rb; rb = rn; rn = t; }
        __syncthreads();
    }
    // (Letting the warps whose rows are all finished skip the column step was tried: no gain -- the step is bound by its
    // dependent chain pivot -> reciprocal -> multiplier -> update -> publish, not by issue slots.)
    // L[r][c] = A_c[r][c] / sqrt(pivot_c) (column c as it stood at step c), X[r][c] = rhs_r[r][c] / sqrt(pivot_r)
    const double ic = (c < wJ) ? rsqrt(pivs[c]) : 1.0;
#pragma unroll
    for (int i = 0; i < RPT; i++) {
        const int r = rr[i];
        if (r < 0) continue;
        if (r < wJ) {
            if (c <= r) Db[(size_t)r * ld + c] = a[i] * ic;
            X[r * BLA_W + c] = (c <= r) ? x[i] * rsqrt(pivs[r]) : 0.0;
        } else {
            X[r * BLA_W + c] = (r == c) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    return ok;
}

// Factor the tall matrix [D; O] (see the header).  Pm = L_{t,t-1} of the previous knot (kp x kp) or null.
// Linv receives the inverses of the 32 x 32 diagonal blocks of L ([bla_ninv][32*32], row-major, lower).
// `flag`: one double of shared / global scratch for the verdict of warp 0.  All threads of the CTA must call.
RBPE_NOINLINE bool chol_tall(int kp, double *D, double *O, const double *Pm, double *Linv, double *flag, double *panel = nullptr) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, nw = nt >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int ntile = kp >> 3;
    PROF_DECL;
    RBPE_STATIC_SMEM(unsigned long long, tma_bar, 1);
    RBPE_STATIC_SMEM(unsigned, tma_phase, 1);
    if (tid == 0) {
        *flag = 0.0;
#if defined(__CUDACC__) && !defined(RBPE_EMU)
        if (panel) { mbar_init(tma_bar, 1); tma_phase[0] = 0; }
#endif
    }
    if (panel) __syncthreads();
    for (int j0 = 0; j0 < kp; j0 += BLA_W) {
        const int wJ = (kp - j0 < BLA_W) ? kp - j0 : BLA_W, ntJ = wJ >> 3;
        double *X = Linv + (size_t)(j0 / BLA_W) * BLA_W * BLA_W;
        // the 32 rows of L every strip of this block column multiplies with: staged once, by TMA, into shared memory
        const double *Bp = D + (size_t)j0 * kp;
        int ldb = kp;
#if defined(__CUDACC__) && !defined(RBPE_EMU)
        const bool d_global = __isGlobal(D);   // small blocks can live in the shared-memory part of the arena: TMA reads global only
#else
        const bool d_global = true;
#endif
        if (panel && j0 > 0 && d_global) {
            if (!stage_panel(panel, D, kp, j0, wJ, j0, tma_bar, tma_phase) && tid == 0) *flag = 1.0;
            Bp = panel; ldb = j0 + BLA_PANEL_PAD;
        }
        // ---- 1. left-looking update of the block column: C -= L[., 0..j0) L[J, 0..j0)'  (+ the previous knot's block).
        // 1a: the strips of the diagonal block (all warps), barrier; 1b: the remaining strips by warps 1.., overlapped
        // with 2: warp 0 factors and inverts the diagonal block (a serial 32-step chain).
        const int nsD = (kp - j0) >> 3, nsO = O ? ntile : 0;
        const bool upd = (j0 > 0 || Pm);
        // Two ways to do the 32 x 32 diagonal block, chosen by size (same-box A/B, tools/gpu_joint.py):
        //   small blocks (kp <= 64: the launch default b = 4) -- the CTA-wide chol32_cta, then all warps do the strips below
        //     (+13..22 % throughput, single-mission latency 102 -> 86 ms: little strip work exists to hide a serial warp);
        //   large blocks (b >= 16) -- warp 0 runs the serial chol32_warp while warps 1.. update the strips below it
        //     (the CTA-wide version was 4 % slower there: with two CTAs per SM the idle warps of one CTA are filled by the other).
        const bool cta_diag = kp <= 64 || nw < 2;
        for (int pass = 0; pass < 2; pass++) {
            const int s_begin = pass == 0 ? 0 : ntJ, s_end = pass == 0 ? ntJ : nsD + nsO;
            const int w = (pass == 0 || cta_diag) ? warp : warp - 1, wn = (pass == 0 || cta_diag) ? nw : nw - 1;
            if (upd && w >= 0)
                for (int s = s_begin + w; s < s_end; s += wn) {
                    const bool isO = s >= nsD;
                    const int i0 = isO ? (s - nsD) * 8 : j0 + s * 8;
                    double *C0 = (isO ? O : D) + (size_t)i0 * kp;
                    const int jt_max = isO ? 3 : (i0 - j0) >> 3;
                    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                    if (j0 > 0) strip_mma(acc, C0, Bp, kp, ldb, j0, ntJ, jt_max);
                    if (!isO && Pm) strip_mma(acc, Pm + (size_t)i0 * kp, Pm + (size_t)j0 * kp, kp, kp, kp, ntJ, jt_max);
#pragma unroll
                    for (int jt = 0; jt < 4; jt++)
                        if (jt < ntJ && jt <= jt_max) {
                            double *c = C0 + (size_t)g * kp + j0 + 8 * jt + 2 * t4;
                            c[0] -= acc[jt][0];
                            c[1] -= acc[jt][1];
                        }
                }
            if (pass == 0) {
                if (upd) __syncthreads();
                PROF(6);
                // ---- 2. diagonal block: factor + invert ----
                if (cta_diag) {
                    double *Dj = D + (size_t)j0 * kp + j0;
                    const bool ok = (nw >= 16) ? chol32_cta_reg<2>(Dj, kp, wJ, X) : ((nw >= 8) ? chol32_cta_reg<4>(Dj, kp, wJ, X) : chol32_cta(Dj, kp, wJ, X));
                    if (!ok && tid == 0) *flag = 1.0;
                    PROF(9);
                } else if (warp == 0) {
                    const bool ok = chol32_warp(D + (size_t)j0 * kp + j0, kp, wJ, X);
                    if (!ok && lane == 0) *flag = 1.0;
                }
            }
        }
        __syncthreads();
        PROF(7);
        // ---- 3. rows below: L21 = A21 L11^-T = A21 X'  (DMMA, k = wJ) ----
        {
            const int r0 = j0 + wJ, nsD = (kp - r0) >> 3, nsO = O ? ntile : 0;
            for (int s = warp; s < nsD + nsO; s += nw) {
                const bool isO = s >= nsD;
                const int i0 = isO ? (s - nsD) * 8 : r0 + s * 8;
                double *C0 = (isO ? O : D) + (size_t)i0 * kp + j0;
                double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
#ifdef RBPE_BLA_STABLE
                // experiment (off by default): one step of refinement of L21 = A21 L11^-T:  R = A21 - L21 L11',  L21 += R X'
                double a21[4][2];
#pragma unroll
                for (int jt = 0; jt < 4; jt++)
                    if (jt < ntJ) { const double *c = C0 + (size_t)g * kp + 8 * jt + 2 * t4; a21[jt][0] = c[0]; a21[jt][1] = c[1]; }
#endif
                strip_mma(acc, C0, X, kp, BLA_W, wJ, ntJ, 3);
#pragma unroll
                for (int jt = 0; jt < 4; jt++)
                    if (jt < ntJ) {
                        double *c = C0 + (size_t)g * kp + 8 * jt + 2 * t4;
                        c[0] = acc[jt][0];
                        c[1] = acc[jt][1];
                    }
#ifdef RBPE_BLA_STABLE
                {
                    __syncwarp();
                    double r[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                    // L21 L11': B[k][n] = L11[n][k] (lower triangular: k <= n), masked because the strict upper triangle of the
                    // diagonal block holds stale values
                    const double *ap = C0 + (size_t)g * kp + t4;
                    for (int k0 = 0; k0 < wJ; k0 += 4) {
                        const double a = ap[k0];
#pragma unroll
                        for (int jt = 0; jt < 4; jt++)
                            if (jt < ntJ) {
                                const int n = 8 * jt + g, k = k0 + t4;
                                const double b = (k <= n) ? D[(size_t)(j0 + n) * kp + j0 + k] : 0.0;
                                dmma884(r[jt][0], r[jt][1], a, b);
                            }
                    }
                    __syncwarp();
#pragma unroll
                    for (int jt = 0; jt < 4; jt++)
                        if (jt < ntJ) {   // the strip now holds the residual R; the first estimate stays in acc
                            double *c = C0 + (size_t)g * kp + 8 * jt + 2 * t4;
                            c[0] = a21[jt][0] - r[jt][0];
                            c[1] = a21[jt][1] - r[jt][1];
                        }
                    __syncwarp();
                    double d[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                    strip_mma(d, C0, X, kp, BLA_W, wJ, ntJ, 3);
                    __syncwarp();
#pragma unroll
                    for (int jt = 0; jt < 4; jt++)
                        if (jt < ntJ) {
                            double *c = C0 + (size_t)g * kp + 8 * jt + 2 * t4;
                            c[0] = acc[jt][0] + d[jt][0];
                            c[1] = acc[jt][1] + d[jt][1];
                        }
                }
#endif
            }
        }
        __syncthreads();
        PROF(8);
    }
    // every thread reads the verdict BEFORE thread 0 of the next call may reset the flag (compute-sanitizer racecheck, r2:
    // a slow thread could otherwise miss a failure and leave the CTA's control flow divergent)
    const bool ok_all = (*flag == 0.0);
    __syncthreads();
    return ok_all;
}

// block tridiagonal Cholesky: Dall (nblk diagonal blocks, lower), Oall (nblk-1 blocks (t+1, t)), ld = kp
RBPE_DEV bool factor_bt_blk(int nblk, int kp, double *Dall, double *Oall, double *Linv, double *flag, double *panel = nullptr) {
    const size_t kk = (size_t)kp * kp, li = (size_t)bla_ninv(kp) * BLA_W * BLA_W;
    bool ok = true;
    for (int t = 0; t < nblk; t++)
        ok = chol_tall(kp, Dall + t * kk, (t < nblk - 1) ? Oall + t * kk : nullptr, (t > 0) ? Oall + (t - 1) * kk : nullptr,
                       Linv + t * li, flag, panel) && ok;
    return ok;
}

// g (nblk blocks of kb, stride kb) <- (L L')^-1 g.  w: nblk*kp work doubles, y: kp + 32 work doubles.
RBPE_NOINLINE void solve_bt_blk(int nblk, int kb, int kp, const double *Dall, const double *Oall, const double *Linv, double *g,
                           double *w, double *y) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, nw = nt >> 5, lane = tid & 31;
    const size_t kk = (size_t)kp * kp, li = (size_t)bla_ninv(kp) * BLA_W * BLA_W;
    for (int i = tid; i < nblk * kp; i += nt) {
        int t = i / kp, r = i % kp;
        w[i] = (r < kb) ? g[t * kb + r] : 0.0;
    }
    __syncthreads();
    // ---- forward: L w = g ----
    for (int t = 0; t < nblk; t++) {
        const double *L = Dall + t * kk, *Li = Linv + t * li;
        double *wt = w + (size_t)t * kp;
        if (t > 0) {   // w_t -= L_{t,t-1} w_{t-1}: one warp per row
            const double *P = Oall + (t - 1) * kk, *wp = w + (size_t)(t - 1) * kp;
            for (int r = warp; r < kp; r += nw) {
                double sm = 0;
                for (int k = lane; k < kp; k += 32) sm += P[(size_t)r * kp + k] * wp[k];
                for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                if (lane == 0) wt[r] -= sm;
            }
            __syncthreads();
        }
        for (int j0 = 0; j0 < kp; j0 += BLA_W) {
            const int wJ = (kp - j0 < BLA_W) ? kp - j0 : BLA_W;
            const double *X = Li + (size_t)(j0 / BLA_W) * BLA_W * BLA_W;
#ifdef RBPE_BLA_STABLE
            // experiment (off by default): substitution inside the 32-block by one warp instead of the product with the
            // explicit inverse -- componentwise backward stable; X only supplies 1 / L_jj
            if (warp == 0) {
                double wv = (lane < wJ) ? wt[j0 + lane] : 0.0;
                for (int j = 0; j < wJ; j++) {
                    const double yj = __shfl_sync(0xffffffffu, wv, j) * X[j * BLA_W + j];
                    if (lane == j) wv = yj;
                    else if (lane > j && lane < wJ) wv -= L[(size_t)(j0 + lane) * kp + j0 + j] * yj;
                }
                y[lane] = wv;
            }
#else
            for (int r = warp; r < wJ; r += nw) {   // y = Linv_J w_J
                double sm = (lane <= r) ? X[r * BLA_W + lane] * wt[j0 + lane] : 0.0;
                for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                if (lane == 0) y[r] = sm;
            }
#endif
            __syncthreads();
            for (int r = j0 + tid; r < kp; r += nt) {
                if (r < j0 + wJ) { wt[r] = y[r - j0]; continue; }
                const double *Lr = L + (size_t)r * kp + j0;
                double s0 = 0, s1 = 0;
                for (int c = 0; c < wJ; c += 2) { s0 += Lr[c] * y[c]; s1 += Lr[c + 1] * y[c + 1]; }
                wt[r] -= s0 + s1;
            }
            __syncthreads();
        }
    }
    // ---- backward: L' u = w ----
    for (int t = nblk - 1; t >= 0; t--) {
        const double *L = Dall + t * kk, *Li = Linv + t * li;
        double *wt = w + (size_t)t * kp;
        if (t < nblk - 1) {   // w_t -= L_{t+1,t}' u_{t+1}: one thread per column, coalesced over the rows of L_{t+1,t}
            const double *P = Oall + t * kk, *un = w + (size_t)(t + 1) * kp;
            for (int c = tid; c < kp; c += nt) {
                double s0 = 0, s1 = 0;
                for (int k = 0; k < kp; k += 2) { s0 += P[(size_t)k * kp + c] * un[k]; s1 += P[(size_t)(k + 1) * kp + c] * un[k + 1]; }
                wt[c] -= s0 + s1;
            }
            __syncthreads();
        }
        for (int j0 = ((kp - 1) / BLA_W) * BLA_W; j0 >= 0; j0 -= BLA_W) {
            const int wJ = (kp - j0 < BLA_W) ? kp - j0 : BLA_W;
            const double *X = Li + (size_t)(j0 / BLA_W) * BLA_W * BLA_W;
#ifdef RBPE_BLA_STABLE
            if (warp == 0) {   // L_JJ' y = w_J by back substitution
                double wv = (lane < wJ) ? wt[j0 + lane] : 0.0;
                for (int j = wJ - 1; j >= 0; j--) {
                    const double yj = __shfl_sync(0xffffffffu, wv, j) * X[j * BLA_W + j];
                    if (lane == j) wv = yj;
                    else if (lane < j) wv -= L[(size_t)(j0 + j) * kp + j0 + lane] * yj;
                }
                y[lane] = wv;
            }
#else
            if (warp == 0) {   // y = Linv_J' w_J
                double sm = 0;
                if (lane < wJ)
                    for (int r = lane; r < wJ; r++) sm += X[r * BLA_W + lane] * wt[j0 + r];
                y[lane] = sm;
            }
#endif
            __syncthreads();
            for (int r = tid; r < j0 + wJ; r += nt) {
                if (r >= j0) { wt[r] = y[r - j0]; continue; }
                double s0 = 0, s1 = 0;
                for (int c = 0; c < wJ; c += 2) { s0 += L[(size_t)(j0 + c) * kp + r] * y[c]; s1 += L[(size_t)(j0 + c + 1) * kp + r] * y[c + 1]; }
                wt[r] -= s0 + s1;
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < nblk * kb; i += nt) {
        int t = i / kb, r = i % kb;
        g[i] = w[(size_t)t * kp + r];
    }
    __syncthreads();
}

// ---- small joint batches (kp <= 64: b <= 7, the launch default b = 4) ------------------------------------------------
// Same substitutions as solve_bt_blk (products with the inverted 32 x 32 diagonal blocks, same operands), organised for
// latency: every step is ONE matrix-vector product spread over the whole CTA -- eight threads per output, each adding
// every eighth term, three shuffle rounds -- followed by one barrier; results go to a second vector instead of being
// copied back, so a knot costs 2 (kp <= 32) or 4 steps per direction.  A lone CTA pays ~6 cycles per instruction and warp:
// the general routine above (warp-per-row products with 5-round butterflies, 32-term serial loops, a copy and two
// barriers per block) took 48 k cycles per solve at b = 4 on a lone CTA, this one 32 k; with many CTAs in flight the
// smaller instruction count gives +12..15 % throughput at b = 4.
// out[o] = (base ? base[o] : 0) - / + sum_k A[o * so + k * sk] v[k],  o < nout, k < nin <= 64.  No barrier inside.
RBPE_DEV void mv8(int nout, int nin, const double *A, int so, int sk, const double *v, const double *base, double *out, bool neg) {
    const int tid = threadIdx.x, nt = blockDim.x, part = tid & 7;
    const int nmine = (nin - part + 7) >> 3;          // terms of this thread: k = part, part + 8, ...
    const int step = 8 * sk;
    for (int o0 = 0; o0 < nout; o0 += nt >> 3) {   // (uniform trip count: whole warps stay in the shuffles)
        const int o = o0 + (tid >> 3);
        double s0 = 0, s1 = 0;
        if (o < nout) {
            const double *a = A + (o * so + part * sk), *vp = v + part;   // 32-bit index arithmetic, running pointers
            int i = 0;
#pragma unroll 1
            for (; i + 2 <= nmine; i += 2, a += 2 * step, vp += 16) { s0 = fma(a[0], vp[0], s0); s1 = fma(a[step], vp[8], s1); }
            if (i < nmine) s0 = fma(a[0], vp[0], s0);
        }
        double sm = s0 + s1;
        sm += __shfl_xor_sync(0xffffffffu, sm, 1);
        sm += __shfl_xor_sync(0xffffffffu, sm, 2);
        sm += __shfl_xor_sync(0xffffffffu, sm, 4);
        if (o < nout && part == 0) {
            const double b = base ? base[o] : 0.0;
            out[o] = neg ? b - sm : b + sm;
        }
    }
}
// g (nblk blocks of kb, stride kb) <- (L L')^-1 g.  w, u: nblk*kp work doubles each.  All threads of the CTA must call.
RBPE_NOINLINE void solve_bt_small(int nblk, int kb, int kp, const double *Dall, const double *Oall, const double *Linv, double *g,
                                  double *w, double *u) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const size_t kk = (size_t)kp * kp, li = (size_t)bla_ninv(kp) * BLA_W * BLA_W;
    for (int i = tid; i < nblk * kp; i += nt) {
        int t = i / kp, r = i % kp;
        w[i] = (r < kb) ? g[t * kb + r] : 0.0;
    }
    __syncthreads();
    // ---- forward: L u = g (right-hand side in w, solution in u) ----
    for (int t = 0; t < nblk; t++) {
        const double *L = Dall + t * kk, *Li = Linv + t * li;
        double *wt = w + (size_t)t * kp, *ut = u + (size_t)t * kp;
        if (t > 0) {   // w_t -= L_{t,t-1} u_{t-1}
            mv8(kp, kp, Oall + (t - 1) * kk, kp, 1, u + (size_t)(t - 1) * kp, wt, wt, true);
            __syncthreads();
        }
        for (int j0 = 0; j0 < kp; j0 += BLA_W) {
            const int wJ = (kp - j0 < BLA_W) ? kp - j0 : BLA_W;
            const double *X = Li + (size_t)(j0 / BLA_W) * BLA_W * BLA_W;
            mv8(wJ, wJ, X, BLA_W, 1, wt + j0, nullptr, ut + j0, false);          // u_J = Linv_J w_J (zeros above the diagonal)
            __syncthreads();
            const int below = kp - j0 - wJ;
            if (below > 0) {                                                     // w_r -= L[r][J] u_J for the rows below
                mv8(below, wJ, L + (size_t)(j0 + wJ) * kp + j0, kp, 1, ut + j0, wt + j0 + wJ, wt + j0 + wJ, true);
                __syncthreads();
            }
        }
    }
    // ---- backward: L' x = u (right-hand side in u, solution in w) ----
    for (int t = nblk - 1; t >= 0; t--) {
        const double *L = Dall + t * kk, *Li = Linv + t * li;
        double *wt = w + (size_t)t * kp, *ut = u + (size_t)t * kp;
        if (t < nblk - 1) {   // u_t -= L_{t+1,t}' x_{t+1}
            mv8(kp, kp, Oall + t * kk, 1, kp, w + (size_t)(t + 1) * kp, ut, ut, true);
            __syncthreads();
        }
        for (int j0 = ((kp - 1) / BLA_W) * BLA_W; j0 >= 0; j0 -= BLA_W) {
            const int wJ = (kp - j0 < BLA_W) ? kp - j0 : BLA_W;
            const double *X = Li + (size_t)(j0 / BLA_W) * BLA_W * BLA_W;
            mv8(wJ, wJ, X, 1, BLA_W, ut + j0, nullptr, wt + j0, false);          // x_J = Linv_J' u_J
            __syncthreads();
            if (j0 > 0) {                                                        // u_r -= L[J][r]' x_J for the rows above
                mv8(j0, wJ, L + (size_t)j0 * kp, 1, kp, wt + j0, ut, ut, true);
                __syncthreads();
            }
        }
    }
    for (int i = tid; i < nblk * kb; i += nt) {
        int t = i / kb, r = i % kb;
        g[i] = w[(size_t)t * kp + r];
    }
    __syncthreads();
}

// ---- one-agent batches (9 x 9 blocks): the whole block tridiagonal system handled by one warp ---------------------
// The one-agent kernel is instruction-issue bound and, after the row presolve, spends most of its instructions here
// (ncu r1), so the routines are organised for few instructions, not few flops:
//   * the factorisation produces, per knot, the INVERSE of the diagonal factor L_tt (in place of D_t) together with
//     L_{t+1,t} (in place of O_t): lanes 0..8 hold the rows of D_t, lanes 9..17 the rows of O_t in a register window
//     that rotates by one column per step (compact rolled code); column j of L is broadcast through `cb` (shared
//     memory) instead of shuffles, and the same broadcast drives the forward substitution L X = I (lane c owns column c
//     of X), so the inverse costs 8 more FMAs per column;
//   * the two solves of an interior-point iteration are then 9 x 9 matrix-vector products (one lane per row, no
//     shuffles, no 9-step substitution chains).
// D_t -= L_{t,t-1} L_{t,t-1}' runs on 27 lanes (3 entries each).  cb: >= 32 doubles of shared memory.
template <int TAG>   // TAG: one private copy per kernel (a copy shared by two kernels changed the one-agent kernel's layout: -12%)
RBPE_NOINLINE bool factor_bt9v(int nblk, double *Dall, double *Oall, double *cb) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool isD = lane < 9, isO = lane >= 9 && lane < 18;
    const int row = isD ? lane : (isO ? lane - 9 : 0);
    bool ok = true;
#pragma unroll 1
    for (int t = 0; t < nblk; t++) {
        double *D = Dall + t * 81, *O = Oall + t * 81;
        const bool hasO = t < nblk - 1;
        if (t > 0) {
            const double *P = Oall + (t - 1) * 81;
            if (lane < 27) {
                const int r = lane / 3, c0 = 3 * (lane % 3);
                double s0 = D[r * 9 + c0], s1 = D[r * 9 + c0 + 1], s2 = D[r * 9 + c0 + 2];
#pragma unroll
                for (int k = 0; k < 9; k++) {
                    const double pk = P[r * 9 + k];
                    s0 -= pk * P[c0 * 9 + k]; s1 -= pk * P[(c0 + 1) * 9 + k]; s2 -= pk * P[(c0 + 2) * 9 + k];
                }
                D[r * 9 + c0] = s0; D[r * 9 + c0 + 1] = s1; D[r * 9 + c0 + 2] = s2;
            }
            __syncwarp();
        }
        double a[9], v[9];
#pragma unroll
        for (int c = 0; c < 9; c++) {
            a[c] = isD ? D[row * 9 + c] : ((isO && hasO) ? O[row * 9 + c] : 0.0);
            v[c] = (c == lane) ? 1.0 : 0.0;
        }
        __syncwarp();   // rows are in registers: D may now be overwritten by the inverse
#pragma unroll 1
        for (int j = 0; j < 9; j++) {
            double piv = __shfl_sync(FULL, a[0], j);
            if (!(piv > 0)) { ok = false; piv = 1.0; }
            const double inv = rsqrt(piv);
            const double a0 = a[0] * inv;                  // L_tt[lane][j] (lanes j..8), L_{t+1,t}[lane-9][j] (lanes 9..17)
            const double xj = v[0] * inv;                  // X[j][lane]
            cb[(lane - j) & 31] = (isD && lane >= j) ? a0 : 0.0;   // cb[k] = L_tt[j+k][j], zero beyond the block
            if (isO && hasO) O[row * 9 + j] = a0;
            if (isD) D[j * 9 + lane] = xj;                 // zero above the diagonal
            __syncwarp();
#pragma unroll
            for (int k = 1; k < 9; k++) {
                const double l = cb[k];
                a[k - 1] = a[k] - a0 * l;
                v[k - 1] = v[k] - l * xj;
            }
            a[8] = 0.0; v[8] = 0.0;
            __syncwarp();
        }
    }
    return ok;
}

// Block tridiagonal factorisation by one warp, latency-oriented variant (written for the latency kernel, where the warp
// is alone on its scheduler): same outputs as factor_bt9v (the INVERSE of
// the diagonal factor block in place of D_t, L_{t+1,t} in place of O_t), organised for the LATENCY of the 9-step pivot
// chain instead of the instruction count:
//   * square-root free inside the chain: column j is eliminated with u = a_j / pivot (MUFU.RCP64H + two Newton steps);
//     the factors 1 / sqrt(pivot) are applied afterwards, for all nine columns at once (one rsqrt latency per knot
//     instead of nine in series);
//   * the next pivot is computed by its own lane from registers and broadcast BEFORE the shared-memory exchange of the
//     column, so the exchange overlaps the reciprocal of the next step;
//   * three lane groups -- rows of D_t (0..8), rows of O_t (9..17), columns of the inverse (18..26) -- run the same
//     update w[k-1] = w[k] - w[0] u[k] on one register window (8 FMAs per lane and column).
// cb: >= 32 doubles, pv: >= 9 doubles of shared memory.
template <int TAG>   // one private copy per kernel, as factor_bt9v
RBPE_NOINLINE bool factor_bt9l(int nblk, double *Dall, double *Oall, double *cb, double *pv) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool isD = lane < 9, isO = lane >= 9 && lane < 18, isX = lane >= 18 && lane < 27;
    const int row = isD ? lane : (isO ? lane - 9 : (isX ? lane - 18 : 0));
    bool ok = true;
#pragma unroll 1
    for (int t = 0; t < nblk; t++) {
        double *D = Dall + t * 81, *O = Oall + t * 81;
        const bool hasO = t < nblk - 1;
        if (t > 0) {   // D_t -= L_{t,t-1} L_{t,t-1}'
            const double *P = Oall + (t - 1) * 81;
            if (lane < 27) {
                const int r = lane / 3, c0 = 3 * (lane % 3);
                double s0 = D[r * 9 + c0], s1 = D[r * 9 + c0 + 1], s2 = D[r * 9 + c0 + 2];
                // rolled in the throughput kernel (TAG 1: instruction-fetch bound, text size counts more than 24 loop
                // instructions: +1.3 %), unrolled in the latency kernel (TAG 2: a lone warp pays per instruction: 2 %)
#pragma unroll (TAG == 1 ? 1 : 9)
                for (int k = 0; k < 9; k++) {
                    const double pk = P[r * 9 + k];
                    s0 -= pk * P[c0 * 9 + k]; s1 -= pk * P[(c0 + 1) * 9 + k]; s2 -= pk * P[(c0 + 2) * 9 + k];
                }
                D[r * 9 + c0] = s0; D[r * 9 + c0 + 1] = s1; D[r * 9 + c0 + 2] = s2;
            }
            __syncwarp();
        }
        double w[9];
        {   // one set of loads through a selected pointer (three-way branches around the loads cost ~40 instructions per knot)
            const bool ld = isD || (isO && hasO);
            const double *src = ((isO && hasO) ? O : D) + row * 9;
#pragma unroll
            for (int c = 0; c < 9; c++) {
                const double v = src[c];
                w[c] = ld ? v : ((isX && c == row) ? 1.0 : 0.0);
            }
        }
        __syncwarp();   // rows are in registers: D may now be overwritten by the inverse
        double piv = __shfl_sync(FULL, w[0], 0);
#pragma unroll 1
        for (int j = 0; j < 9; j++) {
            if (!(piv > 0)) { ok = false; piv = 1.0; }
            const double w0 = w[0];
            const double u = w0 * bla_rcp(piv);                            // a_j / pivot
            const double pn = __shfl_sync(FULL, w[1] - w0 * u, (j + 1) & 31);   // next pivot, from the lane that owns it
            cb[(lane - j) & 31] = (isD && lane >= j) ? u : 0.0;           // cb[k] = u of row j + k, zero beyond the block
            if (lane == j) pv[j] = piv;
            if (isO && hasO) O[row * 9 + j] = w0;                         // unscaled: times 1/sqrt(pivot_j) below
            if (isX) D[j * 9 + row] = w0;
            __syncwarp();
#pragma unroll
            for (int k = 1; k < 9; k++) w[k - 1] = w[k] - w0 * cb[k];
            w[8] = 0.0;
            __syncwarp();
            piv = pn;
        }
        if (lane < 9) pv[lane] = rsqrt(pv[lane]);
        __syncwarp();
        if (lane < 27) {   // 27 lanes x 3 entries, as in the update above: no index division in a loop
            const int r = lane / 3, c0 = 3 * (lane % 3);
            const double sr = pv[r], s0 = pv[c0], s1 = pv[c0 + 1], s2 = pv[c0 + 2];
            double *d = D + r * 9 + c0;
            d[0] *= sr; d[1] *= sr; d[2] *= sr;                  // row r of the inverse
            if (hasO) {
                double *o = O + r * 9 + c0;
                o[0] *= s0; o[1] *= s1; o[2] *= s2;              // columns of L_{t+1,t}
            }
        }
        __syncwarp();
    }
    return ok;
}

// g <- (L L')^-1 g with Dall = inverses of the diagonal factor blocks, Oall = L_{t+1,t} (factor_bt9v / factor_bt9l).
// (Rolling the four 9-term loops for the throughput kernel, as in factor_bt9l's update, measured -2 %: they run 16 x per
// iteration, the rolled loop overhead outweighs the smaller text.)
template <int TAG>
RBPE_NOINLINE void solve_bt9v(int nblk, const double *Dall, const double *Oall, double *g, double *cb) {
    const int lane = threadIdx.x & 31;
    const bool act = lane < 9;
    const int row = act ? lane : 0;
#pragma unroll 1
    for (int t = 0; t < nblk; t++) {   // w_t = X_t (g_t - L_{t,t-1} w_{t-1})
        const double *X = Dall + t * 81;
        double s = g[t * 9 + row];
        if (t > 0) {
            const double *P = Oall + (t - 1) * 81 + row * 9, *wp = g + (t - 1) * 9;
#pragma unroll
            for (int k = 0; k < 9; k++) s -= P[k] * wp[k];
        }
        if (act) cb[row] = s;
        __syncwarp();
        double w = 0;
#pragma unroll
        for (int c = 0; c < 9; c++) w += X[row * 9 + c] * cb[c];
        if (act) g[t * 9 + row] = w;
        __syncwarp();
    }
#pragma unroll 1
    for (int t = nblk - 1; t >= 0; t--) {   // u_t = X_t' (w_t - L_{t+1,t}' u_{t+1})
        const double *X = Dall + t * 81;
        double s = g[t * 9 + row];
        if (t < nblk - 1) {
            const double *P = Oall + t * 81 + row, *un = g + (t + 1) * 9;
#pragma unroll
            for (int k = 0; k < 9; k++) s -= P[k * 9] * un[k];
        }
        if (act) cb[row] = s;
        __syncwarp();
        double u = 0;
#pragma unroll
        for (int r = 0; r < 9; r++) u += X[r * 9 + row] * cb[r];
        if (act) g[t * 9 + row] = u;
        __syncwarp();
    }
}

#endif
}  // namespace rbpe
