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
//               8 x 32 strip, operands straight from L1/L2; the 8-column panels inside are factored in registers.
//   tri_inverse the 32 x 32 diagonal blocks of L_tt are inverted once per factorisation so that the two solves of an
//               interior-point iteration are matrix-vector products (2 barriers per 32 unknowns instead of 64).
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

// acc[jt] (8 x 8 tiles jt < ntj of the strip rows A0.., columns = rows B0 + 8 jt.. of Bm) += A0[.][0..klen) * Bm[.][0..klen)'
// tiles with jt > jt_max are skipped (upper triangle).  Warp-uniform arguments.
RBPE_DEV void strip_mma(double (&acc)[4][2], const double *A0, const double *B0, int ld, int klen, int ntj, int jt_max) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double *ap = A0 + (size_t)g * ld + t;
    const double *bp = B0 + (size_t)g * ld + t;
    const size_t tile = (size_t)8 * ld;
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

// Cholesky of the 8 x 8 tile held packed (lower, row-major: l[r(r+1)/2 + c]) in registers; returns false on a
// non-positive pivot.  On exit l holds L and di[c] = 1 / L[c][c].
RBPE_DEV bool chol8_reg(double (&l)[36], double (&di)[8]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double d = l[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= l[j * (j + 1) / 2 + k] * l[j * (j + 1) / 2 + k];
        if (!(d > 0)) { ok = false; d = 1.0; }
        const double inv = rsqrt(d);
        di[j] = inv;
        l[j * (j + 1) / 2 + j] = d * inv;
#pragma unroll
        for (int r = j + 1; r < 8; r++) {
            double v = l[r * (r + 1) / 2 + j];
#pragma unroll
            for (int k = 0; k < j; k++) v -= l[r * (r + 1) / 2 + k] * l[j * (j + 1) / 2 + k];
            l[r * (r + 1) / 2 + j] = v * inv;
        }
    }
    return ok;
}

// Factor the tall matrix [D; O] (see the header).  Pm = L_{t,t-1} of the previous knot (kp x kp) or null.
// Linv receives the inverses of the 32 x 32 diagonal blocks of L ([bla_ninv][32*32], row-major, lower).
// Every thread factors the same diagonal tiles, so the verdict is uniform.  All threads of the CTA must call.
RBPE_DEV bool chol_tall(int kp, double *D, double *O, const double *Pm, double *Linv) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, nw = nt >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int ntile = kp >> 3;
    bool ok = true;
    for (int j0 = 0; j0 < kp; j0 += BLA_W) {
        const int wJ = (kp - j0 < BLA_W) ? kp - j0 : BLA_W, ntJ = wJ >> 3;
        // ---- 1. left-looking update of the block column ----
        if (j0 > 0 || Pm) {
            const int nsD = (kp - j0) >> 3, nsO = O ? ntile : 0;
            for (int s = warp; s < nsD + nsO; s += nw) {
                const bool isO = s >= nsD;
                const int i0 = isO ? (s - nsD) * 8 : j0 + s * 8;
                double *C0 = (isO ? O : D) + (size_t)i0 * kp;
                const int jt_max = isO ? 3 : (i0 - j0) >> 3;
                double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                if (j0 > 0) strip_mma(acc, C0, D + (size_t)j0 * kp, kp, j0, ntJ, jt_max);
                if (!isO && Pm) strip_mma(acc, Pm + (size_t)i0 * kp, Pm + (size_t)j0 * kp, kp, kp, ntJ, jt_max);
#pragma unroll
                for (int jt = 0; jt < 4; jt++)
                    if (jt < ntJ && jt <= jt_max) {
                        double *c = C0 + (size_t)g * kp + j0 + 8 * jt + 2 * t4;
                        c[0] -= acc[jt][0];
                        c[1] -= acc[jt][1];
                    }
            }
            __syncthreads();
        }
        // ---- 2. the block column itself, 8 columns at a time ----
        for (int s8 = 0; s8 < ntJ; s8++) {
            const int c0 = j0 + 8 * s8;
            double l[36], di[8];
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c <= r; c++) l[r * (r + 1) / 2 + c] = D[(size_t)(c0 + r) * kp + c0 + c];
            if (!chol8_reg(l, di)) ok = false;
            const int nrD = kp - c0 - 8, nrT = nrD + (O ? kp : 0);
            for (int r = tid; r < nrT; r += nt) {   // rows below the tile: v <- v L11^-T
                double *rp = (r < nrD) ? D + (size_t)(c0 + 8 + r) * kp + c0 : O + (size_t)(r - nrD) * kp + c0;
                double v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] = rp[c];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double sm = v[c];
#pragma unroll
                    for (int k = 0; k < c; k++) sm -= v[k] * l[c * (c + 1) / 2 + k];
                    v[c] = sm * di[c];
                }
#pragma unroll
                for (int c = 0; c < 8; c++) rp[c] = v[c];
            }
            __syncthreads();   // every thread has read the diagonal tile
            if (tid == 0) {
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) D[(size_t)(c0 + r) * kp + c0 + c] = (c <= r) ? l[r * (r + 1) / 2 + c] : 0.0;
            }
            // in-panel update of the remaining columns of the block column (k = 8)
            if (s8 + 1 < ntJ) {
                const int r0 = c0 + 8, nsD = (kp - r0) >> 3, nsO = O ? ntile : 0, ntj = ntJ - s8 - 1;
                for (int s = warp; s < nsD + nsO; s += nw) {
                    const bool isO = s >= nsD;
                    const int i0 = isO ? (s - nsD) * 8 : r0 + s * 8;
                    double *C0 = (isO ? O : D) + (size_t)i0 * kp;
                    const int jt_max = isO ? 3 : (i0 - r0) >> 3;
                    double acc[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
                    strip_mma(acc, C0 + c0, D + (size_t)r0 * kp + c0, kp, 8, ntj, jt_max);
#pragma unroll
                    for (int jt = 0; jt < 3; jt++)
                        if (jt < ntj && jt <= jt_max) {
                            double *c = C0 + (size_t)g * kp + r0 + 8 * jt + 2 * t4;
                            c[0] -= acc[jt][0];
                            c[1] -= acc[jt][1];
                        }
                }
            }
            __syncthreads();
        }
    }
    // ---- inverses of the 32 x 32 diagonal blocks: one warp per block, one lane per column ----
    const int ninv = bla_ninv(kp);
    for (int J = warp; J < ninv; J += nw) {
        const int j0 = J * BLA_W, wJ = (kp - j0 < BLA_W) ? kp - j0 : BLA_W;
        double *X = Linv + (size_t)J * BLA_W * BLA_W;
        const int c = lane;
        for (int r = 0; r < BLA_W; r++) {
            double v = 0.0;
            if (c < wJ && r < wJ && r >= c) {
                const double *Lr = D + (size_t)(j0 + r) * kp + j0;
                double sm = (r == c) ? 1.0 : 0.0;
                for (int k = c; k < r; k++) sm -= Lr[k] * X[k * BLA_W + c];
                v = sm / Lr[r];
            }
            X[r * BLA_W + c] = v;   // a lane only ever re-reads its own column
        }
    }
    __syncthreads();
    return ok;
}

// block tridiagonal Cholesky: Dall (nblk diagonal blocks, lower), Oall (nblk-1 blocks (t+1, t)), ld = kp
RBPE_DEV bool factor_bt_blk(int nblk, int kp, double *Dall, double *Oall, double *Linv) {
    const size_t kk = (size_t)kp * kp, li = (size_t)bla_ninv(kp) * BLA_W * BLA_W;
    bool ok = true;
    for (int t = 0; t < nblk; t++)
        ok = chol_tall(kp, Dall + t * kk, (t < nblk - 1) ? Oall + t * kk : nullptr, (t > 0) ? Oall + (t - 1) * kk : nullptr,
                       Linv + t * li) && ok;
    return ok;
}

// g (nblk blocks of kb, stride kb) <- (L L')^-1 g.  w: nblk*kp work doubles, y: kp + 32 work doubles.
RBPE_DEV void solve_bt_blk(int nblk, int kb, int kp, const double *Dall, const double *Oall, const double *Linv, double *g,
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
            for (int r = warp; r < wJ; r += nw) {   // y = Linv_J w_J
                double sm = (lane <= r) ? X[r * BLA_W + lane] * wt[j0 + lane] : 0.0;
                for (int o = 16; o > 0; o >>= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
                if (lane == 0) y[r] = sm;
            }
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
            if (warp == 0) {   // y = Linv_J' w_J
                double sm = 0;
                if (lane < wJ)
                    for (int r = lane; r < wJ; r++) sm += X[r * BLA_W + lane] * wt[j0 + r];
                y[lane] = sm;
            }
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

#endif
}  // namespace rbpe
