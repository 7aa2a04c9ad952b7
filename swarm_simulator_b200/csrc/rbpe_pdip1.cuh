// rbpe_pdip1.cuh -- k2 for one-agent batches (plan/batch_size = 1, the per-agent QP of the north-star):
// ONE WARP PER QP, ONE LANE PER CONTROL POINT.  (Round-2 layout of the code: ONE copy of the row loop with a run-time
// pass mode and a phase machine around it, because the kernel is instruction-FETCH bound -- ncu r1/r2a: icc hit rate 69 %,
// GPC instruction-cache requests at 92 % of peak, `no_instruction` the top stall -- while the B200's L0 / L1.5 instruction
// caches hold ~6 KB / 32 KB: four template copies of the row loop, run by 16 warps in 16 different phases, do not fit.
// The round-1 layout is kept as rbpe_pdip1_v1.cuh for A/B runs, -DRBPE_W1_V1.)
//
// Same algorithm and row order as pdip_kernel (rbpe_kernels.cuh), different mapping:
//   * lane <-> Bernstein control point (6M of them, 32 per "slot"); the lane walks the rows that touch its control
//     point (N-1 RSFC rows against the frozen agents + 6 box rows) and accumulates G'(.) and the 3x3 block sum w g g'
//     in registers -- no shuffle reductions per control point, no block barriers, no idle warps;
//   * row state (h, s, z) lives in an L2-resident arena laid out [slot][neighbour][lane] (coalesced); 1/(s z) is
//     recomputed where needed from MUFU.RCP + two Newton steps instead of an FP64 division or a stored reciprocal;
//   * the 9(M-1) x 9(M-1) block tridiagonal reduced system is factored / solved by the same warp out of registers
//     (factor_bt9l / solve_bt9v; the square-root-free factorisation written for the latency kernel is 0.7 % faster here
//     too -- same-box A/B 2.941 -> 2.962 M agent-QPs/s -- and keeps the two one-agent kernels on the same arithmetic).
// A CTA holds W1_WARPS independent QPs (missions in Gauss-Seidel mode, (mission, agent) pairs in Jacobi mode).
#pragma once

namespace rbpe {

#ifndef RBPE_W1_WARPS
#define RBPE_W1_WARPS 4
#endif
constexpr int W1_WARPS = RBPE_W1_WARPS;   // QPs (warps) per CTA; 16 warps per SM either way (registers, shared memory)
#ifndef RBPE_W1_UNROLL
#define RBPE_W1_UNROLL 1
#endif
constexpr int W1_UNROLL = RBPE_W1_UNROLL;
#ifndef RBPE_W1_FUSE_COR
// 1: no corrector pass over the rows -- its G' product is accumulated in two parts by the affine pass (w1_pass): +10 %
// (same-box A/B, profiles/r2_pdip1_ab.md); 0: four passes per iteration (tools only)
#define RBPE_W1_FUSE_COR 1
#endif
#ifndef RBPE_W1_PF
#define RBPE_W1_PF 2   // software-pipeline depth of the row loop: 2 = rows and normals one iteration ahead, 1 = rows only
#endif   // unroll factor of the row loop (tuning; 1 = smallest code)

#define QROW(a) q_base_int(a, 0), q_base_int(a, 1), q_base_int(a, 2), q_base_int(a, 3), q_base_int(a, 4), q_base_int(a, 5)
#if defined(__CUDACC__)
__constant__ double c_QB[36] = {QROW(0), QROW(1), QROW(2), QROW(3), QROW(4), QROW(5)};   // Q_base (build_Q_base L327-L347)
#else
static const double c_QB[36] = {QROW(0), QROW(1), QROW(2), QROW(3), QROW(4), QROW(5)};
#endif
#undef QROW

// Row state of one warp: blocks of W1_ROWBLK doubles, block (slot, j) = the j-th kept row of every lane of the slot:
// h[32] | s[32] | z[32] | row number e[32] (int).  One running pointer per lane walks a lane's rows (constant offsets
// 0 / 256 / 512 bytes, stride 896 bytes): the separate-array layout of round 1 cost ~40 integer instructions of 64-bit
// address arithmetic per row and pass (ncu r2b), more than the row's floating-point work.
constexpr int W1_ROWBLK = 3 * 32 + 16;
// Per-segment constants staged in shared memory (per warp): CL[3][3] | CR[3][3] | RQ[6][6] | qscale.  The knot-space helpers
// (Z, Z', reduced Hessian, dual residual) read them with 32-bit shared addressing; from global memory every access cost a
// 64-bit LEA pair (ncu r2b: build_W spent ~110 instructions per entry, mostly address arithmetic).
constexpr int SEGC_CL = 0, SEGC_CR = 9, SEGC_RQ = 18, SEGC_QS = 54, SEGC = 56;
__host__ __device__ inline size_t w1_smem_doubles(int M) {  // per warp
    size_t ncp = 6 * (size_t)M, nr = 9 * (size_t)(M > 1 ? M - 1 : 0);
    return al2(6 * 3 * ncp) + al2(6 * ncp) + al2((size_t)(M > 1 ? M - 1 : 1) * 81) + al2((size_t)(M > 2 ? M - 2 : 1) * 81) + al2(nr) + al2(nr > 32 ? nr : 32) + al2((6 * (size_t)M + 31) / 32) + (size_t)M * SEGC;
}
__host__ __device__ inline size_t w1_scratch_doubles(int N, int M) {  // per warp, global arena
    size_t nslot = (6 * (size_t)M + 31) / 32, NR = (size_t)(N > 1 ? N - 1 : 0) + 6;   // + the 6 box rows of a control point
    return nslot * NR * W1_ROWBLK + al2(nslot * 16) + al2((size_t)M * NR * 3) + 8;
}

#if defined(__CUDACC__) || defined(RBPE_EMU)

// max(acc, v) that keeps acc when v is NaN (what fmax does for our accumulators) in DSETP + 2 SEL; fmax() itself expands to
// ~10 instructions of NaN handling, and the row loop holds up to three of them per row
RBPE_DEV double dmax(double acc, double v) { return (v > acc) ? v : acc; }

RBPE_DEV double rcp_nr(double a) {  // 1/a to double rounding: 20-bit hardware seed + two Newton steps
#ifdef RBPE_EMU
    double r = (double)(1.0f / (float)a);
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));   // MUFU.RCP64H: ~20 bits, one XU op, no FP32<->FP64 conversions
#endif
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
#ifndef RBPE_RCP_ONE_STEP
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
#endif
    return r;
}

struct W1 {
    int N, M, NE, NR, ncp, nslot, nr, qa, sequential;
    const double *start, *goal, *radius, *segbox, *segmat;
    const double *segc;   // shared-memory copy of the per-segment constants the iteration needs (SEGC_* below)
    const float *reln;
    const double *ctrl_src;
    // shared memory (per warp); x-space index v = m*18 + k*6 + i
    double *x, *dxa, *dx, *rdx, *vA, *vB, *Dcp, *Wd, *Wo, *sg, *sg2, *dinv;
    const double *QB;
    int *cmax;   // [nslot] largest kept-row count of a slot (warp-uniform loop bound)
    // global arena (per warp)
    // rows of a control point: e < NE the RSFC rows against the other agents, e = NE + 2k + side its box rows
    // (x_k <= ub, -x_k <= -lb) written as ordinary rows with unit normals, so that every pass is ONE loop (code size)
    // Presolve: an RSFC row whose maximal activity over the box of its control point stays below its right-hand side can
    // never be active (bound-based row redundancy) and is not stored; a lane keeps cnt <= NR rows, compacted.
    double *rows;                                           // [slot][j][W1_ROWBLK]: h, s, z, e of the j-th kept row of every lane, j < cnt[slot][lane]
    int *cnt;                                               // [slot][lane]
    double *nrm;                                            // [m][e][3], sign folded in (FP64: no per-row F2F)
};

// Warp all-reduce of (sum, sum, max, max, max) -- ONE non-inlined copy with a rolled butterfly: the inlined templated
// version accounted for a third of the kernel's code (64-bit shuffles are two SHFLs plus packing each), and the kernel is
// instruction-fetch bound (ncu r1: no_instruction is the top stall whenever the text grows past ~125 KB).
struct Red5 { double s1, s2, mx, mx2, mx3; };
RBPE_NOINLINE Red5 warp_reduce5(double s1, double s2, double mx, double mx2, double mx3) {
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        mx = dmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mx2 = dmax(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
        mx3 = dmax(mx3, __shfl_xor_sync(0xffffffffu, mx3, o));
    }
    Red5 r;
    r.s1 = s1; r.s2 = s2; r.mx = mx; r.mx2 = mx2; r.mx3 = mx3;
    return r;
}


RBPE_DEV bool w1_dead(const W1 &c, int cp) { int m = cp / 6, i = cp % 6; return (m == 0 && i < 3) || (m == c.M - 1 && i >= 3); }

// One pass over the kept inequality rows: the lane walks the rows of its control point.  `mode` is warp-uniform, every
// branch on it is a uniform branch; same row algebra as row_eval (rbpe_kernels.cuh), t = 1/(s z) recomputed here:
//   P_INIT   unit weights: vA = G'(h - G x), D = sum g g';                     acc.mx2 = max |h|
//   P_START  z = G x - h, s = -z;                                             acc.mx = max(-s), acc.mx2 = max(-z)
//   P_SHIFT  s += sa, z += sb
//   P_RES    pending step (sa = its sigma mu, sb = its length) fused in; vA = G'z, vB = G' t_aff, D = sum w g g';
//            acc.s1 = s'z, acc.s2 = h'z, acc.mx = max |rg|, acc.mx2 = max z
//   P_AFF    ratio test of the affine direction (max form) in acc.mx; acc.s1, acc.s2 = coefficients of mu_aff(a)
//   P_COR    vA = G'(corrector coefficients), sa = sigma mu   (RBPE_W1_FUSE_COR = 0 only; otherwise P_AFF also leaves
//            vA = G'(coefficients without sigma mu), vB = G'(1/s), and the phase machine forms vA - sigma mu vB)
//   P_STEP   ratio test of the final direction in acc.mx
// Called from ONE place (w1_solve_qp's phase machine) so that a single copy of this loop exists in the kernel text.
RBPE_DEV void w1_pass(const W1 &c, const int mode, const double sa, const double sb, Acc &out) {
    const int lane = threadIdx.x & 31;
    Acc acc;
    acc.s1 = 0; acc.s2 = 0; acc.mx = (mode == P_AFF || mode == P_STEP) ? 0.0 : -1e300; acc.mx2 = -1e300; acc.mn = 1e300;
#pragma unroll 1
    for (int slot = 0; slot < c.nslot; slot++) {
        const int cp = slot * 32 + lane;
        const bool on = cp < c.ncp && !w1_dead(c, cp);
        if (!__any_sync(0xffffffffu, on)) continue;
        const int m = on ? cp / 6 : 0, i = on ? cp % 6 : 0, v0 = m * 18 + i;
        double x0 = 0, x1 = 0, x2 = 0, a0 = 0, a1 = 0, a2 = 0, d0 = 0, d1 = 0, d2 = 0;
        if (on) {
            x0 = c.x[v0]; x1 = c.x[v0 + 6]; x2 = c.x[v0 + 12];
            a0 = c.dxa[v0]; a1 = c.dxa[v0 + 6]; a2 = c.dxa[v0 + 12];
            d0 = c.dx[v0]; d1 = c.dx[v0 + 6]; d2 = c.dx[v0 + 12];
        }
        double vA0 = 0, vA1 = 0, vA2 = 0, vB0 = 0, vB1 = 0, vB2 = 0;
        double Dxx = 0, Dxy = 0, Dxz = 0, Dyy = 0, Dyz = 0, Dzz = 0;
        if (on) {
            const double *nm = c.nrm + (size_t)m * c.NR * 3;
            double *pr = c.rows + (size_t)slot * c.NR * W1_ROWBLK + lane;          // h at pr[0], s at pr[32], z at pr[64]
            const int *pe = (const int *)(c.rows + (size_t)slot * c.NR * W1_ROWBLK + 96) + lane;
            const int cnt = c.cnt[slot * 32 + lane];
            // Software pipeline: the loads of row j + 1 (h, s, z and its normal) are issued at the top of iteration j, the row
            // number e two iterations ahead, so that an L2 round trip (the arena is L2 resident) overlaps the arithmetic of a
            // whole row instead of stalling every iteration (ncu r2b: long_scoreboard was 29 % of the stall slots).
            int e1 = cnt > 0 ? pe[0] : 0;                       // e of row j + 1 (at loop entry: row 0)
            int e2 = cnt > 1 ? pe[2 * W1_ROWBLK] : 0;           // e of row j + 2
            pe += 4 * W1_ROWBLK;
            double hN = 0, sN = 0, zN = 0, m0 = 0, m1 = 0, m2 = 0;
            if (cnt > 0) {
                hN = pr[0]; sN = pr[32]; zN = pr[64];
                const double *ne = nm + e1 * 3;
                m0 = ne[0]; m1 = ne[1]; m2 = ne[2];
            }
#pragma unroll 1
            for (int j = 0; j < cnt; j++, pr += W1_ROWBLK, pe += 2 * W1_ROWBLK) {
#if RBPE_W1_PF == 2
                const double n0 = m0, n1 = m1, n2 = m2, h = hN;
                double s = sN, z = zN;
                e1 = e2;
                if (j + 1 < cnt) {
                    hN = pr[W1_ROWBLK]; sN = pr[W1_ROWBLK + 32]; zN = pr[W1_ROWBLK + 64];
                    const double *ne = nm + e1 * 3;
                    m0 = ne[0]; m1 = ne[1]; m2 = ne[2];
                    e2 = (j + 2 < cnt) ? *pe : 0;
                }
#else   // rows one ahead, normals in the iteration that uses them
                const double h = hN;
                double s = sN, z = zN;
                const double *ne = nm + e1 * 3;
                const double n0 = ne[0], n1 = ne[1], n2 = ne[2];
                e1 = e2;
                if (j + 1 < cnt) {
                    hN = pr[W1_ROWBLK]; sN = pr[W1_ROWBLK + 32]; zN = pr[W1_ROWBLK + 64];
                    e2 = (j + 2 < cnt) ? *pe : 0;
                }
#endif
                double gx = n0 * x0 + n1 * x1 + n2 * x2;
                double cA, w;
                if (mode <= P_SHIFT) {
                    if (mode != P_INIT) {
                        if (mode == P_START) {
                            z = gx - h; s = -z;
                            acc.mx = dmax(acc.mx, -s); acc.mx2 = dmax(acc.mx2, -z);
                        } else {
                            s += sa; z += sb;
                        }
                        pr[32] = s; pr[64] = z;
                        continue;
                    }
                    w = 1.0; cA = h - gx;
                    acc.mx2 = dmax(acc.mx2, fabs(h));
                } else {
                    const double ga = n0 * a0 + n1 * a1 + n2 * a2, gd = n0 * d0 + n1 * d1 + n2 * d2;
                    double t = rcp_nr(s * z), rs = t * z;
                    if (mode == P_RES) {
                        if (sb != 0.0) {   // pending step of the previous iteration, fused into the residual pass
                            double rgo = gx + s - h, wo = z * rs;
                            double dsa = -rgo - ga, dza = -z - wo * dsa;
                            double rc = s * z + dsa * dza - sa;
                            double ds = -rgo - gd, dz = (-rc - z * ds) * rs;
                            s += sb * ds; z += sb * dz;
                            gx += sb * gd;
                            t = rcp_nr(s * z);
                            rs = t * z;
                            pr[32] = s; pr[64] = z;
                        }
                        const double rg = gx + s - h;
                        w = z * rs;
                        cA = z;
                        const double cB = -(w * rg - z);
                        acc.s1 += s * z; acc.s2 += h * z; acc.mx = dmax(acc.mx, fabs(rg)); acc.mx2 = dmax(acc.mx2, z);
                        vB0 += cB * n0; vB1 += cB * n1; vB2 += cB * n2;
                    } else {
                        const double rg = gx + s - h;
                        w = z * rs;
                        const double rz = t * s;
                        const double dsa = -rg - ga, dza = -z - w * dsa;
                        if (mode == P_AFF) {
                            acc.mx = dmax(acc.mx, dmax(-dsa * rs, -dza * rz));
                            acc.s1 += s * dza + z * dsa; acc.s2 += dsa * dza;
#if RBPE_W1_FUSE_COR
                            // The corrector's coefficient -(z rg - rc) / s with rc = s z + dsa dza - sigma mu is affine in
                            // sigma mu, which is only known after this pass's reductions: vA takes the part without it, vB
                            // the sum of g / s; the phase machine forms vA - sigma mu vB.  Saves the fourth pass over the rows.
                            const double c1 = -(z * rg - (s * z + dsa * dza)) * rs;
                            vA0 += c1 * n0; vA1 += c1 * n1; vA2 += c1 * n2;
                            vB0 += rs * n0; vB1 += rs * n1; vB2 += rs * n2;
#endif
                            continue;
                        }
                        const double rc = s * z + dsa * dza - sa;
#if RBPE_W1_FUSE_COR
                        {   // P_STEP
                            const double ds = -rg - gd, dz = (-rc - z * ds) * rs;
                            acc.mx = dmax(acc.mx, dmax(-ds * rs, -dz * rz));
                            continue;
                        }
#else
                        if (mode == P_STEP) {
                            const double ds = -rg - gd, dz = (-rc - z * ds) * rs;
                            acc.mx = dmax(acc.mx, dmax(-ds * rs, -dz * rz));
                            continue;
                        }
                        cA = -(z * rg - rc) * rs;   // P_COR
                        vA0 += cA * n0; vA1 += cA * n1; vA2 += cA * n2;
                        continue;
#endif
                    }
                }
                vA0 += cA * n0; vA1 += cA * n1; vA2 += cA * n2;
                {
                    const double w0 = w * n0, w1 = w * n1, w2 = w * n2;
                    Dxx += w0 * n0; Dxy += w0 * n1; Dxz += w0 * n2; Dyy += w1 * n1; Dyz += w1 * n2; Dzz += w2 * n2;
                }
            }
#if RBPE_W1_FUSE_COR
            if (mode == P_INIT || mode == P_RES || mode == P_AFF) { c.vA[v0] = vA0; c.vA[v0 + 6] = vA1; c.vA[v0 + 12] = vA2; }
            if (mode == P_RES || mode == P_AFF) { c.vB[v0] = vB0; c.vB[v0 + 6] = vB1; c.vB[v0 + 12] = vB2; }
#else
            if (mode == P_INIT || mode == P_RES || mode == P_COR) { c.vA[v0] = vA0; c.vA[v0 + 6] = vA1; c.vA[v0 + 12] = vA2; }
            if (mode == P_RES) { c.vB[v0] = vB0; c.vB[v0 + 6] = vB1; c.vB[v0 + 12] = vB2; }
#endif
            if (mode == P_INIT || mode == P_RES) {
                double *D = c.Dcp + (size_t)cp * 6;
                D[0] = Dxx; D[1] = Dxy; D[2] = Dxz; D[3] = Dyy; D[4] = Dyz; D[5] = Dzz;
            }
        }
    }
    if (mode != P_SHIFT && mode != P_COR) {
        Red5 r = warp_reduce5(acc.s1, acc.s2, acc.mx, acc.mx2, 0.0);
        acc.s1 = r.s1; acc.s2 = r.s2; acc.mx = r.mx; acc.mx2 = r.mx2;
    }
    __syncwarp();
    out = acc;
}
// The helpers below are single non-inlined copies (instruction-cache footprint) and take plain arguments, so that the
// context struct never has its address taken and stays in registers.
// out (nr) = Z' vec (x-space)
RBPE_NOINLINE void w1_Zt(const double *segc, int nr, const double *vec, double *out) {
    #pragma unroll 1
    for (int r = threadIdx.x & 31; r < nr; r += 32) {
        int t = r / 9 + 1, cc = r % 9, k = cc / 3, d = cc % 3;
        const double *CR = segc + (t - 1) * SEGC + SEGC_CR, *CL = segc + t * SEGC + SEGC_CL;
        const double *vl = vec + (t - 1) * 18 + k * 6 + 3, *vr = vec + t * 18 + k * 6;
        double s = 0;
        for (int j = 0; j < 3; j++) s += CR[j * 3 + d] * vl[j] + CL[j * 3 + d] * vr[j];
        out[r] = s;
    }
    __syncwarp();
}
// out (x-space) = Z sg
RBPE_NOINLINE void w1_Z(const double *segc, int M, const double *sg, double *out) {
    const int nv = 18 * M;
    #pragma unroll 1
    for (int v = threadIdx.x & 31; v < nv; v += 32) {
        int m = v / 18, r = v % 18, k = r / 6, i = r % 6;
        double s = 0;
        if (i < 3) {
            if (m > 0) {
                const double *C = segc + m * SEGC + SEGC_CL + i * 3, *g = sg + (m - 1) * 9 + k * 3;
                s = C[0] * g[0] + C[1] * g[1] + C[2] * g[2];
            }
        } else if (m < M - 1) {
            const double *C = segc + m * SEGC + SEGC_CR + (i - 3) * 3, *g = sg + m * 9 + k * 3;
            s = C[0] * g[0] + C[1] * g[1] + C[2] * g[2];
        }
        out[v] = s;
    }
    __syncwarp();
}
RBPE_NOINLINE void w1_build_W(const double *segc, int M, const double *Dcp, double *Wd, double *Wo) {
    // entry idx = r * 9 + cc of a 9 x 9 block, r = (axis k, derivative d), cc = (k2, d2); a lane owns idx = lane, lane + 32,
    // lane + 64 of EVERY knot, so the index arithmetic is done three times per call instead of once per entry
#pragma unroll 1
    for (int idx = threadIdx.x & 31; idx < 81; idx += 32) {
        const int r = idx / 9, cc = idx - 9 * r;
        const int k = r / 3, d = r - 3 * k, k2 = cc / 3, d2 = cc - 3 * k2;
        const int e = sym6(k, k2);
        const bool low = cc <= r, same = k == k2;
#pragma unroll 1
        for (int t = 1; t < M; t++) {
            const double *sl = segc + (t - 1) * SEGC, *sr = segc + t * SEGC;
            double s = 0;
            if (low) {
                const double *CR = sl + SEGC_CR, *CL = sr + SEGC_CL;
                const double *Dl = Dcp + ((size_t)(t - 1) * 6 + 3) * 6 + e, *Dr = Dcp + ((size_t)t * 6) * 6 + e;
                for (int j = 0; j < 3; j++) s += CR[j * 3 + d] * CR[j * 3 + d2] * Dl[j * 6] + CL[j * 3 + d] * CL[j * 3 + d2] * Dr[j * 6];
                if (same) s += sl[SEGC_RQ + (3 + d) * 6 + 3 + d2] + sr[SEGC_RQ + d * 6 + d2];
            }
            Wd[(t - 1) * 81 + idx] = s;
            if (t < M - 1) Wo[(t - 1) * 81 + idx] = same ? sr[SEGC_RQ + (3 + d) * 6 + d2] : 0.0;
        }
    }
    __syncwarp();
}
// dxout = Z (Z'HZ)^-1 Z' r
RBPE_DEV void w1_solve(const W1 &c, const double *r, double *dxout) {
    w1_Zt(c.segc, c.nr, r, c.sg);
    solve_bt9v<1>(c.M - 1, c.Wd, c.Wo, c.sg, c.dinv);
    w1_Z(c.segc, c.M, c.sg, dxout);
}
// rdx = 2 Q x + vA; returns the lane-partial objective and max|Px|
struct ObjMpx { double obj, mpx; };
RBPE_NOINLINE ObjMpx w1_dual(const double *segc, int M, const double *QB, const double *x, const double *vA, double *rdx) {
    const int nv = 18 * M;
    double obj = 0, mpx = 0;
    #pragma unroll 1
    for (int v = threadIdx.x & 31; v < nv; v += 32) {
        int m = v / 18, i = v % 6, b6 = v - i;
        double s = 0;
        for (int j = 0; j < 6; j++) s += QB[i * 6 + j] * x[b6 + j];
        double pxv = 2.0 * segc[m * SEGC + SEGC_QS] * s;
        rdx[v] = pxv + vA[v];
        obj += 0.5 * x[v] * pxv;
        mpx = fmax(mpx, fabs(pxv));
    }
    __syncwarp();
    ObjMpx r;
    r.obj = obj; r.mpx = mpx;
    return r;
}

// returns the number of live (kept, non-constant) inequality rows; dead_viol = largest violation of a row made constant
// by the start / goal equalities (rows on the fixed control points: checked here, never stored)
RBPE_DEV int w1_setup(const W1 &c, double &dead_viol) {
    const int lane = threadIdx.x & 31, M = c.M, N = c.N, nv = 18 * M;
    #pragma unroll 1
    for (int v = lane; v < nv; v += 32) {
        int m = v / 18, r = v % 18, k = r / 6, i = r % 6;
        double xp = 0;
        if (m == 0 && i < 3) {
            const double *C = c.segmat + SEGMAT_CL + i * 3, *st = c.start + (size_t)c.qa * 9 + k;
            xp = C[0] * st[0] + C[1] * st[3] + C[2] * st[6];
        }
        if (m == M - 1 && i >= 3) {
            const double *C = c.segmat + (M - 1) * SEGMAT + SEGMAT_CR + (i - 3) * 3, *gl = c.goal + (size_t)c.qa * 9 + k;
            xp = C[0] * gl[0] + C[1] * gl[3] + C[2] * gl[6];
        }
        c.x[v] = xp; c.dxa[v] = 0; c.dx[v] = 0; c.vA[v] = 0; c.vB[v] = 0; c.rdx[v] = 0;
    }
    // signed normals of the RSFC rows against every other agent (g = sg*n, sg = +1 if qa < qo), then the 6 unit
    // normals of the box rows
    #pragma unroll 1
    for (int idx = lane; idx < M * c.NR; idx += 32) {
        int m = idx / c.NR, e = idx % c.NR;
        double n0, n1, n2;
        if (e < c.NE) {
            int qo = (e < c.qa) ? e : e + 1;
            long it = (c.qa < qo) ? pair_index(N, c.qa, qo) : pair_index(N, qo, c.qa);
            const float *nf = c.reln + ((size_t)it * M + m) * 3;
            double sg = (c.qa < qo) ? 1.0 : -1.0;
            n0 = sg * (double)nf[0]; n1 = sg * (double)nf[1]; n2 = sg * (double)nf[2];
        } else {
            int k = (e - c.NE) >> 1;
            double sg = ((e - c.NE) & 1) ? -1.0 : 1.0;
            n0 = k == 0 ? sg : 0.0; n1 = k == 1 ? sg : 0.0; n2 = k == 2 ? sg : 0.0;
        }
        c.nrm[idx * 3] = n0; c.nrm[idx * 3 + 1] = n1; c.nrm[idx * 3 + 2] = n2;
    }
    __syncwarp();
    int live_rows = 0;
    double dviol = -1e300;
    #pragma unroll 1
    for (int slot = 0; slot < c.nslot; slot++) {
        const int cp = slot * 32 + lane;
        int kept = 0;
        if (cp < c.ncp) {
            const int m = cp / 6, i = cp % 6;
            const bool dead = w1_dead(c, cp);
            const double *box = c.segbox + ((size_t)c.qa * M + m) * 6;
            const double *nm = c.nrm + (size_t)m * c.NR * 3;
            double *pr = c.rows + (size_t)slot * c.NR * W1_ROWBLK + lane;
            const int v0 = m * 18 + i;
            const double xd0 = c.x[v0], xd1 = c.x[v0 + 6], xd2 = c.x[v0 + 12];
            double blo[3], bhi[3];          // bounds of this control point (cp_bounds), once per lane
            #pragma unroll
            for (int k = 0; k < 3; k++) cp_bounds(box, M, m, i, k, blo[k], bhi[k]);
            const double ra = c.radius[c.qa];
            #pragma unroll 1
            for (int e = 0; e < c.NR; e++) {
                double h;
                if (e < c.NE) {
                    int qo = (e < c.qa) ? e : e + 1;
                    const double *co = c.ctrl_src + (size_t)qo * 18 * M + m * 6 + i;
                    // h = sg*n.dummy_other - (r_a + r_other), accumulated in the reference's order (L643-L668)
                    h = -(ra + c.radius[qo]);
                    h += nm[e * 3] * co[0];
                    h += nm[e * 3 + 1] * co[6 * M];
                    h += nm[e * 3 + 2] * co[12 * M];
                    if (!dead) {   // bound-based redundancy: max of g.x over the control point's box
                        double amax = 0;
                        #pragma unroll
                        for (int k = 0; k < 3; k++) {
                            double g = nm[e * 3 + k], a = g * bhi[k], b = g * blo[k];
                            amax += (a > b) ? a : b;
                        }
                        if (amax < h - 1e-9 * fmax(1.0, fabs(h))) continue;
                    }
                } else {   // x_k <= ub ; -x_k <= -lb (L626-L635)
                    const int k = (e - c.NE) >> 1;
                    const double lb = k == 0 ? blo[0] : (k == 1 ? blo[1] : blo[2]), ub = k == 0 ? bhi[0] : (k == 1 ? bhi[1] : bhi[2]);
                    h = ((e - c.NE) & 1) ? -lb : ub;
                }
                if (dead) {   // constant row: only its violation matters (P_DEAD of the round-1 layout)
                    dviol = fmax(dviol, nm[e * 3] * xd0 + nm[e * 3 + 1] * xd1 + nm[e * 3 + 2] * xd2 - h);
                    continue;
                }
                pr[0] = h; pr[32] = 1; pr[64] = 1; ((int *)(pr - lane + 96))[lane] = e;
                pr += W1_ROWBLK;
                kept++;
            }
            if (!dead) live_rows += kept;
            c.cnt[slot * 32 + lane] = kept;
        }
        int mx = kept;
        for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, mx, o); mx = t > mx ? t : mx; }
        if (lane == 0) c.cmax[slot] = mx;
    }
    for (int o = 16; o > 0; o >>= 1) live_rows += __shfl_xor_sync(0xffffffffu, live_rows, o);
    dead_viol = warp_reduce5(0.0, 0.0, dviol, 0.0, 0.0).mx;
    __syncwarp();
    return live_rows;
}


// Phase machine around the single copy of the row loop.  Same sequence of operations as pdip_solve (rbpe_kernels.cuh).
// (A variant with one call site each for the factorisation / solve and a non-inlined axpby helper for the vector updates
// was 4 % slower in the same-box A/B: the extra calls cost more than the 0.8 KB of text they saved.)
enum { PH_INIT = P_INIT, PH_START = P_START, PH_SHIFT = P_SHIFT, PH_RES = P_RES, PH_AFF = P_AFF, PH_COR = P_COR, PH_STEP = P_STEP, PH_DONE = 100 };

RBPE_DEV int w1_solve_qp(const W1 &c, int max_iter, double tol_gap, double tol_res, double *obj_out, int *it_out, double *res_out) {
    const int lane = threadIdx.x & 31;
    Acc acc;
    double dead_viol;
    PROF_DECL;
    const int live_rows = w1_setup(c, dead_viol);
    PROF(0);
    int status = ST_NOT_CONVERGED, it = 0;
    double obj = 0, gap = 0, nrd = 0, nrg = 0, hn = 0;
    int phase = PH_INIT;
    if (dead_viol > PRESOLVE_FEAS_TOL) { status = ST_INFEASIBLE; phase = PH_DONE; }
    if (phase != PH_DONE && c.nr == 0) {
        double o = w1_dual(c.segc, c.M, c.QB, c.x, c.vA, c.rdx).obj;
        obj = warp_reduce5(o, 0.0, 0.0, 0.0, 0.0).s1;
        status = ST_OK; phase = PH_DONE;
    }
    double sa = 0, sb = 0;          // pass scalars of the next pass
    double sigmu = 0, al = 0, mu = 0;
    const double mi = live_rows > 0 ? (double)live_rows : 1.0;
#pragma unroll 1
    while (phase != PH_DONE) {
        w1_pass(c, phase, sa, sb, acc);
        PROF(1);
        if (phase == PH_RES) {
            if (al != 0.0) {
                #pragma unroll 1
                for (int v = lane; v < 18 * c.M; v += 32) c.x[v] += al * c.dx[v];
                __syncwarp();
            }
            mu = acc.s1 / mi;
            const double hz = acc.s2, zmax = acc.mx2;
            nrg = fmax(acc.mx, 0.0);
            ObjMpx om = w1_dual(c.segc, c.M, c.QB, c.x, c.vA, c.rdx);
            double o = om.obj, mpx = om.mpx;
            w1_Zt(c.segc, c.nr, c.rdx, c.sg);
            w1_Zt(c.segc, c.nr, c.vA, c.sg2);
            double mr = 0, mc = 0;
            #pragma unroll 1
            for (int r = lane; r < c.nr; r += 32) { mr = fmax(mr, fabs(c.sg[r])); mc = fmax(mc, fabs(c.sg2[r])); }
            { Red5 r = warp_reduce5(o, 0.0, mpx, mr, mc); o = r.s1; mpx = r.mx; mr = r.mx2; mc = r.mx3; }
            obj = o; nrd = mr;
            gap = mu;
            PROF(2);
            if (!(mu == mu) || !(nrd == nrd)) { status = ST_NOT_CONVERGED; break; }
            {   // acceptance rule of pdip_solve (rbpe_kernels.cuh): strict test, else the round-off floor of the dual residual
                const bool gap_ok = gap <= tol_gap * fmax(1.0, fabs(obj)) && nrg <= tol_res * (1 + hn);
                if (gap_ok && nrd <= tol_res * (1.0 + mpx)) { status = ST_OK; break; }
                if (gap_ok && nrd <= TOL_DUAL_FLOOR * (1.0 + mpx)) { status = ST_OK; break; }
            }
            const double cert = (hz < -PRESOLVE_FEAS_TOL * zmax) ? mc / (-hz) : 1e300;
            if (cert < CERT_RATIO) { status = ST_INFEASIBLE; break; }
            w1_build_W(c.segc, c.M, c.Dcp, c.Wd, c.Wo);
            PROF(3);
            if (!factor_bt9l<1>(c.M - 1, c.Wd, c.Wo, c.dinv, c.sg)) { status = cert < CERT_RATIO_BREAKDOWN ? ST_INFEASIBLE : ST_NOT_CONVERGED; break; }
            PROF(4);
            #pragma unroll 1
            for (int v = lane; v < 18 * c.M; v += 32) c.vB[v] = -c.rdx[v] + c.vB[v];
            __syncwarp();
            w1_solve(c, c.vB, c.dxa);
            PROF(5);
            phase = PH_AFF; sa = 0; sb = 0;
        } else if (phase == PH_AFF) {
            const double aa = (acc.mx > 1.0) ? 1.0 / acc.mx : 1.0;
            const double mua = (mu * mi + aa * acc.s1 + aa * aa * acc.s2) / mi;
            const double sigma = (mu > 0) ? (mua / mu) * (mua / mu) * (mua / mu) : 0.0;
            sigmu = sigma * mu;
#if RBPE_W1_FUSE_COR
            __syncwarp();
            #pragma unroll 1
            for (int v = lane; v < 18 * c.M; v += 32) c.vA[v] = -c.rdx[v] + (c.vA[v] - sigmu * c.vB[v]);
            __syncwarp();
            w1_solve(c, c.vA, c.dx);
            PROF(5);
            phase = PH_STEP; sa = sigmu; sb = 0;
#else
            phase = PH_COR; sa = sigmu; sb = 0;
#endif
        } else if (phase == PH_COR) {
            #pragma unroll 1
            for (int v = lane; v < 18 * c.M; v += 32) c.vA[v] = -c.rdx[v] + c.vA[v];
            __syncwarp();
            w1_solve(c, c.vA, c.dx);
            PROF(5);
            phase = PH_STEP; sa = sigmu; sb = 0;
        } else if (phase == PH_STEP) {
            al = (0.99 < acc.mx) ? 0.99 / acc.mx : 1.0;
            it++;
            if (it >= max_iter) break;
            phase = PH_RES; sa = sigmu; sb = al;
        } else if (phase == PH_INIT) {
            hn = acc.mx2;
            w1_build_W(c.segc, c.M, c.Dcp, c.Wd, c.Wo);
            if (!factor_bt9l<1>(c.M - 1, c.Wd, c.Wo, c.dinv, c.sg)) break;
            w1_dual(c.segc, c.M, c.QB, c.x, c.vA, c.rdx);   // rdx = P x_p + vA
            #pragma unroll 1
            for (int v = lane; v < 18 * c.M; v += 32) c.rdx[v] = 2.0 * c.vA[v] - c.rdx[v];
            __syncwarp();
            w1_solve(c, c.rdx, c.dx);
            #pragma unroll 1
            for (int v = lane; v < 18 * c.M; v += 32) { c.x[v] += c.dx[v]; c.dx[v] = 0; }
            __syncwarp();
            PROF(6);
            phase = PH_START;
        } else if (phase == PH_START) {
            const double ap = acc.mx, ad = acc.mx2;
            sa = ap >= 0 ? 1.0 + ap : 0.0; sb = ad >= 0 ? 1.0 + ad : 0.0;
            phase = PH_SHIFT;
        } else {   // PH_SHIFT
            sa = 0; sb = 0; sigmu = 0; al = 0;
            phase = (max_iter > 0) ? PH_RES : PH_DONE;
        }
    }
    // |Ax - b| for the record
    double mrp = 0;
    #pragma unroll 1
    for (int e = lane; e < 9 * (c.M + 1); e += 32) {
        int t = e / 9, cc = e % 9, k = cc / 3, d = cc % 3;
        double sm = 0;
        if (t < c.M) {
            const double *sp = c.segmat + t * SEGMAT + SEGMAT_AL + d * 6, *xx = c.x + t * 18 + k * 6;
            for (int i = 0; i < 6; i++) sm += sp[i] * xx[i];
        }
        if (t > 0) {
            const double *sp = c.segmat + (t - 1) * SEGMAT + SEGMAT_AR + d * 6, *xx = c.x + (t - 1) * 18 + k * 6;
            for (int i = 0; i < 6; i++) sm += sp[i] * xx[i];
        }
        if (t == 0) sm -= c.start[(size_t)c.qa * 9 + k + 3 * d];
        if (t == c.M) sm -= c.goal[(size_t)c.qa * 9 + k + 3 * d];
        mrp = fmax(mrp, fabs(sm));
    }
    mrp = warp_reduce5(0.0, 0.0, mrp, 0.0, 0.0).mx;
    PROF(7);
    if (lane == 0) {
        *obj_out = obj;
        *it_out = it;
        res_out[0] = gap; res_out[1] = mrp; res_out[2] = nrd; res_out[3] = nrg;
    }
    return status;
}

#ifndef RBPE_W1_MINB
#define RBPE_W1_MINB (16 / RBPE_W1_WARPS)
#endif
__global__ void __launch_bounds__(W1_WARPS * 32, RBPE_W1_MINB) pdip1_kernel(SolveArgs S) {   // blockDim.x = 32 * (QPs per CTA) <= W1_WARPS * 32
    RBPE_DYN_SMEM(smem);
    const int N = S.N, M = S.M, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long unit = (long)blockIdx.x * (blockDim.x >> 5) + warp;   // one QP chain (mode 0) or one QP (mode 1) per warp
    int cidx, l_begin, l_end;
    if (S.mode == 0) {
        cidx = (int)unit; l_begin = 0; l_end = S.nbatch;
    } else {
        int per = S.batch_end - S.batch_begin;
        cidx = (int)(unit / per);
        l_begin = S.batch_begin + (int)(unit % per);
        l_end = l_begin + 1;
    }
    if (cidx >= S.count) return;   // whole warps only; no block-level barrier is used in this kernel
    if (S.status[cidx] != ST_OK && S.mode == 0) return;
    const long P = (long)N * (N - 1) / 2;
    W1 c;
    c.N = N; c.M = M; c.sequential = S.sequential;
    c.NE = S.sequential ? N - 1 : 0;
    c.NR = c.NE + 6;
    c.ncp = 6 * M; c.nslot = (c.ncp + 31) / 32; c.nr = 9 * (M > 1 ? M - 1 : 0);
    c.start = S.start + (size_t)cidx * N * 9;
    c.goal = S.goal + (size_t)cidx * N * 9;
    c.radius = S.radius + (size_t)cidx * N;
    c.segbox = S.segbox + (size_t)cidx * N * M * 6;
    c.segmat = S.segmat + (size_t)cidx * M * SEGMAT;
    c.reln = S.reln + (size_t)cidx * P * M * 3;
    double *ctrl = S.ctrl + (size_t)cidx * N * 18 * M;
    c.ctrl_src = (S.mode == 0) ? ctrl : S.ctrl_frozen + (size_t)cidx * N * 18 * M;
    {   // shared memory of this warp
        double *p = (double *)smem + (size_t)warp * w1_smem_doubles(M);
        const size_t nv = 18 * (size_t)M;
        c.x = p; c.dxa = p + nv; c.dx = p + 2 * nv; c.rdx = p + 3 * nv; c.vA = p + 4 * nv; c.vB = p + 5 * nv;
        p += al2(6 * 3 * c.ncp);
        c.Dcp = p; p += al2(6 * (size_t)c.ncp);
        c.Wd = p; p += al2((size_t)(M > 1 ? M - 1 : 1) * 81);
        c.Wo = p; p += al2((size_t)(M > 2 ? M - 2 : 1) * 81);
        c.sg = p; p += al2(c.nr); c.sg2 = p; c.dinv = p; p += al2(c.nr > 32 ? c.nr : 32);   // the column exchange buffer of the 9x9 routines shares sg2 (dead by then)
        c.QB = c_QB;
        c.cmax = (int *)p; p += al2(((size_t)c.nslot + 1) / 2);
        double *sc = p;
        c.segc = sc;
        #pragma unroll 1
        for (int idx = lane; idx < M * SEGC; idx += 32) {
            const int m = idx / SEGC, o = idx - m * SEGC;
            const double *sm = c.segmat + (size_t)m * SEGMAT;
            sc[idx] = o < SEGC_CR ? sm[SEGMAT_CL + o] : (o < SEGC_RQ ? sm[SEGMAT_CR + o - SEGC_CR] : (o < SEGC_QS ? sm[SEGMAT_RQ + o - SEGC_RQ] : (o == SEGC_QS ? sm[SEGMAT_QS] : 0.0)));
        }
        __syncwarp();
    }
    {   // global arena of this warp
        double *g = S.scratch + (size_t)unit * S.scratch_stride;
        c.rows = g; g += (size_t)c.nslot * c.NR * W1_ROWBLK;
        c.cnt = (int *)g; g += al2((size_t)c.nslot * 16);
        c.nrm = g;
    }
    const int iters = (S.mode == 0) ? S.iteration : 1;
    for (int iter = 0; iter < iters; iter++)
        for (int l = l_begin; l < l_end; l++) {
            c.qa = l;   // batch l of one-agent batches = agent l
            if (c.qa >= N) continue;
            int rec = (S.mode == 0 ? iter * S.nbatch : S.rec_offset) + l;
            int st = w1_solve_qp(c, S.max_iter, S.tol_gap, S.tol_res, S.qp_obj + (size_t)cidx * S.nrec + rec,
                                 S.qp_iters + (size_t)cidx * S.nrec + rec, S.qp_res + ((size_t)cidx * S.nrec + rec) * 4);
            if (lane == 0) {
                S.qp_status[(size_t)cidx * S.nrec + rec] = st;
                if (st != ST_OK) atomicCAS(&S.status[cidx], (int)ST_OK, st);
            }
            if (st != ST_OK && S.mode == 0) return;
            if (st != ST_OK && S.npeer <= 0) continue;
            #pragma unroll 1
            for (int v = lane; v < 18 * M; v += 32) {   // dummy <- vals (L182-L184); with peers: next table of every rank
                int m = v / 18, r = v % 18, k = r / 6, i = r % 6;
                const size_t at = (size_t)c.qa * 18 * M + (size_t)k * 6 * M + m * 6 + i;
#ifdef RBPE_NO_PEER1
                ctrl[at] = c.x[v]; continue;
#endif
                if (S.npeer <= 0) { ctrl[at] = c.x[v]; continue; }
                const double val = (st == ST_OK) ? c.x[v] : c.ctrl_src[at];
                #pragma unroll 1
                for (int p = 0; p < S.npeer; p++) S.peer_ctrl[p][(size_t)cidx * N * 18 * M + at] = val;
            }
            __syncwarp();
        }
    peer_signal_done(S, lane == 0);
}

#endif
}  // namespace rbpe
