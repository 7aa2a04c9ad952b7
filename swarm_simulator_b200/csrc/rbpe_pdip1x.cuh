// rbpe_pdip1x.cuh -- k2c, the LATENCY kernel for one-agent batches: ONE CTA OF SEVERAL WARPS PER QP.
// pdip1_kernel (one warp per QP) is built for throughput: thousands of QPs in flight, every warp a serial chain of FP64
// instructions.  With a handful of missions (the reference's own use: one swarm, planned agent after agent) a warp alone
// on its SM runs that chain at the latency of every instruction: 0.5 ms per QP, 54 % of it in the row passes, 6 % in the
// setup (clock64 phase split, tools/gpu_latency.py with a -DRBPE_PROFILE build, r2).  Here a QP owns X1 warps:
//   * lane <-> control point as in pdip1_kernel, but the rows of a control point are dealt round-robin to the warps
//     (setup: every warp evaluates the rows e = warp, warp + nw, ..., the kept ones are then re-dealt so that every warp
//     holds the same number); a pass runs on all warps at once, partial G'(.) / sum w g g' go through shared memory and
//     are added in warp order (deterministic);
//   * the row state lives in SHARED memory (the kernel is used only where that fits: rbpe_api.cu), so a pass starts
//     without an L2 round trip;
//   * knot-space work (dual residual, Z', reduced Hessian, Z) is spread over the CTA; the 9x9 block tridiagonal
//     factorisation and the substitutions stay on warp 0 (a serial chain either way).
// Same algorithm, acceptance rule and certificate as pdip1_kernel / pdip_solve; summation order differs (rows are dealt
// to warps), so results agree to rounding, not bit for bit.
#pragma once

namespace rbpe {

constexpr int X1_MAXW = 8;        // warps per QP
constexpr int X1_PART = 12;       // partial sums per lane and warp: vA[3] | vB[3] | D[6]

__host__ __device__ inline int x1_cap(int N, int nw) { return ((N > 1 ? N - 1 : 0) + 6 + nw - 1) / nw; }   // rows per (slot, warp)
__host__ __device__ inline size_t x1_smem_doubles(int N, int M, int nw) {
    size_t ncp = 6 * (size_t)M, nr = 9 * (size_t)(M > 1 ? M - 1 : 0), nslot = (ncp + 31) / 32, NR = (size_t)(N > 1 ? N - 1 : 0) + 6;
    size_t t = al2(6 * 3 * ncp) + al2(6 * ncp) + al2((size_t)(M > 1 ? M - 1 : 1) * 81) + al2((size_t)(M > 2 ? M - 2 : 1) * 81) + al2(nr) +
               al2(nr > 32 ? nr : 32) + (size_t)M * SEGC;
    t += (size_t)nw * X1_PART * 32 + 2 * (size_t)X1_MAXW * 8 + 8 + 16;                      // part, scal (two buffers), flags
    t += nslot * nw * (size_t)x1_cap(N, nw) * W1_ROWBLK + al2(nslot * nw * 16) + al2((size_t)M * NR * 3);   // rows, cnt, nrm
    return t;
}

#if defined(__CUDACC__) || defined(RBPE_EMU)

struct X1 {
    int N, M, NE, NR, ncp, nslot, nr, qa, nw, cap;
    const double *start, *goal, *radius, *segbox, *segmat;
    const float *reln;
    const double *ctrl_src;
    double *x, *dxa, *dx, *rdx, *vA, *vB, *Dcp, *Wd, *Wo, *sg, *sg2, *dinv, *segc;
    double *part;   // [warp][X1_PART][32]
    double *scal;   // 2 x [warp][8]
    int *flag;
    double *pv;     // [16] pivots of a 9 x 9 block (factor_bt9l, rbpe_blockla.cuh)
    double *rows;   // [slot][warp][cap][W1_ROWBLK]: h | s | z | e of the warp's rows of every lane
    int *cnt;       // [slot][warp][32]
    double *nrm;    // [m][e][3]
};

struct X1Red { double s1, s2, mx, mx2, mx3; };
// CTA all-reduce of (sum, sum, max, max, max) in two halves with ONE barrier (the caller's) between them:
// x1_red_put -- warp butterfly, one shared-memory slot per warp; x1_red_get -- every thread adds the warp results in
// warp order.  `flip` alternates between two buffers, so that no second barrier is needed before the next reduction.
template <int MASK>   // bit i set: slot i (s1, s2, mx, mx2, mx3) is wanted; the others are not shuffled (their results are garbage)
RBPE_NOINLINE void x1_red_put(double *scal, int flip, double s1, double s2, double mx, double mx2, double mx3) {
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        if (MASK & 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        if (MASK & 2) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        if (MASK & 4) mx = dmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (MASK & 8) mx2 = dmax(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
        if (MASK & 16) mx3 = dmax(mx3, __shfl_xor_sync(0xffffffffu, mx3, o));
    }
    if ((threadIdx.x & 31) == 0) {
        double *p = scal + flip * X1_MAXW * 8 + (threadIdx.x >> 5) * 8;
        p[0] = s1; p[1] = s2; p[2] = mx; p[3] = mx2; p[4] = mx3;
    }
}
template <int MASK = 31>
RBPE_DEV X1Red x1_red_get(const double *scal, int nw, int &flip) {
    const double *b = scal + flip * X1_MAXW * 8;
    flip ^= 1;
    X1Red r;
    r.s1 = (MASK & 1) ? b[0] : 0.0; r.s2 = (MASK & 2) ? b[1] : 0.0; r.mx = (MASK & 4) ? b[2] : 0.0; r.mx2 = (MASK & 8) ? b[3] : 0.0; r.mx3 = (MASK & 16) ? b[4] : 0.0;
#pragma unroll
    for (int w = 1; w < X1_MAXW; w++)   // unrolled: the loads of all warps' slots are in flight together
        if (w < nw) {
            const double *p = b + w * 8;
            if (MASK & 1) r.s1 += p[0];
            if (MASK & 2) r.s2 += p[1];
            if (MASK & 4) r.mx = dmax(r.mx, p[2]);
            if (MASK & 8) r.mx2 = dmax(r.mx2, p[3]);
            if (MASK & 16) r.mx3 = dmax(r.mx3, p[4]);
        }
    return r;
}
template <int MASK = 31>
RBPE_DEV X1Red x1_reduce(double *scal, int nw, int &flip, double s1, double s2, double mx, double mx2, double mx3) {
    x1_red_put<MASK>(scal, flip, s1, s2, mx, mx2, mx3);
    __syncthreads();
    return x1_red_get<MASK>(scal, nw, flip);
}

RBPE_DEV bool x1_dead(const X1 &c, int cp) { int m = cp / 6, i = cp % 6; return (m == 0 && i < 3) || (m == c.M - 1 && i >= 3); }

// One pass over the kept rows, all warps at once (modes and row algebra: w1_pass in rbpe_pdip1.cuh).  Every thread
// returns the same reductions.  The vectors a pass produces (vA, vB, Dcp) are complete only after the CALLER's next
// barrier; in P_COR the pass stores vA - rdx, the right-hand side the corrector solve needs.
// RBPE_X1_FUSE_COR = 1: no corrector pass; the affine pass leaves the two parts of its G' product in vA and vB (see
// RBPE_W1_FUSE_COR in rbpe_pdip1.cuh) and the phase machine forms (vA - sigma mu vB) - rdx.
#ifndef RBPE_X1_FUSE_COR
#define RBPE_X1_FUSE_COR 1
#endif
RBPE_DEV void x1_pass(const X1 &c, const int mode, const double sa, const double sb, int &flip, Acc &out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x, nt = blockDim.x;
    Acc acc;
    acc.s1 = 0; acc.s2 = 0; acc.mx = (mode == P_AFF || mode == P_STEP) ? 0.0 : -1e300; acc.mx2 = -1e300; acc.mn = 1e300;
#if RBPE_X1_FUSE_COR
    const bool vec = (mode == P_INIT || mode == P_RES || mode == P_AFF);
    const int P_NOD = P_AFF;    // the vector-producing pass without the 3 x 3 blocks
#else
    const bool vec = (mode == P_INIT || mode == P_RES || mode == P_COR);
    const int P_NOD = P_COR;
#endif
    PROF_DECL;
#pragma unroll 1
    for (int slot = 0; slot < c.nslot; slot++) {
        const int cp = slot * 32 + lane;
        const bool on = cp < c.ncp && !x1_dead(c, cp);
        const int m = on ? cp / 6 : 0, i = on ? cp % 6 : 0, v0 = m * 18 + i;
        double vA0 = 0, vA1 = 0, vA2 = 0, vB0 = 0, vB1 = 0, vB2 = 0;
        double Dxx = 0, Dxy = 0, Dxz = 0, Dyy = 0, Dyz = 0, Dzz = 0;
        if (on) {
            const double x0 = c.x[v0], x1 = c.x[v0 + 6], x2 = c.x[v0 + 12];
            const double a0 = c.dxa[v0], a1 = c.dxa[v0 + 6], a2 = c.dxa[v0 + 12];
            const double d0 = c.dx[v0], d1 = c.dx[v0 + 6], d2 = c.dx[v0 + 12];
            const double *nm = c.nrm + (size_t)m * c.NR * 3;
            double *pr = c.rows + (size_t)(slot * c.nw + warp) * c.cap * W1_ROWBLK + lane;
            const int *pe = (const int *)(c.rows + (size_t)(slot * c.nw + warp) * c.cap * W1_ROWBLK + 96) + lane;
            const int cnt = c.cnt[(slot * c.nw + warp) * 32 + lane];
#pragma unroll 1   // (2-3 rows per warp and pass: unrolling by 2 measured 3 % slower)
            for (int j = 0; j < cnt; j++, pr += W1_ROWBLK, pe += 2 * W1_ROWBLK) {
                const double h = pr[0];
                double s = pr[32], z = pr[64];
                const double *ne = nm + *pe * 3;
                const double n0 = ne[0], n1 = ne[1], n2 = ne[2];
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
#if RBPE_X1_FUSE_COR
                            const double c1 = -(z * rg - (s * z + dsa * dza)) * rs;
                            vA0 += c1 * n0; vA1 += c1 * n1; vA2 += c1 * n2;
                            vB0 += rs * n0; vB1 += rs * n1; vB2 += rs * n2;
#endif
                            continue;
                        }
                        const double rc = s * z + dsa * dza - sa;
#if RBPE_X1_FUSE_COR
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
        }
        PROF(8);
        if (!vec) continue;   // (mode is uniform over the CTA)
        const bool last = slot == c.nslot - 1;
        {   // partial sums of this warp; added below in warp order
            double *pp = c.part + (size_t)warp * X1_PART * 32 + lane;
            pp[0] = vA0; pp[32] = vA1; pp[64] = vA2;
            if (mode == P_RES || (RBPE_X1_FUSE_COR && mode == P_AFF)) { pp[96] = vB0; pp[128] = vB1; pp[160] = vB2; }
            if (mode != P_NOD) { pp[192] = Dxx; pp[224] = Dxy; pp[256] = Dxz; pp[288] = Dyy; pp[320] = Dyz; pp[352] = Dzz; }
        }
        if (last && mode != P_COR) x1_red_put<15>(c.scal, flip, acc.s1, acc.s2, acc.mx, acc.mx2, 0.0);   // rides on the same barrier
        __syncthreads();
        PROF(9);
#pragma unroll 1
        for (int idx = tid; idx < X1_PART * 32; idx += nt) {
            const int k = idx >> 5, l = idx & 31, cp2 = slot * 32 + l;
            if (cp2 >= c.ncp || x1_dead(c, cp2)) continue;
#if RBPE_X1_FUSE_COR
            if (k >= 6 && mode == P_AFF) continue;
            if (k >= 3 && k < 6 && mode != P_RES && mode != P_AFF) continue;
#else
            if (k >= 3 && mode == P_COR) continue;
            if (k >= 3 && k < 6 && mode != P_RES) continue;
#endif
            const double *pp = c.part + k * 32 + l;
            double sm = pp[0];
#pragma unroll
            for (int w = 1; w < X1_MAXW; w++)
                if (w < c.nw) sm += pp[w * X1_PART * 32];
            const int m2 = cp2 / 6, i2 = cp2 % 6, vv = m2 * 18 + i2;
            if (k < 3) c.vA[vv + 6 * k] = (mode == P_COR) ? sm - c.rdx[vv + 6 * k] : sm;
            else if (k < 6) c.vB[vv + 6 * (k - 3)] = sm;
            else c.Dcp[(size_t)cp2 * 6 + (k - 6)] = sm;
        }
        if (!last) __syncthreads();   // `part` is reused by the next slot
        PROF(10);
    }
    if (mode != P_SHIFT && mode != P_COR) {
        // (ratio tests need the maximum only; P_AFF also the two sums; P_START two maxima)
        X1Red r = vec ? x1_red_get<15>(c.scal, c.nw, flip)
                      : (mode == P_STEP ? x1_reduce<4>(c.scal, c.nw, flip, 0.0, 0.0, acc.mx, 0.0, 0.0)
                                        : (mode == P_AFF ? x1_reduce<7>(c.scal, c.nw, flip, acc.s1, acc.s2, acc.mx, 0.0, 0.0)
                                                         : x1_reduce<12>(c.scal, c.nw, flip, 0.0, 0.0, acc.mx, acc.mx2, 0.0)));
        acc.s1 = r.s1; acc.s2 = r.s2; acc.mx = r.mx; acc.mx2 = r.mx2;
    }
    PROF(11);
    out = acc;
}

// out (nr) = Z' vec, for one or two vectors at once (threads nr .. 2 nr - 1 take the second); no barrier inside
RBPE_DEV void x1_Zt2(const X1 &c, const double *veca, double *outa, const double *vecb, double *outb, double &mxa, double &mxb) {
    const int total = vecb ? 2 * c.nr : c.nr;
#pragma unroll 1
    for (int r2 = threadIdx.x; r2 < total; r2 += blockDim.x) {
        const bool second = r2 >= c.nr;
        const int r = second ? r2 - c.nr : r2;
        const double *vec = second ? vecb : veca;
        int t = r / 9 + 1, cc = r % 9, k = cc / 3, d = cc % 3;
        const double *CR = c.segc + (t - 1) * SEGC + SEGC_CR, *CL = c.segc + t * SEGC + SEGC_CL;
        const double *vl = vec + (t - 1) * 18 + k * 6 + 3, *vr = vec + t * 18 + k * 6;
        double s = 0;
        for (int j = 0; j < 3; j++) s += CR[j * 3 + d] * vl[j] + CL[j * 3 + d] * vr[j];
        if (second) { outb[r] = s; mxb = fmax(mxb, fabs(s)); }
        else { outa[r] = s; mxa = fmax(mxa, fabs(s)); }
    }
}
// out (x-space) = Z sg; no barrier inside
RBPE_DEV void x1_Z(const X1 &c, const double *sg, double *out) {
    const int nv = 18 * c.M, M = c.M;
#pragma unroll 1
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        int m = v / 18, r = v % 18, k = r / 6, i = r % 6;
        double s = 0;
        if (i < 3) {
            if (m > 0) {
                const double *C = c.segc + m * SEGC + SEGC_CL + i * 3, *g = sg + (m - 1) * 9 + k * 3;
                s = C[0] * g[0] + C[1] * g[1] + C[2] * g[2];
            }
        } else if (m < M - 1) {
            const double *C = c.segc + m * SEGC + SEGC_CR + (i - 3) * 3, *g = sg + m * 9 + k * 3;
            s = C[0] * g[0] + C[1] * g[1] + C[2] * g[2];
        }
        out[v] = s;
    }
}
// reduced Hessian, one entry per thread and round; no barrier inside
RBPE_DEV void x1_build_W(const X1 &c) {
    const int total = 81 * (c.M - 1);
#pragma unroll 1
    for (int g = threadIdx.x; g < total; g += blockDim.x) {
        const int t = g / 81 + 1, idx = g - 81 * (t - 1);
        const int r = idx / 9, cc = idx - 9 * r;
        const int k = r / 3, d = r - 3 * k, k2 = cc / 3, d2 = cc - 3 * k2;
        const int e = sym6(k, k2);
        const bool same = k == k2;
        const double *sl = c.segc + (t - 1) * SEGC, *sr = c.segc + t * SEGC;
        double s = 0;
        if (cc <= r) {
            const double *CR = sl + SEGC_CR, *CL = sr + SEGC_CL;
            const double *Dl = c.Dcp + ((size_t)(t - 1) * 6 + 3) * 6 + e, *Dr = c.Dcp + ((size_t)t * 6) * 6 + e;
            for (int j = 0; j < 3; j++) s += CR[j * 3 + d] * CR[j * 3 + d2] * Dl[j * 6] + CL[j * 3 + d] * CL[j * 3 + d2] * Dr[j * 6];
            if (same) s += sl[SEGC_RQ + (3 + d) * 6 + 3 + d2] + sr[SEGC_RQ + d * 6 + d2];
        }
        c.Wd[(t - 1) * 81 + idx] = s;
        if (t < c.M - 1) c.Wo[(t - 1) * 81 + idx] = same ? sr[SEGC_RQ + (3 + d) * 6 + d2] : 0.0;
    }
}
// rdx = 2 Q x + vA (and, with `rhs`, vB <- vB - rdx: the right-hand side of the affine solve); thread-partial objective
// and max |Px|; no barrier inside
RBPE_DEV void x1_dual(const X1 &c, const bool rhs, double &obj, double &mpx) {
    const int nv = 18 * c.M;
#pragma unroll 1
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        int m = v / 18, i = v % 6, b6 = v - i;
        double s = 0;
        for (int j = 0; j < 6; j++) s += c_QB[i * 6 + j] * c.x[b6 + j];
        double pxv = 2.0 * c.segc[m * SEGC + SEGC_QS] * s;
        const double rd = pxv + c.vA[v];
        c.rdx[v] = rd;
        if (rhs) c.vB[v] = -rd + c.vB[v];
        obj += 0.5 * c.x[v] * pxv;
        mpx = fmax(mpx, fabs(pxv));
    }
}
// dxout = Z (Z'HZ)^-1 Z' r; r must be complete (barrier before the call); dxout is complete on return
RBPE_DEV void x1_solve(const X1 &c, const double *r, double *dxout) {
    double d0 = 0, d1 = 0;
    x1_Zt2(c, r, c.sg, nullptr, nullptr, d0, d1);
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) solve_bt9v<2>(c.M - 1, c.Wd, c.Wo, c.sg, c.dinv);
    __syncthreads();
    x1_Z(c, c.sg, dxout);
    __syncthreads();
}

// rows of the QP: every warp evaluates the rows e = warp, warp + nw, ... of its lanes (h in the reference's order, bound
// based redundancy test, constant rows of the fixed control points only checked); the kept rows are then dealt again so
// that the warps of a lane hold equal shares.  Returns the number of live rows; dead_viol as in w1_setup.
RBPE_DEV int x1_setup(const X1 &c, int &flip, double &dead_viol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x, nt = blockDim.x;
    const int M = c.M, N = c.N, nv = 18 * M;
#pragma unroll 1
    for (int v = tid; v < nv; v += nt) {
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
#pragma unroll 1
    for (int idx = tid; idx < M * c.NR; idx += nt) {
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
    __syncthreads();
    int live_rows = 0;
    double dviol = -1e300;
#pragma unroll 1
    for (int slot = 0; slot < c.nslot; slot++) {
        const int cp = slot * 32 + lane;
        int kept = 0;
        // pass 1: the warp's share of the rows, kept ones parked in the s / z fields of the warp's own region
        double *reg = c.rows + (size_t)(slot * c.nw + warp) * c.cap * W1_ROWBLK;
        if (cp < c.ncp) {
            const int m = cp / 6, i = cp % 6;
            const bool dead = x1_dead(c, cp);
            const double *box = c.segbox + ((size_t)c.qa * M + m) * 6;
            const double *nm = c.nrm + (size_t)m * c.NR * 3;
            double *pr = reg + lane;
            const int v0 = m * 18 + i;
            const double xd0 = c.x[v0], xd1 = c.x[v0 + 6], xd2 = c.x[v0 + 12];
            double blo[3], bhi[3];
#pragma unroll
            for (int k = 0; k < 3; k++) cp_bounds(box, M, m, i, k, blo[k], bhi[k]);
            const double ra = c.radius[c.qa];
#pragma unroll 1
            for (int e = warp; e < c.NR; e += c.nw) {
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
                if (dead) {
                    dviol = fmax(dviol, nm[e * 3] * xd0 + nm[e * 3 + 1] * xd1 + nm[e * 3 + 2] * xd2 - h);
                    continue;
                }
                pr[32] = h; pr[64] = (double)e;
                pr += W1_ROWBLK;
                kept++;
            }
            if (!dead) live_rows += kept;
        }
        c.cnt[(slot * c.nw + warp) * 32 + lane] = kept;
        __syncthreads();
        // pass 2: row k of the lane's concatenated list (warp 0's rows, then warp 1's, ...) goes to warp k mod nw
        int total = 0;
#pragma unroll 1
        for (int w = 0; w < c.nw; w++) total += c.cnt[(slot * c.nw + w) * 32 + lane];
        {
            int src_w = 0, src_base = 0, src_cnt = c.cnt[slot * c.nw * 32 + lane];   // source sublist that holds row k
            double *pr = reg + lane;
#pragma unroll 1
            for (int k = warp; k < total; k += c.nw, pr += W1_ROWBLK) {
                while (k >= src_base + src_cnt) { src_base += src_cnt; src_w++; src_cnt = c.cnt[(slot * c.nw + src_w) * 32 + lane]; }
                const double *ps = c.rows + ((size_t)(slot * c.nw + src_w) * c.cap + (k - src_base)) * W1_ROWBLK + lane;
                pr[0] = ps[32];
                ((int *)(pr - lane + 96))[lane] = (int)ps[64];
            }
        }
        __syncthreads();   // every parked row has been read: s / z and the counts may be overwritten
        c.cnt[(slot * c.nw + warp) * 32 + lane] = (total > warp) ? (total - warp + c.nw - 1) / c.nw : 0;
    }
    X1Red r = x1_reduce(c.scal, c.nw, flip, (double)live_rows, 0.0, dviol, 0.0, 0.0);
    dead_viol = r.mx;
    return (int)(r.s1 + 0.5);
}

RBPE_DEV int x1_solve_qp(const X1 &c, int max_iter, double tol_gap, double tol_res, double *obj_out, int *it_out, double *res_out) {
    const int tid = threadIdx.x, nt = blockDim.x, warp = threadIdx.x >> 5, nv = 18 * c.M;
    int flip = 0;
    Acc acc;
    double dead_viol;
    PROF_DECL;
    const int live_rows = x1_setup(c, flip, dead_viol);
    PROF(0);
    int status = ST_NOT_CONVERGED, it = 0;
    double obj = 0, gap = 0, nrd = 0, nrg = 0, hn = 0;
    int phase = PH_INIT;
    if (dead_viol > PRESOLVE_FEAS_TOL) { status = ST_INFEASIBLE; phase = PH_DONE; }
    if (phase != PH_DONE && c.nr == 0) {
        double o = 0, mpx = 0;
        x1_dual(c, false, o, mpx);
        obj = x1_reduce(c.scal, c.nw, flip, o, 0.0, 0.0, 0.0, 0.0).s1;
        status = ST_OK; phase = PH_DONE;
    }
    double sa = 0, sb = 0;
    double sigmu = 0, al = 0, mu = 0;
    const double mi = live_rows > 0 ? (double)live_rows : 1.0;
#pragma unroll 1
    while (phase != PH_DONE) {
        x1_pass(c, phase, sa, sb, flip, acc);
        PROF(1);
        if (phase == PH_RES) {
            if (al != 0.0) {
#pragma unroll 1
                for (int v = tid; v < nv; v += nt) c.x[v] += al * c.dx[v];
            }
            __syncthreads();   // x, and the pass's vA / vB / Dcp
            mu = acc.s1 / mi;
            const double hz = acc.s2, zmax = acc.mx2;
            nrg = fmax(acc.mx, 0.0);
            double o = 0, mpx = 0, mr = 0, mc = 0;
            x1_dual(c, true, o, mpx);
            __syncthreads();
            x1_Zt2(c, c.rdx, c.sg, c.vA, c.sg2, mr, mc);
            { X1Red r = x1_reduce<29>(c.scal, c.nw, flip, o, 0.0, mpx, mr, mc); o = r.s1; mpx = r.mx; mr = r.mx2; mc = r.mx3; }
            obj = o; nrd = mr;
            gap = mu;
            PROF(2);
            if (!(mu == mu) || !(nrd == nrd)) { status = ST_NOT_CONVERGED; break; }
            {   // acceptance rule of pdip_solve (rbpe_kernels.cuh)
                const bool gap_ok = gap <= tol_gap * fmax(1.0, fabs(obj)) && nrg <= tol_res * (1 + hn);
                if (gap_ok && nrd <= tol_res * (1.0 + mpx)) { status = ST_OK; break; }
                if (gap_ok && nrd <= TOL_DUAL_FLOOR * (1.0 + mpx)) { status = ST_OK; break; }
            }
            const double cert = (hz < -PRESOLVE_FEAS_TOL * zmax) ? mc / (-hz) : 1e300;
            if (cert < CERT_RATIO) { status = ST_INFEASIBLE; break; }
            x1_build_W(c);
            __syncthreads();
            PROF(3);
            if (warp == 0) {
                const bool ok = factor_bt9l<2>(c.M - 1, c.Wd, c.Wo, c.dinv, c.pv);
                if ((tid & 31) == 0) *c.flag = ok ? 1 : 0;
            }
            __syncthreads();
            PROF(4);
            if (!*c.flag) { status = cert < CERT_RATIO_BREAKDOWN ? ST_INFEASIBLE : ST_NOT_CONVERGED; break; }
            x1_solve(c, c.vB, c.dxa);
            PROF(5);
            phase = PH_AFF; sa = 0; sb = 0;
        } else if (phase == PH_AFF) {
            const double aa = (acc.mx > 1.0) ? 1.0 / acc.mx : 1.0;
            const double mua = (mu * mi + aa * acc.s1 + aa * aa * acc.s2) / mi;
            const double sigma = (mu > 0) ? (mua / mu) * (mua / mu) * (mua / mu) : 0.0;
            sigmu = sigma * mu;
#if RBPE_X1_FUSE_COR
            __syncthreads();   // the pass's vA / vB
#pragma unroll 1
            for (int v = tid; v < nv; v += nt) c.vA[v] = (c.vA[v] - sigmu * c.vB[v]) - c.rdx[v];
            __syncthreads();
            x1_solve(c, c.vA, c.dx);
            PROF(5);
            phase = PH_STEP; sa = sigmu; sb = 0;
#else
            phase = PH_COR; sa = sigmu; sb = 0;
#endif
        } else if (phase == PH_COR) {
            __syncthreads();   // vA - rdx, stored by the pass
            x1_solve(c, c.vA, c.dx);
            PROF(5);
            phase = PH_STEP; sa = sigmu; sb = 0;
        } else if (phase == PH_STEP) {
            al = (0.99 < acc.mx) ? 0.99 / acc.mx : 1.0;
            it++;
            if (it >= max_iter) break;
            phase = PH_RES; sa = sigmu; sb = al;
        } else if (phase == PH_INIT) {
            hn = acc.mx2;
            __syncthreads();   // the pass's vA / Dcp
            x1_build_W(c);
            __syncthreads();
            if (warp == 0) {
                const bool ok = factor_bt9l<2>(c.M - 1, c.Wd, c.Wo, c.dinv, c.pv);
                if ((tid & 31) == 0) *c.flag = ok ? 1 : 0;
            }
            __syncthreads();
            if (!*c.flag) break;
            {
                double o = 0, mpx = 0;
                x1_dual(c, false, o, mpx);   // rdx = P x_p + vA
            }
#pragma unroll 1
            for (int v = tid; v < nv; v += nt) c.rdx[v] = 2.0 * c.vA[v] - c.rdx[v];   // (the thread that wrote rdx[v])
            __syncthreads();
            x1_solve(c, c.rdx, c.dx);
#pragma unroll 1
            for (int v = tid; v < nv; v += nt) { c.x[v] += c.dx[v]; c.dx[v] = 0; }   // (the thread that wrote dx[v])
            __syncthreads();
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
    __syncthreads();
    // |Ax - b| for the record
    double mrp = 0;
#pragma unroll 1
    for (int e = tid; e < 9 * (c.M + 1); e += nt) {
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
    mrp = x1_reduce(c.scal, c.nw, flip, 0.0, 0.0, mrp, 0.0, 0.0).mx;
    PROF(7);
    if (tid == 0) {
        *obj_out = obj;
        *it_out = it;
        res_out[0] = gap; res_out[1] = mrp; res_out[2] = nrd; res_out[3] = nrg;
    }
    return status;
}

__global__ void __launch_bounds__(X1_MAXW * 32, 1) pdip1x_kernel(SolveArgs S) {   // blockDim.x = 32 * (warps per QP)
    RBPE_DYN_SMEM(smem);
    const int N = S.N, M = S.M, tid = threadIdx.x, nt = blockDim.x;
    const long unit = blockIdx.x;   // one QP chain (mode 0) or one QP (mode 1) per CTA
    int cidx, l_begin, l_end;
    if (S.mode == 0) {
        cidx = (int)unit; l_begin = 0; l_end = S.nbatch;
    } else {
        int per = S.batch_end - S.batch_begin;
        cidx = (int)(unit / per);
        l_begin = S.batch_begin + (int)(unit % per);
        l_end = l_begin + 1;
    }
    if (cidx >= S.count) return;
    if (S.status[cidx] != ST_OK && S.mode == 0) return;
    const long P = (long)N * (N - 1) / 2;
    X1 c;
    c.N = N; c.M = M;
    c.NE = S.sequential ? N - 1 : 0;
    c.NR = c.NE + 6;
    c.nw = nt >> 5; c.cap = (c.NR + c.nw - 1) / c.nw;
    c.ncp = 6 * M; c.nslot = (c.ncp + 31) / 32; c.nr = 9 * (M > 1 ? M - 1 : 0);
    c.start = S.start + (size_t)cidx * N * 9;
    c.goal = S.goal + (size_t)cidx * N * 9;
    c.radius = S.radius + (size_t)cidx * N;
    c.segbox = S.segbox + (size_t)cidx * N * M * 6;
    c.segmat = S.segmat + (size_t)cidx * M * SEGMAT;
    c.reln = S.reln + (size_t)cidx * P * M * 3;
    double *ctrl = S.ctrl + (size_t)cidx * N * 18 * M;
    c.ctrl_src = (S.mode == 0) ? ctrl : S.ctrl_frozen + (size_t)cidx * N * 18 * M;
    {
        double *p = (double *)smem;
        const size_t nv = 18 * (size_t)M;
        c.x = p; c.dxa = p + nv; c.dx = p + 2 * nv; c.rdx = p + 3 * nv; c.vA = p + 4 * nv; c.vB = p + 5 * nv;
        p += al2(6 * 3 * c.ncp);
        c.Dcp = p; p += al2(6 * (size_t)c.ncp);
        c.Wd = p; p += al2((size_t)(M > 1 ? M - 1 : 1) * 81);
        c.Wo = p; p += al2((size_t)(M > 2 ? M - 2 : 1) * 81);
        c.sg = p; p += al2(c.nr);
        c.sg2 = p; c.dinv = p; p += al2(c.nr > 32 ? c.nr : 32);   // the column buffer of the 9x9 routines shares sg2 (dead by then)
        c.segc = p; p += (size_t)M * SEGC;
        c.part = p; p += (size_t)c.nw * X1_PART * 32;
        c.scal = p; p += 2 * (size_t)X1_MAXW * 8;
        c.flag = (int *)p; p += 8;
        c.pv = p; p += 16;
        c.rows = p; p += (size_t)c.nslot * c.nw * c.cap * W1_ROWBLK;
        c.cnt = (int *)p; p += al2((size_t)c.nslot * c.nw * 16);
        c.nrm = p;
#pragma unroll 1
        for (int idx = tid; idx < M * SEGC; idx += nt) {
            const int m = idx / SEGC, o = idx - m * SEGC;
            const double *sm = c.segmat + (size_t)m * SEGMAT;
            c.segc[idx] = o < SEGC_CR ? sm[SEGMAT_CL + o] : (o < SEGC_RQ ? sm[SEGMAT_CR + o - SEGC_CR] : (o < SEGC_QS ? sm[SEGMAT_RQ + o - SEGC_RQ] : (o == SEGC_QS ? sm[SEGMAT_QS] : 0.0)));
        }
        __syncthreads();
    }
    const int iters = (S.mode == 0) ? S.iteration : 1;
    for (int iter = 0; iter < iters; iter++)
        for (int l = l_begin; l < l_end; l++) {
            c.qa = l;
            if (c.qa >= N) continue;
            int rec = (S.mode == 0 ? iter * S.nbatch : S.rec_offset) + l;
            int st = x1_solve_qp(c, S.max_iter, S.tol_gap, S.tol_res, S.qp_obj + (size_t)cidx * S.nrec + rec,
                                 S.qp_iters + (size_t)cidx * S.nrec + rec, S.qp_res + ((size_t)cidx * S.nrec + rec) * 4);
            if (tid == 0) {
                S.qp_status[(size_t)cidx * S.nrec + rec] = st;
                if (st != ST_OK) atomicCAS(&S.status[cidx], (int)ST_OK, st);
            }
            if (st != ST_OK && S.mode == 0) return;
            if (st != ST_OK && S.npeer <= 0) continue;
#pragma unroll 1
            for (int v = tid; v < 18 * M; v += nt) {   // dummy <- vals (L182-L184); with peers: next table of every rank
                int m = v / 18, r = v % 18, k = r / 6, i = r % 6;
                const size_t at = (size_t)c.qa * 18 * M + (size_t)k * 6 * M + m * 6 + i;
                if (S.npeer <= 0) { ctrl[at] = c.x[v]; continue; }
                const double val = (st == ST_OK) ? c.x[v] : c.ctrl_src[at];
#pragma unroll 1
                for (int p = 0; p < S.npeer; p++) S.peer_ctrl[p][(size_t)cidx * N * 18 * M + at] = val;
            }
            __syncthreads();
        }
    __syncthreads();
    peer_signal_done(S, tid == 0);
}

#endif
}  // namespace rbpe
