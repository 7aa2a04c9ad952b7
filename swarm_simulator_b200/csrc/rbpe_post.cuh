// rbpe_post.cuh -- the two pure-arithmetic neighbours of the hot path:
//   rsfc_kernel     Corridor::updateRelBox   (rbp_corridor.hpp L338-L398): RSFC normals from initTraj, FLOAT32 with
//                   octomath::Vector3 operator semantics; every float operation is an explicit round-to-nearest
//                   intrinsic so that nvcc cannot contract a*b+c into an FMA (results are bit-identical to the oracle)
//   metrics_kernel  RBPPublisher checks      (rbp_publisher.hpp L169-L183, L670-L695, L769-L798): sampled positions,
//                   per-sample minimum safety-margin ratio, per-(agent, sample) polyline lengths
#pragma once
#include "rbpe_types.h"

namespace rbpe {

struct RsfcArgs {
    int count, N, M;
    double downwash;
    const float *init_traj;   // [count][N][M+1][3]
    const double *T;          // [count][M+1]
    float *rsfc_n;            // [count][P][M][3]
    double *rsfc_t;           // [count][P][M]
    int *collided;            // [count]
};
struct MetricsArgs {
    int count, N, M, nt_max;
    double downwash, dt;
    const double *coef;       // [count][N][3][6M]
    const double *T;          // [count][M+1]
    const double *radius;     // [count][N]
    double *ratio_t;          // [count][nt_max]        min over pairs at sample i (1e9 beyond the mission's samples)
    double *seglen;           // [count][N][nt_max]     |p(t_{i+1}) - p(t_i)| (0 beyond)
};

#if defined(__CUDACC__)

__device__ __forceinline__ double v3f_norm(float x, float y, float z) {
    float s = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    return sqrt((double)s);
}
__device__ __forceinline__ double v3f_dot(float ax, float ay, float az, float bx, float by, float bz) {
    return (double)__fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

__global__ void rsfc_kernel(RsfcArgs A) {
    const int N = A.N, M = A.M;
    const long per = (long)N * N * M, total = per * A.count;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
        int c = (int)(g / per);
        long w = g - (long)c * per;
        int iter = (int)(w % M) + 1, qj = (int)((w / M) % N), qi = (int)(w / ((long)M * N));
        if (qj <= qi) continue;
        const float *tr = A.init_traj + (size_t)c * N * (M + 1) * 3;
        const float *pi0 = tr + ((size_t)qi * (M + 1) + iter - 1) * 3, *pj0 = tr + ((size_t)qj * (M + 1) + iter - 1) * 3;
        float ax = __fsub_rn(pj0[0], pi0[0]), ay = __fsub_rn(pj0[1], pi0[1]), az = __fsub_rn(pj0[2], pi0[2]);
        float bx = __fsub_rn(pj0[3], pi0[3]), by = __fsub_rn(pj0[4], pi0[4]), bz = __fsub_rn(pj0[5], pi0[5]);
        az = (float)((double)az / A.downwash);
        bz = (float)((double)bz / A.downwash);
        float mx = ax, my = ay, mz = az;
        if (!(ax == bx && ay == by && az == bz)) {
            double dist_min = v3f_norm(ax, ay, az), dist = v3f_norm(bx, by, bz);
            if (dist_min > dist) { mx = bx; my = by; mz = bz; dist_min = dist; }
            float nx = __fsub_rn(bx, ax), ny = __fsub_rn(by, ay), nz = __fsub_rn(bz, az);
            double len = v3f_norm(nx, ny, nz);
            if (len > 0) { float f = (float)len; nx = __fdiv_rn(nx, f); ny = __fdiv_rn(ny, f); nz = __fdiv_rn(nz, f); }
            float f = (float)v3f_dot(ax, ay, az, nx, ny, nz);
            float cx = __fsub_rn(ax, __fmul_rn(nx, f)), cy = __fsub_rn(ay, __fmul_rn(ny, f)), cz = __fsub_rn(az, __fmul_rn(nz, f));
            dist = v3f_norm(cx, cy, cz);
            double dd = v3f_dot(__fsub_rn(cx, ax), __fsub_rn(cy, ay), __fsub_rn(cz, az), __fsub_rn(cx, bx), __fsub_rn(cy, by), __fsub_rn(cz, bz));
            if (dd < 0 && dist_min > dist) { mx = cx; my = cy; mz = cz; }
        }
        double len = v3f_norm(mx, my, mz);
        if (len > 0) { float f = (float)len; mx = __fdiv_rn(mx, f); my = __fdiv_rn(my, f); mz = __fdiv_rn(mz, f); }
        mz = (float)((double)mz / A.downwash);
        if (v3f_norm(mx, my, mz) == 0) A.collided[c] = 1;
        long it = pair_index(N, qi, qj);
        const size_t P = (size_t)N * (N - 1) / 2;
        float *o = A.rsfc_n + (((size_t)c * P + it) * M + iter - 1) * 3;
        o[0] = mx; o[1] = my; o[2] = mz;
        A.rsfc_t[((size_t)c * P + it) * M + iter - 1] = A.T[(size_t)c * (M + 1) + iter];
    }
}

__device__ __forceinline__ void eval_pos_dev(const double *coef, int M, const double *T, double t, double *p) {
    int index = 0;
    double tseg = 0;
    for (int m = 0; m < M; m++) {
        if (T[m] < t) { tseg = T[m]; index = m; } else break;
    }
    tseg = __dsub_rn(t, tseg);
    for (int k = 0; k < 3; k++) {
        const double *c = coef + (size_t)k * 6 * M + index * 6;
        double s = 0, pw = 1;
        for (int j = 0; j < 6; j++) { s = __dadd_rn(s, __dmul_rn(c[5 - j], pw)); pw = __dmul_rn(pw, tseg); }
        p[k] = s;
    }
}

// one CTA per (mission, sample): positions of all agents at t_i and t_{i+1} in shared memory, then pairs / segments
__global__ void metrics_kernel(MetricsArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *p0 = (double *)smem_raw, *p1 = p0 + (size_t)A.N * 3;
    double *red = p1 + (size_t)A.N * 3;
    const int c = blockIdx.y, i = blockIdx.x, N = A.N, M = A.M;
    const double *T = A.T + (size_t)c * (M + 1);
    const int nt = (int)floor(T[M] / A.dt);
    if (i >= nt) {
        if (threadIdx.x == 0) A.ratio_t[(size_t)c * A.nt_max + i] = 1e9;
        for (int q = threadIdx.x; q < N; q += blockDim.x) A.seglen[((size_t)c * N + q) * A.nt_max + i] = 0;
        return;
    }
    for (int q = threadIdx.x; q < N; q += blockDim.x) {
        const double *cf = A.coef + ((size_t)c * N + q) * 18 * M;
        eval_pos_dev(cf, M, T, __dmul_rn((double)i, A.dt), p0 + q * 3);
        eval_pos_dev(cf, M, T, __dmul_rn((double)(i + 1), A.dt), p1 + q * 3);
    }
    __syncthreads();
    const double *rad = A.radius + (size_t)c * N;
    double best = 1e9;
    for (long pr = threadIdx.x; pr < (long)N * N; pr += blockDim.x) {
        int qi = (int)(pr / N), qj = (int)(pr % N);
        if (qj <= qi) continue;
        double dx = __dsub_rn(p0[qi * 3], p0[qj * 3]), dy = __dsub_rn(p0[qi * 3 + 1], p0[qj * 3 + 1]);
        double dz = __dsub_rn(p0[qi * 3 + 2], p0[qj * 3 + 2]) / A.downwash;
        double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        double ratio = sqrt(d2) / __dadd_rn(rad[qi], rad[qj]);
        best = fmin(best, ratio);
    }
    for (int o = 16; o > 0; o >>= 1) best = fmin(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = fmin(best, red[w]);
        A.ratio_t[(size_t)c * A.nt_max + i] = best;
    }
    for (int q = threadIdx.x; q < N; q += blockDim.x) {
        double len = 0;
        if (i + 1 < nt) {
            double dx = __dsub_rn(p1[q * 3], p0[q * 3]), dy = __dsub_rn(p1[q * 3 + 1], p0[q * 3 + 1]), dz = __dsub_rn(p1[q * 3 + 2], p0[q * 3 + 2]);
            len = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        }
        A.seglen[((size_t)c * N + q) * A.nt_max + i] = len;
    }
}
#endif
}  // namespace rbpe
