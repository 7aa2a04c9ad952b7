// lat_probe.cu -- exploratory: dependent-chain latencies of the FP64 building blocks used by the factorisations.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed) {
    __shared__ double sm[2048];
    int lane = threadIdx.x;
    if (threadIdx.x >= 32) { __syncthreads(); return; }   // other warps wait at the CTA barrier, as in chol_tall
    for (int i = lane; i < 2048; i += 32) sm[i] = 1e-3 * i;
    __syncwarp();
    double x = seed + lane;
    long long t0, t1;
    // 1. dependent DFMA chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) x = fma(x, 1.0000001, 1e-9);
    t1 = clock64(); if (lane == 0) cyc[0] = (t1 - t0) / 64;
    // 2. dependent rsqrt chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) x = rsqrt(x + 1.5);
    t1 = clock64(); if (lane == 0) cyc[1] = (t1 - t0) / 16;
    // 3. dependent shfl (double) chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + 1.0;
    t1 = clock64(); if (lane == 0) cyc[2] = (t1 - t0) / 16;
    // 4. STS + syncwarp + LDS round trip chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) { sm[lane] = x; __syncwarp(); x = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
    t1 = clock64(); if (lane == 0) cyc[3] = (t1 - t0) / 16;
    // 5. 31 independent LDS broadcast + DFMA pairs (the column update)
    double a[32];
#pragma unroll
    for (int i = 0; i < 32; i++) a[i] = x + i;
    t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        const double *col = sm + j * 33;
        double a0 = a[0] * 0.5;
#pragma unroll
        for (int kk = 1; kk < 32; kk++) a[kk - 1] = a[kk] - a0 * col[kk * 32];
        a[31] = 0;
    }
    t1 = clock64(); if (lane == 0) cyc[4] = (t1 - t0) / 16;
    // 6. the full factor step: shfl pivot, rsqrt, scale, STS, syncwarp, update
    t0 = clock64();
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        double piv = __shfl_sync(0xffffffffu, a[0], j);
        if (!(piv > 0)) piv = 1.0;
        double inv = rsqrt(piv), a0 = a[0] * inv;
        sm[lane * 32 + j] = lane >= j ? a0 : 0.0;
        if (lane == j) sm[2040] = inv;
        __syncwarp();
        const double *col = sm + j * 33;
#pragma unroll
        for (int kk = 1; kk < 32; kk++) a[kk - 1] = a[kk] - a0 * col[kk * 32];
        a[31] = 0;
    }
    t1 = clock64(); if (lane == 0) cyc[5] = (t1 - t0) / 16;
    // 7. DMMA dependent chain
    double d0 = 0, d1 = 0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 32; i++) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(x), "d"(a[1]));
    t1 = clock64(); if (lane == 0) cyc[6] = (t1 - t0) / 32;
    // 8. global (L2) dependent load chain: pointer chase through `out`
    double s = 0;
    for (int i = 0; i < 32; i++) s += a[i];
    out[lane] = x + s + d0 + d1;
    __syncthreads();
}
__global__ void chase(const int *p, int n, long long *cyc, int *sink) {
    int i = 0; long long t0 = clock64();
    for (int k = 0; k < n; k++) i = p[i];
    long long t1 = clock64(); cyc[7] = (t1 - t0) / n; *sink = i;
}
int main() {
    double *out; long long *cyc, h[8]; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 64);
    for (int nt = 32; nt <= 256; nt *= 8) { k<<<1, nt>>>(out, cyc, 1.25); k<<<1, nt>>>(out, cyc, 1.25); cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost); printf("threads %d: dfma %lld rsqrt %lld shfl %lld sts-lds %lld update %lld factor-step %lld dmma %lld\n", nt, h[0], h[1], h[2], h[3], h[4], h[5], h[6]); }
    // pointer chase over 8 MB (L2 resident after the first pass), stride 4 KB
    int n = 2048, *p, *hp = new int[n * 1024], *sink; 
    for (int i = 0; i < n; i++) hp[i * 1024] = ((i + 1) % n) * 1024;
    cudaMalloc(&p, n * 1024 * 4); cudaMalloc(&sink, 4); cudaMemcpy(p, hp, n * 1024 * 4, cudaMemcpyHostToDevice);
    chase<<<1, 1>>>(p, n, cyc, sink); chase<<<1, 1>>>(p, n, cyc, sink);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    printf("cycles: dfma chain %lld | rsqrt chain %lld | shfl64 chain %lld | sts+sync+lds+sync %lld | 31 lds+dfma %lld | full factor step %lld | dmma chain %lld | L2 load chain %lld\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    return 0;
}
