// fp64_latency_probe.cu -- dependent-issue latencies that bound the latency kernels (one warp alone on its scheduler):
// DFMA, DADD, DMUL, MUFU.RCP64H + 2 Newton steps, 64-bit SHFL, LDS -> DFMA.  clock64 around 2048-long dependent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency_probe fp64_latency_probe.cu && ./fp64_latency_probe
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void probe(double *out, long long *cyc, double a, double b) {
    __shared__ double sm[64];
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncthreads();
    double x = a + threadIdx.x;
    long long t0, t1;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, b, a);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x + a;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) {   // reciprocal: seed + two Newton steps (rcp_nr of the kernels)
        double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
        double e = fma(-x, r, 1.0); r = fma(r, e, r); e = fma(-x, r, 1.0); r = fma(r, e, r);
        x = r + a;
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_xor_sync(0xffffffffu, x, 1);
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { int j = ((int)__double2loint(x)) & 31; x = sm[j] ; }   // dependent LDS (address from the value)
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) x = rsqrt(x) + a;
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = (x > a) ? x : a + x;   // DSETP + select + DADD
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    out[threadIdx.x] = x;
}
int main() {
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 64);
    for (int rep = 0; rep < 2; rep++) probe<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char *nm[8] = {"DFMA", "DADD", "DMUL", "rcp_nr + DADD", "SHFL.64", "LDS.64 (+I2 addr)", "rsqrt + DADD", "DSETP+SEL+DADD"};
    for (int i = 0; i < 8; i++) printf("%-20s %.1f cycles per dependent op\n", nm[i], (double)h[i] / N);
    return 0;
}
