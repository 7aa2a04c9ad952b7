// dmma_probe.cu -- exploratory (not product): FP64 mma.sync fragment layouts and throughput on sm_100a.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

// C (MxN) = A (MxK) * B^T where Bt is N x K row-major ("NT")
__global__ void layout_kernel(const double *A, const double *Bt, double *C884, double *C1688, double *C1684) {
    int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    {   // m8n8k4: A 8x4 (ld 8), Bt 8x4 (ld 8): use k = 0..3
        double d0 = 0, d1 = 0;
        mma884(d0, d1, A[g * 8 + t], Bt[g * 8 + t]);
        C884[g * 8 + 2 * t] = d0; C884[g * 8 + 2 * t + 1] = d1;
    }
    {   // m16n8k8: A 16x8 (ld 8), Bt 8x8
        double c[4] = {0, 0, 0, 0};
        double a[4] = {A[g * 8 + t], A[(g + 8) * 8 + t], A[g * 8 + t + 4], A[(g + 8) * 8 + t + 4]};
        double b[2] = {Bt[g * 8 + t], Bt[g * 8 + t + 4]};
        mma1688(c, a, b);
        C1688[g * 8 + 2 * t] = c[0]; C1688[g * 8 + 2 * t + 1] = c[1];
        C1688[(g + 8) * 8 + 2 * t] = c[2]; C1688[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
    {   // m16n8k4
        double c[4] = {0, 0, 0, 0};
        double a[2] = {A[g * 8 + t], A[(g + 8) * 8 + t]};
        mma1684(c, a, Bt[g * 8 + t]);
        C1684[g * 8 + 2 * t] = c[0]; C1684[g * 8 + 2 * t + 1] = c[1];
        C1684[(g + 8) * 8 + 2 * t] = c[2]; C1684[(g + 8) * 8 + 2 * t + 1] = c[3];
    }
}

template <int SHAPE>
__global__ void thr_kernel(double *out, int iters, double seed) {
    double acc[8][4];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0;
    double a4[4] = {seed, seed * 0.5, seed * 0.25, seed * 2}, b2[2] = {seed, -seed}, a2[2] = {seed, seed * 0.5};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (SHAPE == 0) { mma884(acc[i][0], acc[i][1], a4[0], b2[0]); mma884(acc[i][2], acc[i][3], a4[1], b2[1]); }
            if (SHAPE == 1) mma1688(acc[i], a4, b2);
            if (SHAPE == 2) mma1684(acc[i], a2, b2[0]);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void thr_dfma(double *out, int iters, double seed) {
    double acc[16];
    for (int i = 0; i < 16; i++) acc[i] = i;
    double a = seed, b = 1.0 - 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], b, a);
    }
    double s = 0;
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double hA[128], hB[64], *dA, *dB, *dC;
    for (int i = 0; i < 128; i++) hA[i] = (double)((i * 37) % 11) - 5 + 0.25 * (i % 3);
    for (int i = 0; i < 64; i++) hB[i] = (double)((i * 53) % 7) - 3 + 0.5 * (i % 2);
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dC, 3 * 128 * 8);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    layout_kernel<<<1, 32>>>(dA, dB, dC, dC + 128, dC + 256);
    double hC[384];
    cudaError_t e = cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost);
    printf("layout kernel: %s\n", cudaGetErrorString(e));
    double err884 = 0, err1688 = 0, err1684 = 0;
    for (int i = 0; i < 16; i++) for (int j = 0; j < 8; j++) {
        double r4 = 0, r8 = 0;
        for (int k = 0; k < 8; k++) { double p = hA[i * 8 + k] * hB[j * 8 + k]; r8 += p; if (k < 4) r4 += p; }
        if (i < 8) err884 = fmax(err884, fabs(hC[i * 8 + j] - r4));
        err1688 = fmax(err1688, fabs(hC[128 + i * 8 + j] - r8));
        err1684 = fmax(err1684, fabs(hC[256 + i * 8 + j] - r4));
    }
    printf("layout errors: m8n8k4 %.3g  m16n8k8 %.3g  m16n8k4 %.3g\n", err884, err1688, err1684);
    int dev_sms = 0; cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    double *out; cudaMalloc(&out, (size_t)dev_sms * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int shape = 0; shape < 4; shape++) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (shape == 0) thr_kernel<0><<<dev_sms, warps * 32>>>(out, iters, 1e-3);
                if (shape == 1) thr_kernel<1><<<dev_sms, warps * 32>>>(out, iters, 1e-3);
                if (shape == 2) thr_kernel<2><<<dev_sms, warps * 32>>>(out, iters, 1e-3);
                if (shape == 3) thr_dfma<<<dev_sms, warps * 32>>>(out, iters, 1e-3);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            // flops per warp per iteration
            double fl = shape == 0 ? 16.0 * 512 : shape == 1 ? 8.0 * 2048 : shape == 2 ? 8.0 * 1024 : 16.0 * 64;
            double tf = fl * iters * warps * dev_sms / (best * 1e-3) / 1e12;
            printf("warps/SM %2d %-8s: %.3f ms  %.2f TFLOP/s\n", warps, shape == 0 ? "m8n8k4" : shape == 1 ? "m16n8k8" : shape == 2 ? "m16n8k4" : "dfma", best, tf);
        }
    }
    return 0;
}
