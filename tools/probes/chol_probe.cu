// chol_probe.cu -- exploratory: times chol32_warp / chol_tall of rbpe_blockla.cuh in isolation.
#include <cstdio>
#include <vector>
#include <cmath>
#include "../../swarm_simulator_b200/csrc/rbpe_kernels.cuh"
using namespace rbpe;
__global__ void __launch_bounds__(256, 1) k32(double *A, int ld, int wJ, double *X, long long *cyc, int nthreads_wait) {
    long long t0 = clock64();
    bool ok = true;
    if ((threadIdx.x >> 5) == 0) ok = chol32_warp(A, ld, wJ, X);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = ok; }
}
__global__ void __launch_bounds__(256, 2) ktall(int kp, double *D, double *O, const double *Pm, double *Linv, double *flag, long long *cyc) {
    long long t0 = clock64();
    bool ok = chol_tall(kp, D, O, Pm, Linv, flag);
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = ok; }
}
int main() {
    for (int kp : {40, 144, 288}) {
        std::vector<double> G((size_t)kp * kp), A((size_t)kp * kp), O((size_t)kp * kp), P((size_t)kp * kp);
        srand(1);
        for (auto &v : G) v = (rand() % 2001 - 1000) / 1000.0;
        for (int i = 0; i < kp; i++) for (int j = 0; j < kp; j++) { double s = 0; for (int k = 0; k < kp; k++) s += G[i * kp + k] * G[j * kp + k]; A[i * kp + j] = s + (i == j ? 4.0 * kp : 0); }
        for (auto &v : O) v = (rand() % 2001 - 1000) / 1000.0;
        for (auto &v : P) v = (rand() % 2001 - 1000) / 3000.0;
        double *dA, *dO, *dP, *dX, *dflag; long long *dc, hc[2];
        size_t nb = (size_t)kp * kp * 8;
        cudaMalloc(&dA, nb); cudaMalloc(&dO, nb); cudaMalloc(&dP, nb); cudaMalloc(&dX, 16 * 1024 * 8); cudaMalloc(&dflag, 8); cudaMalloc(&dc, 16);
        for (int rep = 0; rep < 2; rep++) {
            cudaMemcpy(dA, A.data(), nb, cudaMemcpyHostToDevice);
            k32<<<1, 256>>>(dA, kp, 32, dX, dc, 0);
            cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
        }
        printf("kp=%d chol32_warp (256-thread CTA, warp 0 works): %lld cycles ok=%lld (%s)\n", kp, hc[0], hc[1], cudaGetErrorString(cudaGetLastError()));
        for (int rep = 0; rep < 2; rep++) {
            cudaMemcpy(dA, A.data(), nb, cudaMemcpyHostToDevice); cudaMemcpy(dO, O.data(), nb, cudaMemcpyHostToDevice); cudaMemcpy(dP, P.data(), nb, cudaMemcpyHostToDevice);
            ktall<<<1, 256>>>(kp, dA, dO, dP, dX, dflag, dc);
            cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
        }
        std::vector<double> L((size_t)kp * kp);
        cudaMemcpy(L.data(), dA, nb, cudaMemcpyDeviceToHost);
        // check L L' = A - P P' (lower)
        double err = 0, nrm = 0;
        for (int i = 0; i < kp; i++) for (int j = 0; j <= i; j++) {
            double s = 0, pp = 0;
            for (int k = 0; k <= j; k++) s += L[i * kp + k] * L[j * kp + k];
            for (int k = 0; k < kp; k++) pp += P[i * kp + k] * P[j * kp + k];
            err = fmax(err, fabs(s - (A[i * kp + j] - pp))); nrm = fmax(nrm, fabs(A[i * kp + j]));
        }
        double fl = 2.0 * ((double)kp * kp * kp / 3 + (double)kp * kp * kp / 2 * 1 + (double)kp * kp * kp);
        printf("kp=%d chol_tall [D;O] with previous block: %lld cycles (%.1f us) ok=%lld  max|LL'-A|/|A| = %.2e  ~%.1f GFLOP/s\n", kp, hc[0], hc[0] / 1965.0, hc[1], err / nrm,
               fl / (hc[0] / 1.965e9) / 1e9);
    }
    return 0;
}
