// tc5_probe.cu -- the Blackwell-native tensor path the north-star names, tried on the shape the joint-batch factorisation
// needs: one 128 x 128 x 64 block update  C = A B'  (the Schur-complement update of a 144-order reduced-Hessian block is a
// handful of these) on the 5th-generation tensor cores:
//     operands   fp32 in global memory -> TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) -> shared memory
//     MMA        tcgen05.mma.cta_group::1.kind::tf32, issued by one thread, accumulator in tensor memory (TMEM)
//     epilogue   tcgen05.ld (TMEM -> registers) -> global
// in two precisions: plain TF32 (1 MMA chain) and the 3 x TF32 split (A = Ah + Al, B = Bh + Bl;  Ah Bh + Ah Bl + Al Bh),
// and compared with an FP64 reference.  tcgen05 has no FP64 kind; this probe measures what the split buys (about fp32
// accuracy) and what it costs, which is what DESIGN.md section 4 bases its "FP64 tensor pipe (DMMA) for the interior-point
// factorisation" decision on.  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -o tc5_probe tc5_probe.cu
// Every wait is bounded (a lost completion must not hang the device).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

constexpr int TM = 128, TN = 128, TK = 64;        // tile
constexpr int KB = 32;                            // fp32 per 128-byte swizzle row
constexpr int NKB = TK / KB;                      // k blocks per operand
constexpr uint32_t TILE_BYTES = 128 * KB * 4;     // one k block of one operand: 128 rows x 128 B
constexpr int TMEM_COLS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 400000000LL) return false;
    }
    return true;
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B, bits [61,64)
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Args {
    float *C;              // [TM][TN]
    long long *cycles;     // [2]: TMA issue -> data landed, first MMA issue -> commit observed
    int *status;           // 0 ok, 1 TMA timeout, 2 MMA timeout
    int split;             // 0: TF32, 1: 3 x TF32
};

__global__ void __launch_bounds__(128, 1) tc5_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                                                     const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *sAh = smem, *sAl = smem + NKB * TILE_BYTES, *sBh = smem + 2 * NKB * TILE_BYTES, *sBl = smem + 3 * NKB * TILE_BYTES;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_tma)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_mma)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    bool ok = true;
    if (tid == 0) {
        const int nmat = a.split ? 4 : 2;
        t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_tma)), "r"((uint32_t)(nmat * NKB * TILE_BYTES)) : "memory");
        for (int kb = 0; kb < NKB; kb++) {
            const CUtensorMap *maps[4] = {&mAh, &mBh, &mAl, &mBl};
            uint8_t *dst[4] = {sAh, sBh, sAl, sBl};
            for (int m = 0; m < nmat; m++)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(dst[m] + kb * TILE_BYTES)), "l"(maps[m]), "r"(kb * KB), "r"(0), "r"(smem_u32(&bar_tma)) : "memory");
        }
    }
    if (warp == 0) {
        ok = mbar_wait(&bar_tma, 0);
        t1 = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok && lane == 0) {
            const uint32_t idesc = umma_idesc(TM, TN);
            const int npass = a.split ? 3 : 1;
            uint32_t acc = 0;
            t2 = clock64();
            for (int p = 0; p < npass; p++) {
                const uint8_t *A = (p == 2) ? sAl : sAh, *B = (p == 1) ? sBl : sBh;
                for (int kb = 0; kb < NKB; kb++)
                    for (int ks = 0; ks < KB / 8; ks++) {     // UMMA K = 8 for tf32 (32 bytes)
                        const uint64_t da = umma_desc(smem_u32(A + kb * TILE_BYTES + ks * 32));
                        const uint64_t db = umma_desc(smem_u32(B + kb * TILE_BYTES + ks * 32));
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
                        acc = 1;
                    }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
        }
        __syncwarp();
    }
    const bool ok2 = mbar_wait(&bar_mma, 0);
    if (tid == 0) t3 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok2) {
        // epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows of C; 32 columns per tcgen05.ld
        for (int c0 = 0; c0 < TN; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float *row = a.C + (size_t)(warp * 32 + lane) * TN + c0;
            for (int j = 0; j < 32; j++) row[j] = __uint_as_float(v[j]);
        }
    }
    if (tid == 0) {
        a.cycles[0] = t1 - t0;
        a.cycles[1] = t3 - t2;
        *a.status = !ok ? 1 : (!ok2 ? 2 : 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float tf32_round(float x) {   // round to nearest even on the 13 dropped mantissa bits
    uint32_t u; memcpy(&u, &x, 4);
    u += 0xFFFu + ((u >> 13) & 1u);
    u &= 0xFFFFE000u;
    float r; memcpy(&r, &u, 4);
    return r;
}

int main() {
    cudaSetDevice(0);
    cudaFree(0);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qr) != cudaSuccess || !encode) {
        printf("cuTensorMapEncodeTiled not available\n");
        return 1;
    }
    // operands with the dynamic range of a late interior-point factor column scaling: entries spread over 6 decades
    std::vector<double> A64(TM * TK), B64(TN * TK);
    srand(7);
    for (auto &v : A64) v = ((rand() / (double)RAND_MAX) - 0.5) * pow(10.0, (rand() % 7) - 3);
    for (auto &v : B64) v = ((rand() / (double)RAND_MAX) - 0.5) * pow(10.0, (rand() % 7) - 3);
    std::vector<float> Ah(TM * TK), Al(TM * TK), Bh(TN * TK), Bl(TN * TK);
    for (int i = 0; i < TM * TK; i++) { float x = (float)A64[i]; Ah[i] = tf32_round(x); Al[i] = tf32_round((float)(A64[i] - (double)Ah[i])); }
    for (int i = 0; i < TN * TK; i++) { float x = (float)B64[i]; Bh[i] = tf32_round(x); Bl[i] = tf32_round((float)(B64[i] - (double)Bh[i])); }
    std::vector<double> ref(TM * TN, 0.0), mag(TM * TN, 0.0);
    for (int i = 0; i < TM; i++)
        for (int j = 0; j < TN; j++) {
            double s = 0, m = 0;
            for (int k = 0; k < TK; k++) { s += A64[i * TK + k] * B64[j * TK + k]; m += fabs(A64[i * TK + k] * B64[j * TK + k]); }
            ref[i * TN + j] = s; mag[i * TN + j] = m;
        }
    float *dA[4];
    std::vector<float> *hosts[4] = {&Ah, &Al, &Bh, &Bl};
    CUtensorMap maps[4];
    for (int m = 0; m < 4; m++) {
        cudaMalloc(&dA[m], sizeof(float) * TM * TK);
        cudaMemcpy(dA[m], hosts[m]->data(), sizeof(float) * TM * TK, cudaMemcpyHostToDevice);
        cuuint64_t gdim[2] = {(cuuint64_t)TK, (cuuint64_t)TM}, gstr[1] = {(cuuint64_t)TK * 4};
        cuuint32_t box[2] = {KB, 128}, estr[2] = {1, 1};
        CUresult r = encode(&maps[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA[m], gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("tensor map %d: error %d\n", m, (int)r); return 1; }
    }
    Args a;
    cudaMalloc(&a.C, sizeof(float) * TM * TN);
    cudaMalloc(&a.cycles, 16);
    cudaMalloc(&a.status, 4);
    const size_t smem = 4 * NKB * TILE_BYTES + 1024;
    cudaFuncSetAttribute(tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int split = 0; split < 2; split++) {
        a.split = split;
        long long best[2] = {1LL << 60, 1LL << 60};
        int st = -1;
        std::vector<float> C(TM * TN);
        for (int rep = 0; rep < 5; rep++) {
            cudaMemset(a.C, 0, sizeof(float) * TM * TN);
            tc5_kernel<<<1, 128, smem>>>(maps[0], maps[1], maps[2], maps[3], a);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel error: %s\n", cudaGetErrorString(e)); return 1; }
            long long cyc[2];
            cudaMemcpy(cyc, a.cycles, 16, cudaMemcpyDeviceToHost);
            cudaMemcpy(&st, a.status, 4, cudaMemcpyDeviceToHost);
            if (st != 0) break;
            for (int i = 0; i < 2; i++) if (cyc[i] < best[i]) best[i] = cyc[i];
        }
        if (st != 0) { printf("%s: status %d (1 = TMA wait timed out, 2 = MMA wait timed out)\n", split ? "3xTF32" : "TF32", st); continue; }
        cudaMemcpy(C.data(), a.C, sizeof(float) * TM * TN, cudaMemcpyDeviceToHost);
        double emax = 0, erel = 0;
        for (int i = 0; i < TM * TN; i++) {
            double e = fabs((double)C[i] - ref[i]);
            if (e > emax) emax = e;
            if (e / mag[i] > erel) erel = e / mag[i];
        }
        const double flops = 2.0 * TM * TN * TK * (split ? 3 : 1);
        printf("%-7s  max |C - C64| = %.3e   max error / sum|a b| = %.3e   TMA load %lld cycles (%d KB)   MMA chain %lld cycles = %.1f useful GFLOP/s/SM at 1.965 GHz\n",
               split ? "3xTF32" : "TF32", emax, erel, best[0], (int)((split ? 4 : 2) * NKB * TILE_BYTES / 1024), best[1],
               2.0 * TM * TN * TK / (best[1] / 1.965e9) / 1e9);
        (void)flops;
    }
    return 0;
}
