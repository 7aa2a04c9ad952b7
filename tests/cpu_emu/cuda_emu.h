// cuda_emu.h -- TEST INFRASTRUCTURE ONLY (never built into or loaded by the product library).
//
// A tiny fiber emulator of the CUDA execution model so that swarm_simulator_b200/csrc/rbpe_kernels.cuh can be
// compiled with g++ and its control flow / indexing debugged in a container that has no GPU.  Every CUDA thread
// of a block is a ucontext fiber; __syncthreads / __syncwarp / __shfl_*_sync are cooperative barriers.  Blocks run
// one after another.  It is slow (seconds per small QP) and exists to catch logic errors before a GPU call; the
// parity tests proper (-m gpu) run the real kernels through the C ABI.
#pragma once
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define __restrict__ __restrict

namespace emu {
struct Idx { unsigned x, y, z; };
struct Barrier { int count = 0; unsigned gen = 0; int expected = 0; };
struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
    Barrier *wait = nullptr;
    unsigned wait_gen = 0;
    Idx tid{0, 0, 0};
};
struct State {
    std::vector<Fiber> fibers;
    int cur = 0;
    ucontext_t sched;
    Idx blockIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
    Barrier cta;
    std::vector<Barrier> warp;
    std::vector<double> xch;  // 32 slots x 2 doubles per warp
    unsigned char *smem = nullptr;
    std::function<void()> body;
};
inline State &S() { static State s; return s; }

inline void yield_to_sched() {
    State &s = S();
    swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}
inline void barrier_wait(Barrier &b) {
    State &s = S();
    unsigned g = b.gen;
    if (++b.count == b.expected) {
        b.count = 0;
        b.gen++;
        return;
    }
    Fiber &f = s.fibers[s.cur];
    f.wait = &b;
    f.wait_gen = g;
    while (b.gen == g) yield_to_sched();
    f.wait = nullptr;
}
inline void trampoline() {
    State &s = S();
    s.body();
    s.fibers[s.cur].done = true;
    yield_to_sched();
}

template <class F>
void launch(F kernel_body, unsigned grid, unsigned block, size_t smem_bytes) {
    State &s = S();
    s.gridDim = {grid, 1, 1};
    s.blockDim = {block, 1, 1};
    std::vector<unsigned char> smem(smem_bytes + 64);
    s.smem = (unsigned char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    s.body = kernel_body;
    if (s.fibers.size() < block) s.fibers.resize(block);
    for (unsigned b = 0; b < grid; b++) {
        s.blockIdx = {b, 0, 0};
        s.cta = Barrier();
        s.cta.expected = (int)block;
        int nw = (int)((block + 31) / 32);
        s.warp.assign(nw, Barrier());
        for (int w = 0; w < nw; w++) {
            int hi = (int)block - w * 32;
            s.warp[w].expected = hi < 32 ? hi : 32;
        }
        s.xch.assign((size_t)nw * 64, 0.0);
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = s.fibers[t];
            if (f.stack.empty()) f.stack.resize(256 * 1024);
            f.done = false;
            f.wait = nullptr;
            f.tid = {t, 0, 0};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack.data();
            f.ctx.uc_stack.ss_size = f.stack.size();
            f.ctx.uc_link = &s.sched;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
        unsigned alive = block;
        while (alive) {
            unsigned progressed = 0;
            for (unsigned t = 0; t < block; t++) {
                Fiber &f = s.fibers[t];
                if (f.done) continue;
                if (f.wait && f.wait->gen == f.wait_gen) continue;  // still blocked
                s.cur = (int)t;
                swapcontext(&s.sched, &f.ctx);
                progressed++;
                if (f.done) alive--;
            }
            if (!progressed && alive) {
                fprintf(stderr, "cuda_emu: deadlock in block %u (%u threads blocked; a barrier was not reached by all)\n", b, alive);
                abort();
            }
        }
    }
}
}  // namespace emu

#define threadIdx (emu::S().fibers[emu::S().cur].tid)
#define blockIdx (emu::S().blockIdx)
#define blockDim (emu::S().blockDim)
#define gridDim (emu::S().gridDim)
#define RBPE_DYN_SMEM(name) unsigned char *name = emu::S().smem

inline void __syncthreads() { emu::barrier_wait(emu::S().cta); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::barrier_wait(emu::S().warp[threadIdx.x >> 5]); }

template <class T>
inline T __shfl_down_sync(unsigned, T v, int delta) {
    static_assert(sizeof(T) <= 16, "emu shuffle payload");
    emu::State &s = emu::S();
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    char *slot = (char *)&s.xch[(size_t)w * 64];
    memcpy(slot + lane * 16, &v, sizeof(T));
    emu::barrier_wait(s.warp[w]);
    T r = v;
    if (lane + delta < 32) memcpy(&r, slot + (lane + delta) * 16, sizeof(T));
    emu::barrier_wait(s.warp[w]);
    return r;
}

template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) <= 16, "emu shuffle payload");
    emu::State &s = emu::S();
    int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    char *slot = (char *)&s.xch[(size_t)w * 64];
    memcpy(slot + lane * 16, &v, sizeof(T));
    emu::barrier_wait(s.warp[w]);
    T r;
    memcpy(&r, slot + (src & 31) * 16, sizeof(T));
    emu::barrier_wait(s.warp[w]);
    return r;
}

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int mask) {
    return __shfl_sync(0xffffffffu, v, (int)((threadIdx.x & 31) ^ mask));
}
inline int __any_sync(unsigned, int pred) {
    int v = pred ? 1 : 0, r = 0;
    for (int l = 0; l < 32; l++) r |= __shfl_sync(0xffffffffu, v, l);
    return r;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned v = pred ? 1u : 0u, r = 0;
    for (int l = 0; l < 32; l++) r |= __shfl_sync(0xffffffffu, v, l) << l;
    return r;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int atomicCAS(int *addr, int cmp, int val) {
    int old = *addr;
    if (old == cmp) *addr = val;
    return old;
}
inline int atomicAdd(int *addr, int val) {
    int old = *addr;
    *addr = old + val;
    return old;
}
