// Minimal CPU emulation of the CUDA constructs csrc/fem2d.cu uses, so that its KERNELS -- not just their
// arithmetic -- run on the CPU: one std::thread per CUDA thread of a block, pthread barriers for __syncthreads and
// for the warp exchanges, compare-and-swap for atomicAdd(double*).  One block at a time.  TEST INFRASTRUCTURE ONLY
// (oracle/fem2d_emu.cpp); it checks indexing, phase order and barrier placement, not performance, and a data race
// would show up only by chance.
#pragma once
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstring>

struct EmuDim {
    unsigned x, y, z;
};
extern thread_local EmuDim threadIdx;
extern thread_local EmuDim blockIdx;
extern EmuDim blockDim;
extern pthread_barrier_t emu_block_barrier;
extern pthread_barrier_t emu_warp_barrier[32];
extern double emu_warp_buf[32][32];

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n)
#define __shared__

inline void __syncthreads() { pthread_barrier_wait(&emu_block_barrier); }

inline void __syncwarp() { pthread_barrier_wait(&emu_warp_barrier[threadIdx.x >> 5]); }

inline double __shfl_xor_sync(unsigned, double v, int d) {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    emu_warp_buf[w][l] = v;
    pthread_barrier_wait(&emu_warp_barrier[w]);
    const double r = emu_warp_buf[w][l ^ d];
    pthread_barrier_wait(&emu_warp_barrier[w]);
    return r;
}

inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

inline double atomicAdd(double* p, double v) {
    uint64_t* q = reinterpret_cast<uint64_t*>(p);
    uint64_t old = __atomic_load_n(q, __ATOMIC_RELAXED), want;
    double o;
    do {
        std::memcpy(&o, &old, 8);
        const double n = o + v;
        std::memcpy(&want, &n, 8);
    } while (!__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return o;
}
