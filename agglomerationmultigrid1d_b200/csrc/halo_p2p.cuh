// Slab-edge exchange over NVLink peer memory: the halo step of the sharded V-cycle without NCCL.
//
// Every rank keeps the vectors of its sharded levels (x[0], x[1], b, with their ghost slots) and a block of
// 64-bit receive counters in ONE device allocation, the "arena", which both slab neighbours map with CUDA IPC
// (amg1d_finalize).  An exchange is then two tiny kernels instead of an NCCL send / recv group:
//
//   k_halo_push   (producer, after a leg)  stores the rank's ghost_depth edge elements of up to two vectors
//                 straight into the neighbours' ghost slots (plain stores through NVLink), makes them visible
//                 system-wide (__threadfence_system) and adds 1 to the neighbours' receive counter of the channel;
//   k_halo_wait   (consumer, before the leg that reads those ghosts)  spins until its counters of the channel
//                 reach the handle's cycle number (`epoch`, a device counter that k_epoch_wait advances once per
//                 V-cycle - every cycle pushes exactly once on every channel, so counter == epoch means "this
//                 cycle's edge has landed"; the values are monotone, nothing is ever reset, and a CUDA graph
//                 replays the same kernels for every cycle).
//
// Channels: 2 l (after the down leg of sharded level l: pre-smoothed iterate + coarse right-hand side) and
// 2 l + 1 (after its up leg: the corrected iterate, which is also next cycle's incoming iterate on level 0).
// The only write-after-read hazard is on level 0 - a neighbour that is a whole leg ahead would overwrite the
// ghosts of the iterate buffer that this rank's last up leg still reads; deeper levels are ordered by the
// exchanges of the levels above them - and is closed by k_epoch_wait: nobody starts cycle c before both
// neighbours' up-leg push of cycle c - 1 on level 0 has arrived, which they issue after that up leg completed.
//
// A neighbour that never arrives (crashed rank) must not hang the GPU: every spin gives up after
// AMG1D_P2P_TIMEOUT_CYCLES clock ticks and raises the handle's error word, which the host reports as
// AMG1D_ERR_NCCL at the next synchronising call.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#define AMG1D_P2P_TIMEOUT_CYCLES (20LL * 1000 * 1000 * 1000)   // ~10 s at 2 GHz

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void spin_until(const unsigned long long* flag, unsigned long long target, int* err) {
    if (!flag) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < target) {
        if (clock64() - t0 > AMG1D_P2P_TIMEOUT_CYCLES) { atomicExch(err, 1); break; }
        __nanosleep(64);
    }
}

struct HaloPush {
    const double* src[2];        // local vectors (element 0); src[1] may be null
    double* dst_left[2];         // left neighbour's slot for MY first elements  = its element n_left (right ghosts)
    double* dst_right[2];        // right neighbour's slot for MY last elements  = its element -gd   (left ghosts)
    long long n[2];              // owned elements of the vector on this rank
    int m[2];                    // block size
    int gd;                      // ghost depth (elements per edge)
    unsigned long long* flag_left;    // receive counters of this channel in the neighbours' arenas (or null)
    unsigned long long* flag_right;
};

__global__ void k_halo_push(HaloPush a) {
    const int t = threadIdx.x;
    for (int p = 0; p < 2; ++p) {
        if (!a.src[p]) continue;
        const int cnt = a.gd * a.m[p];
        if (a.dst_left[p])
            for (int i = t; i < cnt; i += blockDim.x) a.dst_left[p][i] = a.src[p][i];
        if (a.dst_right[p])
            for (int i = t; i < cnt; i += blockDim.x) a.dst_right[p][i] = a.src[p][(a.n[p] - a.gd) * a.m[p] + i];
    }
    __threadfence_system();          // every writer orders its own stores before the signal
    __syncthreads();
    if (t == 0) {
        if (a.flag_left) atomicAdd_system(a.flag_left, 1ULL);
        if (a.flag_right) atomicAdd_system(a.flag_right, 1ULL);
    }
}

// consumer: both receive counters of a channel must have reached epoch - lag
__global__ void k_halo_wait(const unsigned long long* flag_left, const unsigned long long* flag_right,
                            const unsigned long long* epoch, unsigned long long lag, int* err) {
    const unsigned long long target = *epoch - lag;
    spin_until(flag_left, target, err);
    spin_until(flag_right, target, err);
}

// start of a V-cycle: advance the cycle number, then wait for the neighbours' level-0 up-leg push of the
// previous cycle (see the header: iterate ghosts + the level-0 write-after-read hazard)
__global__ void k_epoch_wait(unsigned long long* epoch, const unsigned long long* flag_left,
                             const unsigned long long* flag_right, int* err) {
    const unsigned long long e = *epoch + 1;
    *epoch = e;
    spin_until(flag_left, e - 1, err);
    spin_until(flag_right, e - 1, err);
}
