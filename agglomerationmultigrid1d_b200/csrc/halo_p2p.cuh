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
// FUSED FORM (the register-resident legs f_down / f_up, kernels_fused.cuh).  The stand-alone kernels above cost a
// launch and a serialisation point per exchange.  The fused legs therefore do both halves themselves (HaloLeg):
// the threads that own one of the ghost_depth edge elements store it to the neighbour as well, fence and add 1 to
// the neighbour's counter - the exchange overlaps the rest of the grid - and the CTAs whose window reaches the
// slab edge poll the counters they depend on right after the programmatic-dependency wait, while every other CTA
// of the leg is already computing.  Counters therefore count ELEMENTS: a channel receives q = ghost_depth (+
// ghost_depth coarse elements on a down leg whose coarse level is sharded too) per cycle and side, and "arrived"
// means counter >= epoch * q.  Levels whose legs are not f_down / f_up (row-per-thread legs) use the kernels above.
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

// Per-launch description of the fused exchange of one leg (all pointers null / on = 0: no exchange).
struct HaloLeg {
    int on;
    int gd;                                   // ghost depth
    // consumer: counters that must reach (epoch - lag) * q before this leg reads its ghosts
    const unsigned long long* w_left[2];
    const unsigned long long* w_right[2];
    unsigned long long wq[2], wlag[2];
    const unsigned long long* epoch;
    int* err;
    // producer: where this leg's edge elements go
    double* x_left;                           // left neighbour's slot for my element 0 of the output iterate
    double* x_right;                          // right neighbour's slot for my element n - gd
    double* c_left;                           // (down leg) the same for the coarse right-hand side
    double* c_right;
    unsigned long long* f_left;               // the neighbours' receive counters of this leg's channel
    unsigned long long* f_right;
    long long nc;                             // owned coarse elements
};

// edge CTAs: one thread polls, the CTA follows (call from all threads of the CTA; the condition is CTA-uniform)
__device__ __forceinline__ void halo_leg_wait(const HaloLeg& hl, bool touch_left, bool touch_right) {
    if (!hl.on || !(touch_left || touch_right)) return;
    if (threadIdx.x == 0) {
        const unsigned long long e = *hl.epoch;
        for (int i = 0; i < 2; ++i) {
            const unsigned long long target = (e - hl.wlag[i]) * hl.wq[i];
            if (touch_left) spin_until(hl.w_left[i], target, hl.err);
            if (touch_right) spin_until(hl.w_right[i], target, hl.err);
        }
    }
    __syncthreads();
}

// the owner of edge element e (0 <= e < n) of a vector with block size M sends it on
template <int M>
__device__ __forceinline__ void halo_leg_push(const double (&v)[M], long long e, long long n, int gd, double* left,
                                              double* right, unsigned long long* f_left, unsigned long long* f_right) {
    if (left && e < gd) {
#pragma unroll
        for (int i = 0; i < M; ++i) left[e * M + i] = v[i];
        __threadfence_system();
        atomicAdd_system(f_left, 1ULL);
    }
    if (right && e >= n - gd) {
#pragma unroll
        for (int i = 0; i < M; ++i) right[(e - (n - gd)) * M + i] = v[i];
        __threadfence_system();
        atomicAdd_system(f_right, 1ULL);
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
    if (t == 0) {                    // counters count elements (see the fused form)
        const unsigned long long q = (unsigned long long)a.gd * ((a.src[0] ? 1 : 0) + (a.src[1] ? 1 : 0));
        if (a.flag_left) atomicAdd_system(a.flag_left, q);
        if (a.flag_right) atomicAdd_system(a.flag_right, q);
    }
}

// consumer: both receive counters of a channel must have reached epoch - lag
__global__ void k_halo_wait(const unsigned long long* flag_left, const unsigned long long* flag_right,
                            const unsigned long long* epoch, unsigned long long lag, unsigned long long q, int* err) {
    const unsigned long long target = (*epoch - lag) * q;
    spin_until(flag_left, target, err);
    spin_until(flag_right, target, err);
}

// start of a V-cycle: advance the cycle number, then wait for the neighbours' level-0 up-leg push of the
// previous cycle (see the header: iterate ghosts + the level-0 write-after-read hazard)
__global__ void k_epoch_wait(unsigned long long* epoch, const unsigned long long* flag_left,
                             const unsigned long long* flag_right, unsigned long long q, int* err) {
    const unsigned long long e = *epoch + 1;
    *epoch = e;
    spin_until(flag_left, (e - 1) * q, err);
    spin_until(flag_right, (e - 1) * q, err);
}
