// Host stand-ins for the warp intrinsics the grouped-lane code (csrc/sb_group.cuh) uses.  TEST
// INFRASTRUCTURE ONLY (see cuda_shim.h).  The G lanes of ONE group are G host threads; every
// shuffle / vote / __syncwarp is a rendezvous on a pthread barrier with the exchange going through
// a small slot array.  Between those points the threads run freely (no lockstep), which is why the
// emulation keeps the per-instance controller record private to each lane
// (-DSB_GROUP_SHARED_CTL=0); what it checks is the cross-lane part: butterfly sums, the LU across
// the lanes, the shared-memory exchange of the evaluation vectors and the driver.  The
// lockstep-dependent sharing of the controller record is verified on the GPU (-DSB_GROUP_CHECK).
#pragma once
#include <pthread.h>

#define __shared__ static

struct EmuGroup {
    pthread_barrier_t bar;
    double dslot[32];
    int islot[32];
    int size;
};
static EmuGroup emu_group;

static inline void emu_rendezvous() { pthread_barrier_wait(&emu_group.bar); }
static inline int emu_rank() { return (int)(threadIdx.x & 31u); }

static inline void __syncwarp(unsigned) { emu_rendezvous(); }
static inline unsigned __activemask() { return 0xffffffffu; }
static inline double __shfl_sync(unsigned, double v, int src, int width) {
    emu_group.dslot[emu_rank()] = v;
    emu_rendezvous();
    const int base = emu_rank() & ~(width - 1);
    const double r = emu_group.dslot[base + (src & (width - 1))];
    emu_rendezvous();
    return r;
}
static inline double __shfl_xor_sync(unsigned, double v, int o, int width) {
    emu_group.dslot[emu_rank()] = v;
    emu_rendezvous();
    const double r = emu_group.dslot[emu_rank() ^ o];
    (void)width;
    emu_rendezvous();
    return r;
}
static inline int __shfl_xor_sync(unsigned, int v, int o, int width) {
    emu_group.islot[emu_rank()] = v;
    emu_rendezvous();
    const int r = emu_group.islot[emu_rank() ^ o];
    (void)width;
    emu_rendezvous();
    return r;
}
static inline int __all_sync(unsigned, int pred) {
    emu_group.islot[emu_rank()] = pred;
    emu_rendezvous();
    int all = 1;
    for (int i = 0; i < emu_group.size; ++i) all = all && emu_group.islot[i];
    emu_rendezvous();
    return all;
}
