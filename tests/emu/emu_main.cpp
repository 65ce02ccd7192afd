// Entry points of the host emulation build: loop over instances instead of launching a grid.
#define SB_HOST_EMULATION 1
#include "cuda_shim.h"
#ifdef SB_HOST_EMULATION_GROUP
#include <thread>
#include <vector>
#include "cuda_shim_group.h"
#define SB_GROUP_SHARED_CTL 0
#ifndef SB_BLOCK
#define SB_BLOCK 32
#endif
#endif
#include "generated_problem.inc"     // generated __device__ functions + SB_NS/SB_NP/SB_ND
#include "sb_kernels.cuh"

#ifdef SB_HOST_EMULATION_GROUP
template <int NBLK>
static void emu_forward_group_t(const SbForwardArgs* a) {
    constexpr int G = sb::GROUP > 1 ? sb::GROUP : 2;
    if (sb::GROUP <= 1) return;
    emu_group.size = G;
    pthread_barrier_init(&emu_group.bar, nullptr, G);
    for (long long i = 0; i < a->B; ++i) {
        std::vector<std::thread> lanes;
        for (int r = 0; r < G; ++r)
            lanes.emplace_back([a, i, r]() {
                threadIdx.x = (unsigned)r; blockDim.x = 32; blockIdx.x = 0;
                sb::forward_instance_group<G, NBLK>(*a, i, true);
            });
        for (auto& t : lanes) t.join();
    }
    pthread_barrier_destroy(&emu_group.bar);
}
#endif

extern "C" {
void emu_forward(const SbForwardArgs* a) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < a->B; ++i) sb::forward_instance(*a, i, true);
}
void emu_forward_sens(const SbForwardArgs* a) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < a->B; ++i) sb::forward_sens_instance(*a, i, true);
}
void emu_tables(const SbTablesArgs* a) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < a->B; ++i)
        for (int idx = 0; idx < a->hist_cap; ++idx) sb::build_table_entry(*a, i, idx);
}
void emu_backward(const SbBackwardArgs* a) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < a->B; ++i) sb::backward_instance(*a, i, true);
}
// one segment of intervals [k0, k1) for every instance (tools/lane_efficiency.py walks a solve
// interval by interval to read the per-interval counters out of the carry buffers)
void emu_backward_flat(const SbBackwardArgs* a) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < a->B; ++i) sb::backward_instance_flat(*a, i, true);
}
#ifdef SB_FUND
void emu_backward_fund(const SbBackwardArgs* a) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < a->B; ++i) sb::backward_fund_instance(*a, i, true);
}
#endif
void emu_backward_unit(const SbBackwardArgs* a, int k0, int k1) {
    #pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < a->B; ++i) sb::backward_unit<false>(*a, i, true, k0, k1);
}
#ifdef SB_HOST_EMULATION_GROUP
// The grouped-lane drivers: one group (= one instance) at a time, its lanes as threads.
int emu_group_size() { return sb::GROUP; }
void emu_forward_group(const SbForwardArgs* a) { emu_forward_group_t<1>(a); }
#if SB_ND > 0
void emu_forward_sens_group(const SbForwardArgs* a) { emu_forward_group_t<1 + SB_ND>(a); }
#endif
void emu_backward_group(const SbBackwardArgs* a) {
    constexpr int G = sb::GROUP > 1 ? sb::GROUP : 2;
    if (sb::GROUP <= 1) return;
    emu_group.size = G;
    pthread_barrier_init(&emu_group.bar, nullptr, G);
    for (long long i = 0; i < a->B; ++i) {
        std::vector<std::thread> lanes;
        for (int r = 0; r < G; ++r)
            lanes.emplace_back([a, i, r]() {
                threadIdx.x = (unsigned)r; blockDim.x = 32; blockIdx.x = 0;
                sb::backward_unit_group<G>(*a, i, true, 0, a->n_t + 1);
            });
        for (auto& t : lanes) t.join();
    }
    pthread_barrier_destroy(&emu_group.bar);
}
#endif
void emu_eval(const SbEvalArgs* a) {
    for (long long i = 0; i < a->n; ++i) sb::eval_instance(*a, i);
}
int emu_hist_stride() { return sb::HIST_STRIDE; }   // what the cubin exports as sb_hist_stride
int emu_sizes(int* ns, int* np, int* nd) { *ns = SB_NS; *np = SB_NP; *nd = SB_ND; return 0; }
// sizeof the argument blocks as the C++ side sees them (checked against the ctypes mirrors)
void emu_arg_sizes(int* fwd, int* tab, int* bwd, int* ev) {
    *fwd = (int)sizeof(SbForwardArgs); *tab = (int)sizeof(SbTablesArgs);
    *bwd = (int)sizeof(SbBackwardArgs); *ev = (int)sizeof(SbEvalArgs);
}
}
