// sb_kernels.cuh -- kernel entry points.  This file is appended (by #include) to the generated
// problem functions (sb_rhs, sb_jac, sb_adj_rhs, sb_adj_jac, sb_quad_rhs and SB_NS/SB_NP/SB_ND,
// see sunode_b200/symode/codegen.py) and compiled as one translation unit for sm_100a, so the
// problem functions are inlined into the integrator -- the device-side replacement of the
// reference's numba @cfunc trampolines (/root/reference/sunode/problem.py:156-383).
//
//   sb_forward   K1: Solver.solve / AdjointSolver.solve_forward   (solver.py:467-527, 682-721)
//   sb_tables        interpolation tables for the stored forward steps (CVODES CV_POLYNOMIAL data)
//   sb_backward  K2: AdjointSolver.solve_backward                  (solver.py:723-784)
//   sb_eval          batched evaluation of the generated functions (as_pytensor.py:160-183 EvalRhs)
//
// One thread integrates one instance; instances are independent, so there is no inter-thread
// communication at all.  Inputs/outputs are instance-major (a batch-1 view is byte-compatible
// with the reference's [n_t, n_s] C-order buffers).
#pragma once
#include "sb_args.h"
#include "sb_bdf.cuh"

#ifndef SB_BLOCK
#define SB_BLOCK 128
#endif
#ifndef SB_MIN_BLOCKS
#define SB_MIN_BLOCKS 1
#endif
#ifndef SB_TAB_PREFETCH_MAX_NS
#define SB_TAB_PREFETCH_MAX_NS 4
#endif
#ifndef SB_GROUP_MIN_BLOCKS
#define SB_GROUP_MIN_BLOCKS (384 / SB_BLOCK)   /* grouped lanes: 12 warps per SM, <= 168 registers */
#endif
#ifndef SB_FLAT_IDLE
#define SB_FLAT_IDLE 16     /* mean passes a lane may wait per interval before the warp goes flat */
#endif

namespace sb {

constexpr int NS = SB_NS;
constexpr int NP = SB_NP;
constexpr int ND = SB_ND;
constexpr int NP_ = SB_NP > 0 ? SB_NP : 1;
constexpr int ND_ = SB_ND > 0 ? SB_ND : 1;
#if defined(SB_NO_GROUP) || (defined(SB_HOST_EMULATION) && !defined(SB_HOST_EMULATION_GROUP))
constexpr int GROUP = 1;
#else
constexpr int GROUP = SB_GROUP_SIZE(SB_NS);
#endif
// Lanes per instance of the forward kernels.  Forward sensitivities (sb_forward_sens: y and ND
// sensitivity blocks, 56 components for the SEIR problem) run in the backward kernels' lane groups
// (measured, SEIR, 32 768 draws: 88.5 ms against 98.1 ms with one lane per instance and 9 kB of
// spills; -DSB_NO_FWD_GROUP restores that).  The plain forward kernel stays one lane per instance:
// its 8-component state spills 1.5 kB but is still faster (4.2 ms against 6.1 ms in groups, where
// the scalar controller code runs once per group of lanes instead of once per 32 instances);
// -DSB_FWD_GROUP builds it in groups.  Constraint builds (cvCheckConstraints is written for one
// lane per instance) never group.
#if defined(SB_FWD_GROUP) && !defined(SB_CONSTRAINTS)
constexpr int FWD_GROUP = GROUP;
#else
constexpr int FWD_GROUP = 1;
#endif
#if defined(SB_NO_FWD_GROUP) || defined(SB_CONSTRAINTS)
constexpr int SENS_GROUP = 1;
#else
constexpr int SENS_GROUP = GROUP;
#endif
constexpr int HIST_STRIDE = SB_HIST_STRIDE(SB_NS);
constexpr int TAB_STRIDE = SB_TAB_STRIDE(SB_NS);

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// ------------------------------------------------------------------------------------ forward
// NBLK = 1: plain forward problem.  NBLK = 1 + ND: y and the ND sensitivity vectors dy/dp_k stacked
// (CVODES forward sensitivity analysis, /root/reference/sunode/solver.py:360-392 with the
// analytic sensitivity right-hand side J s_k + df/dp_k of symode/problem.py:557-583).
template <int NBLK>
struct FwdSysT {
    static constexpr bool TSTOP = false;
#ifdef SB_CONSTRAINTS
    static constexpr bool CONSTR = (NBLK == 1);    // the reference constrains the forward ODE only
#endif
    static constexpr int GROUP = 1, NS_FULL = NS, NQ_FULL = 1;
    static constexpr bool MAT_SHARED = false;
    struct GroupIds {};
    const SbForwardArgs& a;
    double p[NP_];
    double t;
    __device__ __forceinline__ explicit FwdSysT(const SbForwardArgs& a_) : a(a_) {}
    // tolerances / stop time are launch constants: read from the kernel arguments (constant
    // bank) where needed instead of being carried in registers
    __device__ __forceinline__ double rtol() const { return a.rtol; }
    __device__ __forceinline__ double atol(int i) const { return __ldg(a.atol + i); }
    __device__ __forceinline__ double rtolQ() const { return 0.0; }
    __device__ __forceinline__ double atolQ() const { return 1.0; }
    __device__ __forceinline__ double tstop() const { return 0.0; }
    __device__ __forceinline__ void set_time(double t_) { t = t_; }
    __device__ __forceinline__ void rhs(const double* y, double* out) const {
        sb_rhs(t, y, p, out);
        if (NBLK > 1) sb_sens_rhs(t, y, y + NS, p, out + NS);
    }
    __device__ __forceinline__ void jac(const double* y, double* J) const { sb_jac(t, y, p, J); }
    __device__ __forceinline__ void quad(const double*, double*) const {}
};
using FwdSys = FwdSysT<1>;

// Interpolation table entry of the stored interval (idx - 1, idx): Newton divided differences
// through the `order + 1` stored points ending at the interval's right end, scaled by the interval
// length as CVODES does (CVApolynomialGetY); `order` is the BDF order of the step that produced
// the right end point.  `hist` / `tab` are this instance's arrays.
template <int OUT_STRIDE = TAB_STRIDE>
__device__ __forceinline__ void build_table_entry_at(const double* hist, double* tab, int idx) {
    double* e = tab + (long long)idx * OUT_STRIDE;
#ifdef SB_HERMITE
    // CVAhermiteGetY: the cubic through (y, y') at both ends of the interval, as a Newton form
    // with the nodes t_hi, t_hi, t_lo (scaled by powers of the interval length like the
    // polynomial entries, so the backward kernels evaluate it unchanged as an order-3 entry)
    {
        const double* p1 = hist + (long long)idx * HIST_STRIDE;
        const double* p0 = p1 - HIST_STRIDE;
        const double t1 = p1[0], t0 = p0[0];
        const double D = t1 - t0, a = fabs(D);
        const double sgn = (D >= 0.0) ? 1.0 : -1.0;       // a / D
        e[0] = t0; e[1] = t1; e[2] = 3.0; e[3] = 1.0 / a;
        e[4] = t1; e[5] = t1; e[6] = t0; e[7] = 0.0; e[8] = 0.0; e[9] = 0.0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const double y1 = p1[2 + k], y0 = p0[2 + k], yd1 = p1[2 + NS + k], yd0 = p0[2 + NS + k];
            const double dy = y1 - y0;
            e[10 + k] = y1;
            e[10 + NS + k] = a * yd1;
            e[10 + 2 * NS + k] = sgn * (a * yd1 - sgn * dy);
            e[10 + 3 * NS + k] = a * (yd1 + yd0) - 2.0 * sgn * dy;
            e[10 + 4 * NS + k] = 0.0;
            e[10 + 5 * NS + k] = 0.0;
        }
    }
#else
    int order = (int)hist[(long long)idx * HIST_STRIDE + 1];
    if (order > idx) order = idx;
    if (order < 1) order = 1;
    double T[SB_LMAX], Y[SB_LMAX][NS];
#pragma unroll
    for (int j = 0; j < SB_LMAX; ++j) {
        if (j <= order) {
            const double* pnt = hist + (long long)(idx - j) * HIST_STRIDE;
            T[j] = pnt[0];
#pragma unroll
            for (int k = 0; k < NS; ++k) Y[j][k] = pnt[2 + k];
        } else {
            T[j] = 0.0;
#pragma unroll
            for (int k = 0; k < NS; ++k) Y[j][k] = 0.0;
        }
    }
    const double delt = fabs(T[0] - T[1]);
#pragma unroll
    for (int i = 1; i < SB_LMAX; ++i) {
#pragma unroll
        for (int j = SB_LMAX - 1; j >= 1; --j) {
            if (i <= order && j >= i && j <= order) {
                const double factor = delt / (T[j] - T[j - i]);
#pragma unroll
                for (int k = 0; k < NS; ++k) Y[j][k] = factor * (Y[j][k] - Y[j - 1][k]);
            }
        }
    }
    e[0] = T[1];
    e[1] = T[0];
    e[2] = (double)order;
    e[3] = 1.0 / delt;
#pragma unroll
    for (int j = 0; j < SB_LMAX; ++j) e[4 + j] = T[j];
#pragma unroll
    for (int j = 0; j < SB_LMAX; ++j)
#pragma unroll
        for (int k = 0; k < NS; ++k) e[10 + NS * j + k] = Y[j][k];
#endif
}

// `hyd` = zn[1] = h y' of the step that ended at t (CVAhermiteStorePnt keeps zn[1] / h)
__device__ __forceinline__ void store_point(double* hist, int idx, double t, int order, const double* y,
                                            const double* hyd, double h) {
    double* e = hist + (size_t)idx * HIST_STRIDE;
    e[0] = t;
    e[1] = (double)order;
#pragma unroll
    for (int i = 0; i < NS; ++i) e[2 + i] = y[i];
#ifdef SB_HERMITE
#pragma unroll
    for (int i = 0; i < NS; ++i) e[2 + NS + i] = hyd[i] / h;
#endif
}

// Warp-synchronous driver.  ptxas does not re-converge the lanes of a warp after loops whose trip
// count differs per lane (every lane ends up running alone, 1/32 of the issue rate), so the loops
// over internal steps are made warp-uniform by hand: every pass starts with a warp vote, which is
// also the point where lanes that took different paths through the previous pass meet again.
// `valid` is false for the padding lanes of the last warp; they vote and do nothing else.
//
// Forward: the k-loop over output times of the reference (solver.py:503-521, 705-721) is flattened
// into the step loop -- a lane emits every output time it has already stepped past and then
// takes its next step -- so that lanes do not wait for each other at every output time.
template <int NBLK>
__device__ __forceinline__ void forward_instance_t(const SbForwardArgs& a, long long inst, bool valid) {
    using Sys = FwdSysT<NBLK>;
    using Integrator = Bdf<NS, 0, Sys, NBLK>;
    constexpr int NT = NS * NBLK;
#ifdef SB_CTL_ZERO_INIT
    typename Integrator::Ctl ctl{};
    typename Integrator::Mat mat{};
#else
    typename Integrator::Ctl ctl;      // every field is set by reinit() / the first setup before use
    typename Integrator::Mat mat;
#endif
    Integrator bdf(ctl, mat);
    Sys sys(a);
    double y0[NT];
    if (!valid) inst = 0;
#pragma unroll
    for (int i = 0; i < NS; ++i) y0[i] = a.y0[inst * NS + i];
    if (NBLK > 1) {
        const double* s0 = a.sens0_shared ? a.sens0 : a.sens0 + (size_t)inst * (NT - NS);
#pragma unroll
        for (int i = NS; i < NT; ++i) y0[i] = s0[i - NS];
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) sys.p[i] = a.params[inst * NP + i];
    bdf.clear_stats();
    bdf.reinit(a.t0, y0, nullptr);

    double* yo = a.y_out + (size_t)inst * a.n_t * NS;
    double* so = (NBLK > 1) ? a.sens_out + (size_t)inst * a.n_t * (NT - NS) : nullptr;
    double* hist = a.hist ? a.hist + (size_t)inst * a.hist_cap * HIST_STRIDE : nullptr;
    double* tab = (a.hist && a.tab) ? a.tab + (size_t)inst * a.hist_cap * TAB_STRIDE : nullptr;
    int status = SB_SUCCESS;
    int k = 0;          // next output time
    int nloc = 0;       // internal steps taken towards tvals[k] (CVode's nstloc, summed over retries)

    for (;;) {
        bool work = valid && status == SB_SUCCESS && k < a.n_t;
        if (work && !bdf.in_step) {
            // emit the output times that need no further step (never in the middle of a step:
            // a failed pass leaves the history un-restored until the next attempt)
            for (;;) {
                const double tout = a.tvals[k];
                if (tout == a.t0) {
                    // the reference writes row 0 here whatever k is (solver.py:505-508,707)
#pragma unroll
                    for (int i = 0; i < NS; ++i) yo[i] = y0[i];
#pragma unroll
                    for (int i = NS; i < NT; ++i) so[i - NS] = y0[i];
                } else if (bdf.nst > 0 && (bdf.tn - tout) * bdf.h >= 0.0) {
                    double yk[NT];
                    bdf.get_dky(tout, yk);
#pragma unroll
                    for (int i = 0; i < NS; ++i) yo[(size_t)k * NS + i] = yk[i];
#pragma unroll
                    for (int i = NS; i < NT; ++i) so[(size_t)k * (NT - NS) + i - NS] = yk[i];
                } else {
                    break;
                }
                nloc = 0;
                if (++k == a.n_t) { work = false; break; }
            }
        }
        if (work && !bdf.in_step) {
            // what CVode() does before it calls cvStep
            if (bdf.nst == 0) {
#ifdef SB_CONSTRAINTS
                // cvInitialSetup: y0 must satisfy the constraints
                if (Sys::CONSTR && !bdf.constraints_hold(y0)) status = SB_ILL_INPUT;
                else
#endif
                status = bdf.first_call(sys, a.tvals[k]);
                if (status == SB_SUCCESS && hist) store_point(hist, 0, bdf.tn, 0, bdf.zn[0], bdf.zn[1], bdf.h);
            }
            if (status == SB_SUCCESS) {
                if (nloc >= a.max_steps) status = SB_TOO_MUCH_WORK;
                else if (hist && bdf.nst + 1 >= a.hist_cap) status = SB_TOO_MUCH_WORK;
                else status = bdf.pre_step_checks(sys);
            }
            work = status == SB_SUCCESS;
        }
        const unsigned mask = sb_ballot(work);
        if (mask == 0u) break;
        if (work) {
            const int r = bdf.attempt(sys, mask);
            if (r == SB_SUCCESS) {
                nloc++;
                if (hist) {
                    store_point(hist, bdf.nst, bdf.tn, bdf.qu, bdf.zn[0], bdf.zn[1], bdf.h);
                    if (tab) build_table_entry_at(hist, tab, bdf.nst);
                }
            } else if (r != SB_TRY_AGAIN) {
                status = r;
            }
        }
    }
    if (a.steps_total) {
#ifndef SB_HOST_EMULATION
        unsigned n = (valid && status == SB_SUCCESS) ? (unsigned)bdf.nst : 0u;
        n = __reduce_add_sync(0xffffffffu, n);
        if ((threadIdx.x & 31) == 0 && n) atomicAdd(a.steps_total, (unsigned long long)n);
#else
        if (status == SB_SUCCESS) *a.steps_total += (unsigned long long)bdf.nst;
#endif
    }
    if (!valid) return;

    if (status != SB_SUCCESS) {
        // failed instances read as NaN, like the reference's Ops (as_pytensor.py:289-290)
        for (int j = 0; j < a.n_t * NS; ++j) yo[j] = qnan();
        if (NBLK > 1) for (int j = 0; j < a.n_t * (NT - NS); ++j) so[j] = qnan();
    }
    a.status[inst] = status;
    if (a.fail_k) a.fail_k[inst] = (status == SB_SUCCESS) ? -1 : min(k, a.n_t - 1);
    if (a.hist_n) a.hist_n[inst] = (status == SB_SUCCESS) ? bdf.nst + 1 : 0;
    if (a.stats) {
        int* s = a.stats + inst * SB_STATS_STRIDE;
        s[0] = bdf.st.nst; s[1] = bdf.st.nfe; s[2] = bdf.st.nje; s[3] = bdf.st.nsetups;
        s[4] = bdf.st.netf; s[5] = bdf.st.ncfn; s[6] = bdf.st.nni; s[7] = bdf.nst + 1;
    }
}

__device__ __forceinline__ void forward_instance(const SbForwardArgs& a, long long inst, bool valid) {
    forward_instance_t<1>(a, inst, valid);
}
__device__ __forceinline__ void forward_sens_instance(const SbForwardArgs& a, long long inst, bool valid) {
    forward_instance_t<1 + ND>(a, inst, valid);
}

// ------------------------------------------------------------------------------------ tables
// One thread per (instance, interval): the stand-alone version of the table construction (used when
// the forward kernel did not build the tables itself).
__device__ __forceinline__ void build_table_entry(const SbTablesArgs& a, long long inst, int idx) {
    const int np = a.hist_n[inst];
    if (idx < 1 || idx >= np) return;
    build_table_entry_at(a.hist + (size_t)inst * a.hist_cap * HIST_STRIDE,
                         a.tab + (size_t)inst * a.hist_cap * TAB_STRIDE, idx);
}

// ------------------------------------------------------------------------------------ backward
struct BwdSys {
    static constexpr bool TSTOP = true;
#ifdef SB_CONSTRAINTS
    static constexpr bool CONSTR = false;
#endif
    static constexpr int GROUP = 1, NS_FULL = NS, NQ_FULL = ND_;
    // saved Jacobian + Newton matrix in shared memory (2..4 states; measured: LV backward 17.76 ->
    // 16.78 ms, the 16 registers end the spilling; Robertson 208 -> 193 ms).  From 5 states on the
    // grouped build takes over; the one-lane build of such systems (SB_NO_GROUP, > 64 states)
    // keeps them in registers / local memory, a row per lane would not fit in shared memory.
#if defined(SB_HOST_EMULATION) || defined(SB_MAT_IN_REGISTERS)
    static constexpr bool MAT_SHARED = false;
#else
    static constexpr bool MAT_SHARED = NS >= 2 && NS <= 4;
#endif
    struct GroupIds {};
    const SbBackwardArgs& a;
    __device__ __forceinline__ explicit BwdSys(const SbBackwardArgs& a_) : a(a_) {}
    __device__ __forceinline__ double rtol() const { return a.rtol; }
    __device__ __forceinline__ double atol(int) const { return a.atol; }
    __device__ __forceinline__ double rtolQ() const { return a.rtol_q; }
    __device__ __forceinline__ double atolQ() const { return a.atol_q; }
    // CVodeB stops the backward integrator at the start of the checkpoint interval, i.e. the
    // forward problem's initial time; it steps past each t_lower and interpolates back
    __device__ __forceinline__ double tstop() const { return a.t_end; }
#ifdef SB_PARAMS_IN_MEMORY
    const double* p;       // this instance's parameters, read from (L1-cached) global memory
#else
    double p[NP_];
#endif
    const double* tab;     // this instance's table base
    int np;                // stored points; intervals are 1 .. np-1
    int idx;               // current interval (CVODES' ilast)
    double t;
    double yi[NS];         // forward solution interpolated at t

    __device__ __forceinline__ void set_time(double t_) {
        t = t_;
        if constexpr (NS <= SB_TAB_PREFETCH_MAX_NS) {
            // Small systems: the whole table entry of the current interval is requested at once
            // (independent loads, one memory latency), together with the interval bounds; only
            // when t has left the interval (about one backward step in twenty) the position is
            // moved and the entry fetched again.  Rows above `order` are stored as zeros, so the
            // Newton form is evaluated over all SB_QMAX terms without a branch on the order.
            const double* e = tab + (size_t)idx * TAB_STRIDE;
            double lo, hi, inv_delt, T[SB_QMAX], Y[SB_LMAX][NS];
            int order;
            bool went_left = false;
#pragma unroll 1
            for (;;) {
#if !defined(SB_HOST_EMULATION) && !defined(SB_TAB_LOAD64)
                // 16-byte loads: the entry (an even number of doubles, 16-byte aligned) comes in
                // half as many load instructions -- every one of them touches a different cache
                // line per lane, and waiting for them was the largest single stall of the pass
                static_assert(TAB_STRIDE % 2 == 0, "entry length");
                double v[TAB_STRIDE];
#pragma unroll
                for (int i = 0; i < TAB_STRIDE / 2; ++i) {
                    const double2 d = __ldg(reinterpret_cast<const double2*>(e) + i);
                    v[2 * i] = d.x; v[2 * i + 1] = d.y;
                }
                lo = v[0]; hi = v[1];
                order = (int)v[2];
                inv_delt = v[3];
#pragma unroll
                for (int i = 0; i < SB_QMAX; ++i) T[i] = v[4 + i];
#pragma unroll
                for (int j = 0; j < SB_LMAX; ++j)
#pragma unroll
                    for (int k = 0; k < NS; ++k) Y[j][k] = v[10 + NS * j + k];
#else
                lo = __ldg(e); hi = __ldg(e + 1);
                order = (int)__ldg(e + 2);
                inv_delt = __ldg(e + 3);
#pragma unroll
                for (int i = 0; i < SB_QMAX; ++i) T[i] = __ldg(e + 4 + i);
#pragma unroll
                for (int j = 0; j < SB_LMAX; ++j)
#pragma unroll
                    for (int k = 0; k < NS; ++k) Y[j][k] = __ldg(e + 10 + NS * j + k);
#endif
                // CVAfindIndex: keep the interval while t_lo <= t <= t_hi, else walk
                if ((t < lo || (went_left && t <= lo)) && idx > 1) { --idx; e -= TAB_STRIDE; went_left = true; }
                else if (t > hi && idx < np - 1 && !went_left) { ++idx; e += TAB_STRIDE; }
                else break;
            }
#pragma unroll
            for (int k = 0; k < NS; ++k) yi[k] = Y[0][k];
            double c = 1.0;
#pragma unroll
            for (int i = 0; i < SB_QMAX; ++i) {
                c = (i < order) ? c * ((t - T[i]) * inv_delt) : 0.0;
#pragma unroll
                for (int k = 0; k < NS; ++k) yi[k] = fma(c, Y[i + 1][k], yi[k]);
            }
        } else {
            const double* e = tab + (size_t)idx * TAB_STRIDE;
            // CVAfindIndex: keep the interval while t_lo <= t <= t_hi, else walk
            if (t < __ldg(e)) {
                do { --idx; e -= TAB_STRIDE; } while (idx > 1 && t <= __ldg(e));
                if (idx < 1) { idx = 1; e = tab + TAB_STRIDE; }
            } else if (t > __ldg(e + 1)) {
                while (idx < np - 1 && t > __ldg(e + 1)) { ++idx; e += TAB_STRIDE; }
            }
            const int order = (int)__ldg(e + 2);
            const double inv_delt = __ldg(e + 3);
#pragma unroll
            for (int k = 0; k < NS; ++k) yi[k] = __ldg(e + 10 + k);
            double c = 1.0;
            // rolled on purpose: the table lives in memory (dynamic indexing is free there) and
            // the integrator is instruction-cache bound; this body is inlined at every site
#pragma unroll 1
            for (int i = 0; i < order; ++i) {
                c *= (t - __ldg(e + 4 + i)) * inv_delt;
#pragma unroll
                for (int k = 0; k < NS; ++k) yi[k] = fma(c, __ldg(e + 10 + NS * (i + 1) + k), yi[k]);
            }
        }
    }
    __device__ __forceinline__ void rhs(const double* lam, double* out) const { sb_adj_rhs(t, yi, lam, p, out); }
    __device__ __forceinline__ void jac(const double*, double* J) const { sb_adj_jac(t, yi, p, J); }
    __device__ __forceinline__ void quad(const double* lam, double* out) const { sb_quad_rhs(t, yi, lam, p, out); }
};

// Where the one-lane backward integrator keeps its two matrices: registers, or (BwdSys::MAT_SHARED)
// a per-lane slot of shared memory, an odd number of doubles apart (conflict-free).
template <class Mat, bool SHARED>
__device__ __forceinline__ Mat& mat_home(Mat& in_registers) {
#ifndef SB_HOST_EMULATION
    if constexpr (SHARED) {
        struct Slot { Mat mat; double pad; };
        static_assert((sizeof(Slot) / sizeof(double)) % 2 == 1, "odd stride");
        __shared__ Slot slots[SB_BLOCK];
        return slots[threadIdx.x].mat;
    } else
#endif
    return in_registers;
}

// Backward: the reference restarts the backward integrator at every output time
// (CVodeReInitB + CVodeQuadReInitB, solver.py:756-757).  Because of those restarts NOTHING of the
// integrator survives an interval: only lamda, the quadrature, the status and the table position
// do.  backward_unit therefore processes any range [k_begin, k_end) of the n_t + 1 intervals, which
// lets the launcher cut a solve into short work units (see sb_backward) -- the whole range is the
// plain one-warp-per-32-solves mode.
//
// Every lane walks its intervals with its own counter; what differs between the two modes is when
// a lane that has reached the end of its interval starts the next one:
//  * barrier mode: when no lane of the warp is stepping any more, so that the restart (order-1
//    re-initialisation + cvHin, a long divergent block) runs once per interval for all lanes.  Best
//    when the lanes need about the same number of steps per interval (Lotka-Volterra, SEIR: the
//    wait costs 13-19 % of the lane-passes, restarting lane by lane would cost more);
//  * flat mode: at once.  For stiff problems the steps per interval differ wildly between draws
//    (Robertson: the wait idles 51 % of the lane-passes, restarts are rare against ~400 passes per
//    interval); tools/lane_efficiency.py has the numbers.
// a.flat selects: 0 barrier, 1 flat, -1 the warp decides -- it starts in barrier mode and goes flat
// for good once the lanes of an interval waited more than SB_FLAT_IDLE passes each on average.
// The per-lane sequence of operations is the same in every mode, results do not depend on it.
#define SB_UNIT_TIMEOUT (-1005)
template <bool FLAT>
__device__ __forceinline__ void backward_unit(const SbBackwardArgs& a, long long inst, bool valid,
                                              int k_begin, int k_end) {
    using Integrator = Bdf<NS, ND, BwdSys>;
    if (!valid) inst = 0;
    const bool first = k_begin == 0, last = k_end == a.n_t + 1;
    const int np = a.hist_n[inst];

#ifdef SB_CTL_ZERO_INIT
    typename Integrator::Ctl ctl{};
    typename Integrator::Mat mat_regs{};
#else
    typename Integrator::Ctl ctl;      // every field is set by reinit() / the first setup before use
    typename Integrator::Mat mat_regs;
#endif
    typename Integrator::Mat& mat = mat_home<typename Integrator::Mat, BwdSys::MAT_SHARED>(mat_regs);
    Integrator bdf(ctl, mat);
    BwdSys sys(a);
    double lam[NS], quad[ND_];
    int status;
    bdf.clear_stats();
    sys.idx = np > 1 ? np - 1 : 1;
    if (first) {
        status = a.fwd_status ? a.fwd_status[inst] : SB_SUCCESS;
#pragma unroll
        for (int i = 0; i < NS; ++i) lam[i] = 0.0;
#pragma unroll
        for (int i = 0; i < ND_; ++i) quad[i] = 0.0;
    } else {
        const double* cd = a.carry_d + (size_t)inst * (NS + ND_);
        const int* ci = a.carry_i + (size_t)inst * SB_CARRY_INTS;
#pragma unroll
        for (int i = 0; i < NS; ++i) lam[i] = cd[i];
#pragma unroll
        for (int i = 0; i < ND_; ++i) quad[i] = cd[NS + i];
        status = ci[0]; sys.idx = ci[1];
        bdf.st.nst = ci[2]; bdf.st.nfe = ci[3]; bdf.st.nje = ci[4]; bdf.st.nsetups = ci[5];
        bdf.st.netf = ci[6]; bdf.st.ncfn = ci[7]; bdf.st.nni = ci[8];
    }

#ifdef SB_PARAMS_IN_MEMORY
    sys.p = a.params + inst * NP;
#else
#pragma unroll
    for (int i = 0; i < NP; ++i) sys.p[i] = a.params[inst * NP + i];
#endif
    sys.tab = a.tab + (size_t)inst * a.hist_cap * TAB_STRIDE;
    sys.np = np;
    sys.t = 0.0;
    bdf.reinit(a.t_start, lam, quad);

    const double* g_base = a.grads_shared ? a.grads : a.grads + (size_t)inst * a.n_t * NS;
    // the jump at the lower end of interval k < n_t (solver.py:770-781): lamda -= g, optional traces
    auto jump = [&](int k) {
        const double* g = g_base + (size_t)(a.n_t - 1 - k) * NS;
#pragma unroll
        for (int i = 0; i < NS; ++i) lam[i] -= g[i];
        if (a.lamda_all || a.quad_all) {
            const size_t row = (size_t)inst * a.n_t + (size_t)((a.n_t - k) % a.n_t);
            const bool ok = status == SB_SUCCESS;
            if (a.lamda_all)
#pragma unroll
                for (int i = 0; i < NS; ++i) a.lamda_all[row * NS + i] = ok ? lam[i] : qnan();
            if (a.quad_all)
#pragma unroll
                for (int i = 0; i < ND; ++i) a.quad_all[row * ND + i] = ok ? quad[i] : qnan();
        }
    };

    if constexpr (!FLAT) {
        // ts = [t_start] + reversed(tvals) + [t_end]; interval k is (ts[k+1], ts[k]) (solver.py:750-754)
        for (int k = k_begin; k < k_end; ++k) {
            const double t_upper = (k == 0) ? a.t_start : a.tvals[a.n_t - k];
            const double t_lower = (k == a.n_t) ? a.t_end : a.tvals[a.n_t - 1 - k];
            if (t_lower < t_upper) {                        // warp-uniform: tvals are shared
                // an interval to integrate over needs stored forward steps (none exist when every
                // output time equals t0: then, as in the reference, only the jumps are applied)
                if (valid && status == SB_SUCCESS && np < 2) status = SB_ILL_INPUT;
                const bool live = valid && status == SB_SUCCESS;
                if (live) {
                    bdf.reinit(t_upper, lam, quad);         // CVodeReInitB + CVodeQuadReInitB
                    status = bdf.first_call(sys, t_lower);
                }
                int nloc = 0;
                bool reached = false;
#ifdef SB_REJOIN
                bool sat_out = false;
#endif
                for (;;) {
                    bool work = valid && status == SB_SUCCESS && !reached;
                    if (work && !bdf.in_step) {
                        if (nloc >= a.max_steps) status = SB_TOO_MUCH_WORK;
                        else status = bdf.pre_step_checks(sys);
                        work = status == SB_SUCCESS;
                    }
#ifdef SB_REJOIN
                    {
                        // Order selection (three step-size roots, two extra norms: a quarter of a
                        // pass) runs at the end of a step entered with qwait == 1 -- at constant
                        // order every other step, so the lanes of a warp fall into two classes and
                        // the block runs in every pass for about half of them.  A lane of the
                        // smaller class sits one pass out when that class is small (SB_REJOIN 32nds
                        // of the working lanes); after it, its countdown matches the majority's.
                        // Lanes wait for the slowest lane of the interval anyway; the per-lane
                        // sequence of operations, hence every result, is unchanged.
                        const bool sel = work && bdf.qwait == 1 && bdf.etamax != 1.0;
                        const unsigned m_work = sb_ballot(work), m_sel = sb_ballot(sel);
                        const int n_work = __popc(m_work), n_sel = __popc(m_sel);
                        const bool minority = (2 * n_sel <= n_work) ? sel : !sel;
                        const int n_min = (2 * n_sel <= n_work) ? n_sel : n_work - n_sel;
                        const bool sit = work && minority && !sat_out && n_min > 0 && 32 * n_min <= SB_REJOIN * n_work;
                        sat_out = sit;
                        if (sit) work = false;
                    }
#endif
                    const unsigned mask = sb_ballot(work);
                    if (mask == 0u) break;
                    if (work) {
                        const int r = bdf.attempt(sys, mask);
                        if (r == SB_SUCCESS) {
                            nloc++;
                            bdf.snap_to_tstop(sys);
                            if ((bdf.tn - t_lower) * bdf.h >= 0.0) reached = true;
                            else bdf.limit_to_tstop(sys);
                        } else if (r != SB_TRY_AGAIN) {
                            status = r;
                        }
                    }
                }
                if (valid && status == SB_SUCCESS) {
                    bdf.get_dky(t_lower, lam);                // CVodeGetB
                    if (ND > 0) bdf.get_quad(t_lower, quad);  // CVodeGetQuadB, carried into the next interval
                }
            }
            if (valid && k < a.n_t) jump(k);
        }
    } else {
        // ts = [t_start] + reversed(tvals) + [t_end]; interval k is (ts[k+1], ts[k]) (solver.py:750-754)
        int k = valid ? k_begin : k_end;    // padding lanes only vote
        bool fresh = true;                  // between intervals: interval k has not been started yet
        bool flat = a.flat > 0;
        int nloc = 0, idle = 0;
        double t_lower = 0.0;
        for (;;) {
            const unsigned m_fresh = sb_ballot(k < k_end && fresh);
            const unsigned m_step = sb_ballot(k < k_end && !fresh);
            if ((m_fresh | m_step) == 0u) break;
            if (m_fresh != 0u) {
                if (m_step == 0u) {
                    // a common restart: how long did the lanes of this interval wait for each other?
                    if (a.flat < 0 && idle > SB_FLAT_IDLE * __popc(m_fresh)) flat = true;
                    idle = 0;
                } else {
                    idle += __popc(m_fresh);
                }
            }
            if (k < k_end && fresh && (flat || m_step == 0u)) {
                // start intervals until one has to be integrated over
                while (k < k_end) {
                    const double t_upper = (k == 0) ? a.t_start : a.tvals[a.n_t - k];
                    t_lower = (k == a.n_t) ? a.t_end : a.tvals[a.n_t - 1 - k];
                    if (t_lower < t_upper) {
                        // an interval to integrate over needs stored forward steps (none exist when
                        // every output time equals t0: then, as in the reference, only the jumps are
                        // applied)
                        if (status == SB_SUCCESS && np < 2) status = SB_ILL_INPUT;
                        if (status == SB_SUCCESS) {
                            bdf.reinit(t_upper, lam, quad);         // CVodeReInitB + CVodeQuadReInitB
                            status = bdf.first_call(sys, t_lower);
                            nloc = 0;
                            fresh = false;
                            break;
                        }
                    }
                    if (k < a.n_t) jump(k);
                    ++k;
                }
            }
            // one pass of the step loop for the lanes inside an interval
            bool work = k < k_end && !fresh && status == SB_SUCCESS;
            if (work && !bdf.in_step) {
                if (nloc >= a.max_steps) status = SB_TOO_MUCH_WORK;
                else status = bdf.pre_step_checks(sys);
                work = status == SB_SUCCESS;
            }
            const unsigned mask = sb_ballot(work);
            bool reached = false;
            if (work) {
                const int r = bdf.attempt(sys, mask);
                if (r == SB_SUCCESS) {
                    nloc++;
                    bdf.snap_to_tstop(sys);
                    if ((bdf.tn - t_lower) * bdf.h >= 0.0) reached = true;
                    else bdf.limit_to_tstop(sys);
                } else if (r != SB_TRY_AGAIN) {
                    status = r;
                }
            }
            if (k < k_end && !fresh && (reached || status != SB_SUCCESS)) {
                if (status == SB_SUCCESS) {
                    bdf.get_dky(t_lower, lam);                // CVodeGetB
                    if (ND > 0) bdf.get_quad(t_lower, quad);  // CVodeGetQuadB, carried into the next interval
                }
                if (k < a.n_t) jump(k);
                ++k;
                fresh = true;
            }
        }
    }
    if (!valid) return;
    if (!last) {
        double* cd = a.carry_d + (size_t)inst * (NS + ND_);
        int* ci = a.carry_i + (size_t)inst * SB_CARRY_INTS;
#pragma unroll
        for (int i = 0; i < NS; ++i) cd[i] = lam[i];
#pragma unroll
        for (int i = 0; i < ND_; ++i) cd[NS + i] = quad[i];
        ci[0] = status; ci[1] = sys.idx;
        ci[2] = bdf.st.nst; ci[3] = bdf.st.nfe; ci[4] = bdf.st.nje; ci[5] = bdf.st.nsetups;
        ci[6] = bdf.st.netf; ci[7] = bdf.st.ncfn; ci[8] = bdf.st.nni;
        return;
    }
    double* gout = a.grad_out + inst * ND;
    double* lout = a.lamda_out + inst * NS;
    if (status != SB_SUCCESS) {
#pragma unroll
        for (int i = 0; i < NS; ++i) lam[i] = qnan();
#pragma unroll
        for (int i = 0; i < ND_; ++i) quad[i] = qnan();
    }
#pragma unroll
    for (int i = 0; i < ND; ++i) gout[i] = quad[i];
#pragma unroll
    for (int i = 0; i < NS; ++i) lout[i] = lam[i];
    a.status[inst] = status;
    if (a.stats) {
        int* s = a.stats + inst * SB_STATS_STRIDE;
        s[0] = bdf.st.nst; s[1] = bdf.st.nfe; s[2] = bdf.st.nje; s[3] = bdf.st.nsetups;
        s[4] = bdf.st.netf; s[5] = bdf.st.ncfn; s[6] = bdf.st.nni; s[7] = np;
    }
}

__device__ __forceinline__ void backward_instance(const SbBackwardArgs& a, long long inst, bool valid) {
    backward_unit<false>(a, inst, valid, 0, a.n_t + 1);
}
__device__ __forceinline__ void backward_instance_flat(const SbBackwardArgs& a, long long inst, bool valid) {
    backward_unit<true>(a, inst, valid, 0, a.n_t + 1);
}

// ------------------------------------------------------------------------------------ eval
// One evaluation of a generated function (sb_eval): kind 0 rhs, 1 jacobian, 2 adjoint rhs,
// 4 adjoint Jacobian (-J^T, column-major),
// 3 quadrature rhs.
__device__ __forceinline__ void eval_instance(const SbEvalArgs& a, long long i) {
    double y[NS], p[NP_], lam[NS];
    const double t = a.t[i];
#pragma unroll
    for (int k = 0; k < NS; ++k) { y[k] = a.y[i * NS + k]; lam[k] = a.lam ? a.lam[i * NS + k] : 0.0; }
#pragma unroll
    for (int k = 0; k < NP; ++k) p[k] = a.params_shared ? a.params[k] : a.params[i * NP + k];
    if (a.kind == 0) {
        double out[NS];
        sb_rhs(t, y, p, out);
#pragma unroll
        for (int k = 0; k < NS; ++k) a.out[i * NS + k] = out[k];
    } else if (a.kind == 1) {
        double out[NS * NS];
        sb_jac(t, y, p, out);
#pragma unroll
        for (int k = 0; k < NS * NS; ++k) a.out[i * NS * NS + k] = out[k];
    } else if (a.kind == 2) {
        double out[NS];
        sb_adj_rhs(t, y, lam, p, out);
#pragma unroll
        for (int k = 0; k < NS; ++k) a.out[i * NS + k] = out[k];
    } else if (a.kind == 3) {
        double out[ND_];
        sb_quad_rhs(t, y, lam, p, out);
#pragma unroll
        for (int k = 0; k < ND; ++k) a.out[i * ND + k] = out[k];
    } else if (a.kind == 4) {
        double out[NS * NS];
        sb_adj_jac(t, y, p, out);
#pragma unroll
        for (int k = 0; k < NS * NS; ++k) a.out[i * NS * NS + k] = out[k];
    }
}

}  // namespace sb

#ifdef SB_FUND
#include "sb_fund.cuh"
#endif
#ifdef SB_HOST_EMULATION_GROUP
#include "sb_group.cuh"
#endif
#ifndef SB_HOST_EMULATION   // the host emulation calls the *_instance functions directly
#include "sb_group.cuh"

// read by the launcher (instances per warp of the backward kernels = 32 / sb_group_size)
__device__ int sb_group_size = sb::GROUP;
__device__ int sb_hist_stride = sb::HIST_STRIDE;   // doubles per history point (the launcher sizes the buffer)

__device__ int sb_group_size_fwd = sb::FWD_GROUP;    // lanes per instance of sb_forward
__device__ int sb_group_size_sens = sb::SENS_GROUP;  // ... of sb_forward_sens

// a.lanes instances per warp: one per lane, or one per group of sb::FWD_GROUP lanes
extern "C" __global__ void __launch_bounds__(SB_BLOCK, sb::FWD_GROUP > 1 ? SB_GROUP_MIN_BLOCKS : SB_MIN_BLOCKS)
sb_forward(const __grid_constant__ SbForwardArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int slot = lane / sb::FWD_GROUP;
    const long long inst = warp * a.lanes + slot;
    if constexpr (sb::FWD_GROUP > 1) sb::forward_instance_group<sb::FWD_GROUP, 1>(a, inst, slot < a.lanes && inst < a.B);
    else sb::forward_instance(a, inst, slot < a.lanes && inst < a.B);
}

#if SB_ND > 0
// (y and ND sensitivity blocks per lane: from ten components per lane on, 8 warps per SM with 255
// registers instead of 12 with 168)
#define SB_SENS_GROUP_MIN_BLOCKS \
    ((((SB_NS + sb::SENS_GROUP - 1) / sb::SENS_GROUP) * (1 + SB_ND) >= 10) ? (256 / SB_BLOCK) : SB_GROUP_MIN_BLOCKS)
extern "C" __global__ void __launch_bounds__(SB_BLOCK, sb::SENS_GROUP > 1 ? SB_SENS_GROUP_MIN_BLOCKS : SB_MIN_BLOCKS)
sb_forward_sens(const __grid_constant__ SbForwardArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int slot = lane / sb::SENS_GROUP;
    const long long inst = warp * a.lanes + slot;
    if constexpr (sb::SENS_GROUP > 1) sb::forward_instance_group<sb::SENS_GROUP, 1 + sb::ND>(a, inst, slot < a.lanes && inst < a.B);
    else sb::forward_sens_instance(a, inst, slot < a.lanes && inst < a.B);
}
#endif

#ifdef SB_FUND
// restart-free backward pass (sb_fund.cuh): one lane per instance, grid = ceil(B / 32) warps
extern "C" __global__ void __launch_bounds__(SB_BLOCK, SB_MIN_BLOCKS)
sb_backward_fund(const __grid_constant__ SbBackwardArgs a) {
    const long long inst = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    sb::backward_fund_instance(a, inst, inst < a.B);
}
#endif

// sb_tables: one warp builds 32 consecutive table entries of one instance.  The stored points it
// needs (the tile's own and the SB_QMAX before it) are read with coalesced loads into shared memory,
// every lane builds its entry there, and the tile -- 32 entries are one contiguous run of the
// table -- goes out with coalesced stores (a lane writing its 176-byte entry by itself reached a
// quarter of the copy bandwidth).  SB_TAB_TILES warps per instance, each walking the tiles
// t, t + SB_TAB_TILES, ... up to the instance's stored step count: the grid does not depend on the
// history capacity.  Entries are bit-identical to build_table_entry's (same function).
#ifndef SB_TAB_TILES
#define SB_TAB_TILES 4
#endif
#define SB_TAB_WARPS 2
namespace sb {
constexpr int TAB_PAD = TAB_STRIDE | 1;                         // odd stride: conflict-free
constexpr int TAB_HIST_PTS = 32 + SB_QMAX;
constexpr bool TAB_STAGED = (32 * TAB_PAD + TAB_HIST_PTS * HIST_STRIDE) * 8 * SB_TAB_WARPS <= 40 * 1024;
}
extern "C" __global__ void __launch_bounds__(32 * SB_TAB_WARPS)
sb_tables(const SbTablesArgs a) {
    using namespace sb;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long w = (long long)blockIdx.x * SB_TAB_WARPS + wib;
    const long long inst = w / SB_TAB_TILES;
    if (inst >= a.B) return;
    const int np = a.hist_n[inst];
    const double* hist = a.hist + (size_t)inst * a.hist_cap * HIST_STRIDE;
    double* tab = a.tab + (size_t)inst * a.hist_cap * TAB_STRIDE;
    if constexpr (TAB_STAGED) {
        __shared__ double s_out[SB_TAB_WARPS][32 * TAB_PAD];
        __shared__ double s_hist[SB_TAB_WARPS][TAB_HIST_PTS * HIST_STRIDE];
        for (int tile = (int)(w - inst * SB_TAB_TILES); tile * 32 < np; tile += SB_TAB_TILES) {
            const int i0 = tile * 32;                            // entries i0 .. i0 + 31
            const int p0 = max(i0 - SB_QMAX, 0), p1 = min(i0 + 32, np);     // stored points needed
            for (int j = lane; j < (p1 - p0) * HIST_STRIDE; j += 32)
                s_hist[wib][j] = hist[(size_t)p0 * HIST_STRIDE + j];
            __syncwarp();
            const int idx = i0 + lane;
            if (idx >= 1 && idx < np)
                build_table_entry_at<TAB_PAD>(s_hist[wib] - (long long)p0 * HIST_STRIDE,
                                              s_out[wib] - (long long)i0 * TAB_PAD, idx);
            __syncwarp();
            const int e0 = max(i0, 1), e1 = min(i0 + 32, np);    // entries to write
            for (int j = lane; j < (e1 - e0) * TAB_STRIDE; j += 32) {
                const int e = j / TAB_STRIDE, k = j - e * TAB_STRIDE;
                tab[(size_t)e0 * TAB_STRIDE + j] = s_out[wib][(e0 - i0 + e) * TAB_PAD + k];
            }
            __syncwarp();
        }
    } else {
        for (int idx = (int)(w - inst * SB_TAB_TILES) * 32 + lane; idx < np; idx += 32 * SB_TAB_TILES)
            if (idx >= 1) build_table_entry_at(hist, tab, idx);
    }
}

template <bool FLAT>
__device__ __forceinline__ void sb_backward_body(const SbBackwardArgs& a) {
    if (a.steps_total && (*a.steps_total > a.flat_steps) != FLAT) return;   // the other build runs
    // Persistent warps pull work units (group of a.lanes instances x segment of intervals) from a global
    // counter, in segment-major order.  A solve is ~1700 steps long and the batch is only ~1.7
    // waves of resident warps, so with whole solves as units the second wave leaves a quarter of
    // the machine idle for a full solve time; with short units the idle tail shrinks to one unit
    // (n_seg = 1 is the whole-solve mode).  Unit (g, s) needs (g, s - 1): it was handed out
    // n_groups units earlier, to a warp that is running or done (warps only wait on units handed
    // out before their own, so there is no cycle); the wait is bounded all the same.
    const int lane = threadIdx.x & 31;
    const int total = a.n_seg * a.n_groups;
    for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(a.queue, 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        __syncwarp(0xffffffffu);                   // lane 0 re-joins (grouped lanes must run together)
        if (u >= total) break;
        const int seg = u / a.n_groups, grp = u - seg * a.n_groups;
        int timed_out = 0;
        if (seg > 0) {
            // Wait for the predecessor unit.  Lane 0 looks, all lanes follow its verdict: the loop
            // is warp-uniform on purpose -- a spin loop run by lane 0 alone left that lane
            // separated from the rest of its warp long after the loop (observed with grouped
            // lanes, whose shared per-instance state needs the lanes of a group to run together).
            unsigned long long t_begin = 0;        // (lane 0) set when the first look finds the unit not ready
            for (;;) {
                int verdict = 0;                   // 0 not yet, 1 ready, 2 timed out
                if (lane == 0) {
                    const volatile int* done = a.seg_done;
                    if (done[grp] >= seg) {
                        verdict = 1;
                    } else {
                        // safety net only (a predecessor unit is always running or done): 60 s
                        unsigned long long t_now = 0;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
                        if (t_begin == 0) t_begin = t_now;
                        else if (t_now - t_begin > 60000000000ULL) verdict = 2;
                    }
                }
                verdict = __shfl_sync(0xffffffffu, verdict, 0);
                if (verdict != 0) { timed_out = verdict == 2; break; }
                __nanosleep(200);
            }
            __threadfence();
            __syncwarp(0xffffffffu);
        }
        // a.lanes instances per warp: one per lane, or one per group of sb::GROUP lanes
        const int slot = lane / sb::GROUP;
        const long long inst = (long long)grp * a.lanes + slot;
        const bool valid = slot < a.lanes && inst < a.B;
        const bool writer = (lane % sb::GROUP) == 0;
        const int k0 = seg * a.seg_len;
        const int k1 = min(k0 + a.seg_len, a.n_t + 1);
        if (timed_out) {
            if (valid && writer) a.carry_i[(size_t)inst * SB_CARRY_INTS] = SB_UNIT_TIMEOUT;
            if (valid && writer && k1 == a.n_t + 1) a.status[inst] = SB_UNIT_TIMEOUT;
        } else if constexpr (sb::GROUP > 1) {
            // grouped lanes: both builds run the common-restart schedule
            sb::backward_unit_group<sb::GROUP>(a, inst, valid, k0, k1);
        } else {
            sb::backward_unit<FLAT>(a, inst, valid, k0, k1);
        }
        __threadfence();
        __syncwarp(0xffffffffu);
        if (lane == 0) atomicExch(a.seg_done + grp, seg + 1);
    }
}

extern "C" __global__ void __launch_bounds__(SB_BLOCK, sb::GROUP > 1 ? SB_GROUP_MIN_BLOCKS : SB_MIN_BLOCKS)
sb_backward(const __grid_constant__ SbBackwardArgs a) { sb_backward_body<false>(a); }

extern "C" __global__ void __launch_bounds__(SB_BLOCK, sb::GROUP > 1 ? SB_GROUP_MIN_BLOCKS : SB_MIN_BLOCKS)
sb_backward_flat(const __grid_constant__ SbBackwardArgs a) { sb_backward_body<true>(a); }

extern "C" __global__ void __launch_bounds__(256)
sb_eval(const SbEvalArgs a) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    sb::eval_instance(a, i);
}
#endif  // SB_HOST_EMULATION
