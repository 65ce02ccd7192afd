// sb_bdf.cuh -- per-thread variable-order variable-step BDF integrator (device code, sm_100a).
//
// One CUDA thread owns one IVP instance; every array below has compile-time size and is only ever
// indexed with literals after unrolling, so the Nordsieck history, the Newton matrix and the
// controller state live in registers.  The algorithm is the one the reference reaches through
// lib.CVode / lib.CVodeF / lib.CVodeB (/root/reference/sunode/solver.py:511,711,760): SUNDIALS
// CVODES 5.x BDF with the defaults sunode leaves in place -- fixed-leading-coefficient Nordsieck
// form of order 1..5, modified Newton (<= 3 iterations, dense LU of I - gamma*J with partial
// pivoting, Jacobian reuse policy msbp=20 / msbj=50 / dgmax=0.3), WRMS error test, eta-based
// step/order selection, optional quadrature variables with their own error test
// (solver.py:610-615) and an optional stop time.  SURVEY.md Appendix A lists the constants.
//
// `Sys` supplies the problem:
//     void set_time(double t)                      // prepare evaluations at time t
//     void rhs (const double* y, double* ydot)     // at the prepared time
//     void jac (const double* y, double* J)        // column-major d rhs / d y
//     void quad(const double* y, double* qdot)     // only if NQ > 0
//
// Invariant used throughout: rows zn[j], j > q, are identically zero (the saved correction that
// CVODES parks in zn[qmax] lives in `zsave` instead), so the Pascal-triangle and rescale loops can
// run over a fixed range with cheap guards.
#pragma once

#define SB_QMAX 5
#define SB_LMAX 6
#define SB_UROUND 2.220446049250313e-16

// return codes (reference include/cvodes/16_cvodes.h:45-106)
#define SB_SUCCESS 0
#define SB_TOO_MUCH_WORK (-1)
#define SB_TOO_MUCH_ACC (-2)
#define SB_ERR_FAILURE (-3)
#define SB_CONV_FAILURE (-4)
#define SB_LSETUP_FAIL (-6)
#define SB_RHSFUNC_FAIL (-8)
#define SB_FIRST_RHSFUNC_ERR (-9)
#define SB_REPTD_RHSFUNC_ERR (-10)
#define SB_UNREC_RHSFUNC_ERR (-11)
#define SB_ILL_INPUT (-22)
#define SB_CONSTR_FAIL (-15)
#define SB_BAD_T (-25)
#define SB_TOO_CLOSE (-27)
#define SB_GETY_BADT (-107)
#define SB_TRY_AGAIN 5          /* internal: attempt() wants another pass */

namespace sb {

#ifdef SB_CONSTRAINTS
// Solver(constraints=...) / AdjointSolver(constraints=...): one flag per state, a build option
// of the kernels (-DSB_CONSTRAINTS=c_0,c_1,...); see Bdf::check_constraints
__device__ constexpr double sb_constraints[SB_NS] = {SB_CONSTRAINTS};
#endif

constexpr double ETAMX1 = 10000.0, ETAMX2 = 10.0, ETAMX3 = 10.0, ETAMXF = 0.2, ETAMIN = 0.1;
constexpr double ETACF = 0.25, ADDON = 1e-6, BIAS1 = 6.0, BIAS2 = 6.0, BIAS3 = 10.0;
constexpr double THRESH = 1.5, CRDOWN = 0.3, RDIV = 2.0, DGMAX = 0.3, LS_DGMAX = 0.2;
constexpr double NLSCOEF = 0.1, HLB_FACTOR = 100.0, HUB_FACTOR = 0.1, H_BIAS = 0.5, FUZZ = 100.0;
constexpr int MXNEF1 = 3, SMALL_NEF = 2, SMALL_NST = 10, LONG_WAIT = 10, MXNCF = 10, MXNEF = 7;
constexpr int NLS_MAXCOR = 3, MSBP = 20, MSBJ = 50, HIN_ITERS = 4;

// Warp-level convergence points.  ptxas does not re-converge lanes after data-dependent loops, and
// a lane that falls behind (Jacobian refresh, extra Newton iteration, failed error test) would
// otherwise run the rest of the pass on its own.  `mask` is the set of lanes that entered the
// current pass together (a ballot taken by the kernel's step loop); every one of them executes
// every sb_sync / sb_any below exactly once per pass -- the integrator's pass is single-exit for
// that reason.
#ifndef SB_HOST_EMULATION
__device__ __forceinline__ void sb_sync(unsigned mask) { __syncwarp(mask); }
__device__ __forceinline__ bool sb_any(unsigned mask, bool pred) { return __any_sync(mask, pred) != 0; }
__device__ __forceinline__ unsigned sb_ballot(bool pred) { return __ballot_sync(0xffffffffu, pred); }
#else
inline void sb_sync(unsigned) {}
inline bool sb_any(unsigned, bool pred) { return pred; }
inline unsigned sb_ballot(bool pred) { return pred ? 1u : 0u; }
#endif

enum NFlag { FIRST_CALL = 0, PREV_CONV_FAIL = 1, PREV_ERR_FAIL = 2 };
enum ConvFail { NO_FAILURES = 0, FAIL_BAD_J = 1, FAIL_OTHER = 2 };

// Compile-time loop.  Wherever a loop index is compared for EQUALITY with a runtime quantity
// (current order, pivot row) and also indexes a register array, a plain `#pragma unroll` loop is
// not enough: before unrolling, the optimiser rewrites `(j == q) ? a[j] : x` into `a[q]`, the
// index stays dynamic after unrolling and the whole integrator state is demoted from registers
// to local memory.  With static_for the index is a constant from the start.
template <int V> struct IC { static constexpr int value = V; };
template <int I, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < E) { f(IC<I>{}); static_for<I + 1, E>(f); }
}
#define SB_IDX(J) decltype(J)::value

// Division, square root and the controller's k-th roots.  The compiler's IEEE sequences for these
// carry a special-case slow path each (~18-27 instructions per division with the call); with
// dozens of divisions per step they were a third of all executed instructions and blew the
// instruction cache (ncu: `no_inst` was a third of all stalls).  The integrator's operands are
// well-scaled normal numbers, so it uses branch-free Newton sequences seeded by the SFU
// approximations instead: results are within an ulp of the IEEE ones.  -DSB_EXACT_MATH restores
// the IEEE operations (out of line, to keep the code small).
#ifdef SB_EXACT_MATH
#define SB_EXACT_DIV 1
#define SB_EXACT_SQRT 1
#define SB_EXACT_ROOT 1
#endif
#ifdef SB_HOST_EMULATION
#define SB_EXACT_FN __device__ __forceinline__
// host stand-ins for the SFU seeds: the exact value with the low 32 mantissa bits cleared
static inline double sb_seed_trunc(double v) {
    unsigned long long u; std::memcpy(&u, &v, 8); u &= 0xffffffff00000000ULL; std::memcpy(&v, &u, 8); return v;
}
static inline double sb_rcp_seed(double b) { return sb_seed_trunc(1.0 / b); }
static inline double sb_rsqrt_seed(double x) { return sb_seed_trunc(1.0 / std::sqrt(x)); }
#else
#define SB_EXACT_FN __device__ __noinline__
__device__ __forceinline__ double sb_rcp_seed(double b) {
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b)); return r;       // MUFU.RCP64H
}
__device__ __forceinline__ double sb_rsqrt_seed(double x) {
    double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r;     // MUFU.RSQ64H
}
#endif
#ifdef SB_EXACT_DIV
SB_EXACT_FN double sb_div(double a, double b) { return a / b; }
#else
__device__ __forceinline__ double sb_div(double a, double b) {
    double r = sb_rcp_seed(b);                                  // relative error e <= 2^-20
#ifndef SB_DIV_QUADRATIC
    // one third-order step r (1 + e + e^2), e = 1 - b r: what is left is e^3 < 2^-60.  Three
    // dependent operations instead of the four of two Newton steps -- the kernels wait on
    // dependent FP64 results most of the time, and a step makes ~40 divisions (measured, LV
    // backward: 16.75 -> 16.13 ms with identical step counts)
    const double e = fma(-b, r, 1.0);
    r = fma(r, fma(e, e, e), r);                                // 1/b to ~1 ulp
#else
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);                                           // 1/b to ~1 ulp
#endif
    const double q = a * r;
#ifdef SB_DIV_NO_RESIDUAL
    return q;                                                   // ~2 ulp
#else
    return fma(fma(-b, q, a), r, q);                            // residual correction
#endif
}
#endif
#ifdef SB_EXACT_SQRT
SB_EXACT_FN double sb_sqrt(double a) { return sqrt(a); }
#else
__device__ __forceinline__ double sb_sqrt(double x) {
    double r = sb_rsqrt_seed(x);
    const double hx = 0.5 * x;
    r = fma(r, fma(-hx * r, r, 0.5), r);
    r = fma(r, fma(-hx * r, r, 0.5), r);                        // 1/sqrt(x) to ~1 ulp
    const double s = x * r;
    const double y = fma(fma(-s, s, x), 0.5 * r, s);
    // the SFU seed flushes denormals: below 1e-290 return 0 (absolute error < 1e-145, irrelevant
    // for the weighted norms this is used on); NaN propagates
    return (x >= 1e-290) ? y : ((x == x) ? 0.0 : x);
}
#endif

// The step-size controller's ratio  eta = 1 / (x^(1/k) + ADDON),  k = 1..7  (CVODES:
// 1 / (SUNRpowerR(x, 1/k) + ADDON) in cvCompleteStep / cvDoErrorTest / cvChooseEta), computed from
// the SQUARE x2 = x^2 -- the integrator carries squared error norms so that no step needs a square
// root.  One branch-free sequence for every k: z0 ~ x2^(-1/(2k)) from the single-precision SFU
// log2/exp2 (relative error < 1e-6), the exact factor (1 + r)^(-1/(2k)), r = x2 z0^(2k) - 1, from
// its series to second order (the r^3 term is below 1e-16), then eta = z / (1 + ADDON z).  Good to
// a few ulp, like sqrt() + pow() + a division, at a sixth of the instructions; out of line because
// it has four call sites.
#ifdef SB_HOST_EMULATION
static const double sb_rk_table[8] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7};
#define SB_ROOT_FN inline
#else
__constant__ double sb_rk_table[8] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7};
#define SB_ROOT_FN __device__ __noinline__
#endif
__device__ __forceinline__ double sb_ipow(double z, int k) {   // z^k, 0 <= k <= 7
    const double z2 = z * z, z4 = z2 * z2;
    return ((k & 1) ? z : 1.0) * ((k & 2) ? z2 : 1.0) * ((k & 4) ? z4 : 1.0);
}
SB_ROOT_FN double eta_root2(double x2, int k) {
#ifdef SB_EXACT_ROOT
    return 1.0 / (pow(x2, 0.5 / (double)k) + ADDON);
#else
    if (!(x2 > 1e-36 && x2 < 1e36))                                  // 0, inf, nan, extreme: rare
        return sb_div(1.0, pow(x2, 0.5 / (double)k) + ADDON);
    const double a = 0.5 * sb_rk_table[k];
#ifdef SB_HOST_EMULATION
    const double z0 = (double)powf((float)x2, -0.5f / (float)k);
#else
    const double z0 = (double)exp2f(-__log2f((float)x2) * (float)a);
#endif
    const double r = fma(x2, sb_ipow(z0 * z0, k), -1.0);
    // (1 + r)^(-a) = 1 - a r + a (a + 1) / 2 r^2 - ...,  a = 1/(2k)
    const double z = z0 * fma(r * a, fma(0.5 * (1.0 + a), r, -1.0), 1.0);
    return sb_div(z, fma(ADDON, z, 1.0));
#endif
}

// Three roots at once (the order selection's eta_q, eta_{q-1}, eta_{q+1}): the same sequence as
// eta_root2 per root, written side by side so that the three dependent chains (~25 operations
// each) overlap in the FP64 pipe instead of running one after the other.  Bit-identical to three
// calls of eta_root2.
__device__ __forceinline__ void eta_root2x3(const double* x2, const int* k, double* out) {
#if defined(SB_EXACT_ROOT) || defined(SB_ETA_SERIAL)
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = eta_root2(x2[i], k[i]);
#else
    bool regular = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) regular = regular && (x2[i] > 1e-36 && x2[i] < 1e36);
    if (!regular) {                                                  // 0, inf, nan, extreme: rare
#pragma unroll
        for (int i = 0; i < 3; ++i) out[i] = eta_root2(x2[i], k[i]);
        return;
    }
    double a[3], z0[3], r[3], z[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        a[i] = 0.5 * sb_rk_table[k[i]];
#ifdef SB_HOST_EMULATION
        z0[i] = (double)powf((float)x2[i], -0.5f / (float)k[i]);
#else
        z0[i] = (double)exp2f(-__log2f((float)x2[i]) * (float)a[i]);
#endif
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] = fma(x2[i], sb_ipow(z0[i] * z0[i], k[i]), -1.0);
#pragma unroll
    for (int i = 0; i < 3; ++i) z[i] = z0[i] * fma(r[i] * a[i], fma(0.5 * (1.0 + a[i]), r[i], -1.0), 1.0);
#pragma unroll
    for (int i = 0; i < 3; ++i) out[i] = sb_div(z[i], fma(ADDON, z[i], 1.0));
#endif
}

template <int N>
__device__ __forceinline__ bool all_finite(const double* v) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) s = fma(v[i], 0.0, s);
    return s == 0.0;
}

// mean of the squared weighted components (wrms = sqrt of it)
template <int N>
__device__ __forceinline__ double wms(const double* v, const double* w) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { const double x = v[i] * w[i]; s = fma(x, x, s); }
    return s * (1.0 / N);
}

template <int N>
__device__ __forceinline__ double wrms(const double* v, const double* w) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) { const double x = v[i] * w[i]; s = fma(x, x, s); }
    return sb_sqrt(s * (1.0 / N));
}

// ---- dense LU with partial pivoting ---------------------------------------------------------------
// a is column-major a[i + N*j]; the row permutation is recorded as N small ints; the diagonal of
// the factor holds the RECIPROCAL pivots.  Two implementations behind one interface:
//  * N <= SB_LU_UNROLL_MAX: fully unrolled, rows swapped with selects so that every index is a
//    literal and the matrix lives in registers;
//  * larger N: plain loops with run-time indices; the matrix then lives in (L1-resident,
//    lane-interleaved) local memory.  Measured on the 8-state SEIR problem the unrolled version is
//    still the faster one (backward 219 ms vs 248 ms, forward 4.0 vs 10.7 ms), so the loop version
//    only takes over beyond 8 states, where the unrolled select network (O(N^3)) stops compiling
//    in reasonable time.
#ifndef SB_LU_UNROLL_MAX
#define SB_LU_UNROLL_MAX 8
#endif
template <int N>
__device__ __forceinline__ bool lu_factor(double* a, int* piv) {
    bool ok = true;
    if constexpr (N <= SB_LU_UNROLL_MAX) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            int l = k;
            double best = fabs(a[k + N * k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const double c = fabs(a[i + N * k]);
                if (c > best) { best = c; l = i; }
            }
            piv[k] = l;
            if (!(best > 0.0)) ok = false;
            // swap rows k and l (l >= k) with selects so indices stay literal
            static_for<0, N>([&](auto I_) {
                constexpr int i = SB_IDX(I_);
                const bool sw = (l == i) && (i > k);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const double u = a[k + N * j], v = a[i + N * j];
                    a[k + N * j] = sw ? v : u;
                    a[i + N * j] = sw ? u : v;
                }
            });
            const double mult = sb_div(1.0, a[k + N * k]);
            a[k + N * k] = mult;
#pragma unroll
            for (int i = k + 1; i < N; ++i) a[i + N * k] *= mult;
#pragma unroll
            for (int j = k + 1; j < N; ++j) {
                const double akj = a[k + N * j];
#pragma unroll
                for (int i = k + 1; i < N; ++i) a[i + N * j] = fma(-akj, a[i + N * k], a[i + N * j]);
            }
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            int l = k;
            double best = fabs(a[k + N * k]);
#pragma unroll 1
            for (int i = k + 1; i < N; ++i) {
                const double c = fabs(a[i + N * k]);
                if (c > best) { best = c; l = i; }
            }
            piv[k] = l;
            if (!(best > 0.0)) ok = false;
            if (l != k) {
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    const double u = a[k + N * j];
                    a[k + N * j] = a[l + N * j];
                    a[l + N * j] = u;
                }
            }
            const double mult = sb_div(1.0, a[k + N * k]);
            a[k + N * k] = mult;
#pragma unroll 1
            for (int i = k + 1; i < N; ++i) a[i + N * k] *= mult;
#pragma unroll 1
            for (int j = k + 1; j < N; ++j) {
                const double akj = a[k + N * j];
#pragma unroll 1
                for (int i = k + 1; i < N; ++i) a[i + N * j] = fma(-akj, a[i + N * k], a[i + N * j]);
            }
        }
    }
    return ok;
}

template <int N>
__device__ __forceinline__ void lu_solve(const double* a, const int* piv, double* b) {
    if constexpr (N <= SB_LU_UNROLL_MAX) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int l = piv[k];
            static_for<0, N>([&](auto I_) {
                constexpr int i = SB_IDX(I_);
                const bool sw = (l == i) && (i > k);
                const double u = b[k], v = b[i];
                b[k] = sw ? v : u;
                b[i] = sw ? u : v;
            });
        }
#pragma unroll
        for (int k = 0; k < N - 1; ++k)
#pragma unroll
            for (int i = k + 1; i < N; ++i) b[i] = fma(-a[i + N * k], b[k], b[i]);
#pragma unroll
        for (int k = N - 1; k >= 0; --k) {
            b[k] *= a[k + N * k];
#pragma unroll
            for (int i = 0; i < k; ++i) b[i] = fma(-a[i + N * k], b[k], b[i]);
        }
    } else {
        // the right-hand side is copied into a run-time indexed scratch vector and back
        double x[N];
#pragma unroll
        for (int i = 0; i < N; ++i) x[i] = b[i];
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            const int l = piv[k];
            const double u = x[k];
            x[k] = x[l];
            x[l] = u;
        }
#pragma unroll 1
        for (int k = 0; k < N - 1; ++k) {
            const double xk = x[k];
#pragma unroll 1
            for (int i = k + 1; i < N; ++i) x[i] = fma(-a[i + N * k], xk, x[i]);
        }
#pragma unroll 1
        for (int k = N - 1; k >= 0; --k) {
            const double xk = x[k] * a[k + N * k];
            x[k] = xk;
#pragma unroll 1
            for (int i = 0; i < k; ++i) x[i] = fma(-a[i + N * k], xk, x[i]);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) b[i] = x[i];
    }
}

struct Stats {
    int nst, nfe, nje, nsetups, netf, ncfn, nni;
};

// Fixed-size array with value semantics (so that it can be a member that is either held by value
// or bound by reference, see Bdf::Mem); indices are literals after unrolling, as for plain arrays.
template <class T, int N_>
struct Arr {
    T v[N_];
    __device__ __forceinline__ T& operator[](int i) { return v[i]; }
    __device__ __forceinline__ const T& operator[](int i) const { return v[i]; }
};

// The per-INSTANCE part of the integrator state: everything that is not a vector component.  With
// one lane per instance it lives in that lane's registers (`Bdf` holds these members by value);
// with a group of lanes per instance it is stored once per group in shared memory and `Bdf` binds
// its members to it by reference -- replicated in every lane's registers it would cost ~110
// registers per lane and cap the kernel at 8 warps (32 instances) per SM.
template <int PS>
struct BdfCtl {
    Arr<double, SB_LMAX + 1> tau;
    Arr<double, SB_LMAX> l;
    Arr<double, 6> tq;
    double h, hprime, hscale, eta, etamax, tn, hu;
    double rl1, gamma, gammap, gamrat, crate2, delp2, acnrm2, saved_tq5;
    double step_t0;
    Arr<int, PS> piv;
    int q, qprime, qwait, L, qu;
    int nst, nstlp, nstlj;
    Stats st;
    int ncf, nef, nefQ, nflag, pend;
    bool jcur, in_step;
};
// The matrix rows a lane holds: the saved Jacobian and the Newton matrix I - gamma*J.  With grouped
// lanes they live in shared memory too (rows padded to an odd number of doubles: conflict-free),
// where the cross-lane LU can index them at run time and read the pivot row of another lane.
template <int MSA>
struct BdfMat {
    Arr<double, MSA> savedJ, M;
};
template <bool REF, class T> struct MemT { using type = T; };
// 0: grouped lanes keep private copies of the per-instance record (the host emulation of the group
// code, whose "lanes" are threads that do not run in lockstep between the exchange points)
#ifndef SB_GROUP_SHARED_CTL
#define SB_GROUP_SHARED_CTL 1
#endif
template <class T> struct MemT<true, T> { using type = T&; };

// NM = dimension of the ODE (and of the Newton matrix), NBLK = number of NM-sized blocks that are
// integrated together: 1 for a plain solve, 1 + n_sens for CVODES' simultaneous forward
// sensitivity analysis, where block 0 is y and block 1 + k the sensitivity dy/dp_k.  All blocks
// share the step size, the order and the iteration matrix I - gamma*J; norms are the maximum of
// the per-block WRMS norms (sensitivity error control on, as the reference sets it,
// /root/reference/sunode/solver.py:391-392).  NQ quadrature variables ride along (adjoint only).
template <int NM, int NQ, class Sys, int NBLK = 1>
struct Bdf {
    static constexpr int N = NM * NBLK;          // total length of the state vector
    static constexpr int NQ_ = NQ > 0 ? NQ : 1;
    static constexpr bool QUAD = NQ > 0;

    // Lane groups (Sys::GROUP = G > 1): G adjacent lanes of a warp integrate ONE instance together.
    // Lane r of the group holds component r of every vector (NM = 1; NQ = the quadrature
    // components r, r + G, ... it owns), row r of the Jacobian and of the Newton matrix; the
    // scalar controller state (h, q, tau, l, tq, counters ...) is replicated, so the lanes of a
    // group take every branch together.  What differs from the one-lane-per-instance build is
    // confined to the hooks below: sums / maxima / votes over the group, and the linear algebra.
    static constexpr int G = Sys::GROUP;
    static constexpr int MS = (G > 1) ? NM * Sys::NS_FULL : NM * NM;   // matrix entries a lane holds
    static constexpr int PS = (G > 1) ? Sys::NS_FULL : NM;             // pivot record
    // one-lane build: the two matrices may live in shared memory as well (Sys::MAT_SHARED), which
    // frees 4 NM^2 registers where the integrator state no longer fits (3-4 states)
    static constexpr bool MAT_REF = (G > 1) || Sys::MAT_SHARED;
    static constexpr int MSA = (G > 1) ? ((MS + NM) | 1) : (Sys::MAT_SHARED ? (MS | 1) : MS);   // allocated length
    typename Sys::GroupIds gid;                  // grouped lanes: lane mask + rank (empty otherwise)
    __device__ __forceinline__ double gsum(double x) const { if constexpr (G > 1) return Sys::gsum(x, gid); else return x; }
    __device__ __forceinline__ double gmax(double x) const { if constexpr (G > 1) return Sys::gmax(x, gid); else return x; }
    __device__ __forceinline__ bool gall(bool b) const { if constexpr (G > 1) return Sys::gall(b, gid); else return b; }
    // mean of the squared weighted components of a state-sized block / of a quadrature vector
    __device__ __forceinline__ double ms_y(const double* v, const double* w) const {
        if constexpr (G > 1) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < NM; ++i) { const double x = v[i] * w[i]; s = fma(x, x, s); }
            return gsum(s) * (1.0 / Sys::NS_FULL);
        } else return wms<NM>(v, w);
    }
    __device__ __forceinline__ double ms_q(const double* v) const {
        if constexpr (G > 1) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < NQ_; ++i) { const double x = v[i] * ewtQ[i]; s = fma(x, x, s); }
            return gsum(s) * (1.0 / Sys::NQ_FULL);
        } else return wms<NQ_>(v, ewtQ);
    }
    __device__ __forceinline__ bool finite_y(const double* v) const { return gall(all_finite<N>(v)); }
    __device__ __forceinline__ bool finite_q(const double* v) const { return gall(all_finite<NQ_>(v)); }

    // max over blocks of the weighted RMS norm (one block: the plain WRMS norm)
    __device__ __forceinline__ double norm(const double* v) const { return sb_sqrt(norm2(v)); }
    // Its square.  Every test a step makes on a norm (Newton convergence, error test, tolsf) and
    // the step-size ratios are monotone in the norm, so the step loop works on squares throughout
    // and takes no square roots (names ending in 2: del2, delp2, crate2, acnrm2, dsm2; tq[1..4]
    // hold the squares of CVODES' test quantities).  Only cvHin still needs norms proper.
    __device__ __forceinline__ double norm2(const double* v) const {
        double r = ms_y(v, ewt);
#pragma unroll
        for (int b = 1; b < NBLK; ++b) r = fmax(r, ms_y(v + b * NM, ewt + b * NM));
        return r;
    }

    // Nordsieck arrays
    double zn[SB_LMAX][N], zsave[N], acor[N], ewt[N];
    double znQ[SB_LMAX][NQ_], zsaveQ[NQ_], acorQ[NQ_], ewtQ[NQ_];
    double ycur[N];                 // zn[0] + acor after the nonlinear solve
    // per-instance state: by value (registers) with one lane per instance, references into the
    // group's shared-memory record otherwise (see BdfCtl)
    using Ctl = BdfCtl<PS>;
    template <class T> using Mem = typename MemT<(G > 1) && (SB_GROUP_SHARED_CTL != 0), T>::type;
    template <class T> using MemM = typename MemT<MAT_REF, T>::type;      // the matrix rows
    // step / order control
    Mem<Arr<double, SB_LMAX + 1>> tau;
    Mem<Arr<double, SB_LMAX>> l;
    Mem<Arr<double, 6>> tq;
    Mem<double> h, hprime, hscale, eta, etamax, tn, hu;
    Mem<double> rl1, gamma, gammap, gamrat, crate2, delp2, acnrm2, saved_tq5;
    Mem<int> q, qprime, qwait, L, qu;
    Mem<bool> jcur;
    // linear solver
    using Mat = BdfMat<MSA>;
    MemM<Arr<double, MSA>> savedJ, M;
    Mem<Arr<int, PS>> piv;
    // counters
    Mem<int> nst, nstlp, nstlj;
    Mem<Stats> st;
    // a step in flight (cvStep's locals): one call of attempt() is one pass of cvStep's retry loop,
    // so that the lanes of a warp can be re-converged between passes by the caller
    Mem<double> step_t0;
    Mem<int> ncf, nef, nefQ, nflag;
    Mem<bool> in_step;
    // History manipulations requested by the previous pass (failed pass: restore + rescale, maybe
    // an order drop; new step with a new step size: rescale, maybe an order change).  They are
    // carried out at ONE place, the top of the next attempt(), so that the bulky restore /
    // rescale / order-change code exists once instead of once per failure branch.
    Mem<int> pend;

    __device__ __forceinline__ Bdf(Ctl& c, Mat& mat)
        : savedJ(mat.savedJ), M(mat.M), tau(c.tau), l(c.l), tq(c.tq), h(c.h), hprime(c.hprime), hscale(c.hscale), eta(c.eta),
          etamax(c.etamax), tn(c.tn), hu(c.hu), rl1(c.rl1), gamma(c.gamma), gammap(c.gammap),
          gamrat(c.gamrat), crate2(c.crate2), delp2(c.delp2), acnrm2(c.acnrm2),
          saved_tq5(c.saved_tq5), q(c.q), qprime(c.qprime), qwait(c.qwait), L(c.L), qu(c.qu),
          jcur(c.jcur), piv(c.piv), nst(c.nst), nstlp(c.nstlp), nstlj(c.nstlj), st(c.st),
          step_t0(c.step_t0), ncf(c.ncf), nef(c.nef), nefQ(c.nefQ), nflag(c.nflag),
          in_step(c.in_step), pend(c.pend) {}

    // ------------------------------------------------------------------ (re)initialisation
    // CVodeReInit (+ CVodeQuadReInit): order 1, fresh controller state, counters cleared
    __device__ __forceinline__ void reinit(double t0, const double* y0, const double* q0) {
        tn = t0;
#pragma unroll
        for (int j = 0; j < SB_LMAX; ++j) {
#pragma unroll
            for (int i = 0; i < N; ++i) zn[j][i] = 0.0;
#pragma unroll
            for (int i = 0; i < NQ_; ++i) znQ[j][i] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) { zn[0][i] = y0[i]; zsave[i] = 0.0; acor[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NQ_; ++i) { znQ[0][i] = QUAD ? q0[i] : 0.0; zsaveQ[i] = 0.0; acorQ[i] = 0.0; }
#pragma unroll
        for (int j = 0; j <= SB_LMAX; ++j) tau[j] = 0.0;
        q = 1; L = 2; qwait = 2; qprime = 1; etamax = ETAMX1; qu = 0; hu = 0.0;
        nst = 0; nstlp = 0; nstlj = 0;
        saved_tq5 = 0.0; jcur = false; crate2 = 1.0; delp2 = 0.0; acnrm2 = 0.0;
        h = hprime = hscale = 0.0; eta = 1.0; gamma = gammap = gamrat = 1.0; rl1 = 1.0;
        in_step = false; step_t0 = t0; ncf = nef = nefQ = 0; nflag = FIRST_CALL; pend = 0;
    }

    __device__ __forceinline__ void clear_stats() {
        st.nst = st.nfe = st.nje = st.nsetups = st.netf = st.ncfn = st.nni = 0;
    }

    __device__ __forceinline__ bool set_ewt(const Sys& sys) {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double d = fma(sys.rtol(), fabs(zn[0][i]), sys.atol(i));   // per stacked component
            ok = ok && (d > 0.0);
            ewt[i] = sb_div(1.0, d);
        }
        if (QUAD) {
#pragma unroll
            for (int i = 0; i < NQ_; ++i) {
                const double d = fma(sys.rtolQ(), fabs(znQ[0][i]), sys.atolQ());
                ok = ok && (d > 0.0);
                ewtQ[i] = sb_div(1.0, d);
            }
        }
        return gall(ok);
    }

    // ------------------------------------------------------------------ first step: cvHin
    __device__ __forceinline__ int ydd_norm(Sys& sys, double hg, double* yddnrm) {
        double y[N], f[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = fma(hg, zn[1][i], zn[0][i]);
        sys.set_time(tn + hg);
        sys.rhs(y, f); st.nfe++;
        if (!finite_y(f)) return 1;
        const double rhg = sb_div(1.0, hg);
#pragma unroll
        for (int i = 0; i < N; ++i) f[i] = (f[i] - zn[1][i]) * rhg;
        double nrm = norm(f);
        if (QUAD) {
            double fq[NQ_];
            sys.quad(y, fq);
            if (!finite_q(fq)) return 1;
#pragma unroll
            for (int i = 0; i < NQ_; ++i) fq[i] = (fq[i] - znQ[1][i]) * rhg;
            nrm = fmax(nrm, sb_sqrt(ms_q(fq)));
        }
        *yddnrm = nrm;
        return 0;
    }

    // zn[1] (and znQ[1]) hold the unscaled derivatives on entry
    __device__ __forceinline__ int hin(Sys& sys, double tout) {
        const double tdiff = tout - tn;
        if (tdiff == 0.0) return SB_TOO_CLOSE;
        const double sign = (tdiff > 0.0) ? 1.0 : -1.0;
        const double tdist = fabs(tdiff);
        const double tround = SB_UROUND * fmax(fabs(tn), fabs(tout));
        if (tdist < 2.0 * tround) return SB_TOO_CLOSE;
        const double hlb = HLB_FACTOR * tround;
        // upper bound
        double hub_inv = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double d = fma(HUB_FACTOR, fabs(zn[0][i]), sb_div(1.0, ewt[i]));
            hub_inv = fmax(hub_inv, sb_div(fabs(zn[1][i]), d));
        }
        if (QUAD) {
#pragma unroll
            for (int i = 0; i < NQ_; ++i) {
                const double d = fma(HUB_FACTOR, fabs(znQ[0][i]), sb_div(1.0, ewtQ[i]));
                hub_inv = fmax(hub_inv, sb_div(fabs(znQ[1][i]), d));
            }
        }
        hub_inv = gmax(hub_inv);
        double hub = HUB_FACTOR * tdist;
        if (hub * hub_inv > 1.0) hub = sb_div(1.0, hub_inv);
        double hg = sb_sqrt(hlb * hub);
        if (hub < hlb) { h = sign * hg; return SB_SUCCESS; }

        double hs = hg, hnew = hg, yddnrm = 0.0;
        for (int count1 = 1; count1 <= HIN_ITERS; ++count1) {
            bool hgOK = false;
            for (int count2 = 1; count2 <= HIN_ITERS; ++count2) {
                const int r = ydd_norm(sys, hg * sign, &yddnrm);
                if (r == 0) { hgOK = true; break; }
                hg *= 0.2;
            }
            if (!hgOK) {
                if (count1 <= 2) return SB_REPTD_RHSFUNC_ERR;
                hnew = hs;
                break;
            }
            hs = hg;
            hnew = (yddnrm * hub * hub > 2.0) ? sb_sqrt(sb_div(2.0, yddnrm)) : sb_sqrt(hg * hub);
            if (count1 == HIN_ITERS) break;
            const double hrat = sb_div(hnew, hg);
            if (hrat > 0.5 && hrat < 2.0) break;
            if (count1 > 1 && hrat > 2.0) { hnew = hg; break; }
            hg = hnew;
        }
        double h0 = H_BIAS * hnew;
        h0 = fmin(fmax(h0, hlb), hub);
        h = sign * h0;
        return SB_SUCCESS;
    }

    // The part of CVode() that runs when nst == 0: f(t0, y0), initial h, scale zn[1].
    __device__ __forceinline__ int first_call(Sys& sys, double tout) {
        if (!set_ewt(sys)) return SB_ILL_INPUT;
        sys.set_time(tn);
        sys.rhs(zn[0], zn[1]); st.nfe++;
        if (!finite_y(zn[1])) return SB_FIRST_RHSFUNC_ERR;
        if (QUAD) {
            sys.quad(zn[0], znQ[1]);
            if (!finite_q(znQ[1])) return SB_RHSFUNC_FAIL;
        }
        if (Sys::TSTOP && (sys.tstop() - tn) * (tout - tn) <= 0.0) return SB_ILL_INPUT;
        double tout_hin = tout;
        if (Sys::TSTOP && (tout - tn) * (tout - sys.tstop()) > 0.0) tout_hin = sys.tstop();
        const int hflag = hin(sys, tout_hin);
        if (hflag != SB_SUCCESS) return hflag;
        if (Sys::TSTOP && (tn + h - sys.tstop()) * h > 0.0) h = (sys.tstop() - tn) * (1.0 - 4.0 * SB_UROUND);
        hscale = h; hprime = h;
#pragma unroll
        for (int i = 0; i < N; ++i) zn[1][i] *= h;
        if (QUAD) {
#pragma unroll
            for (int i = 0; i < NQ_; ++i) znQ[1][i] *= h;
        }
        return SB_SUCCESS;
    }

    // ------------------------------------------------------------------ history manipulation
    __device__ __forceinline__ void rescale() {
        // rows above q are zero, so scaling them too is harmless and keeps every index literal
        double factor = eta;
#pragma unroll
        for (int j = 1; j < SB_LMAX; ++j) {
#pragma unroll
            for (int i = 0; i < N; ++i) zn[j][i] *= factor;
            if (QUAD) {
#pragma unroll
                for (int i = 0; i < NQ_; ++i) znQ[j][i] *= factor;
            }
            factor *= eta;
        }
        h = hscale * eta;
        hscale = h;
    }

    __device__ __forceinline__ void predict(const Sys& sys) {
        tn += h;
        if (Sys::TSTOP && (tn - sys.tstop()) * h > 0.0) tn = sys.tstop();
#pragma unroll
        for (int k = 1; k < SB_LMAX; ++k)
#pragma unroll
            for (int j = SB_LMAX - 1; j >= k; --j) {
#pragma unroll
                for (int i = 0; i < N; ++i) zn[j - 1][i] += zn[j][i];
                if (QUAD) {
#pragma unroll
                    for (int i = 0; i < NQ_; ++i) znQ[j - 1][i] += znQ[j][i];
                }
            }
    }

    __device__ __forceinline__ void restore(double saved_t) {
        tn = saved_t;
#pragma unroll
        for (int k = 1; k < SB_LMAX; ++k)
#pragma unroll
            for (int j = SB_LMAX - 1; j >= k; --j) {
#pragma unroll
                for (int i = 0; i < N; ++i) zn[j - 1][i] -= zn[j][i];
                if (QUAD) {
#pragma unroll
                    for (int i = 0; i < NQ_; ++i) znQ[j - 1][i] -= znQ[j][i];
                }
            }
    }

    // cvIncreaseBDF: build row L = q+1 from the saved correction
    __device__ __forceinline__ void increase_order() {
        double lc[SB_LMAX];
#pragma unroll
        for (int i = 0; i < SB_LMAX; ++i) lc[i] = 0.0;
        lc[2] = 1.0;
        double alpha1 = 1.0, prod = 1.0, xiold = 1.0, alpha0 = -1.0, hsum = hscale;
#pragma unroll
        for (int j = 1; j < SB_QMAX; ++j) {
            if (j < q) {
                hsum += tau[j + 1];
                const double xi = sb_div(hsum, hscale);
                prod *= xi;
                alpha0 -= 1.0 / (double)(j + 1);
                alpha1 += sb_div(1.0, xi);
#pragma unroll
                for (int i = SB_LMAX - 1; i >= 2; --i)
                    if (i <= j + 2) lc[i] = fma(lc[i], xiold, lc[i - 1]);
                xiold = xi;
            }
        }
        const double A1 = sb_div(-alpha0 - alpha1, prod);
        double znew[N], znewQ[NQ_];
#pragma unroll
        for (int i = 0; i < N; ++i) znew[i] = A1 * zsave[i];
#pragma unroll
        for (int i = 0; i < NQ_; ++i) znewQ[i] = A1 * zsaveQ[i];
        // rows 2..q get lc[j]*znew added, row q+1 becomes znew; written as value selects so that
        // no index ever depends on q (a conditional `zn[q+1] = ...` would be turned into a
        // dynamically indexed store and push the whole state into local memory)
        static_for<2, SB_LMAX>([&](auto J_) {
            constexpr int j = SB_IDX(J_);
            const bool upd = (j <= q), fresh = (j == q + 1);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double v = fma(lc[j], znew[i], zn[j][i]);
                zn[j][i] = fresh ? znew[i] : (upd ? v : zn[j][i]);
            }
            if (QUAD) {
#pragma unroll
                for (int i = 0; i < NQ_; ++i) {
                    const double v = fma(lc[j], znewQ[i], znQ[j][i]);
                    znQ[j][i] = fresh ? znewQ[i] : (upd ? v : znQ[j][i]);
                }
            }
        });
    }

    // cvDecreaseBDF: fold row q into rows 2..q-1, then drop it (zero, to keep the invariant)
    __device__ __forceinline__ void decrease_order() {
        double lc[SB_LMAX];
#pragma unroll
        for (int i = 0; i < SB_LMAX; ++i) lc[i] = 0.0;
        lc[2] = 1.0;
        double hsum = 0.0;
#pragma unroll
        for (int j = 1; j <= SB_QMAX - 2; ++j) {
            if (j <= q - 2) {
                hsum += tau[j];
                const double xi = sb_div(hsum, hscale);
#pragma unroll
                for (int i = SB_LMAX - 1; i >= 2; --i)
                    if (i <= j + 2) lc[i] = fma(lc[i], xi, lc[i - 1]);
            }
        }
        double zq[N], zqQ[NQ_];
#pragma unroll
        for (int i = 0; i < N; ++i) zq[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NQ_; ++i) zqQ[i] = 0.0;
        static_for<2, SB_LMAX>([&](auto J_) {
            constexpr int j = SB_IDX(J_);
            const bool is_q = (j == q);
#pragma unroll
            for (int i = 0; i < N; ++i) zq[i] = is_q ? zn[j][i] : zq[i];
#pragma unroll
            for (int i = 0; i < NQ_; ++i) zqQ[i] = is_q ? znQ[j][i] : zqQ[i];
        });
        static_for<2, SB_LMAX>([&](auto J_) {
            constexpr int j = SB_IDX(J_);
            const bool upd = (j < q), is_q = (j == q);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double v = fma(-lc[j], zq[i], zn[j][i]);
                zn[j][i] = is_q ? 0.0 : (upd ? v : zn[j][i]);
            }
            if (QUAD) {
#pragma unroll
                for (int i = 0; i < NQ_; ++i) {
                    const double v = fma(-lc[j], zqQ[i], znQ[j][i]);
                    znQ[j][i] = is_q ? 0.0 : (upd ? v : znQ[j][i]);
                }
            }
        });
    }

    // cvAdjustOrder(-1) as used by the error-test failure path (q -> q-1)
    __device__ __forceinline__ void drop_order() {
        if (q == 2) {
            // CVODES skips the polynomial adjustment at q == 2; the row is simply abandoned
#pragma unroll
            for (int i = 0; i < N; ++i) zn[2][i] = 0.0;
#pragma unroll
            for (int i = 0; i < NQ_; ++i) znQ[2][i] = 0.0;
        } else {
            decrease_order();
        }
    }

    static constexpr int PEND_RESTORE = 1, PEND_ORDER = 2, PEND_RESCALE = 4, PEND_RELOAD = 8;

    // cvAdjustParams / the tails of cvHandleNFlag and cvDoErrorTest, in their original order:
    // restore, order change (towards qprime), rescale by eta; or the order-1 reload.
    __device__ __forceinline__ int apply_pending(Sys& sys) {
        if (pend & PEND_RESTORE) restore(step_t0);
        if (pend & PEND_ORDER) {
            if (qprime > q) increase_order();
            else drop_order();
            q = qprime; L = q + 1; qwait = L;
        }
        if (pend & PEND_RESCALE) rescale();
        int ret = SB_SUCCESS;
        if (pend & PEND_RELOAD) {
            // order 1 and still failing: reload the first derivative from scratch
            h *= eta; hscale = h; qwait = LONG_WAIT;
            double f[N];
            sys.set_time(tn);
            sys.rhs(zn[0], f); st.nfe++;
            if (!finite_y(f)) ret = SB_UNREC_RHSFUNC_ERR;
#pragma unroll
            for (int i = 0; i < N; ++i) zn[1][i] = h * f[i];
            if (QUAD) {
                double fq[NQ_];
                sys.quad(zn[0], fq);
                if (!finite_q(fq)) ret = SB_RHSFUNC_FAIL;
#pragma unroll
                for (int i = 0; i < NQ_; ++i) znQ[1][i] = h * fq[i];
            }
        }
        pend = 0;
        return ret;
    }

    // ------------------------------------------------------------------ cvSetBDF + cvSetTqBDF
    // Same quantities, same operands per operation as CVODES' cvSetBDF / cvSetTqBDF, arranged for
    // the FP64 pipe: a division is a chain of eight dependent operations and this routine makes
    // up to fourteen, so (i) every ratio h / (h + tau_1 + ... + tau_{j-1}) the routine can need is
    // formed up front -- five independent chains that overlap -- and picked by order afterwards,
    // (ii) the order / qwait cases are selects instead of branches, which lets the remaining
    // independent divisions (tq[2], tq[5], C, C', C'', 1/l1) overlap as well.  Values that a case
    // does not use may be inf / NaN (zero denominators at order 1); they are never selected.
    __device__ __forceinline__ void set_coeffs() {
#ifdef SB_COEFFS_BRANCHY
        set_coeffs_branchy();
#else
        // x[j] = h / hs_j,  hs_j = h + tau[1] + ... + tau[j-1]   (j = 2 .. 6)
        double hs = h, x[SB_LMAX + 1];
#pragma unroll
        for (int j = 2; j <= SB_LMAX; ++j) { hs += tau[j - 1]; x[j] = sb_div(h, hs); }
        // xi_inv of the alpha0_hat term: x[q] (order > 1); of the tq[3] term: x[q + 1]
        double xq = 1.0, xq1 = 1.0, lq = 1.0, alpha0 = -1.0;
        static_for<1, SB_LMAX>([&](auto J_) {
            constexpr int j = SB_IDX(J_);
            if (j >= 2) xq = (j == q) ? x[j] : xq;
            xq1 = (j == q) ? x[j + 1] : xq1;
        });
#pragma unroll
        for (int i = 0; i < SB_LMAX; ++i) l[i] = 0.0;
        l[0] = l[1] = 1.0;
#pragma unroll
        for (int j = 2; j < SB_QMAX; ++j) {
            const bool on = j < q;
            alpha0 = on ? alpha0 - 1.0 / (double)j : alpha0;
#pragma unroll
            for (int i = SB_QMAX; i >= 1; --i)
                if (i <= j) l[i] = on ? fma(l[i - 1], x[j], l[i]) : l[i];
        }
        const bool hi = q > 1;
        alpha0 = hi ? alpha0 - sb_rk_table[q] : alpha0;
        const double xistar_inv = hi ? -l[1] - alpha0 : 1.0;
        const double xi_inv = hi ? xq : 1.0;
        const double alpha0_hat = hi ? -l[1] - xi_inv : -1.0;
#pragma unroll
        for (int i = SB_QMAX; i >= 1; --i) l[i] = (hi && i <= q) ? fma(l[i - 1], xistar_inv, l[i]) : l[i];
        static_for<1, SB_LMAX>([&](auto J_) {
            constexpr int j = SB_IDX(J_);
            lq = (j == q) ? l[j] : lq;
        });
        const double A1 = 1.0 - alpha0_hat + alpha0;
        const double A2 = 1.0 + q * A1;
        const double tq2 = sb_div(A1, alpha0 * A2);
        const double tq5 = fabs(sb_div(A2 * xistar_inv, lq * xi_inv));
        // the qwait == 1 quantities (used by the order selection of this step)
        const double C = sb_div(xistar_inv, lq);
        const double A3 = alpha0 + sb_rk_table[q];
        const double A4 = alpha0_hat + xi_inv;
        const double Cpinv = sb_div(1.0 - A4 + A3, A3);
        const double A5 = alpha0 - sb_rk_table[q + 1];
        const double A6 = alpha0_hat - xq1;
        const double Cppinv = sb_div(1.0 - A6 + A5, A2);
        const double tq3 = sb_div(Cppinv, xq1 * (q + 2) * A5);
        const double rl = sb_div(1.0, l[1]);
        tq[2] = tq2 * tq2;
        tq[5] = tq5;
        if (qwait == 1) {
            tq[1] = hi ? (C * Cpinv) * (C * Cpinv) : 1.0;
            tq[3] = tq3 * tq3;
        }
        tq[4] = tq[2] * (1.0 / (NLSCOEF * NLSCOEF));   // (tq[2] / nlscoef)^2 = 1 / CVODES' tq[4]^2, a factor
        rl1 = rl;
        gamma = h * rl1;
        if (nst == 0) gammap = gamma;
        gamrat = (nst > 0) ? sb_div(gamma, gammap) : 1.0;
#endif
    }

    __device__ __forceinline__ void set_coeffs_branchy() {
        double xi_inv = 1.0, xistar_inv = 1.0, alpha0 = -1.0, alpha0_hat = -1.0, hsum = h;
#pragma unroll
        for (int i = 0; i < SB_LMAX; ++i) l[i] = 0.0;
        l[0] = l[1] = 1.0;
        if (q > 1) {
#pragma unroll
            for (int j = 2; j < SB_QMAX; ++j) {
                if (j < q) {
                    hsum += tau[j - 1];
                    xi_inv = sb_div(h, hsum);
                    alpha0 -= 1.0 / (double)j;
#pragma unroll
                    for (int i = SB_QMAX; i >= 1; --i)
                        if (i <= j) l[i] = fma(l[i - 1], xi_inv, l[i]);
                }
            }
            alpha0 -= sb_rk_table[q];
            xistar_inv = -l[1] - alpha0;
            double tau_qm1 = 0.0;
            static_for<1, SB_LMAX>([&](auto J_) {
                constexpr int j = SB_IDX(J_);
                tau_qm1 = (j == q - 1) ? tau[j] : tau_qm1;
            });
            hsum += tau_qm1;
            xi_inv = sb_div(h, hsum);
            alpha0_hat = -l[1] - xi_inv;
#pragma unroll
            for (int i = SB_QMAX; i >= 1; --i)
                if (i <= q) l[i] = fma(l[i - 1], xistar_inv, l[i]);
        }
        double lq = 1.0, tau_q = 0.0;
        static_for<1, SB_LMAX>([&](auto J_) {
            constexpr int j = SB_IDX(J_);
            lq = (j == q) ? l[j] : lq;
            tau_q = (j == q) ? tau[j] : tau_q;
        });
        const double A1 = 1.0 - alpha0_hat + alpha0;
        const double A2 = 1.0 + q * A1;
        const double tq2 = sb_div(A1, alpha0 * A2);
        tq[2] = tq2 * tq2;
        tq[5] = fabs(sb_div(A2 * xistar_inv, lq * xi_inv));
        if (qwait == 1) {
            if (q > 1) {
                const double C = sb_div(xistar_inv, lq);
                const double A3 = alpha0 + sb_rk_table[q];
                const double A4 = alpha0_hat + xi_inv;
                const double Cpinv = sb_div(1.0 - A4 + A3, A3);
                tq[1] = (C * Cpinv) * (C * Cpinv);
            } else tq[1] = 1.0;
            hsum += tau_q;
            xi_inv = sb_div(h, hsum);
            const double A5 = alpha0 - sb_rk_table[q + 1];
            const double A6 = alpha0_hat - xi_inv;
            const double Cppinv = sb_div(1.0 - A6 + A5, A2);
            const double tq3 = sb_div(Cppinv, xi_inv * (q + 2) * A5);
            tq[3] = tq3 * tq3;
        }
        tq[4] = tq[2] * (1.0 / (NLSCOEF * NLSCOEF));   // (tq[2] / nlscoef)^2 = 1 / CVODES' tq[4]^2, a factor
        rl1 = sb_div(1.0, l[1]);
        gamma = h * rl1;
        if (nst == 0) gammap = gamma;
        gamrat = (nst > 0) ? sb_div(gamma, gammap) : 1.0;
    }

    // ------------------------------------------------------------------ linear setup
    // returns 0 ok, 1 recoverable
    __device__ __forceinline__ int lsetup(Sys& sys, int convfail, const double* ypred) {
        // |gamma/gammap - 1|: set_coeffs computed that ratio for this pass already
        const double dgamma = fabs(gamrat - 1.0);
        const bool jbad = (nst == 0) || (nst > nstlj + MSBJ) ||
                          (convfail == FAIL_BAD_J && dgamma < LS_DGMAX) || (convfail == FAIL_OTHER);
        if (jbad) {
            st.nje++; nstlj = nst; jcur = true;
            sys.jac(ypred, savedJ.v);
            if (!gall(all_finite<MS>(savedJ.v))) return 1;
        } else {
            jcur = false;
        }
#pragma unroll
        for (int k = 0; k < MS; ++k) M[k] = -gamma * savedJ[k];
        if constexpr (G > 1) {
            sys.add_identity(M.v);
            return sys.lu_factor(M.v, piv.v) ? 0 : 1;
        } else {
#pragma unroll
            for (int i = 0; i < NM; ++i) M[i + NM * i] += 1.0;
            return lu_factor<NM>(M.v, piv.v) ? 0 : 1;
        }
    }

    // ------------------------------------------------------------------ nonlinear solve
    // cvNls + SUNNonlinSol_Newton + cvNlsConvTest.  `go` = this lane takes part (lanes whose pass
    // already failed still walk through for the convergence points).  Returns 0 converged, 1
    // recoverable failure.  The control flow is the reference's -- up to three passes (stale
    // Jacobian -> fresh factorisation -> fresh Jacobian), up to NLS_MAXCOR Newton iterations per
    // pass -- but every loop is closed by a warp vote so that lanes re-join after each iteration.
    __device__ __forceinline__ int nls(Sys& sys, int nflag_, unsigned mask, bool go) {
        int convfail = (nflag_ == FIRST_CALL || nflag_ == PREV_ERR_FAIL) ? NO_FAILURES : FAIL_OTHER;
        bool callSetup = (nflag_ == PREV_CONV_FAIL) || (nflag_ == PREV_ERR_FAIL) || (nst == 0) ||
                         (nst >= nstlp + MSBP) || (fabs(gamrat - 1.0) > DGMAX);
        double delta[N], f[N];
        int retval = 1;
        bool done = !go;
        if (go) sys.set_time(tn);
        // (rolled: the kernels are instruction-fetch bound and the loop bodies are long)
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
            bool live = !done;
            if (live) {
#pragma unroll
                for (int i = 0; i < N; ++i) { acor[i] = 0.0; ycur[i] = zn[0][i]; }
                sys.rhs(ycur, f); st.nfe++;
                // a failure before the Newton loop (residual or setup) is returned without a retry
                if (!finite_y(f)) { done = true; live = false; }
            }
            if (live && callSetup) {
                const int r = lsetup(sys, convfail, ycur);
                st.nsetups++;
                callSetup = false;
                gamrat = 1.0; gammap = gamma; crate2 = 1.0; nstlp = nst;
                if (r != 0) { done = true; live = false; }
            }
            sb_sync(mask);
            bool run = live;
            if (run) {
#pragma unroll
                for (int i = 0; i < N; ++i) delta[i] = fma(gamma, f[i], -fma(rl1, zn[1][i], acor[i]));
                // delta now holds -(rl1*zn1 + acor - gamma*f) = -G
            }
#pragma unroll 1
            for (int m = 0; m < NLS_MAXCOR; ++m) {
                if (run) {
                    st.nni++;
#pragma unroll
                    for (int b = 0; b < NBLK; ++b) {
                        if constexpr (G > 1) sys.lu_solve(M.v, piv.v, delta + b * NM);
                        else lu_solve<NM>(M.v, piv.v, delta + b * NM);
                    }
                    if (gamrat != 1.0) {
                        const double sc = sb_div(2.0, 1.0 + gamrat);
#pragma unroll
                        for (int i = 0; i < N; ++i) delta[i] *= sc;
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) { acor[i] += delta[i]; ycur[i] = zn[0][i] + acor[i]; }
                    const double del2 = norm2(delta);
                    if (m > 0) crate2 = fmax((CRDOWN * CRDOWN) * crate2, sb_div(del2, delp2));
                    const double dcon2 = del2 * fmin(1.0, crate2) * tq[4];
                    if (dcon2 <= 1.0) {
                        acnrm2 = (m == 0) ? del2 : norm2(acor);
                        jcur = false;
                        retval = 0; done = true; run = false;
                    } else if (!(dcon2 > 1.0)) {                      // NaN
                        run = false;
                    } else if (m >= 1 && del2 > (RDIV * RDIV) * delp2) {   // diverging
                        run = false;
                    } else {
                        delp2 = del2;
                        if (m + 1 >= NLS_MAXCOR) {
                            run = false;
                        } else {
                            sys.rhs(ycur, f); st.nfe++;
                            if (!finite_y(f)) run = false;
#pragma unroll
                            for (int i = 0; i < N; ++i)
                                delta[i] = fma(gamma, f[i], -fma(rl1, zn[1][i], acor[i]));
                        }
                    }
                }
                if (!sb_any(mask, run)) break;
            }
            // recoverable failure with a stale Jacobian: one more try with a fresh one
            bool retry = false;
            if (live && !done) {
                if (!jcur) { callSetup = true; convfail = FAIL_BAD_J; retry = true; }
                else done = true;
            }
            if (!sb_any(mask, retry)) break;
        }
        return retval;
    }

    // ------------------------------------------------------------------ after a successful step
    __device__ __forceinline__ void complete_step() {
        nst++; st.nst++;
        hu = h; qu = q;
#pragma unroll
        for (int i = SB_QMAX; i >= 2; --i) tau[i] = (i <= q) ? tau[i - 1] : tau[i];
        if (q == 1 && nst > 1) tau[2] = tau[1];
        tau[1] = h;
        // l[j] == 0 and zn[j] == 0 for j > q, so the update runs over all rows
#pragma unroll
        for (int j = 0; j < SB_LMAX; ++j) {
#pragma unroll
            for (int i = 0; i < N; ++i) zn[j][i] = fma(l[j], acor[i], zn[j][i]);
            if (QUAD) {
#pragma unroll
                for (int i = 0; i < NQ_; ++i) znQ[j][i] = fma(l[j], acorQ[i], znQ[j][i]);
            }
        }
        qwait--;
        if (qwait == 1 && q != SB_QMAX) {
#pragma unroll
            for (int i = 0; i < N; ++i) zsave[i] = acor[i];
            if (QUAD) {
#pragma unroll
                for (int i = 0; i < NQ_; ++i) zsaveQ[i] = acorQ[i];
            }
            saved_tq5 = tq[5];
        }
    }

    __device__ __forceinline__ void set_eta() {
        if (eta < THRESH) { eta = 1.0; hprime = h; }
        else { eta = fmin(eta, etamax); hprime = h * eta; }
    }

    __device__ __forceinline__ void prepare_next_step(double dsm2) {
        if (etamax == 1.0) {
            qwait = max(qwait, 2);
            qprime = q; hprime = h; eta = 1.0;
            return;
        }
        // the lanes of a warp that select the order in this step first form the two extra error
        // norms; then ALL lanes take the three roots together (lanes that need only eta_q pass
        // dummies): one call site, three overlapping chains
        const bool select = qwait == 0;
        const bool need_m1 = select && q > 1;
        const bool need_p1 = select && q != SB_QMAX && saved_tq5 != 0.0;
        double x2[3] = {(BIAS2 * BIAS2) * dsm2, 1.0, 1.0}, etas[3];
        int ks[3] = {L, 1, 1};
        if (need_m1) {
            double zq[N], zqQ[NQ_];
#pragma unroll
            for (int i = 0; i < N; ++i) zq[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NQ_; ++i) zqQ[i] = 0.0;
            static_for<2, SB_LMAX>([&](auto J_) {
                constexpr int j = SB_IDX(J_);
                const bool is_q = (j == q);
#pragma unroll
                for (int i = 0; i < N; ++i) zq[i] = is_q ? zn[j][i] : zq[i];
#pragma unroll
                for (int i = 0; i < NQ_; ++i) zqQ[i] = is_q ? znQ[j][i] : zqQ[i];
            });
            double ddn2 = norm2(zq);
            if (QUAD) ddn2 = fmax(ddn2, ms_q(zqQ));
            ddn2 *= tq[1];
            x2[1] = (BIAS1 * BIAS1) * ddn2; ks[1] = q;
        }
        if (need_p1) {
            const double r = sb_div(h, tau[2]);
            double rp = r;
#pragma unroll
            for (int j = 2; j <= SB_LMAX; ++j) if (j <= L) rp *= r;
            const double cquot = sb_div(tq[5], saved_tq5) * rp;
            double tmp[N];
#pragma unroll
            for (int i = 0; i < N; ++i) tmp[i] = fma(-cquot, zsave[i], acor[i]);
            double dup2 = norm2(tmp);
            if (QUAD) {
                double tmpq[NQ_];
#pragma unroll
                for (int i = 0; i < NQ_; ++i) tmpq[i] = fma(-cquot, zsaveQ[i], acorQ[i]);
                dup2 = fmax(dup2, ms_q(tmpq));
            }
            dup2 *= tq[3];
            x2[2] = (BIAS3 * BIAS3) * dup2; ks[2] = L + 1;
        }
        eta_root2x3(x2, ks, etas);
        const double etaq = etas[0];
        if (!select) { eta = etaq; qprime = q; set_eta(); return; }
        qwait = 2;
        const double etaqm1 = need_m1 ? etas[1] : 0.0, etaqp1 = need_p1 ? etas[2] : 0.0;
        const double etam = fmax(etaqm1, fmax(etaq, etaqp1));
        if (etam < THRESH) { eta = 1.0; qprime = q; }
        else if (etam == etaq) { eta = etaq; qprime = q; }
        else if (etam == etaqm1) { eta = etaqm1; qprime = q - 1; }
        else {
            eta = etaqp1; qprime = q + 1;
#pragma unroll
            for (int i = 0; i < N; ++i) zsave[i] = acor[i];
            if (QUAD) {
#pragma unroll
                for (int i = 0; i < NQ_; ++i) zsaveQ[i] = acorQ[i];
            }
        }
        set_eta();
    }

    // Tail of a failed error test (cvDoErrorTest after the `dsm > 1` branch): decides how the
    // step is retried; the history manipulation itself is queued in `pend`.
    // returns 0 = try again, <0 = fatal
    __device__ __forceinline__ int error_test_failed(double dsm2, int nef_) {
        st.netf++;
        pend = PEND_RESTORE;
        if (nef_ == MXNEF) return SB_ERR_FAILURE;
        etamax = 1.0;
        if (nef_ <= MXNEF1) {
            eta = eta_root2((BIAS2 * BIAS2) * dsm2, L);
            eta = fmax(ETAMIN, eta);
            if (nef_ >= SMALL_NEF) eta = fmin(eta, ETAMXF);
            pend |= PEND_RESCALE;
            return 0;
        }
        eta = ETAMIN;
        if (q > 1) {
            qprime = q - 1;                 // cvAdjustOrder(-1); L = q; q--; qwait = L
            pend |= PEND_ORDER | PEND_RESCALE;
        } else {
            pend |= PEND_RELOAD;
        }
        return 0;
    }

#ifdef SB_CONSTRAINTS
    // ------------------------------------------------------------------ inequality constraints
    // CVodeSetConstraints (/root/reference/sunode/solver.py:268-271, 568-572): SB_CONSTRAINTS is
    // the list of per-component flags, 0 none, +-1 y >= 0 / <= 0, +-2 y > 0 / < 0.
    // N_VConstrMask: is component i of y in violation?
    __device__ static __forceinline__ bool constr_violated(int i, double y) {
        const double c = sb_constraints[i];
        return (fabs(c) > 1.5) ? (y * c <= 0.0) : (fabs(c) > 0.5) ? (y * c < 0.0) : false;
    }
    __device__ __forceinline__ bool constraints_hold(const double* y) const {
        bool ok = true;
#pragma unroll
        for (int i = 0; i < NM; ++i) ok = ok && !constr_violated(i, y[i]);
        return ok;
    }
    // cvCheckConstraints on the converged iterate ycur = zn[0] + acor.  A small violation is
    // projected away (acor -= v) and the step goes on to the error test; otherwise eta is set for
    // a smaller step and true is returned.
    __device__ __forceinline__ bool check_constraints() {
        double v[N];
        bool any = false;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool m = i < NM && constr_violated(i, ycur[i]);
            const double c = sb_constraints[i < NM ? i : 0];
            const double ac = (i < NM && fabs(c) > 1.5) ? c : 0.0;      // a * c: the strict ones
            v[i] = m ? ycur[i] - 0.1 * sb_div(ac, ewt[i]) : 0.0;
            any = any || m;
        }
        if (!any) return false;
        if (norm2(v) * tq[4] <= 1.0) {          // ||v|| <= CVODES' tq[4]
#pragma unroll
            for (int i = 0; i < N; ++i) acor[i] -= v[i];
            return false;
        }
        double mq = 1.7976931348623157e308;     // N_VMinQuotient(zn[0], mm * (zn[0] - y))
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            const double d = zn[0][i] - ycur[i];
            if (constr_violated(i, ycur[i]) && d != 0.0) mq = fmin(mq, sb_div(zn[0][i], d));
        }
        eta = fmax(0.9 * mq, 0.1);
        return true;
    }
#endif

    // ------------------------------------------------------------------ one internal step (cvStep)
    // One pass of cvStep's predict / solve / test loop, entered together by the lanes in `mask`.
    // Returns SB_SUCCESS when the step is complete, SB_TRY_AGAIN when the pass failed recoverably
    // (the retry is queued in `pend`; call again), < 0 on a fatal failure.  Single exit: a lane
    // whose pass has failed keeps walking (doing nothing) through the remaining convergence points.
    __device__ __forceinline__ int attempt(Sys& sys, unsigned mask) {
        if (!in_step) {
            step_t0 = tn;
            ncf = 0; nef = 0; nefQ = 0; nflag = FIRST_CALL;
            pend = 0;
            if (nst > 0 && hprime != h) pend = PEND_RESCALE | ((qprime != q) ? PEND_ORDER : 0);
            in_step = true;
        }
        int result = SB_SUCCESS;
        if (pend != 0) {
            const int pr = apply_pending(sys);
            if (pr != SB_SUCCESS) { in_step = false; result = pr; }
        }
        bool go = (result == SB_SUCCESS);
        sb_sync(mask);
        if (go) {
            predict(sys);
            set_coeffs();
        }
        const int nr = nls(sys, nflag, mask, go);
        double dsm2 = 0.0;
        if (go) {
            if (nr != 0) {
                // cvHandleNFlag
                st.ncfn++; ncf++;
                etamax = 1.0;
                go = false;
                if (ncf == MXNCF) { in_step = false; result = SB_CONV_FAILURE; }
                else {
                    eta = ETACF;
                    nflag = PREV_CONV_FAIL;
                    pend = PEND_RESTORE | PEND_RESCALE;
                    result = SB_TRY_AGAIN;
                }
            } else {
#ifdef SB_CONSTRAINTS
                bool cfail = false;
                if constexpr (Sys::CONSTR) cfail = check_constraints();
                if (cfail) {
                    // CONSTR_RECVR in cvHandleNFlag: counted like a convergence failure, but the
                    // step-size ratio is the one cvCheckConstraints chose
                    st.ncfn++; ncf++;
                    etamax = 1.0;
                    go = false;
                    if (ncf == MXNCF) { in_step = false; result = SB_CONSTR_FAIL; }
                    else {
                        nflag = PREV_CONV_FAIL;
                        pend = PEND_RESTORE | PEND_RESCALE;
                        result = SB_TRY_AGAIN;
                    }
                } else
#endif
                {
                    dsm2 = acnrm2 * tq[2];
                    if (!(dsm2 <= 1.0)) {
                        nef++;
                        nflag = PREV_ERR_FAIL;
                        const int r = error_test_failed(dsm2, nef);
                        go = false;
                        if (r < 0) { in_step = false; result = r; }
                        else result = SB_TRY_AGAIN;
                    }
                }
            }
        }
        if (QUAD) {
            sb_sync(mask);
            if (go) {
                ncf = 0; nef = 0;
                double fq[NQ_];
                sys.quad(ycur, fq);
                if (!finite_q(fq)) {
                    st.ncfn++; ncf++;
                    etamax = 1.0;
                    go = false;
                    if (ncf == MXNCF) { in_step = false; result = SB_REPTD_RHSFUNC_ERR; }
                    else {
                        eta = ETACF;
                        nflag = PREV_CONV_FAIL;
                        pend = PEND_RESTORE | PEND_RESCALE;
                        result = SB_TRY_AGAIN;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NQ_; ++i) acorQ[i] = rl1 * fma(h, fq[i], -znQ[1][i]);
                    const double dsmQ2 = ms_q(acorQ) * tq[2];
                    if (!(dsmQ2 <= 1.0)) {
                        nefQ++;
                        nflag = PREV_ERR_FAIL;
                        const int r = error_test_failed(dsmQ2, nefQ);
                        go = false;
                        if (r < 0) { in_step = false; result = r; }
                        else result = SB_TRY_AGAIN;
                    } else {
                        dsm2 = fmax(dsm2, dsmQ2);
                    }
                }
            }
        }
        sb_sync(mask);
        if (go) {
            complete_step();
            prepare_next_step(dsm2);
            etamax = (nst <= SMALL_NST) ? ETAMX2 : ETAMX3;
            // (CVODES rescales acor by tq[2] here to expose the local error estimate; nothing on
            // this path reads it before the next step overwrites it, so it is not materialised.)
            in_step = false;
        }
        return result;
    }

    // ------------------------------------------------------------------ dense output
    __device__ __forceinline__ void get_dky(double t, double* out) const {
        const double s = sb_div(t - tn, h);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = SB_LMAX - 1; j >= 0; --j) acc = fma(acc, s, zn[j][i]);
            out[i] = acc;
        }
    }

    __device__ __forceinline__ void get_quad(double t, double* out) const {
        const double s = sb_div(t - tn, h);
#pragma unroll
        for (int i = 0; i < NQ_; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = SB_LMAX - 1; j >= 0; --j) acc = fma(acc, s, znQ[j][i]);
            out[i] = acc;
        }
    }

    // Per-step bookkeeping CVode() does around cvStep for nst > 0.  Returns <0 on failure.
    __device__ __forceinline__ int pre_step_checks(const Sys& sys) {
        if (nst > 0 && !set_ewt(sys)) return SB_ILL_INPUT;
        // tolsf = uround * ||y||_wrms > 1, tested on the squares (no square roots needed)
        double nrm2 = norm2(zn[0]);
        if (QUAD) nrm2 = fmax(nrm2, ms_q(znQ[0]));
        if ((SB_UROUND * SB_UROUND) * nrm2 > 1.0) return SB_TOO_MUCH_ACC;
        return SB_SUCCESS;
    }

    // tstop handling after a successful step (CVode loop, "tstop" blocks).  Returns true when the
    // integration reached tstop.
    __device__ __forceinline__ void snap_to_tstop(const Sys& sys) {
        if (Sys::TSTOP) {
            const double troundoff = FUZZ * SB_UROUND * (fabs(tn) + fabs(h));
            if (fabs(tn - sys.tstop()) <= troundoff) tn = sys.tstop();
        }
    }
    __device__ __forceinline__ void limit_to_tstop(const Sys& sys) {
        if (Sys::TSTOP && (tn + hprime - sys.tstop()) * h > 0.0) {
            hprime = (sys.tstop() - tn) * (1.0 - 4.0 * SB_UROUND);
            eta = sb_div(hprime, h);
        }
    }
};

}  // namespace sb
