// sb_group.cuh -- lane groups: G adjacent lanes of a warp integrate ONE instance together.
//
// For systems of more than a few states the one-lane-per-instance build runs out of registers: the
// 8-state SEIR adjoint keeps ~2.5 kB of integrator state per instance, the compiler spills 2 kB of
// it per thread, the unrolled 8x8 LU select network alone is 280 kB of code, and ncu shows the
// kernel waiting on local-memory loads (37 % of the stall samples) and instruction fetch (35 %).
// Here lane r of a group holds component r of every vector, row r of the Jacobian and of the
// Newton matrix I - gamma*J; the dense LU (partial pivoting, implicit row exchange) and its
// triangular solves run across the lanes with warp shuffles, norms are butterfly sums, and the
// scalar controller state is replicated so that the lanes of a group branch together.  The
// algorithm -- CVODES' BDF as the reference reaches it through lib.CVodeB
// (/root/reference/sunode/solver.py:756-760) -- is the one of sb_bdf.cuh: this file only supplies
// the hooks `Bdf` calls when Sys::GROUP > 1 and the backward driver for grouped lanes.
#pragma once

namespace sb {

#define SB_GROUP_SPLIT (-1006)     /* SB_GROUP_CHECK builds: the lanes of a group were found separated */

template <int G>
struct LaneGroup {
    static_assert(G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "group size");
    __device__ __forceinline__ static int rank() { return (int)(threadIdx.x & (G - 1)); }
    __device__ __forceinline__ static unsigned mask() {
        if constexpr (G == 32) return 0xffffffffu;
        else return ((1u << G) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(G - 1));
    }
    // The group's lane mask and the lane's rank are computed once and kept in registers (`Ids`,
    // laundered through an asm so that the compiler does not re-derive them from %tid before
    // every shuffle: ncu showed 800 S2R/shift sequences, each in front of a shuffle).
    struct Ids { unsigned gm; int r; };
    __device__ __forceinline__ static Ids ids() {
        Ids v;
#ifdef SB_HOST_EMULATION
        v.gm = mask(); v.r = rank();
#else
        asm volatile("mov.u32 %0, %2;\n\tmov.u32 %1, %3;" : "=r"(v.gm), "=r"(v.r) : "r"(mask()), "r"(rank()));
#endif
        return v;
    }
    __device__ __forceinline__ static double sum(double x, unsigned gm) {
        // xor butterfly: every lane ends with the bitwise identical sum
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) x += __shfl_xor_sync(gm, x, o, G);
        return x;
    }
    __device__ __forceinline__ static double max(double x, unsigned gm) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(gm, x, o, G));
        return x;
    }
    __device__ __forceinline__ static bool all(bool b, unsigned gm) { return __all_sync(gm, b) != 0; }
    // The per-instance state in shared memory is updated by every lane of the group with the same
    // value in the same instruction; updates like `nst++` or `tn += h` are only correct while the
    // lanes of a group execute TOGETHER.  Group-uniform branches keep them together; after every
    // lane-dependent branch (and wherever the lanes may have been separated before the group code
    // starts) they are joined again here.  -DSB_GROUP_CHECK turns the assumption into a test: a
    // group that is found split reports SB_GROUP_SPLIT through the instance status.
    __device__ __forceinline__ static bool converge(unsigned gm) {
        __syncwarp(gm);
#ifdef SB_GROUP_CHECK
        return (__activemask() & gm) == gm;
#else
        return true;
#endif
    }
    // SB_GROUP_CHECK: are the lanes of the group together right now (no joining)?
    __device__ __forceinline__ static bool together(unsigned gm) {
#ifdef SB_GROUP_CHECK
        return (__activemask() & gm) == gm;
#else
        (void)gm;
        return true;
#endif
    }

    // register array element `rank` (0 for lanes past the end) without dynamic indexing
    template <int N_>
    __device__ __forceinline__ static double pick(const double* v, int idx) {
        double r = 0.0;
        static_for<0, N_>([&](auto J_) { constexpr int j = SB_IDX(J_); r = (idx == j) ? v[j] : r; });
        return r;
    }

    // ---- dense LU across the lanes ------------------------------------------------------------
    // Lane r holds the C rows r*C .. r*C + C - 1 of the N_ x N_ matrix in SHARED memory: local row
    // c at m[c * N_ .. c * N_ + N_ - 1], and in m[C * N_ + c] the elimination step at which that
    // row was the pivot row (-1: never -- rows past N_).  The blocks of the lanes of a warp follow
    // each other with a stride of STRIDE doubles (odd: conflict-free), so the block of lane w is at
    // m + (w - r) * STRIDE and every loop below may use run-time indices (compact code: this
    // kernel is instruction-fetch bound).  Partial pivoting without moving rows: piv[k] is the
    // (global) row that is the k-th pivot row.  After the factorisation a row pivoted at step s
    // holds the multipliers of steps < s in columns 0..s-1, the RECIPROCAL pivot in column s and
    // its U entries in columns s+1.. .
    template <int N_, int C, int STRIDE>
    __device__ __forceinline__ static bool lu_factor(double* m, int* piv, Ids id) {
        const int r = id.r;
        const unsigned gm = id.gm;
        bool ok = true;
        bool used[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            used[c] = r * C + c >= N_;             // padding rows are never pivots
            m[C * N_ + c] = -1.0;
        }
#pragma unroll 1
        for (int k = 0; k < N_; ++k) {
            __syncwarp(gm);                        // the updates of step k - 1 are visible
            double cand = -1.0;
            int who = r * C;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const double v = used[c] ? -1.0 : fabs(m[c * N_ + k]);
                if (v > cand) { cand = v; who = r * C + c; }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const double oc = __shfl_xor_sync(gm, cand, o, G);
                const int ow = __shfl_xor_sync(gm, who, o, G);
                const bool take = (oc > cand) || (oc == cand && ow < who);
                cand = take ? oc : cand;
                who = take ? ow : who;
            }
            piv[k] = who;
            if (!(cand > 0.0)) ok = false;
            const int wlane = who / C, wrow = who - wlane * C;
            const double* prow = m + (wlane - r) * STRIDE + wrow * N_;
            const double rp = sb_div(1.0, prow[k]);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const bool me = (r * C + c == who);
                if (!used[c] && !me) {
                    const double mult = m[c * N_ + k] * rp;
#ifndef SB_GROUP_UNROLL_ELIM
                    // run-time trip count: only the columns right of the pivot (round 2: SEIR
                    // backward 123.5 -> 121.3 ms against the fixed trip count with literal offsets
                    // and a predicate, which executed the whole row at every step)
#pragma unroll 1
                    for (int j = k + 1; j < N_; ++j) m[c * N_ + j] = fma(-mult, prow[j], m[c * N_ + j]);
#else
#pragma unroll
                    for (int j = 1; j < N_; ++j)
                        if (j > k) m[c * N_ + j] = fma(-mult, prow[j], m[c * N_ + j]);
#endif
                }
            }
            __syncwarp(gm);                        // everybody has read the pivot row
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const bool me = (r * C + c == who);
                if (!used[c] && !me) m[c * N_ + k] *= rp;
                if (me) { m[c * N_ + k] = rp; m[C * N_ + c] = (double)k; }
                used[c] = used[c] || me;
            }
        }
        __syncwarp(gm);
        return ok;
    }

    // b[c]: entry r*C + c of the right-hand side on entry, of the solution on return
    template <int N_, int C>
    __device__ __forceinline__ static void lu_solve(const double* m, const int* piv, double* b, Ids id) {
        const int r = id.r;
        const unsigned gm = id.gm;
        int mystep[C];
        double v[C], x[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { mystep[c] = (int)m[C * N_ + c]; v[c] = b[c]; x[c] = 0.0; }
        // (both sweeps rolled by default: the grouped kernel is instruction-fetch bound -- ncu: 43 %
        // of the stall samples are no_inst -- and unrolled they were 955 of its 7 700 instructions)
#ifdef SB_GROUP_UNROLL_SOLVE
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int k = 0; k < N_; ++k) {             // L y = P b
            const int p = piv[k], pl = p / C, pc = p - pl * C;
            const double vk = __shfl_sync(gm, pick<C>(v, pc), pl, G);
#pragma unroll
            for (int c = 0; c < C; ++c) v[c] = (mystep[c] > k) ? fma(-m[c * N_ + k], vk, v[c]) : v[c];
        }
#ifdef SB_GROUP_UNROLL_SOLVE
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int k = N_ - 1; k >= 0; --k) {        // U x = y
            const int p = piv[k], pl = p / C, pc = p - pl * C;
            double mine = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c) mine = (pc == c) ? v[c] * m[c * N_ + k] : mine;
            const double xk = __shfl_sync(gm, mine, pl, G);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                v[c] = (mystep[c] >= 0 && mystep[c] < k) ? fma(-m[c * N_ + k], xk, v[c]) : v[c];
                x[c] = (r * C + c == k) ? xk : x[c];
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) b[c] = x[c];
    }
};

// ------------------------------------------------------------------------------------ backward
template <int G>
struct BwdSysG {
    using LG = LaneGroup<G>;
    static constexpr bool TSTOP = true;
#ifdef SB_CONSTRAINTS
    static constexpr bool CONSTR = false;
#endif
    static constexpr int GROUP = G, NS_FULL = NS, NQ_FULL = ND_;
    static constexpr bool MAT_SHARED = true;
    static constexpr int NQL = (ND + G - 1) / G > 0 ? (ND + G - 1) / G : 1;   // quadrature components per lane
    static constexpr int C = (NS + G - 1) / G;     // state components (and matrix rows) per lane:
                                                   // lane r owns r*C .. r*C + C - 1
    const SbBackwardArgs& a;
    using GroupIds = typename LG::Ids;
    GroupIds id;           // group mask and rank of this lane
    __device__ __forceinline__ explicit BwdSysG(const SbBackwardArgs& a_) : a(a_), id(LG::ids()) {}
    __device__ __forceinline__ double rtol() const { return a.rtol; }
    __device__ __forceinline__ double atol(int) const { return a.atol; }
    __device__ __forceinline__ double rtolQ() const { return a.rtol_q; }
    __device__ __forceinline__ double atolQ() const { return a.atol_q; }
    __device__ __forceinline__ double tstop() const { return a.t_end; }
    __device__ __forceinline__ static double gsum(double x, GroupIds g) { return LG::sum(x, g.gm); }
    __device__ __forceinline__ static double gmax(double x, GroupIds g) { return LG::max(x, g.gm); }
    __device__ __forceinline__ static bool gall(bool b, GroupIds g) { return LG::all(b, g.gm); }
    static constexpr int ROW = (C * NS + C) | 1;   // allocated doubles per lane and matrix (Bdf::MSA)
    __device__ __forceinline__ void add_identity(double* m) const {
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (id.r * C + c < NS) m[c * NS + id.r * C + c] += 1.0;
    }
    __device__ __forceinline__ bool lu_factor(double* m, int* piv) const {
        return LG::template lu_factor<NS, C, 2 * ROW + 1>(m, piv, id);
    }
    __device__ __forceinline__ void lu_solve(const double* m, const int* piv, double* b) const {
        LG::template lu_solve<NS, C>(m, piv, b, id);
    }

    const double* p;       // this instance's parameters (global memory, read where needed)
    const double* tab;     // this instance's table base
    int np;                // stored points; intervals are 1 .. np-1
    int idx;               // current interval (CVODES' ilast)
    double t;
    double* yi;            // forward solution interpolated at t: the group's [NS] shared-memory slot
    double* lamv;          // the vector an evaluation is made at, all components: another [NS] slot

    // as BwdSys::set_time; the lanes of a group share t and idx, read the knots together (one
    // broadcast load) and one column of the divided differences each (one 8*NS-byte segment per
    // row and group), then exchange the interpolated components
    __device__ __forceinline__ void set_time(double t_) {
        t = t_;
        const int r = id.r;
        int col[C];
#pragma unroll
        for (int c = 0; c < C; ++c) col[c] = r * C + c < NS ? r * C + c : 0;
        const double* e = tab + (size_t)idx * TAB_STRIDE;
        double lo, hi, inv_delt, T[SB_QMAX], Y[SB_LMAX][C];
        int order;
        bool went_left = false;
#pragma unroll 1
        for (;;) {
            lo = __ldg(e); hi = __ldg(e + 1);
            order = (int)__ldg(e + 2);
            inv_delt = __ldg(e + 3);
#pragma unroll
            for (int i = 0; i < SB_QMAX; ++i) T[i] = __ldg(e + 4 + i);
#pragma unroll
            for (int j = 0; j < SB_LMAX; ++j)
#pragma unroll
                for (int c = 0; c < C; ++c) Y[j][c] = __ldg(e + 10 + NS * j + col[c]);
            // CVAfindIndex: keep the interval while t_lo <= t <= t_hi, else walk
            if ((t < lo || (went_left && t <= lo)) && idx > 1) { --idx; e -= TAB_STRIDE; went_left = true; }
            else if (t > hi && idx < np - 1 && !went_left) { ++idx; e += TAB_STRIDE; }
            else break;
        }
        double mine[C];
#pragma unroll
        for (int c = 0; c < C; ++c) mine[c] = Y[0][c];
        double w = 1.0;
#pragma unroll
        for (int i = 0; i < SB_QMAX; ++i) {
            w = (i < order) ? w * ((t - T[i]) * inv_delt) : 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c) mine[c] = fma(w, Y[i + 1][c], mine[c]);
        }
        __syncwarp(id.gm);                        // the previous values have been consumed
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (r * C + c < NS) yi[r * C + c] = mine[c];
        __syncwarp(id.gm);
    }
    // every lane contributes its component; the generated functions then read what they need
    __device__ __forceinline__ void gather(const double* mine) const {
        __syncwarp(id.gm);                        // the previous contents have been consumed
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (id.r * C + c < NS) lamv[id.r * C + c] = mine[c];
        __syncwarp(id.gm);
    }
    // every lane evaluates the whole (cheap) function and keeps its own component
    __device__ __forceinline__ void rhs(const double* lam_mine, double* out_mine) const {
        double out[NS];
        gather(lam_mine);
        sb_adj_rhs(t, yi, lamv, p, out);
#pragma unroll
        for (int c = 0; c < C; ++c) out_mine[c] = LG::template pick<NS>(out, id.r * C + c);
    }
    __device__ __forceinline__ void jac(const double*, double* Jrow) const {
        double J[NS * NS];
        sb_adj_jac(t, yi, p, J);
        const int r = id.r;
#pragma unroll
        for (int c = 0; c < C; ++c)
            static_for<0, NS>([&](auto J_) {
                constexpr int j = SB_IDX(J_);
                Jrow[c * NS + j] = LG::template pick<NS>(J + NS * j, r * C + c);     // column-major: J[row + NS*j]
            });
    }
    __device__ __forceinline__ void quad(const double* lam_mine, double* out_mine) const {
        double out[ND_];
        gather(lam_mine);
#pragma unroll
        for (int i = 0; i < ND_; ++i) out[i] = 0.0;
        sb_quad_rhs(t, yi, lamv, p, out);
#pragma unroll
        for (int c = 0; c < NQL; ++c) out_mine[c] = LG::template pick<ND_>(out, id.r + G * c);
    }
};

// The backward driver of kernels.cuh (backward_unit<false>: lanes walk the intervals together) for
// grouped lanes.  `inst` / `valid` are per group; a lane owns lamda[rank] and the quadrature
// components rank, rank + G, ...
template <int G>
__device__ __forceinline__ void backward_unit_group(const SbBackwardArgs& a, long long inst, bool valid,
                                                    int k_begin, int k_end) {
    using Sys = BwdSysG<G>;
    using LG = LaneGroup<G>;
    constexpr int NQL = Sys::NQL;
    constexpr int C = Sys::C;
    using Integrator = Bdf<C, (ND > 0 ? NQL : 0), Sys>;
    if (!valid) inst = 0;
    Sys sys(a);
    const int r = sys.id.r;
    const unsigned gm = sys.id.gm;
    bool has_y[C];                              // the state components this lane owns
#pragma unroll
    for (int c = 0; c < C; ++c) has_y[c] = r * C + c < NS;
    const bool first = k_begin == 0, last = k_end == a.n_t + 1;
    const int np = a.hist_n[inst];

    // per-instance state in shared memory, one record per group (see BdfCtl)
    struct Shared { typename Integrator::Ctl ctl; double yi[NS]; double lamv[NS]; };
    __shared__ Shared sh_all[(SB_BLOCK / 32) * (32 / G)];
    Shared& sh = sh_all[(threadIdx.x >> 5) * (32 / G) + ((threadIdx.x & 31) / G)];
    // per-lane matrix rows, also in shared memory; consecutive lanes 2*ROW + 1 doubles apart
    struct Rows { typename Integrator::Mat mat; double pad; };
    static_assert(sizeof(Rows) == (2 * Sys::ROW + 1) * sizeof(double), "row stride");
    __shared__ Rows rows_all[SB_BLOCK];
#if SB_GROUP_SHARED_CTL
    Integrator bdf(sh.ctl, rows_all[threadIdx.x].mat);
#else
    typename Integrator::Ctl ctl_private;
    Integrator bdf(ctl_private, rows_all[threadIdx.x].mat);
#endif
    bdf.gid = sys.id;
    sys.yi = sh.yi;
    sys.lamv = sh.lamv;
    double lam[C], quad[NQL];
    int status;
    bdf.clear_stats();
    sys.idx = np > 1 ? np - 1 : 1;
    if (first) {
        status = a.fwd_status ? a.fwd_status[inst] : SB_SUCCESS;
#pragma unroll
        for (int c = 0; c < C; ++c) lam[c] = 0.0;
#pragma unroll
        for (int c = 0; c < NQL; ++c) quad[c] = 0.0;
    } else {
        const double* cd = a.carry_d + (size_t)inst * (NS + ND_);
        const int* ci = a.carry_i + (size_t)inst * SB_CARRY_INTS;
#pragma unroll
        for (int c = 0; c < C; ++c) lam[c] = has_y[c] ? cd[r * C + c] : 0.0;
#pragma unroll
        for (int c = 0; c < NQL; ++c) quad[c] = (r + G * c < ND) ? cd[NS + r + G * c] : 0.0;
        status = ci[0]; sys.idx = ci[1];
        bdf.st.nst = ci[2]; bdf.st.nfe = ci[3]; bdf.st.nje = ci[4]; bdf.st.nsetups = ci[5];
        bdf.st.netf = ci[6]; bdf.st.ncfn = ci[7]; bdf.st.nni = ci[8];
    }
    sys.p = a.params + inst * NP;
    sys.tab = a.tab + (size_t)inst * a.hist_cap * TAB_STRIDE;
    sys.np = np;
    sys.t = 0.0;
    bool joined = LG::converge(gm);                       // (the carry loads above are lane-dependent)
    bdf.reinit(a.t_start, lam, quad);

    const double* g_base = a.grads_shared ? a.grads : a.grads + (size_t)inst * a.n_t * NS;
    for (int k = k_begin; k < k_end; ++k) {
        const double t_upper = (k == 0) ? a.t_start : a.tvals[a.n_t - k];
        const double t_lower = (k == a.n_t) ? a.t_end : a.tvals[a.n_t - 1 - k];
        joined = LG::converge(gm) && joined;
        if (t_lower < t_upper) {                        // warp-uniform: tvals are shared
            if (valid && status == SB_SUCCESS && np < 2) status = SB_ILL_INPUT;
            const bool live = valid && status == SB_SUCCESS;
            if (live) {
                bdf.reinit(t_upper, lam, quad);         // CVodeReInitB + CVodeQuadReInitB
                status = bdf.first_call(sys, t_lower);
                joined = LG::together(gm) && joined;      // (the restart ran with the group together)
            }
            int nloc = 0;
            bool reached = false;
            for (;;) {
                bool work = valid && status == SB_SUCCESS && !reached;
                if (work && !bdf.in_step) {
                    if (nloc >= a.max_steps) status = SB_TOO_MUCH_WORK;
                    else status = bdf.pre_step_checks(sys);
                    work = status == SB_SUCCESS;
                }
                const unsigned mask = sb_ballot(work);
                if (mask == 0u) break;
                // (the pass that just ended must have left the group together: its shared-memory
                // updates were made after the previous join)
                joined = LG::together(gm) && joined;
                joined = LG::converge(gm) && joined;
                if (work) {
                    const int rr = bdf.attempt(sys, mask);
                    if (rr == SB_SUCCESS) {
                        nloc++;
                        bdf.snap_to_tstop(sys);
                        if ((bdf.tn - t_lower) * bdf.h >= 0.0) reached = true;
                        else bdf.limit_to_tstop(sys);
                    } else if (rr != SB_TRY_AGAIN) {
                        status = rr;
                    }
                }
            }
            if (valid && status == SB_SUCCESS) {
                bdf.get_dky(t_lower, lam);                // CVodeGetB
                if (ND > 0) bdf.get_quad(t_lower, quad);  // CVodeGetQuadB, carried into the next interval
            }
        }
        if (valid && k < a.n_t) {
            const double* g = g_base + (size_t)(a.n_t - 1 - k) * NS;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (has_y[c]) lam[c] -= g[r * C + c];
            if (a.lamda_all || a.quad_all) {
                const size_t row = (size_t)inst * a.n_t + (size_t)((a.n_t - k) % a.n_t);
                const bool ok = status == SB_SUCCESS;
                if (a.lamda_all)
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (has_y[c]) a.lamda_all[row * NS + r * C + c] = ok ? lam[c] : qnan();
                if (a.quad_all)
#pragma unroll
                    for (int c = 0; c < NQL; ++c)
                        if (r + G * c < ND) a.quad_all[row * ND + r + G * c] = ok ? quad[c] : qnan();
            }
        }
    }
    joined = LG::converge(gm) && joined;
    if (!joined && status == SB_SUCCESS) status = SB_GROUP_SPLIT;    // SB_GROUP_CHECK builds only
    if (!valid) return;
    if (!last) {
        double* cd = a.carry_d + (size_t)inst * (NS + ND_);
        int* ci = a.carry_i + (size_t)inst * SB_CARRY_INTS;
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (has_y[c]) cd[r * C + c] = lam[c];
#pragma unroll
        for (int c = 0; c < NQL; ++c)
            if (r + G * c < ND) cd[NS + r + G * c] = quad[c];
        if (r == 0) {
            ci[0] = status; ci[1] = sys.idx;
            ci[2] = bdf.st.nst; ci[3] = bdf.st.nfe; ci[4] = bdf.st.nje; ci[5] = bdf.st.nsetups;
            ci[6] = bdf.st.netf; ci[7] = bdf.st.ncfn; ci[8] = bdf.st.nni;
        }
        return;
    }
    const bool bad = status != SB_SUCCESS;
#pragma unroll
    for (int c = 0; c < C; ++c)
        if (has_y[c]) a.lamda_out[inst * NS + r * C + c] = bad ? qnan() : lam[c];
#pragma unroll
    for (int c = 0; c < NQL; ++c)
        if (r + G * c < ND) a.grad_out[inst * ND + r + G * c] = bad ? qnan() : quad[c];
    if (r == 0) {
        a.status[inst] = status;
        if (a.stats) {
            int* s = a.stats + inst * SB_STATS_STRIDE;
            s[0] = bdf.st.nst; s[1] = bdf.st.nfe; s[2] = bdf.st.nje; s[3] = bdf.st.nsetups;
            s[4] = bdf.st.netf; s[5] = bdf.st.ncfn; s[6] = bdf.st.nni; s[7] = np;
        }
    }
}


// ------------------------------------------------------------------------------------ forward
// The forward kernels (sb_forward, sb_forward_sens) in lane groups: the same split of the state
// over the G lanes of a group as in the backward kernel -- lane r owns the components
// r*C .. r*C + C - 1 of y (and of every sensitivity block) and the same rows of the saved Jacobian
// and of I - gamma*J.  Every lane evaluates the whole generated right-hand side on the group's
// copy of the evaluation vector (shared memory) and keeps its components.
template <int G, int NBLK>
struct FwdSysG {
    using LG = LaneGroup<G>;
    static constexpr bool TSTOP = false;
#ifdef SB_CONSTRAINTS
    static constexpr bool CONSTR = false;          // constraint builds run one lane per instance
#endif
    static constexpr int GROUP = G, NS_FULL = NS, NQ_FULL = 1;
    static constexpr bool MAT_SHARED = true;
    static constexpr int C = (NS + G - 1) / G;     // state components (and matrix rows) per lane
    const SbForwardArgs& a;
    using GroupIds = typename LG::Ids;
    GroupIds id;
    __device__ __forceinline__ explicit FwdSysG(const SbForwardArgs& a_) : a(a_), id(LG::ids()) {}
    __device__ __forceinline__ double rtol() const { return a.rtol; }
    // i = b*C + c: component c of block b held by this lane (padding rows use the last state's)
    __device__ __forceinline__ double atol(int i) const {
        const int b = i / C, c = i - b * C;
        const int g = id.r * C + c;
        return __ldg(a.atol + b * NS + (g < NS ? g : NS - 1));
    }
    __device__ __forceinline__ double rtolQ() const { return 0.0; }
    __device__ __forceinline__ double atolQ() const { return 1.0; }
    __device__ __forceinline__ double tstop() const { return 0.0; }
    __device__ __forceinline__ static double gsum(double x, GroupIds g) { return LG::sum(x, g.gm); }
    __device__ __forceinline__ static double gmax(double x, GroupIds g) { return LG::max(x, g.gm); }
    __device__ __forceinline__ static bool gall(bool b, GroupIds g) { return LG::all(b, g.gm); }
    static constexpr int ROW = (C * NS + C) | 1;
    __device__ __forceinline__ void add_identity(double* m) const {
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (id.r * C + c < NS) m[c * NS + id.r * C + c] += 1.0;
    }
    __device__ __forceinline__ bool lu_factor(double* m, int* piv) const {
        return LG::template lu_factor<NS, C, 2 * ROW + 1>(m, piv, id);
    }
    __device__ __forceinline__ void lu_solve(const double* m, const int* piv, double* b) const {
        LG::template lu_solve<NS, C>(m, piv, b, id);
    }

    const double* p;       // this instance's parameters (global memory)
    double t;
    double* yv;            // the vector an evaluation is made at, all blocks: the group's [NS * NBLK] slot
    __device__ __forceinline__ void set_time(double t_) { t = t_; }
    __device__ __forceinline__ void gather(const double* mine) const {
        __syncwarp(id.gm);                        // the previous contents have been consumed
#pragma unroll
        for (int b = 0; b < NBLK; ++b)
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (id.r * C + c < NS) yv[b * NS + id.r * C + c] = mine[b * C + c];
        __syncwarp(id.gm);
    }
    __device__ __forceinline__ void rhs(const double* y_mine, double* out_mine) const {
        double out[NS];
        gather(y_mine);
        sb_rhs(t, yv, p, out);
#pragma unroll
        for (int c = 0; c < C; ++c) out_mine[c] = LG::template pick<NS>(out, id.r * C + c);
        if constexpr (NBLK > 1) {
            double so[(NBLK - 1) * NS];
            sb_sens_rhs(t, yv, yv + NS, p, so);
#pragma unroll
            for (int b = 1; b < NBLK; ++b)
#pragma unroll
                for (int c = 0; c < C; ++c)
                    out_mine[b * C + c] = LG::template pick<NS>(so + (b - 1) * NS, id.r * C + c);
        }
    }
    __device__ __forceinline__ void jac(const double* y_mine, double* Jrow) const {
        double J[NS * NS];
        gather(y_mine);
        sb_jac(t, yv, p, J);
        const int r = id.r;
#pragma unroll
        for (int c = 0; c < C; ++c)
            static_for<0, NS>([&](auto J_) {
                constexpr int j = SB_IDX(J_);
                Jrow[c * NS + j] = LG::template pick<NS>(J + NS * j, r * C + c);     // column-major: J[row + NS*j]
            });
    }
    __device__ __forceinline__ void quad(const double*, double*) const {}
};

// forward_instance_t of sb_kernels.cuh for grouped lanes.  `inst` / `valid` are per group; the
// output-time loop is flattened into the step loop as there, and all lanes of a group take every
// branch together (the controller record is the group's, in shared memory).
template <int G, int NBLK>
__device__ __forceinline__ void forward_instance_group(const SbForwardArgs& a, long long inst, bool valid) {
    using Sys = FwdSysG<G, NBLK>;
    using LG = LaneGroup<G>;
    constexpr int C = Sys::C;
    constexpr int NL = C * NBLK;                  // components this lane holds
    using Integrator = Bdf<C, 0, Sys, NBLK>;
    if (!valid) inst = 0;
    Sys sys(a);
    const int r = sys.id.r;
    const unsigned gm = sys.id.gm;
    bool has_y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) has_y[c] = r * C + c < NS;
    const bool writer = r == 0;

    struct Shared { typename Integrator::Ctl ctl; double yv[NS * NBLK]; };
    __shared__ Shared sh_all[(SB_BLOCK / 32) * (32 / G)];
    Shared& sh = sh_all[(threadIdx.x >> 5) * (32 / G) + ((threadIdx.x & 31) / G)];
    struct Rows { typename Integrator::Mat mat; double pad; };
    static_assert(sizeof(Rows) == (2 * Sys::ROW + 1) * sizeof(double), "row stride");
    __shared__ Rows rows_all[SB_BLOCK];
#if SB_GROUP_SHARED_CTL
    Integrator bdf(sh.ctl, rows_all[threadIdx.x].mat);
#else
    typename Integrator::Ctl ctl_private;
    Integrator bdf(ctl_private, rows_all[threadIdx.x].mat);
#endif
    bdf.gid = sys.id;
    sys.yv = sh.yv;
    sys.p = a.params + inst * NP;
    sys.t = 0.0;

    double y0[NL];
#pragma unroll
    for (int c = 0; c < C; ++c) y0[c] = has_y[c] ? a.y0[inst * NS + r * C + c] : 0.0;
    if constexpr (NBLK > 1) {
        const double* s0 = a.sens0_shared ? a.sens0 : a.sens0 + (size_t)inst * (NBLK - 1) * NS;
#pragma unroll
        for (int b = 1; b < NBLK; ++b)
#pragma unroll
            for (int c = 0; c < C; ++c) y0[b * C + c] = has_y[c] ? s0[(b - 1) * NS + r * C + c] : 0.0;
    }
    bool joined = LG::converge(gm);               // (the loads above are lane-dependent)
    bdf.clear_stats();
    bdf.reinit(a.t0, y0, nullptr);

    double* yo = a.y_out + (size_t)inst * a.n_t * NS;
    double* so = (NBLK > 1) ? a.sens_out + (size_t)inst * a.n_t * (NBLK - 1) * NS : nullptr;
    double* hist = a.hist ? a.hist + (size_t)inst * a.hist_cap * HIST_STRIDE : nullptr;
    // (t, order, y[, y']) of a stored point: every lane its components, lane 0 the header
    auto store = [&](int idx, int order) {
        double* e = hist + (size_t)idx * HIST_STRIDE;
        if (writer) { e[0] = bdf.tn; e[1] = (double)order; }
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (has_y[c]) {
                e[2 + r * C + c] = bdf.zn[0][c];
#ifdef SB_HERMITE
                e[2 + NS + r * C + c] = bdf.zn[1][c] / bdf.h;
#endif
            }
    };
    auto emit = [&](int k, const double* v) {     // row k of the outputs from this lane's components
#pragma unroll
        for (int c = 0; c < C; ++c)
            if (has_y[c]) {
                yo[(size_t)k * NS + r * C + c] = v[c];
                if constexpr (NBLK > 1) {
#pragma unroll
                    for (int b = 1; b < NBLK; ++b)
                        so[((size_t)k * (NBLK - 1) + (b - 1)) * NS + r * C + c] = v[b * C + c];
                }
            }
    };
    int status = SB_SUCCESS;
    int k = 0;          // next output time
    int nloc = 0;       // internal steps taken towards tvals[k]

    for (;;) {
        joined = LG::converge(gm) && joined;
        bool work = valid && status == SB_SUCCESS && k < a.n_t;
        if (work && !bdf.in_step) {
            for (;;) {
                const double tout = a.tvals[k];
                if (tout == a.t0) {
                    emit(0, y0);                  // the reference writes row 0 here whatever k is
                } else if (bdf.nst > 0 && (bdf.tn - tout) * bdf.h >= 0.0) {
                    double yk[NL];
                    bdf.get_dky(tout, yk);
                    emit(k, yk);
                } else {
                    break;
                }
                nloc = 0;
                if (++k == a.n_t) { work = false; break; }
            }
        }
        joined = LG::converge(gm) && joined;
        if (work && !bdf.in_step) {
            if (bdf.nst == 0) {
                status = bdf.first_call(sys, a.tvals[k]);
                if (status == SB_SUCCESS && hist) store(0, 0);
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
        joined = LG::together(gm) && joined;
        joined = LG::converge(gm) && joined;
        if (work) {
            const int rr = bdf.attempt(sys, mask);
            if (rr == SB_SUCCESS) {
                nloc++;
                if (hist) store(bdf.nst, bdf.qu);
            } else if (rr != SB_TRY_AGAIN) {
                status = rr;
            }
        }
    }
    joined = LG::converge(gm) && joined;
    if (!joined && status == SB_SUCCESS) status = SB_GROUP_SPLIT;    // SB_GROUP_CHECK builds only
    const int nst_total = bdf.nst;
    if (a.steps_total) {
#ifndef SB_HOST_EMULATION
        unsigned n = (valid && writer && status == SB_SUCCESS) ? (unsigned)nst_total : 0u;
        n = __reduce_add_sync(0xffffffffu, n);
        if ((threadIdx.x & 31) == 0 && n) atomicAdd(a.steps_total, (unsigned long long)n);
#else
        if (writer && status == SB_SUCCESS) *a.steps_total += (unsigned long long)nst_total;
#endif
    }
    if (!valid) return;
    if (status != SB_SUCCESS) {
        for (int j = 0; j < a.n_t; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (has_y[c]) {
                    yo[(size_t)j * NS + r * C + c] = qnan();
                    if constexpr (NBLK > 1) {
#pragma unroll
                        for (int b = 1; b < NBLK; ++b)
                            so[((size_t)j * (NBLK - 1) + (b - 1)) * NS + r * C + c] = qnan();
                    }
                }
    }
    if (!writer) return;
    a.status[inst] = status;
    if (a.fail_k) a.fail_k[inst] = (status == SB_SUCCESS) ? -1 : min(k, a.n_t - 1);
    if (a.hist_n) a.hist_n[inst] = (status == SB_SUCCESS) ? nst_total + 1 : 0;
    if (a.stats) {
        int* s = a.stats + inst * SB_STATS_STRIDE;
        s[0] = bdf.st.nst; s[1] = bdf.st.nfe; s[2] = bdf.st.nje; s[3] = bdf.st.nsetups;
        s[4] = bdf.st.netf; s[5] = bdf.st.ncfn; s[6] = bdf.st.nni; s[7] = nst_total + 1;
    }
}

}  // namespace sb
