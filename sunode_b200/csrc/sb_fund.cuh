// sb_fund.cuh -- restart-free backward pass, an opt-in build of the kernels
// (-DSB_FUND, AdjointSolver(backward='fundamental')).  Included at the end of sb_kernels.cuh.
//
// The reference restarts the backward integrator at every output time because lamda jumps there
// (/root/reference/sunode/solver.py:750-784: CVodeReInitB + `lamda -= g` per tval): 49 restarts
// from order 1 make ~1 660 backward steps out of what a smooth problem would do in ~200.  The
// adjoint equation is LINEAR in lamda, so the jumps can be taken out of the integration exactly:
// with the fundamental matrix
//     Psi' = -J(t, y(t))^T Psi,  Psi(t_r) = I,       W' = Psi^T df/dp,  W(t_r) = 0
// (n_s columns that share one Newton matrix -- the integrator's NBLK blocks -- and an
// n_d x n_s quadrature) the solution between two output times is lamda(t) = Psi(t) a with a
// CONSTANT coefficient vector a, the quadrature grows by sum_b a_b dW_b, and the jump
// lamda -= g at an output time t_k becomes a -= Psi(t_k)^{-1} g: a small dense solve on the dense
// output of an integration that never restarts.  Psi loses conditioning where the problem
// contracts (a stiff forward problem makes the backward fundamental matrix collapse onto its
// slow directions), so the pivots of that solve are watched and the block is re-based (Psi = I at
// t_k, a = lamda) when their ratio exceeds SB_FUND_COND: for stiff problems this degenerates into
// the reference's schedule, for smooth ones the restarts disappear.
//
// Same mathematics, different discretisation: results agree with the reference schedule to the
// tolerances, not step for step (tests/test_options.py pins both against a 1e-12 solve).
#pragma once

#ifndef SB_FUND_COND
#define SB_FUND_COND 1.0e3      /* largest / smallest pivot of Psi(t_k) before the block is re-based */
#endif

namespace sb {

struct FundSys : BwdSys {
    static constexpr int NQ_FULL = ND_ * NS;
    __device__ __forceinline__ explicit FundSys(const SbBackwardArgs& a_) : BwdSys(a_) {}
    __device__ __forceinline__ void rhs(const double* psi, double* out) const {
#pragma unroll
        for (int b = 0; b < NS; ++b) sb_adj_rhs(t, yi, psi + b * NS, p, out + b * NS);
    }
    __device__ __forceinline__ void quad(const double* psi, double* out) const {
#pragma unroll
        for (int b = 0; b < NS; ++b) sb_quad_rhs(t, yi, psi + b * NS, p, out + b * ND_);
    }
};

// One lane = one instance (like the forward kernel; the output-time loop is flattened into the
// step loop so that lanes of a warp do not wait for each other at the output times).
__device__ __forceinline__ void backward_fund_instance(const SbBackwardArgs& a, long long inst, bool valid) {
    constexpr int NN = NS * NS;
    constexpr int NQF = ND * NS;
    constexpr int NQF_ = NQF > 0 ? NQF : 1;
    using Integrator = Bdf<NS, NQF, FundSys, NS>;
    if (!valid) inst = 0;
    const int np = a.hist_n[inst];
    typename Integrator::Ctl ctl;
    typename Integrator::Mat mat_regs;
    typename Integrator::Mat& mat = mat_home<typename Integrator::Mat, FundSys::MAT_SHARED>(mat_regs);
    Integrator bdf(ctl, mat);
    bdf.in_step = false;        // read by the driver before the first reinit() (the record is
    bdf.nst = 0;                // deliberately left uninitialised otherwise)
    FundSys sys(a);
    double lam[NS], coef[NS], quad[ND_], wprev[NQF_];
    int status = a.fwd_status ? a.fwd_status[inst] : SB_SUCCESS;
#pragma unroll
    for (int i = 0; i < NS; ++i) { lam[i] = 0.0; coef[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < ND_; ++i) quad[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NQF_; ++i) wprev[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) sys.p[i] = a.params[inst * NP + i];
    sys.tab = a.tab + (size_t)inst * a.hist_cap * TAB_STRIDE;
    sys.np = np;
    sys.idx = np > 1 ? np - 1 : 1;
    sys.t = 0.0;
    bdf.clear_stats();
    const double* g_base = a.grads_shared ? a.grads : a.grads + (size_t)inst * a.n_t * NS;

    // optional traces (solver.py:778-781): lamda / quadrature right after the jump at the lower end
    // of interval k < n_t, same row convention as backward_unit
    auto trace = [&](int k, const double* lam_now) {
        const size_t row = (size_t)inst * a.n_t + (size_t)((a.n_t - k) % a.n_t);
        if (a.lamda_all)
#pragma unroll
            for (int i = 0; i < NS; ++i) a.lamda_all[row * NS + i] = lam_now[i];
        if (a.quad_all)
#pragma unroll
            for (int i = 0; i < ND; ++i) a.quad_all[row * ND + i] = quad[i];
    };
    const bool tracing = a.lamda_all || a.quad_all;

    // ts = [t_start] + reversed(tvals) + [t_end]; interval k is (ts[k+1], ts[k]) (solver.py:750-754)
    const int k_end = a.n_t + 1;
    int k = valid ? 0 : k_end;
    bool open = false;          // a block (Psi, W) is being integrated; lamda = Psi coef
    int nloc = 0, nrebase = 0;
    for (;;) {
        if (k < k_end && status == SB_SUCCESS && !bdf.in_step) {
            // everything that needs no further step: empty intervals, output times already
            // stepped past (their jumps, re-basing), opening a block
            for (;;) {
                const double t_upper = (k == 0) ? a.t_start : a.tvals[a.n_t - k];
                const double t_lower = (k == a.n_t) ? a.t_end : a.tvals[a.n_t - 1 - k];
                const double* g = (k < a.n_t) ? g_base + (size_t)(a.n_t - 1 - k) * NS : nullptr;
                if (!open) {
                    if (t_lower < t_upper) {
                        if (np < 2) { status = SB_ILL_INPUT; break; }
                        double eye[NN], zero[NQF_];
#pragma unroll
                        for (int i = 0; i < NN; ++i) eye[i] = (i % (NS + 1) == 0) ? 1.0 : 0.0;
#pragma unroll
                        for (int i = 0; i < NQF_; ++i) { zero[i] = 0.0; wprev[i] = 0.0; }
#pragma unroll
                        for (int i = 0; i < NS; ++i) coef[i] = lam[i];
                        bdf.reinit(t_upper, eye, zero);
                        status = bdf.first_call(sys, t_lower);
                        open = true; nloc = 0;
                        break;                       // needs steps (or failed)
                    }
                    // an empty interval: only the jump (solver.py:770-776)
                    if (g) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) lam[i] -= g[i];
                        if (tracing) trace(k, lam);
                    }
                } else {
                    if (!(bdf.nst > 0 && (bdf.tn - t_lower) * bdf.h >= 0.0)) break;   // step on
                    double psi[NN];
                    bdf.get_dky(t_lower, psi);
                    if (ND > 0) {
                        double w[NQF_];
                        bdf.get_quad(t_lower, w);
#pragma unroll
                        for (int b = 0; b < NS; ++b)
#pragma unroll
                            for (int j = 0; j < ND; ++j)
                                quad[j] = fma(coef[b], w[b * ND_ + j] - wprev[b * ND_ + j], quad[j]);
#pragma unroll
                        for (int i = 0; i < NQF_; ++i) wprev[i] = w[i];
                    }
                    bool solved = false;
                    if (g) {
                        // lamda -= g in the basis Psi(t_lower): coef -= Psi^{-1} g
                        double lu[NN], x[NS];
                        int piv[NS];
#pragma unroll
                        for (int i = 0; i < NN; ++i) lu[i] = psi[i];
#pragma unroll
                        for (int i = 0; i < NS; ++i) x[i] = g[i];
                        const bool ok = lu_factor<NS>(lu, piv);
                        double dmin = 1.7976931348623157e308, dmax = 0.0;   // reciprocal pivots
#pragma unroll
                        for (int i = 0; i < NS; ++i) {
                            const double d = fabs(lu[i + NS * i]);
                            dmin = fmin(dmin, d); dmax = fmax(dmax, d);
                        }
                        if (ok && dmax <= SB_FUND_COND * dmin) {
                            lu_solve<NS>(lu, piv, x);
#pragma unroll
                            for (int i = 0; i < NS; ++i) coef[i] -= x[i];
                            solved = true;
                        }
                    }
                    // lamda itself is needed at the end, and where Psi(t_lower) has become too
                    // ill-conditioned for the solve: there the jump is applied to lamda and the
                    // next interval starts a new block from it
                    if (k == a.n_t || (g && !solved)) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) {
                            double s = 0.0;
#pragma unroll
                            for (int b = 0; b < NS; ++b) s = fma(psi[b * NS + i], coef[b], s);
                            lam[i] = s;
                        }
                        if (g) {
#pragma unroll
                            for (int i = 0; i < NS; ++i) lam[i] -= g[i];
                            nrebase++;
                            if (tracing) trace(k, lam);
                        }
                        open = false;
                    } else if (g && tracing) {
                        double lam_now[NS];
#pragma unroll
                        for (int i = 0; i < NS; ++i) {
                            double s = 0.0;
#pragma unroll
                            for (int b = 0; b < NS; ++b) s = fma(psi[b * NS + i], coef[b], s);
                            lam_now[i] = s;
                        }
                        trace(k, lam_now);
                    }
                }
                nloc = 0;
                if (++k == k_end) break;
            }
        }
        bool work = k < k_end && status == SB_SUCCESS && open;
        if (work && !bdf.in_step) {
            if (nloc >= a.max_steps) status = SB_TOO_MUCH_WORK;
            else status = bdf.pre_step_checks(sys);
            work = status == SB_SUCCESS;
        }
        const unsigned mask = sb_ballot(work);
        if (mask == 0u) break;      // every lane is through its intervals (or has failed)
        if (work) {
            const int r = bdf.attempt(sys, mask);
            if (r == SB_SUCCESS) {
                nloc++;
                bdf.snap_to_tstop(sys);
                bdf.limit_to_tstop(sys);
            } else if (r != SB_TRY_AGAIN) {
                status = r;
            }
        }
    }
    if (!valid) return;
    double* gout = a.grad_out + inst * ND;
    double* lout = a.lamda_out + inst * NS;
    const bool ok = status == SB_SUCCESS;
#pragma unroll
    for (int i = 0; i < ND; ++i) gout[i] = ok ? quad[i] : qnan();
#pragma unroll
    for (int i = 0; i < NS; ++i) lout[i] = ok ? lam[i] : qnan();
    if (!ok && tracing) {
        // a failed instance reads as NaN everywhere (as_pytensor.py:339-341)
        for (int j = 0; a.lamda_all && j < a.n_t * NS; ++j) a.lamda_all[(size_t)inst * a.n_t * NS + j] = qnan();
        for (int j = 0; a.quad_all && j < a.n_t * ND; ++j) a.quad_all[(size_t)inst * a.n_t * ND + j] = qnan();
    }
    a.status[inst] = status;
    if (a.stats) {
        int* s = a.stats + inst * SB_STATS_STRIDE;
        s[0] = bdf.st.nst; s[1] = bdf.st.nfe; s[2] = bdf.st.nje; s[3] = bdf.st.nsetups;
        s[4] = bdf.st.netf; s[5] = bdf.st.ncfn; s[6] = bdf.st.nni; s[7] = nrebase;
    }
}

}  // namespace sb
