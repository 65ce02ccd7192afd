/*
 * oracle/cvodes_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the algorithm the reference's hot path runs: sunode's
 * Solver.solve / AdjointSolver.solve_forward / AdjointSolver.solve_backward
 * (/root/reference/sunode/solver.py:467-527, 682-721, 723-784) on top of SUNDIALS CVODES 5.x
 * (conda-forge `sundials<6.0`, /root/reference/.github/workflows/main.yml:36).  CVODES itself is
 * an un-vendored third-party dependency of the reference (no sources under /root/reference, no
 * library in this image), so the integrator below restates its published algorithm --
 * variable-order variable-step fixed-leading-coefficient BDF in Nordsieck form, modified Newton
 * with a dense LU of I - gamma*J, WRMS error control, quadrature variables with error control,
 * adjoint data store with variable-degree polynomial interpolation -- with CVODES' default
 * constants, as configured by the reference's call sites:
 *
 *   solver.py:221-235,328-334  CVodeCreate(CV_BDF), dense linear solver, analytic Jacobian
 *   solver.py:565-622          the same for the adjoint solver's forward problem
 *   solver.py:588              CVodeAdjInit(steps=500000, CV_POLYNOMIAL)  -> one data segment
 *   solver.py:592-615          backward problem: BDF, dense LS, analytic -J^T, tolerances 1e-10,
 *                              quadrature with error control
 *   solver.py:503-521          forward loop: CVode(CV_NORMAL) per tval, <=5 retries on TOO_MUCH_WORK
 *   solver.py:705-721          adjoint forward loop: CVodeF per tval
 *   solver.py:750-784          backward loop: interval list, CVodeReInitB + CVodeQuadReInitB at
 *                              every observation time, lamda -= g jumps, quadrature carry-over
 *
 * PARITY STATUS: "parity unpinned" for the integrator internals -- the reference's tests hold no
 * numeric assertions for this path (sunode/test_solve.py asserts nothing numeric) and CVODES
 * cannot be run here.  The oracle is instead pinned (tests/test_oracle.py) against the one
 * recorded CVODES run the reference ships (notebooks/from_sympy.ipynb:240-242), closed-form
 * solutions, SciPy DOP853 at 1e-13 and finite differences.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NMAX 96          /* max length of the (stacked) state vector / quadratures handled by the oracle */
#define QMAX 5
#define LMAX (QMAX + 1)

/* return codes (reference include/cvodes/16_cvodes.h:45-106) */
#define CV_SUCCESS 0
#define CV_TSTOP_RETURN 1
#define CV_TOO_MUCH_WORK (-1)
#define CV_TOO_MUCH_ACC (-2)
#define CV_ERR_FAILURE (-3)
#define CV_CONV_FAILURE (-4)
#define CV_LSETUP_FAIL (-6)
#define CV_RHSFUNC_FAIL (-8)
#define CV_FIRST_RHSFUNC_ERR (-9)
#define CV_REPTD_RHSFUNC_ERR (-10)
#define CV_UNREC_RHSFUNC_ERR (-11)
#define CV_CONSTR_FAIL (-15)
#define CV_ILL_INPUT (-22)
#define CV_BAD_T (-25)
#define CV_TOO_CLOSE (-27)
#define CV_GETY_BADT (-107)

#define CV_NORMAL 1
#define CV_ONE_STEP 2

/* CVODES constants */
#define ETAMX1 10000.0
#define ETAMX2 10.0
#define ETAMX3 10.0
#define ETAMXF 0.2
#define ETAMIN 0.1
#define ETACF 0.25
#define ADDON 0.000001
#define BIAS1 6.0
#define BIAS2 6.0
#define BIAS3 10.0
#define THRESH 1.5
#define ONEPSM 1.000001
#define MXNEF1 3
#define SMALL_NEF 2
#define SMALL_NST 10
#define LONG_WAIT 10
#define MXNCF 10
#define MXNEF 7
#define NLS_MAXCOR 3
#define CRDOWN 0.3
#define RDIV 2.0
#define DGMAX 0.3
#define MSBP 20
#define MSBJ 50
#define CVLS_DGMAX 0.2
#define NLSCOEF 0.1
#define HLB_FACTOR 100.0
#define HUB_FACTOR 0.1
#define H_BIAS 0.5
#define HIN_MAX_ITERS 4
#define FUZZ_FACTOR 100.0
#define UROUND DBL_EPSILON

/* internal flags */
#define FIRST_CALL 101
#define PREV_CONV_FAIL 102
#define PREV_ERR_FAIL 103
#define DO_ERROR_TEST 2
#define PREDICT_AGAIN 3
#define TRY_AGAIN 5
#define CONV_FAIL 4
#define RHSFUNC_RECVR 9
#define CONSTR_RECVR 10
#define QRHSFUNC_RECVR 11
#define NO_FAILURES 0
#define FAIL_BAD_J 1
#define FAIL_OTHER 2

/* ---- problem callbacks: the calling convention of the generated host module ------------------ */
typedef int (*fn3_t)(double t, const double* y, const double* p, double* out);
typedef int (*fn4_t)(double t, const double* y, const double* v, const double* p, double* out);

typedef struct {
    int ns, np, nd;
    fn3_t rhs;       /* f(t, y, p) */
    fn3_t jac;       /* column-major df/dy */
    fn4_t adj_rhs;   /* -J^T lam */
    fn3_t adj_jac;   /* column-major -J^T */
    fn4_t quad_rhs;  /* lam^T df/dp */
    fn4_t sens_rhs;  /* (t, y, s[nd][ns], p, out[nd][ns]): J s_k + df/dp_k (may be NULL) */
} oracle_problem;

/* ---- adjoint data store ---------------------------------------------------------------------- */
typedef struct {
    int np;          /* number of stored points */
    int cap;
    double* t;       /* [cap] */
    double* y;       /* [cap][ns] */
    double* yd;      /* [cap][ns] y' = zn[1] / h (CV_HERMITE only, else NULL) */
    int* order;      /* [cap] */
    int ns;
    int hermite;     /* CV_HERMITE instead of CV_POLYNOMIAL (solver.py:581-586) */
    double Y0h[NMAX], Y1h[NMAX];   /* CVAhermiteGetY's cached Y[0], Y[1] */
    /* interpolation cache */
    int ilast, newdata;
    int ord_cached;
    double T[LMAX];
    double Y[LMAX][NMAX];
    double delt;
} hist_t;

/* ---- integrator memory ----------------------------------------------------------------------- */
struct cv_mem;
typedef int (*sys_rhs_t)(struct cv_mem*, double t, const double* y, double* ydot);
typedef int (*sys_jac_t)(struct cv_mem*, double t, const double* y, double* J);
typedef int (*sys_quad_t)(struct cv_mem*, double t, const double* y, double* qdot);

typedef struct cv_mem {
    int N, NQ;
    /* forward sensitivity analysis (CV_SIMULTANEOUS, sensitivity error control on;
     * /root/reference/sunode/solver.py:360-392): y and the sensitivity vectors are stacked,
     * N = NM * nblk; all blocks share h, q and the iteration matrix (dimension NM); norms are the
     * max over the per-block WRMS norms.  nblk = 1: plain problem (NM == N). */
    int NM, nblk;
    sys_rhs_t f;
    sys_jac_t jacfn;
    sys_quad_t fQ;
    const oracle_problem* prob;
    const double* p;
    hist_t* hist;            /* backward problems interpolate the forward solution from here */
    const double* constraints;   /* CVodeSetConstraints flags per component (0, +-1, +-2) or NULL */

    double reltol, abstol[NMAX];
    double reltolQ, abstolQ;
    int quadr, errconQ;

    double zn[LMAX][NMAX], ewt[NMAX], acor[NMAX], y[NMAX], tempv[NMAX], ftemp[NMAX];
    double znQ[LMAX][NMAX], ewtQ[NMAX], acorQ[NMAX], yQ[NMAX], tempvQ[NMAX];

    int q, qprime, next_q, qwait, L, qu;
    double hin, h, hprime, next_h, eta, hscale, tn, tretlast, hu, h0u;
    double tau[LMAX + 1], tq[6], l[LMAX];
    double rl1, gamma, gammap, gamrat, crate, delp, acnrm, nlscoef;
    double etaqm1, etaq, etaqp1, etamax, saved_tq5, tolsf;
    int mxstep;
    int tstopset; double tstop;
    double hmin, hmax_inv;

    long nst, nfe, nfQe, ncfn, netf, netfQ, nni, nsetups, nje, nstlp, nstlj, nhnil;
    int jcur, forceSetup, convfail;

    double savedJ[24 * 24], M[24 * 24];
    int piv[NMAX];
} cv_mem;

/* ---- small vector helpers ---------------------------------------------------------------------*/
static double wrms(const double* v, const double* w, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) { double x = v[i] * w[i]; s += x * x; }
    return sqrt(s / n);
}

/* max over blocks of the WRMS norm (N_VWrmsNorm of CVODES' sensitivity wrapper vector) */
static double vnorm(const cv_mem* m, const double* v, const double* w) {
    double r = 0.0;
    for (int b = 0; b < m->nblk; ++b) {
        double x = wrms(v + b * m->NM, w + b * m->NM, m->NM);
        if (b == 0 || x > r) r = x;
    }
    return r;
}

static int ewt_set(cv_mem* m, const double* ycur, double* w) {
    for (int i = 0; i < m->N; ++i) {
        double d = m->reltol * fabs(ycur[i]) + m->abstol[i];
        if (d <= 0.0) return -1;
        w[i] = 1.0 / d;
    }
    return 0;
}

static int ewtQ_set(cv_mem* m, const double* qcur, double* w) {
    for (int i = 0; i < m->NQ; ++i) {
        double d = m->reltolQ * fabs(qcur[i]) + m->abstolQ;
        if (d <= 0.0) return -1;
        w[i] = 1.0 / d;
    }
    return 0;
}

/* dense LU with partial pivoting, column-major a[i + n*j] */
static int lu_factor(double* a, int n, int* piv) {
    for (int k = 0; k < n; ++k) {
        int l = k;
        for (int i = k + 1; i < n; ++i) if (fabs(a[i + n * k]) > fabs(a[l + n * k])) l = i;
        piv[k] = l;
        if (a[l + n * k] == 0.0) return k + 1;
        if (l != k) for (int j = 0; j < n; ++j) { double t = a[l + n * j]; a[l + n * j] = a[k + n * j]; a[k + n * j] = t; }
        double mult = 1.0 / a[k + n * k];
        for (int i = k + 1; i < n; ++i) a[i + n * k] *= mult;
        for (int j = k + 1; j < n; ++j) {
            double akj = a[k + n * j];
            if (akj != 0.0) for (int i = k + 1; i < n; ++i) a[i + n * j] -= akj * a[i + n * k];
        }
    }
    return 0;
}

static void lu_solve(const double* a, int n, const int* piv, double* b) {
    for (int k = 0; k < n; ++k) { int pk = piv[k]; if (pk != k) { double t = b[k]; b[k] = b[pk]; b[pk] = t; } }
    for (int k = 0; k < n - 1; ++k) for (int i = k + 1; i < n; ++i) b[i] -= a[i + n * k] * b[k];
    for (int k = n - 1; k > 0; --k) { b[k] /= a[k + n * k]; for (int i = 0; i < k; ++i) b[i] -= a[i + n * k] * b[k]; }
    b[0] /= a[0];
}

/* ---- (re)initialisation ---------------------------------------------------------------------- */
static void cv_reinit(cv_mem* m, double t0, const double* y0) {
    m->tn = t0;
    for (int i = 0; i < m->N; ++i) m->zn[0][i] = y0[i];
    m->q = 1; m->L = 2; m->qwait = m->L; m->etamax = ETAMX1;
    m->qu = 0; m->hu = 0.0; m->tolsf = 1.0; m->forceSetup = 0;
    m->nst = m->nfe = m->ncfn = m->netf = m->nni = m->nsetups = m->nje = m->nstlp = m->nstlj = m->nhnil = 0;
    m->nfQe = m->netfQ = 0;
    m->h0u = 0.0; m->next_h = 0.0; m->next_q = 0;
    m->hin = 0.0; m->hmin = 0.0; m->hmax_inv = 0.0;
    m->tstopset = 0;
    m->nlscoef = NLSCOEF;
    m->saved_tq5 = 0.0;
    m->jcur = 0;
    m->crate = 1.0;
}

static void cv_quad_reinit(cv_mem* m, const double* q0) {
    for (int i = 0; i < m->NQ; ++i) m->znQ[0][i] = q0[i];
    m->nfQe = 0; m->netfQ = 0;
}

/* ---- initial step size (cvHin, cvUpperBoundH0, cvYddNorm) ------------------------------------ */
static double upper_bound_h0(cv_mem* m, double tdist) {
    double hub_inv = 0.0;
    for (int i = 0; i < m->N; ++i) {
        double d = HUB_FACTOR * fabs(m->zn[0][i]) + 1.0 / m->ewt[i];
        double r = fabs(m->zn[1][i]) / d;
        if (r > hub_inv) hub_inv = r;
    }
    if (m->quadr && m->errconQ) {
        for (int i = 0; i < m->NQ; ++i) {
            double d = HUB_FACTOR * fabs(m->znQ[0][i]) + 1.0 / m->ewtQ[i];
            double r = fabs(m->znQ[1][i]) / d;
            if (r > hub_inv) hub_inv = r;
        }
    }
    double hub = HUB_FACTOR * tdist;
    if (hub * hub_inv > 1.0) hub = 1.0 / hub_inv;
    return hub;
}

static int ydd_norm(cv_mem* m, double hg, double* yddnrm) {
    for (int i = 0; i < m->N; ++i) m->y[i] = hg * m->zn[1][i] + m->zn[0][i];
    int r = m->f(m, m->tn + hg, m->y, m->tempv); m->nfe++;
    if (r < 0) return CV_RHSFUNC_FAIL;
    if (r > 0) return RHSFUNC_RECVR;
    if (m->quadr && m->errconQ) {
        r = m->fQ(m, m->tn + hg, m->y, m->tempvQ); m->nfQe++;
        if (r < 0) return CV_RHSFUNC_FAIL;
        if (r > 0) return QRHSFUNC_RECVR;
    }
    for (int i = 0; i < m->N; ++i) m->tempv[i] = (m->tempv[i] - m->zn[1][i]) / hg;
    *yddnrm = vnorm(m, m->tempv, m->ewt);
    if (m->quadr && m->errconQ) {
        for (int i = 0; i < m->NQ; ++i) m->tempvQ[i] = (m->tempvQ[i] - m->znQ[1][i]) / hg;
        double nq = wrms(m->tempvQ, m->ewtQ, m->NQ);
        if (nq > *yddnrm) *yddnrm = nq;
    }
    return CV_SUCCESS;
}

static int cv_hin(cv_mem* m, double tout) {
    double tdiff = tout - m->tn;
    if (tdiff == 0.0) return CV_TOO_CLOSE;
    int sign = (tdiff > 0.0) ? 1 : -1;
    double tdist = fabs(tdiff);
    double tround = UROUND * fmax(fabs(m->tn), fabs(tout));
    if (tdist < 2.0 * tround) return CV_TOO_CLOSE;

    double hlb = HLB_FACTOR * tround;
    double hub = upper_bound_h0(m, tdist);
    double hg = sqrt(hlb * hub);
    if (hub < hlb) { m->h = (sign == -1) ? -hg : hg; return CV_SUCCESS; }

    int hnewOK = 0; (void)hnewOK;
    double hs = hg, hnew = hg, yddnrm = 0.0;
    for (int count1 = 1; count1 <= HIN_MAX_ITERS; ++count1) {
        int hgOK = 0;
        for (int count2 = 1; count2 <= HIN_MAX_ITERS; ++count2) {
            double hgs = hg * sign;
            int r = ydd_norm(m, hgs, &yddnrm);
            if (r < 0) return CV_RHSFUNC_FAIL;
            if (r == CV_SUCCESS) { hgOK = 1; break; }
            hg *= 0.2;
        }
        if (!hgOK) {
            if (count1 <= 2) return CV_REPTD_RHSFUNC_ERR;
            hnew = hs;
            break;
        }
        hs = hg;
        hnew = (yddnrm * hub * hub > 2.0) ? sqrt(2.0 / yddnrm) : sqrt(hg * hub);
        if (count1 == HIN_MAX_ITERS) break;
        double hrat = hnew / hg;
        if (hrat > 0.5 && hrat < 2.0) break;
        if (count1 > 1 && hrat > 2.0) { hnew = hg; break; }
        hg = hnew;
    }
    double h0 = H_BIAS * hnew;
    if (h0 < hlb) h0 = hlb;
    if (h0 > hub) h0 = hub;
    if (sign == -1) h0 = -h0;
    m->h = h0;
    return CV_SUCCESS;
}

/* ---- Nordsieck array manipulation ------------------------------------------------------------ */
static void cv_rescale(cv_mem* m) {
    double factor = m->eta;
    for (int j = 1; j <= m->q; ++j) {
        for (int i = 0; i < m->N; ++i) m->zn[j][i] *= factor;
        if (m->quadr) for (int i = 0; i < m->NQ; ++i) m->znQ[j][i] *= factor;
        factor *= m->eta;
    }
    m->h = m->hscale * m->eta;
    m->next_h = m->h;
    m->hscale = m->h;
}

static void cv_predict(cv_mem* m) {
    m->tn += m->h;
    if (m->tstopset && (m->tn - m->tstop) * m->h > 0.0) m->tn = m->tstop;
    for (int k = 1; k <= m->q; ++k)
        for (int j = m->q; j >= k; --j) {
            for (int i = 0; i < m->N; ++i) m->zn[j - 1][i] += m->zn[j][i];
            if (m->quadr) for (int i = 0; i < m->NQ; ++i) m->znQ[j - 1][i] += m->znQ[j][i];
        }
}

static void cv_restore(cv_mem* m, double saved_t) {
    m->tn = saved_t;
    for (int k = 1; k <= m->q; ++k)
        for (int j = m->q; j >= k; --j) {
            for (int i = 0; i < m->N; ++i) m->zn[j - 1][i] -= m->zn[j][i];
            if (m->quadr) for (int i = 0; i < m->NQ; ++i) m->znQ[j - 1][i] -= m->znQ[j][i];
        }
}

static void cv_increase_bdf(cv_mem* m) {
    double alpha0, alpha1, prod, xi, xiold, hsum, A1;
    for (int i = 0; i <= QMAX; ++i) m->l[i] = 0.0;
    m->l[2] = alpha1 = prod = xiold = 1.0;
    alpha0 = -1.0;
    hsum = m->hscale;
    if (m->q > 1) {
        for (int j = 1; j < m->q; ++j) {
            hsum += m->tau[j + 1];
            xi = hsum / m->hscale;
            prod *= xi;
            alpha0 -= 1.0 / (j + 1);
            alpha1 += 1.0 / xi;
            for (int i = j + 2; i >= 2; --i) m->l[i] = m->l[i] * xiold + m->l[i - 1];
            xiold = xi;
        }
    }
    A1 = (-alpha0 - alpha1) / prod;
    int L = m->L;
    for (int i = 0; i < m->N; ++i) m->zn[L][i] = A1 * m->zn[QMAX][i];
    for (int j = 2; j <= m->q; ++j) for (int i = 0; i < m->N; ++i) m->zn[j][i] += m->l[j] * m->zn[L][i];
    if (m->quadr) {
        for (int i = 0; i < m->NQ; ++i) m->znQ[L][i] = A1 * m->znQ[QMAX][i];
        for (int j = 2; j <= m->q; ++j) for (int i = 0; i < m->NQ; ++i) m->znQ[j][i] += m->l[j] * m->znQ[L][i];
    }
}

static void cv_decrease_bdf(cv_mem* m) {
    double hsum = 0.0, xi;
    for (int i = 0; i <= QMAX; ++i) m->l[i] = 0.0;
    m->l[2] = 1.0;
    for (int j = 1; j <= m->q - 2; ++j) {
        hsum += m->tau[j];
        xi = hsum / m->hscale;
        for (int i = j + 2; i >= 2; --i) m->l[i] = m->l[i] * xi + m->l[i - 1];
    }
    for (int j = 2; j < m->q; ++j) for (int i = 0; i < m->N; ++i) m->zn[j][i] -= m->l[j] * m->zn[m->q][i];
    if (m->quadr)
        for (int j = 2; j < m->q; ++j) for (int i = 0; i < m->NQ; ++i) m->znQ[j][i] -= m->l[j] * m->znQ[m->q][i];
}

static void cv_adjust_order(cv_mem* m, int deltaq) {
    if (m->q == 2 && deltaq != 1) return;
    if (deltaq == 1) cv_increase_bdf(m);
    else if (deltaq == -1) cv_decrease_bdf(m);
}

static void cv_adjust_params(cv_mem* m) {
    if (m->qprime != m->q) {
        cv_adjust_order(m, m->qprime - m->q);
        m->q = m->qprime;
        m->L = m->q + 1;
        m->qwait = m->L;
    }
    cv_rescale(m);
}

/* ---- method coefficients (cvSetBDF, cvSetTqBDF) ---------------------------------------------- */
static void cv_set(cv_mem* m) {
    double alpha0, alpha0_hat, xi_inv, xistar_inv, hsum;
    int q = m->q;
    m->l[0] = m->l[1] = xi_inv = xistar_inv = 1.0;
    for (int i = 2; i <= q; ++i) m->l[i] = 0.0;
    alpha0 = alpha0_hat = -1.0;
    hsum = m->h;
    if (q > 1) {
        for (int j = 2; j < q; ++j) {
            hsum += m->tau[j - 1];
            xi_inv = m->h / hsum;
            alpha0 -= 1.0 / j;
            for (int i = j; i >= 1; --i) m->l[i] += m->l[i - 1] * xi_inv;
        }
        alpha0 -= 1.0 / q;
        xistar_inv = -m->l[1] - alpha0;
        hsum += m->tau[q - 1];
        xi_inv = m->h / hsum;
        alpha0_hat = -m->l[1] - xi_inv;
        for (int i = q; i >= 1; --i) m->l[i] += m->l[i - 1] * xistar_inv;
    }
    /* test quantities */
    double A1 = 1.0 - alpha0_hat + alpha0;
    double A2 = 1.0 + q * A1;
    m->tq[2] = fabs(A1 / (alpha0 * A2));
    m->tq[5] = fabs(A2 * xistar_inv / (m->l[q] * xi_inv));
    if (m->qwait == 1) {
        if (q > 1) {
            double C = xistar_inv / m->l[q];
            double A3 = alpha0 + 1.0 / q;
            double A4 = alpha0_hat + xi_inv;
            double Cpinv = (1.0 - A4 + A3) / A3;
            m->tq[1] = fabs(C * Cpinv);
        } else m->tq[1] = 1.0;
        hsum += m->tau[q];
        xi_inv = m->h / hsum;
        double A5 = alpha0 - (1.0 / (q + 1));
        double A6 = alpha0_hat - xi_inv;
        double Cppinv = (1.0 - A6 + A5) / A2;
        m->tq[3] = fabs(Cppinv / (xi_inv * (q + 2) * A5));
    }
    m->tq[4] = m->nlscoef / m->tq[2];

    m->rl1 = 1.0 / m->l[1];
    m->gamma = m->h * m->rl1;
    if (m->nst == 0) m->gammap = m->gamma;
    m->gamrat = (m->nst > 0) ? m->gamma / m->gammap : 1.0;
}

/* ---- linear solver setup (cvLsSetup with dense matrix + analytic Jacobian) --------------------- */
static int ls_setup(cv_mem* m, int convfail, const double* ypred) {
    int n = m->NM;
    double dgamma = fabs((m->gamma / m->gammap) - 1.0);
    int jbad = (m->nst == 0) || (m->nst > m->nstlj + MSBJ) ||
               ((convfail == FAIL_BAD_J) && (dgamma < CVLS_DGMAX)) || (convfail == FAIL_OTHER);
    if (!jbad) {
        m->jcur = 0;
    } else {
        m->nje++; m->nstlj = m->nst; m->jcur = 1;
        int r = m->jacfn(m, m->tn, ypred, m->savedJ);
        if (r < 0) return -1;
        if (r > 0) return 1;
    }
    for (int k = 0; k < n * n; ++k) m->M[k] = -m->gamma * m->savedJ[k];
    for (int i = 0; i < n; ++i) m->M[i + n * i] += 1.0;
    if (lu_factor(m->M, n, m->piv) != 0) return 1;   /* singular: recoverable */
    return 0;
}

/* ---- nonlinear solve (cvNls + SUNNonlinSol_Newton + cvNlsConvTest) ---------------------------- */
static int cv_check_constraints(cv_mem* m);

static int cv_nls(cv_mem* m, int nflag) {
    int n = m->N;
    int callSetup;
    m->convfail = (nflag == FIRST_CALL || nflag == PREV_ERR_FAIL) ? NO_FAILURES : FAIL_OTHER;
    callSetup = (nflag == PREV_CONV_FAIL) || (nflag == PREV_ERR_FAIL) || (m->nst == 0) ||
                (m->nst >= m->nstlp + MSBP) || (fabs(m->gamrat - 1.0) > DGMAX);
    if (m->forceSetup) { callSetup = 1; m->convfail = FAIL_OTHER; }

    for (int i = 0; i < n; ++i) m->acor[i] = 0.0;
    double delta[NMAX];
    int retval;

    for (;;) {
        /* residual at the predictor */
        for (int i = 0; i < n; ++i) m->y[i] = m->zn[0][i] + m->acor[i];
        retval = m->f(m, m->tn, m->y, m->ftemp); m->nfe++;
        if (retval < 0) { retval = CV_RHSFUNC_FAIL; goto done; }
        if (retval > 0) { retval = RHSFUNC_RECVR; goto done; }   /* Sys failure before the Newton loop: no retry */
        for (int i = 0; i < n; ++i) delta[i] = m->rl1 * m->zn[1][i] + m->acor[i] - m->gamma * m->ftemp[i];

        if (callSetup) {
            retval = ls_setup(m, m->convfail, m->y);
            m->nsetups++;
            callSetup = 0;
            m->forceSetup = 0;
            m->gamrat = 1.0; m->gammap = m->gamma; m->crate = 1.0; m->nstlp = m->nst;
            if (retval < 0) { retval = CV_LSETUP_FAIL; goto done; }
            if (retval > 0) { retval = CONV_FAIL; goto done; }
        }

        int curiter = 0;
        for (;;) {
            m->nni++;
            for (int i = 0; i < n; ++i) delta[i] = -delta[i];
            for (int b = 0; b < m->nblk; ++b) lu_solve(m->M, m->NM, m->piv, delta + b * m->NM);
            if (m->gamrat != 1.0) { double s = 2.0 / (1.0 + m->gamrat); for (int i = 0; i < n; ++i) delta[i] *= s; }
            for (int i = 0; i < n; ++i) m->acor[i] += delta[i];

            /* convergence test */
            double del = vnorm(m, delta, m->ewt);
            if (curiter > 0) m->crate = fmax(CRDOWN * m->crate, del / m->delp);
            double dcon = del * fmin(1.0, m->crate) / m->tq[4];
            if (dcon <= 1.0) {
                m->acnrm = (curiter == 0) ? del : vnorm(m, m->acor, m->ewt);
                m->jcur = 0;
                for (int i = 0; i < n; ++i) m->y[i] = m->zn[0][i] + m->acor[i];
                if (m->constraints) return cv_check_constraints(m);
                return CV_SUCCESS;
            }
            if (!(dcon <= 1.0) && !(dcon > 1.0)) { retval = CONV_FAIL; break; }   /* NaN guard */
            if (curiter >= 1 && del > RDIV * m->delp) { retval = CONV_FAIL; break; }
            m->delp = del;
            curiter++;
            if (curiter >= NLS_MAXCOR) { retval = CONV_FAIL; break; }

            for (int i = 0; i < n; ++i) m->y[i] = m->zn[0][i] + m->acor[i];
            retval = m->f(m, m->tn, m->y, m->ftemp); m->nfe++;
            if (retval < 0) { retval = CV_RHSFUNC_FAIL; goto done; }
            if (retval > 0) { retval = RHSFUNC_RECVR; break; }
            for (int i = 0; i < n; ++i) delta[i] = m->rl1 * m->zn[1][i] + m->acor[i] - m->gamma * m->ftemp[i];
        }
        /* recoverable failure with stale Jacobian: retry once with a fresh one */
        if (retval > 0 && !m->jcur) {
            callSetup = 1;
            m->convfail = FAIL_BAD_J;
            for (int i = 0; i < n; ++i) m->acor[i] = 0.0;
            continue;
        }
        break;
    }
done:
    for (int i = 0; i < n; ++i) m->y[i] = m->zn[0][i] + m->acor[i];
    return retval;
}

/* N_VConstrMask for one component: does x violate its constraint flag c? */
static int constr_violated(double c, double x) {
    if (fabs(c) > 1.5) return x * c <= 0.0;
    if (fabs(c) > 0.5) return x * c < 0.0;
    return 0;
}

/* cvCheckConstraints (called at the end of cvNls after a converged solve): a violation whose
 * correction v is small (||v|| <= tq[4]) is projected away by changing acor; otherwise eta is set
 * for a smaller step and CONSTR_RECVR returned. */
static int cv_check_constraints(cv_mem* m) {
    int n = m->N, any = 0;
    double* mm = m->ftemp; double* tmp = m->tempv;
    for (int i = 0; i < n; ++i) { mm[i] = constr_violated(m->constraints[i], m->y[i]) ? 1.0 : 0.0; any |= (mm[i] != 0.0); }
    if (!any) return CV_SUCCESS;
    for (int i = 0; i < n; ++i) {
        double a = (fabs(m->constraints[i]) >= 1.5) ? 1.0 : 0.0;
        tmp[i] = mm[i] * (m->y[i] - 0.1 * (a * m->constraints[i] / m->ewt[i]));
    }
    double vnorm = wrms(tmp, m->ewt, n);
    if (vnorm <= m->tq[4]) {
        for (int i = 0; i < n; ++i) m->acor[i] -= tmp[i];
        return CV_SUCCESS;
    }
    if (fabs(m->h) <= m->hmin * ONEPSM) return CV_CONSTR_FAIL;
    double mq = DBL_MAX;
    for (int i = 0; i < n; ++i) {
        double d = mm[i] * (m->zn[0][i] - m->y[i]);
        if (d != 0.0) mq = fmin(mq, m->zn[0][i] / d);
    }
    m->eta = fmax(0.9 * mq, 0.1);
    return CONSTR_RECVR;
}

static int cv_handle_nflag(cv_mem* m, int* nflagPtr, double saved_t, int* ncfPtr, long* ncfnPtr) {
    int nflag = *nflagPtr;
    if (nflag == CV_SUCCESS) return DO_ERROR_TEST;
    (*ncfnPtr)++;
    cv_restore(m, saved_t);
    if (nflag < 0) return nflag;
    (*ncfPtr)++;
    m->etamax = 1.0;
    if (fabs(m->h) <= m->hmin * ONEPSM || *ncfPtr == MXNCF) {
        if (nflag == CONV_FAIL) return CV_CONV_FAILURE;
        if (nflag == CONSTR_RECVR) return CV_CONSTR_FAIL;
        return CV_REPTD_RHSFUNC_ERR;
    }
    /* for CONSTR_RECVR eta was already set by cv_check_constraints */
    if (nflag != CONSTR_RECVR) m->eta = fmax(ETACF, m->hmin / fabs(m->h));
    *nflagPtr = PREV_CONV_FAIL;
    cv_rescale(m);
    return PREDICT_AGAIN;
}

static int cv_do_error_test(cv_mem* m, int* nflagPtr, double saved_t, double acor_nrm,
                            int* nefPtr, long* netfPtr, double* dsmPtr) {
    double dsm = acor_nrm * m->tq[2];
    *dsmPtr = dsm;
    if (dsm <= 1.0) return CV_SUCCESS;

    (*nefPtr)++; (*netfPtr)++;
    *nflagPtr = PREV_ERR_FAIL;
    cv_restore(m, saved_t);
    if (fabs(m->h) <= m->hmin * ONEPSM || *nefPtr == MXNEF) return CV_ERR_FAILURE;
    m->etamax = 1.0;
    if (*nefPtr <= MXNEF1) {
        m->eta = 1.0 / (pow(BIAS2 * dsm, 1.0 / m->L) + ADDON);
        m->eta = fmax(ETAMIN, fmax(m->eta, m->hmin / fabs(m->h)));
        if (*nefPtr >= SMALL_NEF) m->eta = fmin(m->eta, ETAMXF);
        cv_rescale(m);
        return TRY_AGAIN;
    }
    if (m->q > 1) {
        m->eta = fmax(ETAMIN, m->hmin / fabs(m->h));
        cv_adjust_order(m, -1);
        m->L = m->q; m->q--; m->qwait = m->L;
        cv_rescale(m);
        return TRY_AGAIN;
    }
    /* order 1: reload zn from scratch */
    m->eta = fmax(ETAMIN, m->hmin / fabs(m->h));
    m->h *= m->eta; m->next_h = m->h; m->hscale = m->h; m->qwait = LONG_WAIT;
    int r = m->f(m, m->tn, m->zn[0], m->tempv); m->nfe++;
    if (r < 0) return CV_RHSFUNC_FAIL;
    if (r > 0) return CV_UNREC_RHSFUNC_ERR;
    for (int i = 0; i < m->N; ++i) m->zn[1][i] = m->h * m->tempv[i];
    if (m->quadr) {
        r = m->fQ(m, m->tn, m->zn[0], m->tempvQ); m->nfQe++;
        if (r != 0) return CV_RHSFUNC_FAIL;
        for (int i = 0; i < m->NQ; ++i) m->znQ[1][i] = m->h * m->tempvQ[i];
    }
    return TRY_AGAIN;
}

static void cv_complete_step(cv_mem* m) {
    m->nst++;
    m->hu = m->h; m->qu = m->q;
    for (int i = m->q; i >= 2; --i) m->tau[i] = m->tau[i - 1];
    if (m->q == 1 && m->nst > 1) m->tau[2] = m->tau[1];
    m->tau[1] = m->h;
    for (int j = 0; j <= m->q; ++j) {
        for (int i = 0; i < m->N; ++i) m->zn[j][i] += m->l[j] * m->acor[i];
        if (m->quadr) for (int i = 0; i < m->NQ; ++i) m->znQ[j][i] += m->l[j] * m->acorQ[i];
    }
    m->qwait--;
    if (m->qwait == 1 && m->q != QMAX) {
        for (int i = 0; i < m->N; ++i) m->zn[QMAX][i] = m->acor[i];
        if (m->quadr) for (int i = 0; i < m->NQ; ++i) m->znQ[QMAX][i] = m->acorQ[i];
        m->saved_tq5 = m->tq[5];
    }
}

static void cv_set_eta(cv_mem* m) {
    if (m->eta < THRESH) { m->eta = 1.0; m->hprime = m->h; }
    else {
        m->eta = fmin(m->eta, m->etamax);
        m->eta /= fmax(1.0, fabs(m->h) * m->hmax_inv * m->eta);
        m->hprime = m->h * m->eta;
    }
}

static void cv_prepare_next_step(cv_mem* m, double dsm) {
    if (m->etamax == 1.0) {
        m->qwait = (m->qwait > 2) ? m->qwait : 2;
        m->qprime = m->q; m->hprime = m->h; m->eta = 1.0;
        return;
    }
    m->etaq = 1.0 / (pow(BIAS2 * dsm, 1.0 / m->L) + ADDON);
    if (m->qwait != 0) { m->eta = m->etaq; m->qprime = m->q; cv_set_eta(m); return; }

    m->qwait = 2;
    /* eta at order q-1 */
    m->etaqm1 = 0.0;
    if (m->q > 1) {
        double ddn = vnorm(m, m->zn[m->q], m->ewt);
        if (m->quadr && m->errconQ) { double dq = wrms(m->znQ[m->q], m->ewtQ, m->NQ); if (dq > ddn) ddn = dq; }
        ddn *= m->tq[1];
        m->etaqm1 = 1.0 / (pow(BIAS1 * ddn, 1.0 / m->q) + ADDON);
    }
    /* eta at order q+1 */
    m->etaqp1 = 0.0;
    if (m->q != QMAX && m->saved_tq5 != 0.0) {
        double cquot = (m->tq[5] / m->saved_tq5) * pow(m->h / m->tau[2], (double)m->L);
        for (int i = 0; i < m->N; ++i) m->tempv[i] = m->acor[i] - cquot * m->zn[QMAX][i];
        double dup = vnorm(m, m->tempv, m->ewt);
        if (m->quadr && m->errconQ) {
            for (int i = 0; i < m->NQ; ++i) m->tempvQ[i] = m->acorQ[i] - cquot * m->znQ[QMAX][i];
            double dq = wrms(m->tempvQ, m->ewtQ, m->NQ); if (dq > dup) dup = dq;
        }
        dup *= m->tq[3];
        m->etaqp1 = 1.0 / (pow(BIAS3 * dup, 1.0 / (m->L + 1)) + ADDON);
    }
    /* choose */
    double etam = fmax(m->etaqm1, fmax(m->etaq, m->etaqp1));
    if (etam < THRESH) { m->eta = 1.0; m->qprime = m->q; }
    else if (etam == m->etaq) { m->eta = m->etaq; m->qprime = m->q; }
    else if (etam == m->etaqm1) { m->eta = m->etaqm1; m->qprime = m->q - 1; }
    else {
        m->eta = m->etaqp1; m->qprime = m->q + 1;
        for (int i = 0; i < m->N; ++i) m->zn[QMAX][i] = m->acor[i];
        if (m->quadr && m->errconQ) for (int i = 0; i < m->NQ; ++i) m->znQ[QMAX][i] = m->acorQ[i];
    }
    cv_set_eta(m);
}

/* ---- one internal step (cvStep) -------------------------------------------------------------- */
static int cv_step(cv_mem* m) {
    double saved_t = m->tn, dsm = 0.0, dsmQ = 0.0;
    int ncf = 0, nef = 0, nefQ = 0, nflag = FIRST_CALL, kflag, eflag;

    if (m->nst > 0 && m->hprime != m->h) cv_adjust_params(m);

    for (;;) {
        cv_predict(m);
        cv_set(m);
        nflag = cv_nls(m, nflag);
        kflag = cv_handle_nflag(m, &nflag, saved_t, &ncf, &m->ncfn);
        if (kflag == PREDICT_AGAIN) continue;
        if (kflag != DO_ERROR_TEST) return kflag;

        eflag = cv_do_error_test(m, &nflag, saved_t, m->acnrm, &nef, &m->netf, &dsm);
        if (eflag == TRY_AGAIN) continue;
        if (eflag != CV_SUCCESS) return eflag;

        if (m->quadr) {
            ncf = nef = 0;
            /* cvQuadNls */
            int r = m->fQ(m, m->tn, m->y, m->acorQ); m->nfQe++;
            if (r < 0) return CV_RHSFUNC_FAIL;
            if (r > 0) {
                nflag = QRHSFUNC_RECVR;
                kflag = cv_handle_nflag(m, &nflag, saved_t, &ncf, &m->ncfn);
                if (kflag == PREDICT_AGAIN) continue;
                return kflag;
            }
            for (int i = 0; i < m->NQ; ++i) {
                m->acorQ[i] = m->rl1 * (m->h * m->acorQ[i] - m->znQ[1][i]);
                m->yQ[i] = m->znQ[0][i] + m->acorQ[i];
            }
            if (m->errconQ) {
                double acnrmQ = wrms(m->acorQ, m->ewtQ, m->NQ);
                eflag = cv_do_error_test(m, &nflag, saved_t, acnrmQ, &nefQ, &m->netfQ, &dsmQ);
                if (eflag == TRY_AGAIN) continue;
                if (eflag != CV_SUCCESS) return eflag;
                if (dsmQ > dsm) dsm = dsmQ;
            }
        }
        break;
    }
    cv_complete_step(m);
    cv_prepare_next_step(m, dsm);
    m->etamax = (m->nst <= SMALL_NST) ? ETAMX2 : ETAMX3;
    for (int i = 0; i < m->N; ++i) m->acor[i] *= m->tq[2];
    if (m->quadr) for (int i = 0; i < m->NQ; ++i) m->acorQ[i] *= m->tq[2];
    return CV_SUCCESS;
}

/* ---- dense output (CVodeGetDky with k = 0; CVodeGetQuadDky) ----------------------------------- */
static int cv_get_dky(const cv_mem* m, double t, double* dky) {
    double tfuzz = FUZZ_FACTOR * UROUND * (fabs(m->tn) + fabs(m->hu));
    if (m->hu < 0.0) tfuzz = -tfuzz;
    double tp = m->tn - m->hu - tfuzz, tn1 = m->tn + tfuzz;
    if ((t - tp) * (t - tn1) > 0.0) return CV_BAD_T;
    double s = (t - m->tn) / m->h;
    for (int i = 0; i < m->N; ++i) {
        double acc = m->zn[m->q][i];
        for (int j = m->q - 1; j >= 0; --j) acc = acc * s + m->zn[j][i];
        dky[i] = acc;
    }
    return CV_SUCCESS;
}

static int cv_get_quad(const cv_mem* m, double t, double* dky) {
    double s = (t - m->tn) / m->h;
    for (int i = 0; i < m->NQ; ++i) {
        double acc = m->znQ[m->q][i];
        for (int j = m->q - 1; j >= 0; --j) acc = acc * s + m->znQ[j][i];
        dky[i] = acc;
    }
    return CV_SUCCESS;
}

/* ---- main driver (CVode) ---------------------------------------------------------------------- */
static int cv_solve(cv_mem* m, double tout, double* yout, double* tret, int itask) {
    int istate, ier;
    double troundoff;

    if (m->nst == 0) {
        m->tretlast = *tret = m->tn;
        /* cvInitialSetup: y0 must satisfy the constraints */
        if (m->constraints)
            for (int i = 0; i < m->N; ++i)
                if (constr_violated(m->constraints[i], m->zn[0][i])) return CV_ILL_INPUT;
        if (ewt_set(m, m->zn[0], m->ewt)) return CV_ILL_INPUT;
        if (m->quadr && m->errconQ && ewtQ_set(m, m->znQ[0], m->ewtQ)) return CV_ILL_INPUT;
        int r = m->f(m, m->tn, m->zn[0], m->zn[1]); m->nfe++;
        if (r < 0) return CV_RHSFUNC_FAIL;
        if (r > 0) return CV_FIRST_RHSFUNC_ERR;
        if (m->quadr) {
            r = m->fQ(m, m->tn, m->zn[0], m->znQ[1]); m->nfQe++;
            if (r != 0) return CV_RHSFUNC_FAIL;
        }
        if (m->tstopset && (m->tstop - m->tn) * (tout - m->tn) <= 0.0) return CV_ILL_INPUT;

        m->h = m->hin;
        if (m->h != 0.0 && (tout - m->tn) * m->h < 0.0) return CV_ILL_INPUT;
        if (m->h == 0.0) {
            double tout_hin = tout;
            if (m->tstopset && (tout - m->tn) * (tout - m->tstop) > 0.0) tout_hin = m->tstop;
            int hflag = cv_hin(m, tout_hin);
            if (hflag != CV_SUCCESS) return hflag;
        }
        double rh = fabs(m->h) * m->hmax_inv;
        if (rh > 1.0) m->h /= rh;
        if (fabs(m->h) < m->hmin) m->h *= m->hmin / fabs(m->h);
        if (m->tstopset && (m->tn + m->h - m->tstop) * m->h > 0.0)
            m->h = (m->tstop - m->tn) * (1.0 - 4.0 * UROUND);
        m->hscale = m->h; m->h0u = m->h; m->hprime = m->h;
        for (int i = 0; i < m->N; ++i) m->zn[1][i] *= m->h;
        if (m->quadr) for (int i = 0; i < m->NQ; ++i) m->znQ[1][i] *= m->h;
    }

    if (m->nst > 0) {
        troundoff = FUZZ_FACTOR * UROUND * (fabs(m->tn) + fabs(m->h));
        if (itask == CV_NORMAL && (m->tn - tout) * m->h >= 0.0) {
            m->tretlast = *tret = tout;
            ier = cv_get_dky(m, tout, yout);
            if (ier != CV_SUCCESS) return CV_ILL_INPUT;
            return CV_SUCCESS;
        }
        if (itask == CV_ONE_STEP && fabs(m->tn - m->tretlast) > troundoff) {
            m->tretlast = *tret = m->tn;
            for (int i = 0; i < m->N; ++i) yout[i] = m->zn[0][i];
            return CV_SUCCESS;
        }
        if (m->tstopset) {
            if (fabs(m->tn - m->tstop) <= troundoff) {
                ier = cv_get_dky(m, m->tstop, yout);
                if (ier != CV_SUCCESS) return CV_ILL_INPUT;
                m->tretlast = *tret = m->tstop;
                m->tstopset = 0;
                return CV_TSTOP_RETURN;
            }
            if ((m->tn + m->hprime - m->tstop) * m->h > 0.0) {
                m->hprime = (m->tstop - m->tn) * (1.0 - 4.0 * UROUND);
                m->eta = m->hprime / m->h;
            }
        }
    }

    long nstloc = 0;
    for (;;) {
        m->next_h = m->h; m->next_q = m->q;
        if (m->nst > 0) {
            if (ewt_set(m, m->zn[0], m->ewt)) { istate = CV_ILL_INPUT; m->tretlast = *tret = m->tn; for (int i = 0; i < m->N; ++i) yout[i] = m->zn[0][i]; break; }
            if (m->quadr && m->errconQ && ewtQ_set(m, m->znQ[0], m->ewtQ)) { istate = CV_ILL_INPUT; break; }
        }
        if (m->mxstep > 0 && nstloc >= m->mxstep) {
            istate = CV_TOO_MUCH_WORK;
            m->tretlast = *tret = m->tn;
            for (int i = 0; i < m->N; ++i) yout[i] = m->zn[0][i];
            break;
        }
        double nrm = vnorm(m, m->zn[0], m->ewt);
        if (m->quadr && m->errconQ) { double nq = wrms(m->znQ[0], m->ewtQ, m->NQ); if (nq > nrm) nrm = nq; }
        m->tolsf = UROUND * nrm;
        if (m->tolsf > 1.0) {
            istate = CV_TOO_MUCH_ACC;
            m->tretlast = *tret = m->tn;
            for (int i = 0; i < m->N; ++i) yout[i] = m->zn[0][i];
            m->tolsf *= 2.0;
            break;
        } else m->tolsf = 1.0;
        if (m->tn + m->h == m->tn) m->nhnil++;

        int kflag = cv_step(m);
        if (kflag != CV_SUCCESS) {
            istate = kflag;
            m->tretlast = *tret = m->tn;
            for (int i = 0; i < m->N; ++i) yout[i] = m->zn[0][i];
            break;
        }
        nstloc++;

        if (m->tstopset) {
            troundoff = FUZZ_FACTOR * UROUND * (fabs(m->tn) + fabs(m->h));
            if (fabs(m->tn - m->tstop) <= troundoff) m->tn = m->tstop;
        }
        if (itask == CV_NORMAL && (m->tn - tout) * m->h >= 0.0) {
            istate = CV_SUCCESS;
            m->tretlast = *tret = tout;
            (void)cv_get_dky(m, tout, yout);
            m->next_q = m->qprime; m->next_h = m->hprime;
            break;
        }
        if (m->tstopset) {
            troundoff = FUZZ_FACTOR * UROUND * (fabs(m->tn) + fabs(m->h));
            if (fabs(m->tn - m->tstop) <= troundoff) {
                (void)cv_get_dky(m, m->tstop, yout);
                m->tretlast = *tret = m->tstop;
                m->tstopset = 0;
                istate = CV_TSTOP_RETURN;
                break;
            }
            if ((m->tn + m->hprime - m->tstop) * m->h > 0.0) {
                m->hprime = (m->tstop - m->tn) * (1.0 - 4.0 * UROUND);
                m->eta = m->hprime / m->h;
            }
        }
        if (itask == CV_ONE_STEP) {
            istate = CV_SUCCESS;
            m->tretlast = *tret = m->tn;
            for (int i = 0; i < m->N; ++i) yout[i] = m->zn[0][i];
            m->next_q = m->qprime; m->next_h = m->hprime;
            break;
        }
    }
    return istate;
}

/* ================================================================================================
 * Problem adapters
 * ============================================================================================== */
static int fwd_rhs(cv_mem* m, double t, const double* y, double* ydot) { return m->prob->rhs(t, y, m->p, ydot); }
static int fwd_jac(cv_mem* m, double t, const double* y, double* J) { return m->prob->jac(t, y, m->p, J); }

/* ---- adjoint data store: CVApolynomialStorePnt / CVAfindIndex / CVApolynomialGetY -------------- */
static int hist_store(hist_t* H, long idx, double t, const double* y, int order, const double* yd) {
    /* CVodeF writes the point of step nst at dt_mem[nst] (cvodea.c: "Load next point in dt_mem");
     * a CVode(CV_ONE_STEP) call that only reports an already-taken step rewrites the same slot. */
    if (idx >= H->cap) {
        int ncap = H->cap ? 2 * H->cap : 256;
        while (ncap <= idx) ncap *= 2;
        H->t = (double*)realloc(H->t, sizeof(double) * ncap);
        H->y = (double*)realloc(H->y, sizeof(double) * ncap * H->ns);
        H->order = (int*)realloc(H->order, sizeof(int) * ncap);
        if (!H->t || !H->y || !H->order) return -1;
        if (H->hermite) {
            H->yd = (double*)realloc(H->yd, sizeof(double) * ncap * H->ns);
            if (!H->yd) return -1;
        }
        H->cap = ncap;
    }
    H->t[idx] = t;
    memcpy(H->y + (size_t)idx * H->ns, y, sizeof(double) * H->ns);
    if (H->hermite) memcpy(H->yd + (size_t)idx * H->ns, yd, sizeof(double) * H->ns);
    H->order[idx] = order;
    if (idx + 1 > H->np) H->np = (int)idx + 1;
    return 0;
}

static int hist_get_y(hist_t* H, double t, double* y) {
    /* forward integration direction is +t or -t: sign as in CVAfindIndex */
    int ns = H->ns;
    if (H->np < 2) { memcpy(y, H->y, sizeof(double) * ns); return 0; }
    int sign = (H->t[H->np - 1] - H->t[0] > 0.0) ? 1 : -1;
    int newpoint = 0, indx;
    if (H->newdata) { H->ilast = H->np - 1; newpoint = 1; H->newdata = 0; }
    int to_left = sign * (t - H->t[H->ilast - 1]) < 0.0;
    int to_right = sign * (t - H->t[H->ilast]) > 0.0;
    if (to_left) {
        newpoint = 1;
        indx = H->ilast;
        for (;;) {
            if (indx == 0) break;
            if (sign * (t - H->t[indx - 1]) <= 0.0) indx--;
            else break;
        }
        H->ilast = (indx == 0) ? 1 : indx;
        if (indx == 0 && fabs(t - H->t[0]) > FUZZ_FACTOR * UROUND) return CV_GETY_BADT;
    } else if (to_right) {
        newpoint = 1;
        indx = H->ilast;
        for (;;) {
            if (indx >= H->np - 1) break;     /* guard: do not run past the stored data */
            if (sign * (t - H->t[indx]) > 0.0) indx++;
            else break;
        }
        H->ilast = indx;
    } else indx = H->ilast;

    if (indx == 0) { memcpy(y, H->y, sizeof(double) * ns); return 0; }

    if (H->hermite) {
        /* CVAhermiteGetY: cubic through (y, y') at the two ends of the stored step */
        double t0 = H->t[indx - 1], t1 = H->t[indx], delta = t1 - t0;
        const double* y0 = H->y + (size_t)(indx - 1) * ns;
        const double* yd0 = H->yd + (size_t)(indx - 1) * ns;
        if (newpoint) {
            const double* y1 = H->y + (size_t)indx * ns;
            const double* yd1 = H->yd + (size_t)indx * ns;
            for (int k = 0; k < ns; ++k) {
                H->Y1h[k] = -2.0 * y1[k] + 2.0 * y0[k] + delta * yd1[k] + delta * yd0[k];
                H->Y0h[k] = y1[k] - y0[k] - delta * yd0[k];
            }
        }
        double factor1 = t - t0;
        double factor2 = factor1 / delta;
        factor2 = factor2 * factor2;
        double factor3 = factor2 * (t - t1) / delta;
        for (int k = 0; k < ns; ++k)
            y[k] = y0[k] + factor1 * yd0[k] + factor2 * H->Y0h[k] + factor3 * H->Y1h[k];
        return 0;
    }

    double delt = fabs(H->t[indx] - H->t[indx - 1]);
    int base, order;
    if (sign == 1) {
        base = indx;
        order = H->order[base];
        if (indx < order) base += order - indx;
    } else {
        base = indx - 1;
        order = H->order[base];
        if (H->np - indx > order) base -= indx + order - H->np;
    }
    if (base > H->np - 1) { order -= base - (H->np - 1); base = H->np - 1; if (order < 0) order = 0; }

    if (newpoint) {
        for (int j = 0; j <= order; ++j) {
            int src = (sign == 1) ? base - j : base - 1 + j;
            H->T[j] = H->t[src];
            memcpy(H->Y[j], H->y + (size_t)src * ns, sizeof(double) * ns);
        }
        for (int i = 1; i <= order; ++i)
            for (int j = order; j >= i; --j) {
                double factor = delt / (H->T[j] - H->T[j - i]);
                for (int k = 0; k < ns; ++k) H->Y[j][k] = factor * (H->Y[j][k] - H->Y[j - 1][k]);
            }
        H->ord_cached = order;
        H->delt = delt;
    }
    order = H->ord_cached;
    double cvals[LMAX];
    cvals[0] = 1.0;
    for (int i = 0; i < order; ++i) cvals[i + 1] = cvals[i] * (t - H->T[i]) / H->delt;
    for (int k = 0; k < ns; ++k) {
        double acc = 0.0;
        for (int i = 0; i <= order; ++i) acc += cvals[i] * H->Y[i][k];
        y[k] = acc;
    }
    return 0;
}

static int bwd_rhs(cv_mem* m, double t, const double* lam, double* out) {
    double y[NMAX];
    if (hist_get_y(m->hist, t, y)) return -1;
    return m->prob->adj_rhs(t, y, lam, m->p, out);
}
static int bwd_jac(cv_mem* m, double t, const double* lam, double* J) {
    double y[NMAX]; (void)lam;
    if (hist_get_y(m->hist, t, y)) return -1;
    return m->prob->adj_jac(t, y, m->p, J);
}
static int bwd_quad(cv_mem* m, double t, const double* lam, double* out) {
    double y[NMAX];
    if (hist_get_y(m->hist, t, y)) return -1;
    return m->prob->quad_rhs(t, y, lam, m->p, out);
}

/* ================================================================================================
 * Exported API: the reference's solve loops, one instance at a time, and batch drivers
 * ============================================================================================== */
typedef struct {
    double rtol;
    const double* atol; int n_atol;       /* scalar (n_atol == 1) or per state */
    double rtol_b, atol_b, rtol_q, atol_q;
    int mxstep, max_retries, mxstep_b, max_retries_b;
    int hermite;                          /* AdjointSolver(interpolation='hermite') */
    const double* constraints;            /* [ns] CVodeSetConstraints flags of the forward ODE, or NULL */
    const double* pbar;                   /* [nd] CVodeSetSensParams scaling factors, or NULL (= 1) */
} oracle_options;

static void set_fwd_tols(cv_mem* m, const oracle_options* o) {
    m->reltol = o->rtol;
    for (int i = 0; i < m->N; ++i) {
        m->abstol[i] = (o->n_atol == 1) ? o->atol[0] : o->atol[i % m->NM];
        /* cvSensEwtSetEE: ewtS = pbar / (rtol |pbar yS| + atol), i.e. atolS = atol / |pbar| */
        if (o->pbar && i >= m->NM) m->abstol[i] /= fabs(o->pbar[i / m->NM - 1]);
    }
}

/* stats layout: nst, nfe, nje, nsetups, netf, ncfn, nni, (bwd) nst, nfe, nje, nsetups, netf(+Q), ncfn, nni */
#define NSTATS 16

/* Solver.solve (solver.py:467-527): CVode(CV_NORMAL) per tval. Returns status (0 ok). */
int oracle_solve_forward(const oracle_problem* prob, const oracle_options* opt,
                         double t0, const double* tvals, int n_t, const double* y0,
                         const double* p, double* y_out, long* stats)
{
    cv_mem m; memset(&m, 0, sizeof(m));
    m.N = prob->ns; m.NM = prob->ns; m.nblk = 1; m.NQ = 0; m.f = fwd_rhs; m.jacfn = fwd_jac; m.prob = prob; m.p = p;
    set_fwd_tols(&m, opt);
    m.mxstep = opt->mxstep;
    m.constraints = opt->constraints;
    cv_reinit(&m, t0, y0);
    int status = 0;
    double tret = t0, ybuf[NMAX];
    for (int i = 0; i < n_t; ++i) {
        if (tvals[i] == t0) { memcpy(y_out, y0, sizeof(double) * m.N); continue; }  /* row 0, solver.py:505 */
        int ok = 0;
        for (int retry = 0; retry < opt->max_retries; ++retry) {
            int r = cv_solve(&m, tvals[i], ybuf, &tret, CV_NORMAL);
            if (r == 0) { ok = 1; break; }
            if (r != CV_TOO_MUCH_WORK) { status = r; break; }
        }
        if (status) break;
        if (!ok) { status = CV_TOO_MUCH_WORK; break; }
        memcpy(y_out + (size_t)i * m.N, ybuf, sizeof(double) * m.N);
    }
    if (stats) { stats[0] = m.nst; stats[1] = m.nfe; stats[2] = m.nje; stats[3] = m.nsetups; stats[4] = m.netf; stats[5] = m.ncfn; stats[6] = m.nni; }
    return status;
}

/* AdjointSolver.solve_forward (solver.py:682-721): CVodeF per tval, storing (t, y, order). */
static int adjoint_forward(const oracle_problem* prob, const oracle_options* opt, double t0,
                           const double* tvals, int n_t, const double* y0, const double* p,
                           double* y_out, hist_t* H, long* stats)
{
    cv_mem m; memset(&m, 0, sizeof(m));
    m.N = prob->ns; m.NM = prob->ns; m.nblk = 1; m.NQ = 0; m.f = fwd_rhs; m.jacfn = fwd_jac; m.prob = prob; m.p = p;
    set_fwd_tols(&m, opt);
    m.mxstep = opt->mxstep;
    m.constraints = opt->constraints;
    cv_reinit(&m, t0, y0);
    H->np = 0; H->ns = prob->ns; H->newdata = 1; H->hermite = opt->hermite;
    int first = 1, status = 0;
    double tret = t0, ybuf[NMAX], ydbuf[NMAX];
    for (int i = 0; i < n_t && !status; ++i) {
        double tout = tvals[i];
        if (tout == t0) { memcpy(y_out, y0, sizeof(double) * m.N); continue; }
        /* CVodeF */
        if (first) {
            /* CVAhermiteStorePnt: y' of the initial point is f(t0, y0), later ones zn[1] / h */
            if (H->hermite && prob->rhs(m.tn, m.zn[0], p, ydbuf)) return CV_FIRST_RHSFUNC_ERR;
            if (hist_store(H, 0, m.tn, m.zn[0], m.qu, ydbuf)) return -20;
            first = 0;
        } else if ((m.tn - tout) * m.h >= 0.0) {
            if (cv_get_dky(&m, tout, ybuf)) { status = CV_BAD_T; break; }
            memcpy(y_out + (size_t)i * m.N, ybuf, sizeof(double) * m.N);
            continue;
        }
        for (;;) {
            int r = cv_solve(&m, tout, ybuf, &tret, CV_ONE_STEP);
            if (r < 0) { status = r; break; }
            if (H->hermite) for (int j = 0; j < m.N; ++j) ydbuf[j] = m.zn[1][j] / m.h;
            if (hist_store(H, m.nst, m.tn, m.zn[0], m.qu, ydbuf)) return -20;
            if ((tret - tout) * m.h >= 0.0) {
                (void)cv_get_dky(&m, tout, ybuf);
                m.tretlast = tout;
                break;
            }
        }
        if (status) break;
        memcpy(y_out + (size_t)i * m.N, ybuf, sizeof(double) * m.N);
    }
    H->newdata = 1;
    if (stats) { stats[0] = m.nst; stats[1] = m.nfe; stats[2] = m.nje; stats[3] = m.nsetups; stats[4] = m.netf; stats[5] = m.ncfn; stats[6] = m.nni; }
    return status;
}

/* AdjointSolver.solve_backward (solver.py:723-784).  tB0 = the reference's `t0` argument (last
 * time), tend = initial time; grads[n_t][ns]. */
static int adjoint_backward(const oracle_problem* prob, const oracle_options* opt, double tB0,
                            double tend, const double* tvals, int n_t, const double* grads,
                            const double* p, hist_t* H, double* grad_out, double* lamda_out,
                            long* stats)
{
    int ns = prob->ns, nd = prob->nd;
    cv_mem m; memset(&m, 0, sizeof(m));
    m.N = ns; m.NM = ns; m.nblk = 1; m.NQ = nd; m.f = bwd_rhs; m.jacfn = bwd_jac; m.fQ = bwd_quad; m.prob = prob; m.p = p;
    m.hist = H;
    m.reltol = opt->rtol_b; for (int i = 0; i < ns; ++i) m.abstol[i] = opt->atol_b;
    m.reltolQ = opt->rtol_q; m.abstolQ = opt->atol_q;
    m.quadr = (nd > 0); m.errconQ = (nd > 0);
    m.mxstep = opt->mxstep_b;

    double lam[NMAX], quad[NMAX], quad_out[NMAX];
    for (int i = 0; i < ns; ++i) lam[i] = 0.0;
    for (int i = 0; i < nd; ++i) quad[i] = quad_out[i] = 0.0;
    long tot[7] = {0, 0, 0, 0, 0, 0, 0};
    int status = 0;

    /* ts = [tB0] + reversed(tvals) + [tend]; intervals (ts[k+1], ts[k]); cotangents g[n_t-1], ..., g[0], None */
    int n_int = n_t + 1;
    for (int k = 0; k < n_int && !status; ++k) {
        double t_upper = (k == 0) ? tB0 : tvals[n_t - k];
        double t_lower = (k == n_t) ? tend : tvals[n_t - 1 - k];
        const double* g = (k < n_t) ? grads + (size_t)(n_t - 1 - k) * ns : NULL;
        if (t_lower < t_upper) {
            cv_reinit(&m, t_upper, lam);            /* CVodeReInitB */
            if (nd) cv_quad_reinit(&m, quad);       /* CVodeQuadReInitB */
            double tret = t_upper;
            int ok = 0;
            for (int retry = 0; retry < opt->max_retries_b; ++retry) {
                /* CVodeB sets the stop time of the backward integrator to the start of the current
                 * checkpoint interval (ck_t0), which with one data segment (solver.py:588) is the
                 * forward problem's initial time -- not tBout: the integrator steps past t_lower
                 * and CV_NORMAL interpolates back. */
                m.tstop = tend; m.tstopset = 1;
                int r = cv_solve(&m, t_lower, lam, &tret, CV_NORMAL);
                if (r >= 0) { ok = 1; break; }
                if (r != CV_TOO_MUCH_WORK) { status = r; break; }
            }
            tot[0] += m.nst; tot[1] += m.nfe; tot[2] += m.nje; tot[3] += m.nsetups;
            tot[4] += m.netf + m.netfQ; tot[5] += m.ncfn; tot[6] += m.nni;
            if (status) break;
            if (!ok) { status = CV_TOO_MUCH_WORK; break; }
            if (nd) { cv_get_quad(&m, tret, quad_out); memcpy(quad, quad_out, sizeof(double) * nd); }
        }
        if (g) for (int i = 0; i < ns; ++i) lam[i] -= g[i];
    }
    memcpy(grad_out, quad_out, sizeof(double) * nd);
    memcpy(lamda_out, lam, sizeof(double) * ns);
    if (stats) for (int i = 0; i < 7; ++i) stats[7 + i] = tot[i];
    return status;
}

static void hist_free(hist_t* H) { free(H->t); free(H->y); free(H->yd); free(H->order); }

/* ---- forward sensitivities: Solver(sens_mode=...).solve (solver.py:467-527 with 483-488, 523-527)
 * y_out[n_t][ns], sens_out[n_t][nd][ns]; sens0[nd][ns]. */
static int fsa_rhs(cv_mem* m, double t, const double* y, double* ydot) {
    int ns = m->NM;
    int r = m->prob->rhs(t, y, m->p, ydot);
    if (r) return r;
    return m->prob->sens_rhs(t, y, y + ns, m->p, ydot + ns);
}

int oracle_solve_forward_sens(const oracle_problem* prob, const oracle_options* opt,
                              double t0, const double* tvals, int n_t, const double* y0,
                              const double* p, const double* sens0, double* y_out,
                              double* sens_out, long* stats)
{
    int ns = prob->ns, nd = prob->nd, nt = ns * (1 + nd);
    if (nt > NMAX || !prob->sens_rhs) return CV_ILL_INPUT;
    cv_mem m; memset(&m, 0, sizeof(m));
    m.N = nt; m.NM = ns; m.nblk = 1 + nd; m.NQ = 0; m.f = fsa_rhs; m.jacfn = fwd_jac; m.prob = prob; m.p = p;
    set_fwd_tols(&m, opt);
    m.mxstep = opt->mxstep;
    double yy[NMAX];
    memcpy(yy, y0, sizeof(double) * ns);
    memcpy(yy + ns, sens0, sizeof(double) * ns * nd);
    cv_reinit(&m, t0, yy);
    int status = 0;
    double tret = t0, ybuf[NMAX];
    for (int i = 0; i < n_t; ++i) {
        if (tvals[i] == t0) {               /* row 0, solver.py:505-508 */
            memcpy(y_out, y0, sizeof(double) * ns);
            memcpy(sens_out, sens0, sizeof(double) * ns * nd);
            continue;
        }
        int ok = 0;
        for (int retry = 0; retry < opt->max_retries; ++retry) {
            int r = cv_solve(&m, tvals[i], ybuf, &tret, CV_NORMAL);
            if (r == 0) { ok = 1; break; }
            if (r != CV_TOO_MUCH_WORK) { status = r; break; }
        }
        if (status) break;
        if (!ok) { status = CV_TOO_MUCH_WORK; break; }
        memcpy(y_out + (size_t)i * ns, ybuf, sizeof(double) * ns);
        memcpy(sens_out + (size_t)i * ns * nd, ybuf + ns, sizeof(double) * ns * nd);
    }
    if (stats) { stats[0] = m.nst; stats[1] = m.nfe; stats[2] = m.nje; stats[3] = m.nsetups; stats[4] = m.netf; stats[5] = m.ncfn; stats[6] = m.nni; }
    return status;
}

int oracle_solve_forward_sens_batch(const oracle_problem* prob, const oracle_options* opt, long B,
                                    double t0, const double* tvals, int n_t, const double* y0,
                                    const double* p, const double* sens0, int sens0_shared,
                                    double* y_out, double* sens_out, int* status, long* stats,
                                    int n_threads)
{
    int ns = prob->ns, np_ = prob->np, nd = prob->nd;
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
    #pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < B; ++b) {
        long st[16] = {0};
        const double* s0 = sens0_shared ? sens0 : sens0 + (size_t)b * nd * ns;
        int r = oracle_solve_forward_sens(prob, opt, t0, tvals, n_t, y0 + b * ns, p + b * np_, s0,
                                          y_out + (size_t)b * n_t * ns,
                                          sens_out + (size_t)b * n_t * nd * ns, st);
        if (r) {
            for (size_t k = 0; k < (size_t)n_t * ns; ++k) y_out[(size_t)b * n_t * ns + k] = NAN;
            for (size_t k = 0; k < (size_t)n_t * nd * ns; ++k) sens_out[(size_t)b * n_t * nd * ns + k] = NAN;
        }
        if (status) status[b] = r;
        if (stats) memcpy(stats + b * 16, st, sizeof(st));
    }
    return 0;
}

/* One forward + one backward solve (the notebook's unit of work, from_sympy.ipynb:178-179). */
int oracle_solve_adjoint(const oracle_problem* prob, const oracle_options* opt, double t0,
                         const double* tvals, int n_t, const double* y0, const double* p,
                         const double* grads, double* y_out, double* grad_out, double* lamda_out,
                         long* stats, long* n_hist)
{
    hist_t* H = (hist_t*)calloc(1, sizeof(hist_t));
    int status = adjoint_forward(prob, opt, t0, tvals, n_t, y0, p, y_out, H, stats);
    if (n_hist) *n_hist = H->np;
    if (status == 0)
        status = adjoint_backward(prob, opt, tvals[n_t - 1], t0, tvals, n_t, grads, p, H,
                                  grad_out, lamda_out, stats);
    hist_free(H); free(H);
    return status;
}

/* history export for tests: runs the adjoint forward pass and copies out (t, order, y) */
int oracle_forward_history(const oracle_problem* prob, const oracle_options* opt, double t0,
                           const double* tvals, int n_t, const double* y0, const double* p,
                           double* y_out, int cap, double* ht, int* horder, double* hy, int* np_out)
{
    hist_t* H = (hist_t*)calloc(1, sizeof(hist_t));
    int status = adjoint_forward(prob, opt, t0, tvals, n_t, y0, p, y_out, H, NULL);
    int n = H->np < cap ? H->np : cap;
    for (int i = 0; i < n; ++i) {
        ht[i] = H->t[i]; horder[i] = H->order[i];
        memcpy(hy + (size_t)i * prob->ns, H->y + (size_t)i * prob->ns, sizeof(double) * prob->ns);
    }
    *np_out = H->np;
    hist_free(H); free(H);
    return status;
}

/* ---- batch drivers (instances independent; OpenMP over instances) ------------------------------
 * y0[B][ns], p[B][np], grads[B][n_t][ns] (or shared if grads_shared), outputs [B][...]; failed
 * instances are NaN-filled like the reference's Ops do (wrappers/as_pytensor.py:289-290,339-341). */
static void nan_fill(double* a, size_t n) { for (size_t i = 0; i < n; ++i) a[i] = NAN; }

int oracle_solve_forward_batch(const oracle_problem* prob, const oracle_options* opt, long B,
                               double t0, const double* tvals, int n_t, const double* y0,
                               const double* p, double* y_out, int* status, long* stats,
                               int n_threads)
{
    int ns = prob->ns, np_ = prob->np;
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
    #pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < B; ++b) {
        long st[NSTATS] = {0};
        int r = oracle_solve_forward(prob, opt, t0, tvals, n_t, y0 + b * ns, p + b * np_,
                                     y_out + (size_t)b * n_t * ns, st);
        if (r) nan_fill(y_out + (size_t)b * n_t * ns, (size_t)n_t * ns);
        if (status) status[b] = r;
        if (stats) memcpy(stats + b * NSTATS, st, sizeof(st));
    }
    return 0;
}

int oracle_solve_adjoint_batch(const oracle_problem* prob, const oracle_options* opt, long B,
                               double t0, const double* tvals, int n_t, const double* y0,
                               const double* p, const double* grads, int grads_shared,
                               double* y_out, double* grad_out, double* lamda_out, int* status,
                               long* stats, int n_threads)
{
    int ns = prob->ns, np_ = prob->np, nd = prob->nd;
#ifdef _OPENMP
    omp_set_num_threads(n_threads > 0 ? n_threads : omp_get_num_procs());
#endif
    #pragma omp parallel for schedule(dynamic, 16)
    for (long b = 0; b < B; ++b) {
        long st[NSTATS] = {0}; long nh = 0;
        const double* g = grads_shared ? grads : grads + (size_t)b * n_t * ns;
        int r = oracle_solve_adjoint(prob, opt, t0, tvals, n_t, y0 + b * ns, p + b * np_, g,
                                     y_out + (size_t)b * n_t * ns, grad_out + b * nd,
                                     lamda_out + b * ns, st, &nh);
        st[14] = nh;
        if (r) {
            nan_fill(y_out + (size_t)b * n_t * ns, (size_t)n_t * ns);
            nan_fill(grad_out + b * nd, nd); nan_fill(lamda_out + b * ns, ns);
        }
        if (status) status[b] = r;
        if (stats) memcpy(stats + b * NSTATS, st, sizeof(st));
    }
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
