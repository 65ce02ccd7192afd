// sb_args.h -- kernel argument blocks shared by the host launcher (sb_api.cpp) and the device
// code (sb_kernels.cuh, compiled by NVRTC).  Plain C structs; 8-byte members first so host and
// device agree on the layout without packing pragmas.
#pragma once

#define SB_STATS_STRIDE 8     /* ints per instance: nst nfe nje nsetups netf ncfn nni aux */
#define SB_CARRY_INTS 10      /* status, table index, 7 counters, pad */

/* history point: (t, order, y[NS]) (+ y'[NS] with SB_HERMITE) -> NS + 2 (2 NS + 2) doubles
 * interpolation table entry for the interval (t_lo, t_hi):
 *   [0] t_lo [1] t_hi [2] order [3] 1/delt [4..9] T[0..5] [10 + NS*j + k] Y[j][k]  -> 10 + 6*NS */
/* Lanes of a warp that integrate one instance together (sb_group.cuh): one for small systems, from
 * SB_GROUP_MIN_NS states on a power-of-two group whose lanes own ceil(ns / lanes) state components
 * (and matrix rows) each -- two per lane by default (measured on the 8-state SEIR adjoint, 32 768
 * draws: 8 lanes x 1 component 145 ms, 4 x 2 139 ms, 2 x 4 151 ms); beyond 64 states one lane per
 * instance again (loop-based LU in local memory). */
#ifndef SB_GROUP_MIN_NS
#define SB_GROUP_MIN_NS 5
#endif
#ifdef SB_GROUP_LANES   /* experiment: a fixed group size, ceil(ns / lanes) components per lane */
#define SB_GROUP_SIZE(ns) ((ns) < SB_GROUP_MIN_NS ? 1 : SB_GROUP_LANES)
#else
#define SB_GROUP_SIZE(ns) ((ns) < SB_GROUP_MIN_NS ? 1 : (ns) <= 4 ? 2 : (ns) <= 8 ? 4 : (ns) <= 16 ? 8 : (ns) <= 32 ? 16 : (ns) <= 64 ? 32 : 1)
#endif

/* -DSB_HERMITE (AdjointSolver(interpolation='hermite'), CV_HERMITE of
 * /root/reference/sunode/solver.py:581-586): every history point also carries y' = zn[1] / h, and
 * the table entry of an interval is the cubic through (y, y') at its two ends, written in the same
 * Newton form (order 3, nodes t_hi, t_hi, t_lo) -- the backward kernels do not change. */
#ifdef SB_HERMITE
#define SB_HIST_STRIDE(ns) (2 * (ns) + 2)
#else
#define SB_HIST_STRIDE(ns) ((ns) + 2)
#endif
#define SB_TAB_STRIDE(ns) (10 + 6 * (ns))

typedef struct SbForwardArgs {
    double t0;
    double rtol;
    const double* tvals;      /* [n_t] */
    const double* y0;         /* [B][NS] */
    const double* params;     /* [B][NP] */
    const double* atol;       /* [NS * (1 + ND)]: states, then the sensitivity blocks (atol / |pbar_k|) */
    double* y_out;            /* [B][n_t][NS] */
    double* hist;             /* [B][hist_cap][NS+2] or NULL */
    int* hist_n;              /* [B] number of stored points */
    int* status;              /* [B] */
    int* stats;               /* [B][SB_STATS_STRIDE] or NULL */
    long long B;
    int n_t;
    int hist_cap;
    int max_steps;            /* internal steps allowed per output time */
    int sens0_shared;         /* sens0 is [ND][NS] for all instances instead of [B][ND][NS] */
    int lanes;                /* lanes of a warp that carry an instance (32) */
    int pad_;
    /* forward sensitivities (sb_forward_sens only; NULL otherwise) */
    const double* sens0;      /* [B][ND][NS] or [ND][NS] */
    double* sens_out;         /* [B][n_t][ND][NS] */
    /* when set (with hist): the interpolation table entry of every stored step is built by the
     * forward kernel itself, right after the step, and the separate sb_tables launch is skipped */
    double* tab;              /* [B][hist_cap][10 + 6*NS] or NULL */
    /* when set: the accepted steps of all successful instances are added up here; the backward
     * kernels choose their interval schedule from it (see SbBackwardArgs.steps_total) */
    unsigned long long* steps_total;
    /* when set: -1, or for a failed instance the index of the output time it was integrating to
     * (the `time=` of the reference's error message, solver.py:516-519) */
    int* fail_k;              /* [B] or NULL */
} SbForwardArgs;

typedef struct SbTablesArgs {
    const double* hist;
    const int* hist_n;
    double* tab;              /* [B][hist_cap][10 + 6*NS] */
    long long B;
    int hist_cap;
    int pad_;
} SbTablesArgs;

typedef struct SbBackwardArgs {
    double rtol, atol, rtol_q, atol_q;
    double t_start;           /* the reference's `t0` argument of solve_backward: the LAST time */
    double t_end;             /* the reference's `tend`: the initial time */
    const double* tvals;      /* [n_t] */
    const double* params;     /* [B][NP] */
    const double* grads;      /* [B][n_t][NS], or [n_t][NS] when grads_shared */
    const double* tab;
    const int* hist_n;
    const int* fwd_status;    /* [B] or NULL */
    double* grad_out;         /* [B][ND] */
    double* lamda_out;        /* [B][NS] */
    int* status;              /* [B] */
    int* stats;               /* [B][SB_STATS_STRIDE] or NULL */
    long long B;
    int n_t;
    int hist_cap;
    int max_steps;            /* internal steps allowed per interval */
    int grads_shared;
    /* optional traces (solver.py:778-781): lamda / quadrature right after the jump at each output
     * time; row (n_t - k) % n_t for the k-th jump, as the reference's `lamda_all_out[-i]` indexing */
    double* lamda_all;        /* [B][n_t][NS] or NULL */
    double* quad_all;         /* [B][n_t][ND] or NULL */
    /* segmented execution (n_seg > 1): the n_t + 1 intervals of a solve are cut into n_seg
     * segments of seg_len intervals; a work unit = one segment of one group of 32 instances,
     * handed out by a global counter to persistent warps (see sb_backward) */
    int* queue;               /* [1] next unit, zeroed by the launcher */
    int* seg_done;            /* [n_groups] segments completed per group, zeroed by the launcher */
    double* carry_d;          /* [B][NS + max(ND,1)] lamda, quadrature between segments */
    int* carry_i;             /* [B][SB_CARRY_INTS] status, table position, counters */
    int n_seg;
    int seg_len;
    int n_groups;             /* ceil(B / lanes) */
    int lanes;                /* lanes of a warp that carry an instance (32; fewer only as an
                               * experiment, see lanes_per_warp in sb_api.cpp) */
    int flat;                 /* sb_backward_flat: when a lane starts its next interval: 0 with the
                               * other lanes of the warp, 1 at once, -1 the warp decides */
    int pad2_;
    /* Two builds of the backward kernel are launched back to back, sb_backward (lanes of a warp walk
     * the intervals together) and sb_backward_flat (every lane on its own); the one whose schedule
     * does not suit the stored forward solve returns at once: flat runs iff
     * *steps_total > flat_steps (forward steps of the whole batch; NULL: the plain kernel runs). */
    const unsigned long long* steps_total;
    unsigned long long flat_steps;
} SbBackwardArgs;

typedef struct SbEvalArgs {
    const double* t;          /* [n] */
    const double* y;          /* [n][NS] */
    const double* params;     /* [n][NP] or [NP] when params_shared */
    const double* lam;        /* [n][NS] or NULL */
    double* out;              /* kind 0: [n][NS] rhs; 1: [n][NS*NS] jac; 2: [n][NS] adj; 3: [n][ND] quad; 4: [n][NS*NS] adj jac */
    long long n;
    int kind;
    int params_shared;
} SbEvalArgs;
