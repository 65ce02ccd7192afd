/*
 * sunode_b200.h -- C ABI of libsunode_b200.so, the drop-in boundary of the B200 engine.
 *
 * This library takes the place of the reference's cffi extension `_sundials_cvodes`
 * (/root/reference/sunode/build_cvodes.py:60-75, loaded at /root/reference/sunode/__init__.py:1)
 * for the solve path.  It is NOT a SUNDIALS re-implementation: instead of ~40 fine-grained
 * CVode* calls per solve it exports one call per reference *driver loop*, batched over
 * independent instances.  Each entry point below names the reference call sites it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are C-contiguous float64 / int32;
 *   - return value 0 = ok, < 0 = library error (text via sb_last_error());
 *     per-instance integrator outcomes are CVODES flag values in `status[B]`
 *     (/root/reference/include/cvodes/16_cvodes.h:45-106), failed instances are NaN-filled like
 *     the reference's Ops do (/root/reference/sunode/wrappers/as_pytensor.py:289-290,339-341);
 *   - `mem` says where the data pointers of that call live: SB_MEM_HOST (numpy; the library
 *     stages through device buffers it owns and synchronises before returning) or SB_MEM_DEVICE
 *     (e.g. torch `data_ptr()`; launches are asynchronous on `stream`).  `tvals` is always a host
 *     pointer.
 *   - a handle is single-stream and not re-entrant, like a reference Solver object
 *     (one CVODES memory + one user_data buffer, /root/reference/sunode/solver.py:226-227).
 */
#ifndef SUNODE_B200_H
#define SUNODE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_MEM_HOST 0
#define SB_MEM_DEVICE 1

#define SB_OK 0
#define SB_ERR_CUDA (-1001)      /* driver call failed / no device / libcuda missing */
#define SB_ERR_NVRTC (-1002)     /* JIT compilation failed (see log) */
#define SB_ERR_ARG (-1003)       /* invalid argument */
#define SB_ERR_STATE (-1004)     /* call sequence error, e.g. backward without stored forward */

#define SB_STATS_PER_INSTANCE 8  /* nst nfe nje nsetups netf ncfn nni n_points */

typedef struct sb_problem sb_problem;

int sb_version(void);
const char* sb_last_error(void);
int sb_device_count(int* count);
/* Version of the NVRTC that sb_compile uses (the toolkit's, /usr/local/cuda/lib64, unless
 * SUNODE_B200_NVRTC names another library); part of the cubin cache key of the Python layer. */
int sb_nvrtc_version(int* major, int* minor);

/* JIT-compile the generated problem functions (CUDA flavour of SympyProblem.generated) together
 * with the integrator kernels embedded in this library into a cubin for `arch` ("sm_100a").
 * Replaces: numba.njit/cfunc compilation of the callbacks in
 * /root/reference/sunode/problem.py:156-383 and /root/reference/sunode/symode/lambdify.py:203-270.
 * Needs no GPU.  *cubin and *log are malloc'd; release with sb_free(). */
int sb_compile(const char* generated_src, const char* arch, int block_threads, int min_blocks,
               void** cubin, size_t* cubin_size, char** log);
void sb_free(void* p);

/* Create a solver handle on `device` from a cubin produced by sb_compile.
 * Replaces: CVodeCreate/CVodeInit/CVodeSetUserData/CVodeSetLinearSolver/CVodeSetJacFn and the
 * adjoint setup CVodeAdjInit/CVodeCreateB/CVodeInitB/CVodeQuadInitB/...
 * (/root/reference/sunode/solver.py:221-235, 565-622). */
int sb_problem_create(sb_problem** out, int n_states, int n_params, int n_deriv,
                      const void* cubin, size_t cubin_size, int device);
int sb_problem_destroy(sb_problem* p);

/* CVodeSStolerances / CVodeSVtolerances (solver.py:394-417, 624-635). n_atol is 1 or n_states. */
int sb_set_tolerances(sb_problem* p, double rtol, const double* atol, int n_atol);
/* CVodeSetSensParams(ode, NULL, pbar, NULL) (solver.py:381-387, `scaling_factors`): with
 * CVodeSensEEtolerances (solver.py:389) the sensitivity block k is integrated with the absolute
 * tolerance atol / |pbar[k]|.  n = n_deriv; pbar = NULL restores pbar = 1. */
int sb_set_sens_scaling(sb_problem* p, const double* pbar, int n);
/* CVodeSStolerancesB (solver.py:599; README.md:246) */
int sb_set_tolerances_b(sb_problem* p, double rtol, double atol);
/* CVodeQuadSStolerancesB (solver.py:614; README.md:247) */
int sb_set_quad_tolerances_b(sb_problem* p, double rtol, double atol);
/* CVodeSetMaxNumSteps + the Python retry loops (solver.py:510-519, 759-768; README.md:248-249):
 * an instance may take mxstep * max_retries internal steps per output time / interval. */
int sb_set_max_num_steps(sb_problem* p, int mxstep, int max_retries);
int sb_set_max_num_steps_b(sb_problem* p, int mxstep, int max_retries);
/* CVodeAdjInit(steps, CV_POLYNOMIAL) (solver.py:588): stored forward steps per instance. */
int sb_set_history_capacity(sb_problem* p, int n_steps);

/* Solver.solve (solver.py:467-527) with store_history = 0, AdjointSolver.solve_forward
 * (solver.py:682-721) with store_history = 1; batched.
 *   y0[B][n_states], params[B][n_params] (all parameters, declaration order),
 *   y_out[B][n_t][n_states], status[B], stats[B][8] or NULL. */
int sb_solve_forward(sb_problem* p, int64_t B, double t0, const double* tvals, int n_t,
                     const double* y0, const double* params, double* y_out, int32_t* status,
                     int32_t* stats, int store_history, int mem, void* stream);

/* Solver.solve with forward sensitivities (Solver(sens_mode=...), solver.py:360-392, 483-527):
 * y and the n_deriv sensitivity vectors dy/dp_k are integrated together (CVODES simultaneous
 * corrector, sensitivity error control on, analytic sensitivity right-hand side).
 *   sens0[B][n_deriv][n_states] (or [n_deriv][n_states] if sens0_shared),
 *   sens_out[B][n_t][n_deriv][n_states]  (the reference's sens_out[n_t, n_params, n_states]). */
int sb_solve_forward_sens(sb_problem* p, int64_t B, double t0, const double* tvals, int n_t,
                          const double* y0, const double* params, const double* sens0,
                          int sens0_shared, double* y_out, double* sens_out, int32_t* status,
                          int32_t* stats, int mem, void* stream);

/* AdjointSolver.solve_backward (solver.py:723-784), batched, on the history stored by the last
 * sb_solve_forward(store_history = 1) of this handle (same B, n_t, tvals).
 *   t_start = the reference's `t0` argument (the LAST time), t_end = `tend` (the initial time);
 *   grads[B][n_t][n_states] (or [n_t][n_states] if grads_shared);
 *   grad_out[B][n_deriv], lamda_out[B][n_states]. */
int sb_solve_backward(sb_problem* p, int64_t B, double t_start, double t_end,
                      const double* tvals, int n_t, const double* params, const double* grads,
                      int grads_shared, double* grad_out, double* lamda_out, int32_t* status,
                      int32_t* stats, int mem, void* stream);
/* Optional traces for the NEXT sb_solve_backward / sb_solve_adjoint call of this handle (cleared by
 * it): lamda_all[B][n_t][n_states], quad_all[B][n_t][n_deriv] = lamda / quadrature right after the
 * jump at each output time (`lamda_all_out`, `quad_all_out` of solver.py:723-724,778-781; same row
 * convention).  The pointers live where that call's `mem` says.  Either may be NULL. */
int sb_set_backward_trace(sb_problem* p, double* lamda_all, double* quad_all);

/* One forward + one backward solve in a single call: the reference's unit of work
 * `solver.solve_forward(...); solver.solve_backward(tvals[-1], t0, tvals, grads, ...)`
 * (/root/reference/notebooks/from_sympy.ipynb:178-179; the Ops in
 * /root/reference/sunode/wrappers/as_pytensor.py:324-344 run exactly this pair).  Inputs are
 * uploaded once and only the results come back.  `status[B]` is the backward status, or the
 * forward status for instances whose forward solve failed; stats_fwd / stats_bwd may be NULL. */
int sb_solve_adjoint(sb_problem* p, int64_t B, double t0, const double* tvals, int n_t,
                     const double* y0, const double* params, const double* grads,
                     int grads_shared, double* y_out, double* grad_out, double* lamda_out,
                     int32_t* status, int32_t* stats_fwd, int32_t* stats_bwd, int mem,
                     void* stream);

/* Bounded memory (the reference's `checkpoint_n`, /root/reference/sunode/solver.py:533,588): the
 * step history and interpolation tables of one launch take hist_cap * (12 + 7 n_states) * 8 bytes
 * per instance; sb_solve_adjoint cuts the batch into chunks whose store fits `bytes` (0 = 4/5 of the
 * free device memory, the default) and runs forward + backward chunk by chunk.  sb_last_chunks:
 * how many chunks the last sb_solve_adjoint call of this handle needed. */
int sb_set_workspace_limit(sb_problem* p, size_t bytes);
int sb_last_chunks(sb_problem* p);

/* For the instances of the last forward solve: -1, or for a failed instance the index of the output
 * time it was integrating to -- the `time=` of the reference's error text (solver.py:516-519,
 * 716-719).  out[B] is host memory; synchronises the device. */
int sb_forward_fail_index(sb_problem* p, int64_t B, int32_t* out);

/* Batched evaluation of the generated functions: kind 0 rhs, 1 jacobian (column-major),
 * 2 adjoint rhs, 3 quadrature rhs, 4 adjoint jacobian -J^T (column-major).  Replaces calling the numba functions from Python
 * (/root/reference/sunode/wrappers/as_pytensor.py:160-183). */
int sb_eval(sb_problem* p, int kind, int64_t n, const double* t, const double* y,
            const double* params, int params_shared, const double* lam, double* out, int mem,
            void* stream);

int sb_synchronize(sb_problem* p);

/* Device time (CUDA events on the launching stream) of the most recent forward / table /
 * backward kernels, in milliseconds; -1 where no such launch was recorded. */
int sb_last_kernel_ms(sb_problem* p, float* forward_ms, float* tables_ms, float* backward_ms);
/* Number of kernels this handle has launched. */
int64_t sb_launch_count(sb_problem* p);
/* Registers per thread / max resident blocks per SM of the forward and backward kernels. */
int sb_kernel_info(sb_problem* p, int* regs_fwd, int* regs_bwd, int* blocks_per_sm_fwd,
                   int* blocks_per_sm_bwd, int* block_threads, int* sm_count);

/* Pinned host memory for callers that want asynchronous H2D/D2H staging. */
int sb_host_alloc(void** ptr, size_t bytes);
int sb_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* SUNODE_B200_H */
