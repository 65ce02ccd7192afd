"""Solver drivers with the reference's surface, running on the B200 engine.

Counterpart of ``sunode/solver.py`` of the reference:

* :class:`Solver` -- ``Solver.__init__`` (solver.py:242-317) and ``Solver.solve`` (:467-527);
* :class:`AdjointSolver` -- ``AdjointSolver.__init__`` (:531-622), ``solve_forward`` (:682-721)
  and ``solve_backward`` (:723-784).

The reference drives SUNDIALS CVODES one ``CVode``/``CVodeF``/``CVodeB`` call per output time from
Python.  Here each of those loops is ONE call into ``libsunode_b200.so`` which runs the whole loop
inside an sm_100a kernel for a batch of independent instances (``*_batch`` methods); the
reference-shaped batch-1 methods are the same call with ``B = 1`` and raise
:class:`~sunode_b200.basic.SolverError` with the reference's messages.  Nothing here falls back
to a CPU integrator: without a CUDA device every solve raises.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import _lib
from ._engine import Engine
from .basic import CV_TOO_MUCH_WORK, ERRORS, SolverError
from .problem import Problem

__all__ = ['Solver', 'AdjointSolver', 'SolverError']


def _as_dict(data: np.ndarray) -> Dict[str, Any]:
    if data.dtype.fields is not None:
        return {name: _as_dict(data[name]) for name in data.dtype.names}
    if data.shape == ():
        return data.item()
    return data


def _is_torch(x: Any) -> bool:
    return type(x).__module__.split('.')[0] == 'torch'


class _ParamsMixin:
    """Parameter plumbing shared by both solvers (reference solver.py:428-465, 650-680)."""
    _problem: Problem
    _user_data: np.ndarray

    def as_xarray(self, tvals, out, sens_out=None, unstack_state=True, unstack_params=True):
        return self._problem.solution_to_xarray(
            tvals, out, self._user_data, sensitivity=sens_out,
            unstack_state=unstack_state, unstack_params=unstack_params)

    @property
    def params_dtype(self):
        return self._problem.params_dtype

    @property
    def derivative_params_dtype(self):
        return self._problem.params_subset.subset_dtype

    @property
    def remainder_params_dtype(self):
        return self._problem.params_subset.remainder.subset_dtype

    def set_params(self, params):
        self._problem.update_params(self._user_data, params)

    def get_params(self):
        return self._problem.extract_params(self._user_data)

    def set_params_dict(self, params):
        data = self.get_params()
        self._problem.params_subset.from_dict(params, data)
        self.set_params(data)

    def get_params_dict(self):
        return _as_dict(self.get_params())

    def set_derivative_params(self, params):
        self._problem.update_subset_params(self._user_data, params)

    def set_remaining_params(self, params):
        self._problem.update_remaining_params(self._user_data, params)

    # tokens for the README-style raw pokes (sunode_b200._cvodes.lib)
    @property
    def _ode(self):
        from ._cvodes import OdeToken
        return OdeToken(self, False)

    @property
    def _odeB(self):
        from ._cvodes import OdeToken
        return OdeToken(self, True)

    # ---- helpers for the batch entry points
    def _flat_params(self) -> np.ndarray:
        return self._problem.flat_params(self._user_data)

    def _batch_params(self, params, B: int):
        """``params`` for a batch call: None -> the solver's current parameters for every
        instance; a structured array of ``params_dtype`` with shape (B,); or a float64
        ``[B, n_params_total]`` array / CUDA tensor in declaration order."""
        n_all = self._problem.n_params_total
        if params is None:
            return np.ascontiguousarray(np.broadcast_to(self._flat_params(), (B, n_all)))
        if _is_torch(params):
            return params
        params = np.asarray(params)
        if params.dtype == self._problem.params_dtype and params.dtype.fields is not None:
            params = np.ascontiguousarray(params).reshape(-1)
            if params.dtype.itemsize:
                params = params.view(np.float64).reshape(params.shape[0], n_all)
            else:
                params = np.zeros((params.shape[0], 0))
        params = np.asarray(params, dtype=np.float64)
        if params.ndim == 1:
            params = np.broadcast_to(params, (B, n_all))
        return np.ascontiguousarray(params)

    def _flat_state(self, y0) -> np.ndarray:
        if _is_torch(y0):
            return y0
        y0 = np.asarray(y0)
        if y0.dtype == self._problem.state_dtype and y0.dtype.fields is not None:
            y0 = np.ascontiguousarray(y0).reshape(-1).view(np.float64).reshape(-1, self._problem.n_states)
            return y0
        return y0


def _raise_forward(status: int, tvals, fail_k: int) -> None:
    """Reproduce the reference's error text (solver.py:516-519, 716-719): ``time`` is the output
    time the integrator was heading for (``fail_k`` from the forward kernel)."""
    t_fail = tvals[fail_k] if 0 <= fail_k < len(tvals) else (tvals[-1] if len(tvals) else float('nan'))
    if status == CV_TOO_MUCH_WORK:
        raise SolverError(f"Too many solver retries before time={t_fail}.")
    error = ERRORS.get(int(status), 'UNKNOWN')
    raise SolverError(f"Solving ode failed before time={t_fail}: {error} ({int(status)})")


def _constraint_defines(constraints, n_states) -> Tuple[Optional[np.ndarray], Tuple[str, ...]]:
    """``constraints`` of the reference (solver.py:268-271, 566-572: broadcast to the states and
    handed to ``CVodeSetConstraints``; flags 0 none, +-1 ``y >= 0`` / ``<= 0``, +-2 strict) as a
    build option of the kernels: the flags are compile-time constants of the forward integrator
    (``Bdf::check_constraints`` in csrc/sb_bdf.cuh)."""
    if constraints is None:
        return None, ()
    c = np.broadcast_to(np.asarray(constraints, dtype=np.float64), (n_states,)).copy()
    if not np.isin(c, (-2.0, -1.0, 0.0, 1.0, 2.0)).all():
        raise ValueError('Bad return code from sundials: CV_ILL_INPUT (-22)')   # CVodeSetConstraints
    if not c.any():
        return c, ()
    return c, ('SB_CONSTRAINTS=' + ','.join('%.1f' % v for v in c),)


class Solver(_ParamsMixin):
    """Forward solver (reference ``Solver``, solver.py:213-527), dense-BDF path.

    Options of the reference that select other SUNDIALS modules are accepted for signature
    compatibility and rejected with ``NotImplementedError`` when they would change the
    algorithm (ADAMS, non-dense linear solvers).  ``constraints`` (CVodeSetConstraints flags per
    state) are a build option of the forward kernel.

    ``sens_mode`` = ``"simultaneous"`` or ``"staggered"`` turns on forward sensitivity analysis
    (reference solver.py:360-392): y and dy/dp_k for the derivative parameters are integrated
    together with the analytic sensitivity right-hand side and sensitivity error control, as the
    reference configures CVODES.  The engine implements the simultaneous corrector; "staggered"
    is accepted and runs the same kernel (the two differ in the order of the corrector
    iterations, not in what is computed or in the error control)."""

    def __init__(self, problem: Problem, *, abstol=1e-10, reltol=1e-10,
                 sens_mode: Optional[str] = None, scaling_factors: Optional[np.ndarray] = None,
                 constraints: Optional[np.ndarray] = None, solver='BDF', linear_solver='dense',
                 linear_solver_kwargs=None, device: Optional[int] = None,
                 block_threads: Optional[int] = None, min_blocks: Optional[int] = None):
        if linear_solver_kwargs is None:
            linear_solver_kwargs = {}
        if solver not in ('BDF', 'ADAMS'):
            raise ValueError(f'Unknown solver {solver}.')
        if solver != 'BDF':
            raise NotImplementedError('Only the BDF method is implemented on the B200 engine.')
        if linear_solver != 'dense':
            raise NotImplementedError(
                'Only linear_solver="dense" (in-register LU with the analytic Jacobian) is implemented.')
        if sens_mode == 'staggered1':
            raise ValueError('staggered1 not implemented.')
        if sens_mode not in (None, 'simultaneous', 'staggered'):
            raise ValueError('sens_mode must be one of "simultaneous" and "staggered".')
        if sens_mode is not None and problem.n_params == 0:
            raise ValueError('Forward sensitivities need at least one derivative parameter.')
        if scaling_factors is not None:
            scaling_factors = np.asarray(scaling_factors, dtype=np.float64)
            if scaling_factors.shape != (problem.n_params,):
                raise ValueError('Invalid shape of scaling_factors.')
            if sens_mode is not None and not np.all(scaling_factors != 0.0):
                # CVodeSetSensParams: "pbar has zero component(s) (illegal)"
                raise ValueError('Bad return code from sundials: CV_ILL_INPUT (-22)')
        constraints, self._defines = _constraint_defines(constraints, problem.n_states)
        if self._defines and sens_mode is not None:
            # CVODES rejects constraints together with the simultaneous corrector (the one the
            # engine implements) at the first solve: CV_ILL_INPUT
            raise NotImplementedError('Constraints together with forward sensitivities are not implemented.')
        self._problem = problem
        self._user_data = problem.make_user_data()
        self._constraints = constraints
        self._abstol = abstol
        self._reltol = reltol
        self._linear_solver_kind = linear_solver
        self._linear_solver_kwargs = linear_solver_kwargs
        self._sens_mode = sens_mode
        self._scaling_factors = scaling_factors
        self._solver_kind = solver
        self._device = device
        self._launch_cfg = (block_threads, min_blocks)
        self._mxstep = 500
        self._state_names = ['_problem', '_user_data', '_constraints', '_abstol', '_reltol',
                             '_linear_solver_kind', '_linear_solver_kwargs', '_sens_mode',
                             '_scaling_factors',
                             '_solver_kind', '_device', '_launch_cfg', '_mxstep', '_defines',
                             '_state_names']
        self._init_engine()

    def _init_engine(self) -> None:
        self._compute_sens = self._sens_mode is not None
        self._engine = Engine(self._problem.generated, device=self._device,
                              block_threads=self._launch_cfg[0], min_blocks=self._launch_cfg[1],
                              defines=self._defines)
        if self._compute_sens and self._scaling_factors is not None:
            # CVodeSetSensParams(pbar) + CVodeSensEEtolerances (solver.py:381-389)
            self._engine.set_sens_scaling(self._scaling_factors)
        self._set_tolerances(self._abstol, self._reltol)

    # pickling like the reference (solver.py:319-324): configuration only, handle re-created
    def __getstate__(self):
        return {name: self.__dict__[name] for name in self._state_names}

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._init_engine()

    def _set_tolerances(self, atol=None, rtol=None):
        atol = np.array(atol, dtype=np.float64)
        rtol = np.array(rtol, dtype=np.float64)
        n_states = self._problem.n_states
        if rtol.ndim != 0:
            raise NotImplementedError('Vector rtol is not implemented; use a scalar rtol.')
        if atol.ndim == 1:
            assert atol.shape == (n_states,)
        elif atol.ndim != 0:
            raise ValueError('Invalid tolerance.')
        self._engine.set_tolerances(float(rtol), atol)
        self._atol = atol
        self._rtol = rtol

    def set_max_num_steps(self, mxstep: int) -> None:
        """``lib.CVodeSetMaxNumSteps(solver._ode, mxstep)`` of the reference README (:248)."""
        self._mxstep = int(mxstep)

    def make_output_buffers(self, tvals):
        n_states, n_params = self._problem.n_states, self._problem.n_params
        y_vals = np.zeros((len(tvals), n_states))
        if self._compute_sens:
            return y_vals, np.zeros((len(tvals), n_params, n_states))
        return y_vals

    # ------------------------------------------------------------------ batch-1, reference API
    def solve(self, t0, tvals, y0, y_out, *, sens0=None, sens_out=None, max_retries=5):
        if self._compute_sens and (sens0 is None or sens_out is None):
            raise ValueError('"sens_out" and "sens0" are required when computin sensitivities.')
        n_states = self._problem.n_states
        y0 = np.asarray(y0)
        if y0.dtype == self._problem.state_dtype and y0.dtype.fields is not None:
            y0 = y0[None].view(np.float64)
        if y0.shape != (n_states,):
            raise ValueError(f"y0 should have shape {(n_states,)} but has shape {y0.shape}.")
        tvals = np.asarray(tvals, dtype=np.float64)
        if self._compute_sens:
            sens0 = np.ascontiguousarray(sens0, dtype=np.float64)
            out, sens, status = self.solve_sens_batch(t0, tvals, y0[None, :], None, sens0,
                                                      max_retries=max_retries)
            if status[0] != 0:
                _raise_forward(int(status[0]), tvals, int(self._engine.forward_fail_index(1)[0]))
            y_out[...] = out[0]
            sens_out[...] = sens[0]
            return
        out, status = self.solve_batch(t0, tvals, y0[None, :], None, max_retries=max_retries)
        if status[0] != 0:
            _raise_forward(int(status[0]), tvals, int(self._engine.forward_fail_index(1)[0]))
        y_out[...] = out[0]

    # ------------------------------------------------------------------ batched
    def solve_batch(self, t0, tvals, y0, params=None, y_out=None, *, status=None, stats=None,
                    max_retries=5, stream=None) -> Tuple[Any, Any]:
        """Solve ``B`` independent instances: ``y0[B, n_states]``, ``params[B, n_params_total]``
        (or None for the solver's current parameters).  Returns ``(y_out[B, n_t, n_states],
        status[B])``; failed instances are NaN rows with a CVODES flag in ``status``."""
        y0 = self._flat_state(y0)
        B = int(y0.shape[0])
        tvals = np.asarray(tvals, dtype=np.float64)
        params = self._batch_params(params, B)
        y_out, status = _alloc_like(y0, y_out, (B, len(tvals), self._problem.n_states), status, B)
        self._engine.set_max_num_steps(self._mxstep, max_retries)
        self._engine.forward(float(t0), tvals, y0, params, y_out, status, stats,
                             store_history=False, stream=stream)
        return y_out, status


def _solve_sens_batch(self, t0, tvals, y0, params, sens0, y_out=None, sens_out=None, *,
                      status=None, stats=None, max_retries=5, stream=None):
    """Batched solve with forward sensitivities: ``sens0`` is ``[n_deriv, n_states]`` (shared by
    all instances) or ``[B, n_deriv, n_states]``.  Returns ``(y_out[B, n_t, n_s],
    sens_out[B, n_t, n_deriv, n_s], status[B])``."""
    if not self._compute_sens:
        raise ValueError('The solver was created without sens_mode.')
    y0 = self._flat_state(y0)
    B = int(y0.shape[0])
    tvals = np.asarray(tvals, dtype=np.float64)
    params = self._batch_params(params, B)
    if not _is_torch(sens0):
        sens0 = np.ascontiguousarray(sens0, dtype=np.float64)
    n_s, n_d = self._problem.n_states, self._problem.n_params
    y_out, status = _alloc_like(y0, y_out, (B, len(tvals), n_s), status, B)
    sens_out = _alloc_out(y0, sens_out, (B, len(tvals), n_d, n_s))
    self._engine.set_max_num_steps(self._mxstep, max_retries)
    self._engine.forward_sens(float(t0), tvals, y0, params, sens0, y_out, sens_out, status, stats,
                              stream=stream)
    return y_out, sens_out, status


Solver.solve_sens_batch = _solve_sens_batch


def _alloc_like(like, out, shape, status, B):
    if _is_torch(like) and like.is_cuda:
        import torch
        if out is None:
            out = torch.empty(shape, dtype=torch.float64, device=like.device)
        if status is None:
            status = torch.empty((B,), dtype=torch.int32, device=like.device)
    else:
        if out is None:
            out = np.empty(shape)
        if status is None:
            status = np.empty((B,), dtype=np.int32)
    return out, status


def _alloc_out(like, out, shape):
    if out is not None:
        return out
    if _is_torch(like) and like.is_cuda:
        import torch
        return torch.empty(shape, dtype=torch.float64, device=like.device)
    return np.empty(shape)


class AdjointSolver(_ParamsMixin):
    """Forward + adjoint solver (reference ``AdjointSolver``, solver.py:530-784).

    As in the reference the backward problem and its quadrature always run with tolerances
    1e-10 (solver.py:599,614) unless changed through :meth:`set_backward_tolerances` /
    :meth:`set_quad_tolerances` -- the counterparts of the raw ``lib.CVodeSStolerancesB`` /
    ``lib.CVodeQuadSStolerancesB`` pokes shown in the reference README (:243-249)."""

    def __init__(self, problem: Problem, *, abstol=1e-10, reltol=1e-10, checkpoint_n=500_000,
                 interpolation='polynomial', constraints=None, solver='BDF', adjoint_solver='BDF',
                 backward: str = 'reference',
                 device: Optional[int] = None, history_capacity: Optional[int] = None,
                 block_threads: Optional[int] = None, min_blocks: Optional[int] = None):
        if solver not in ('BDF', 'ADAMS'):
            raise ValueError(f'Unknown solver {solver}.')
        if adjoint_solver not in ('BDF', 'ADAMS'):
            raise ValueError(f'Unknown solver {adjoint_solver}.')
        if solver != 'BDF' or adjoint_solver != 'BDF':
            raise NotImplementedError('Only the BDF method is implemented on the B200 engine.')
        if interpolation not in ('polynomial', 'hermite'):
            assert False
        constraints, constraint_defines = _constraint_defines(constraints, problem.n_states)
        self._problem = problem
        self._user_data = problem.make_user_data()
        self._constraints = constraints
        self._interpolation = interpolation
        # CV_HERMITE (solver.py:581-586): the forward kernel also stores y' at every step and the
        # table kernel writes cubic Hermite entries; a build option of the kernels (csrc/sb_args.h)
        defines = (('SB_HERMITE',) if interpolation == 'hermite' else ()) + constraint_defines
        # backward='fundamental' (not a reference option): the restart-free
        # backward pass of csrc/sb_fund.cuh -- the fundamental matrix of the adjoint equation is
        # integrated without restarts and the jumps at the output times become small dense
        # solves.  Same results to the tolerances, ~6x fewer backward steps on smooth problems;
        # the default keeps the reference's restart-per-output-time schedule.
        if backward not in ('reference', 'fundamental'):
            raise ValueError(f'Unknown backward schedule {backward}.')
        if backward == 'fundamental':
            if problem.n_states > 4:
                raise NotImplementedError(
                    'backward="fundamental" integrates n_states^2 + n_states * n_params components '
                    'per lane; it is implemented for up to 4 states.')
            defines = defines + ('SB_FUND',)
        self._backward = backward
        self._engine = Engine(problem.generated, device=device, block_threads=block_threads,
                              min_blocks=min_blocks, defines=defines)
        # the reference keeps every forward step of one solve in memory (checkpoint_n = 500 000
        # steps per checkpoint, solver.py:533,588); here the per-instance capacity is explicit
        self._history_capacity = int(history_capacity or min(int(checkpoint_n), 1024))
        self._engine.set_history_capacity(self._history_capacity)
        # Without an explicit capacity the store grows like the reference's does (its checkpoints
        # hold 500 000 steps): a host-memory solve whose instances ran out of history slots is
        # repeated with four times the capacity, up to checkpoint_n steps / _HISTORY_BYTES_MAX.
        self._history_auto = history_capacity is None
        self._history_limit = int(checkpoint_n)
        self._set_tolerances(abstol, reltol)
        self._engine.set_tolerances_b(1e-10, 1e-10)        # solver.py:599
        self._engine.set_quad_tolerances_b(1e-10, 1e-10)   # solver.py:614
        self._mxstep = 500
        self._mxstep_b = 500
        self._last_forward: Optional[Tuple[int, int]] = None

    def _set_tolerances(self, atol=None, rtol=None):
        atol = np.array(atol, dtype=np.float64)
        rtol = np.array(rtol, dtype=np.float64)
        if not (atol.ndim in (0, 1) and rtol.ndim == 0):
            raise ValueError('Invalid tolerance.')
        self._engine.set_tolerances(float(rtol), atol)
        self._atol = atol
        self._rtol = rtol

    def set_backward_tolerances(self, reltol: float, abstol: float) -> None:
        """``lib.CVodeSStolerancesB(solver._ode, solver._odeB, reltol, abstol)`` (README :246)."""
        self._engine.set_tolerances_b(reltol, abstol)

    def set_quad_tolerances(self, reltol: float, abstol: float) -> None:
        """``lib.CVodeQuadSStolerancesB(...)`` (README :247)."""
        self._engine.set_quad_tolerances_b(reltol, abstol)

    def set_max_num_steps(self, mxstep: int) -> None:
        self._mxstep = int(mxstep)

    def set_max_num_steps_backward(self, mxstep: int) -> None:
        self._mxstep_b = int(mxstep)

    def set_history_capacity(self, n_steps: int) -> None:
        self._history_capacity = int(n_steps)
        self._history_auto = False
        self._engine.set_history_capacity(self._history_capacity)

    def set_workspace_limit(self, n_bytes: int) -> None:
        """Upper bound in bytes for the step history + interpolation tables of one launch (default:
        4/5 of the free device memory).  :meth:`solve_adjoint_batch` cuts a batch whose store would
        not fit into chunks and runs forward + backward chunk by chunk -- the bounded-memory
        counterpart of the reference's ``checkpoint_n`` (solver.py:533,588)."""
        self._engine.set_workspace_limit(int(n_bytes))

    _HISTORY_BYTES_MAX = 32 << 30

    def _grow_history(self, B: int, status, stats) -> bool:
        """True when instances of a host-memory solve ran out of history slots (CV_TOO_MUCH_WORK
        with every slot used) and the capacity could be raised; the caller then solves again."""
        if not self._history_auto or _is_torch(status) or stats is None:
            return False
        full = (np.asarray(status) == CV_TOO_MUCH_WORK) & (np.asarray(stats)[:, 7] >= self._history_capacity)
        if not full.any():
            return False
        n_s = self._problem.n_states
        new = min(self._history_capacity * 4, max(self._history_limit, self._history_capacity))
        per_step = 8 * ((2 * n_s + 2) + (10 + 6 * n_s))       # history point + table entry
        if new <= self._history_capacity or B * new * per_step > self._HISTORY_BYTES_MAX:
            return False
        self._history_capacity = new
        self._engine.set_history_capacity(new)
        return True

    def make_output_buffers(self, tvals):
        y_vals = np.zeros((len(tvals), self._problem.n_states))
        grad_out = np.zeros(self._problem.n_params)
        lamda_out = np.zeros(self._problem.n_states)
        return y_vals, grad_out, lamda_out

    # ------------------------------------------------------------------ batch-1, reference API
    def solve_forward(self, t0, tvals, y0, y_out, *, max_retries=5):
        y0 = np.asarray(y0)
        if y0.dtype == self._problem.state_dtype and y0.dtype.fields is not None:
            y0 = y0[None].view(np.float64)
        y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(1, self._problem.n_states)
        tvals = np.asarray(tvals, dtype=np.float64)
        out, status = self.solve_forward_batch(t0, tvals, y0, None, max_retries=max_retries)
        if status[0] != 0:
            _raise_forward(int(status[0]), tvals, int(self._engine.forward_fail_index(1)[0]))
        y_out[...] = out[0]

    def solve_backward(self, t0, tend, tvals, grads, grad_out, lamda_out,
                       lamda_all_out=None, quad_all_out=None, max_retries=50):
        tvals = np.asarray(tvals, dtype=np.float64)
        n_s, n_d = self._problem.n_states, self._problem.n_params
        grads = np.ascontiguousarray(grads, dtype=np.float64).reshape(1, len(tvals), n_s)
        lam_all = np.empty((1, len(tvals), n_s)) if lamda_all_out is not None else None
        quad_all = np.empty((1, len(tvals), n_d)) if quad_all_out is not None else None
        g, lam, status = self.solve_backward_batch(t0, tend, tvals, grads, None,
                                                   max_retries=max_retries,
                                                   lamda_all_out=lam_all, quad_all_out=quad_all)
        if status[0] != 0:
            code = int(status[0])
            if code == CV_TOO_MUCH_WORK:
                raise SolverError(f"Too many solver retries between time {t0} and {tend}.")
            raise SolverError(f"Solving ode failed between time {t0} and {tend}: "
                              f"{ERRORS.get(code, 'UNKNOWN')} ({code})")
        grad_out[:] = g[0]
        lamda_out[:] = lam[0]
        if lamda_all_out is not None:
            lamda_all_out[...] = lam_all[0]
        if quad_all_out is not None:
            quad_all_out[...] = quad_all[0]

    # ------------------------------------------------------------------ batched
    def solve_forward_batch(self, t0, tvals, y0, params=None, y_out=None, *, status=None,
                            stats=None, max_retries=5, stream=None):
        """Batched ``solve_forward``; the step history of every instance stays on the device
        for a following :meth:`solve_backward_batch`.

        History capacity: with host (numpy) arrays and no explicit ``history_capacity`` the store
        grows like the reference's checkpoints do -- instances that ran out of slots
        (``CV_TOO_MUCH_WORK`` with every slot used) make the solve repeat with four times the
        capacity, up to ``checkpoint_n``.  With device (torch CUDA) arrays the call is asynchronous
        and the status is not inspected: the capacity is what the solver was created with (default
        1 024 steps), such instances report ``CV_TOO_MUCH_WORK``; pass ``history_capacity=`` for
        problems that take more forward steps."""
        y0 = self._flat_state(y0)
        B = int(y0.shape[0])
        tvals = np.asarray(tvals, dtype=np.float64)
        params = self._batch_params(params, B)
        y_out, status = _alloc_like(y0, y_out, (B, len(tvals), self._problem.n_states), status, B)
        self._engine.set_max_num_steps(self._mxstep, max_retries)
        if stats is None and self._history_auto and not _is_torch(status):
            stats = np.empty((B, 8), dtype=np.int32)
        while True:
            self._engine.forward(float(t0), tvals, y0, params, y_out, status, stats,
                                 store_history=True, stream=stream)
            if not self._grow_history(B, status, stats):
                break
        self._last_forward = (B, len(tvals))
        return y_out, status

    def solve_backward_batch(self, t0, tend, tvals, grads, params=None, grad_out=None,
                             lamda_out=None, *, status=None, stats=None, max_retries=50,
                             lamda_all_out=None, quad_all_out=None, stream=None):
        """Batched ``solve_backward`` on the stored forward pass.  ``t0`` is the LAST time and
        ``tend`` the initial time, as in the reference (solver.py:723-724).  ``grads`` is
        ``[B, n_t, n_states]`` or ``[n_t, n_states]`` (shared by all instances).  ``params=None``
        reuses the parameters of the stored forward pass.  Returns
        ``(grad_out[B, n_deriv], lamda_out[B, n_states], status[B])``."""
        if self._last_forward is None:
            raise SolverError('solve_backward called before solve_forward.')
        B, n_t = self._last_forward
        tvals = np.asarray(tvals, dtype=np.float64)
        if params is not None:
            params = self._batch_params(params, B)
        if not _is_torch(grads):
            grads = np.ascontiguousarray(grads, dtype=np.float64)
        n_s, n_d = self._problem.n_states, self._problem.n_params
        grad_out = _alloc_out(grads, grad_out, (B, n_d))
        lamda_out, status = _alloc_like(grads, lamda_out, (B, n_s), status, B)
        self._engine.set_max_num_steps_b(self._mxstep_b, max_retries)
        self._engine.backward(float(t0), float(tend), tvals, params, grads, grad_out, lamda_out,
                              status, stats, lamda_all=lamda_all_out, quad_all=quad_all_out,
                              stream=stream)
        return grad_out, lamda_out, status

    def solve_adjoint_batch(self, t0, tvals, y0, params, grads, *, y_out=None, grad_out=None,
                            lamda_out=None, status=None, stats_fwd=None, stats_bwd=None,
                            max_retries=5, max_retries_backward=50, stream=None):
        """``solve_forward`` followed by ``solve_backward(tvals[-1], t0, tvals, grads, ...)`` for a
        batch, as one library call (inputs uploaded once).  Returns
        ``(y_out, grad_out, lamda_out, status)``."""
        y0 = self._flat_state(y0)
        B = int(y0.shape[0])
        tvals = np.asarray(tvals, dtype=np.float64)
        params = self._batch_params(params, B)
        if not _is_torch(grads):
            grads = np.ascontiguousarray(grads, dtype=np.float64)
        n_s, n_d = self._problem.n_states, self._problem.n_params
        y_out, status = _alloc_like(y0, y_out, (B, len(tvals), n_s), status, B)
        grad_out = _alloc_out(y0, grad_out, (B, n_d))
        lamda_out = _alloc_out(y0, lamda_out, (B, n_s))
        self._engine.set_max_num_steps(self._mxstep, max_retries)
        self._engine.set_max_num_steps_b(self._mxstep_b, max_retries_backward)
        if stats_fwd is None and self._history_auto and not _is_torch(status):
            stats_fwd = np.empty((B, 8), dtype=np.int32)
        while True:
            self._engine.adjoint(float(t0), tvals, y0, params, grads, y_out, grad_out, lamda_out,
                                 status, stats_fwd, stats_bwd, stream=stream)
            if not self._grow_history(B, status, stats_fwd):
                break
        self._last_forward = (B, len(tvals))
        return y_out, grad_out, lamda_out, status
