"""PyTensor Ops over the B200 solvers -- the counterpart of the reference's
``sunode/wrappers/as_pytensor.py`` (``solve_ivp`` :20-137, ``EvalRhs`` :140-183,
``SolveODEAdjoint`` :266-308, ``SolveODEAdjointBackward`` :311-344).

Same entry point, argument meaning and return tuples as the reference for
``derivatives='adjoint'`` and ``derivatives='forward'`` (``SolveODE`` :186-263).

PyTensor is optional at import time: the Ops' numeric bodies (``perform``) only need numpy and
are exercised directly by the tests; building a graph (``solve_ivp``, ``Op.__call__``,
``Op.grad``) needs PyTensor and raises ``ImportError`` without it.

In addition to the reference's batch-1 Ops there is :class:`SolveODEAdjointBatch`, which solves a
whole batch of draws per call (``y0[B, n_s]``, ``params[B, n_deriv]``) -- the form in which the
GPU engine is actually fast.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import numpy as np

from .. import basic
from ..dtypesubset import as_flattened
from ..solver import AdjointSolver, Solver, SolverError
from ..symode.problem import SympyProblem

try:  # pragma: no cover - pytensor is not part of the build image
    import pytensor.tensor as pt
    from pytensor.gradient import grad_not_implemented
    from pytensor.graph.basic import Constant, Variable
    from pytensor.graph.op import Op
    HAVE_PYTENSOR = True
except ImportError:  # numeric bodies stay usable and testable
    pt = None
    HAVE_PYTENSOR = False

    class Op:  # type: ignore[no-redef]
        """Stand-in base class: keeps ``perform`` callable without PyTensor."""

        def __call__(self, *args, **kwargs):
            raise ImportError('pytensor is required to build graphs with sunode_b200 Ops')


def _need_pytensor() -> None:
    if not HAVE_PYTENSOR:
        raise ImportError('pytensor is required for sunode_b200.wrappers.as_pytensor.solve_ivp')


def solve_ivp(
    t0: float,
    y0: Any,
    params: Dict[str, Any],
    tvals: Any,
    rhs: Callable[..., Dict[str, Any]],
    derivatives: str = 'adjoint',
    coords: Optional[Dict[str, Any]] = None,
    make_solver=None,
    derivative_subset=None,
    solver_kwargs=None,
    simplify=None,
) -> Any:
    """Build the ODE solution as a PyTensor graph (reference as_pytensor.py:20-137).

    ``y0`` / ``params`` are (nested) dicts whose leaves are ``(tensor, shape_or_dims)`` tuples or
    plain arrays; parameters given as non-constant PyTensor variables become derivative
    parameters.  Returns ``(solution_dict, flat_solution, problem, solver, y0_flat,
    params_subs_flat)``."""
    _need_pytensor()
    if derivatives not in ('adjoint', 'forward'):
        raise ValueError('derivatives must be "adjoint" or "forward"')
    solver_kwargs = dict(solver_kwargs or {})
    if derivatives == 'forward':
        # sensitivities w.r.t. the initial values ride along as pseudo-parameters (:37-39)
        params = dict(params)
        params['__initial_values'] = y0
    dtype = basic.data_dtype

    def leaf(val):
        return val if isinstance(val, tuple) else (val, None)

    def dims_of(vals, name=None):
        if isinstance(vals, dict):
            return {key: dims_of(item, key) for key, item in vals.items()}
        tensor, dim_names = leaf(vals)
        if dim_names is None:
            dim_names = pt.as_tensor_variable(tensor, dtype='float64').type.shape
            if any(d is None for d in dim_names):
                raise ValueError('Shapes of tensors need to be statically known or given explicitly.')
        if isinstance(dim_names, (str, int)):
            dim_names = (dim_names,)
        tensor = pt.as_tensor_variable(tensor, dtype='float64')
        if tensor.ndim != len(dim_names):
            raise ValueError(f'Dimension mismatch for {name}: Value has rank {tensor.ndim}, '
                             f'but {len(dim_names)} was specified.')
        if np.dtype(tensor.dtype) != dtype:
            raise ValueError(f'Dtype mismatch for {name}: Got {tensor.dtype} but expected {dtype}.')
        return dim_names

    y0_dims = dims_of(y0)
    params_dims = dims_of(params)
    flat_params = as_flattened(params)
    if derivative_subset is None:
        derivative_subset = [
            path for path, val in flat_params.items()
            if isinstance(leaf(val)[0], Variable) and not isinstance(leaf(val)[0], Constant)]

    problem = SympyProblem(params_dims, y0_dims, rhs, derivative_subset, coords=coords,
                           simplify=simplify)

    def concat(table, paths):
        parts = [pt.as_tensor_variable(leaf(table[p])[0], dtype='float64').reshape((-1,))
                 for p in paths]
        return pt.concatenate(parts) if parts else pt.as_tensor_variable(np.zeros(0), dtype='float64')

    params_subs_flat = concat(flat_params, problem.params_subset.subset_paths)
    params_remaining_flat = concat(flat_params, problem.params_subset.remainder.subset_paths)
    y0_flat = concat(as_flattened(y0), problem.state_subset.paths)
    t0 = pt.as_tensor_variable(t0, dtype='float64')
    tvals = pt.as_tensor_variable(tvals, dtype='float64')

    if derivatives == 'forward':
        if 'sens_mode' not in solver_kwargs:
            raise ValueError('When `derivatives=True`, the `solver_kwargs` must contain one of '
                             '`sens_mode={"simultaneous" | "staggered"}`.')
        sol = make_solver(problem, **solver_kwargs) if make_solver else Solver(problem, **solver_kwargs)
        wrapper = SolveODE(sol)
        flat_solution, flat_sens = wrapper(y0_flat, params_subs_flat, params_remaining_flat, t0, tvals)
        solution = problem.flat_solution_as_dict(flat_solution)
        return (solution, flat_solution, problem, sol, y0_flat, params_subs_flat, flat_sens, wrapper)
    sol = make_solver(problem, **solver_kwargs) if make_solver else AdjointSolver(problem, **solver_kwargs)
    wrapper = SolveODEAdjoint(sol)
    flat_solution = wrapper(y0_flat, params_subs_flat, params_remaining_flat, t0, tvals)
    solution = problem.flat_solution_as_dict(flat_solution)
    return solution, flat_solution, problem, sol, y0_flat, params_subs_flat


class _SolverOp(Op):
    """Shared plumbing: identity by solver (reference: ``__props__ = ('_solver_id',)``) and the
    scatter of the flat derivative / remaining parameter vectors into the solver."""
    __props__ = ('_solver_id',)

    def __init__(self, solver):
        self._solver = solver
        self._solver_id = id(solver)
        self._deriv_dtype = solver.derivative_params_dtype
        self._fixed_dtype = solver.remainder_params_dtype

    def _set_params(self, params, params_fixed) -> None:
        params = np.ascontiguousarray(params, dtype=np.float64)
        params_fixed = np.ascontiguousarray(params_fixed, dtype=np.float64)
        if self._deriv_dtype.itemsize:
            self._solver.set_derivative_params(params.view(self._deriv_dtype)[0])
        if self._fixed_dtype.itemsize:
            self._solver.set_remaining_params(params_fixed.view(self._fixed_dtype)[0])


class EvalRhs(_SolverOp):
    """rhs(t_i, y_i) for every output time: the factor of d solution / d tvals
    (reference :140-183).  Evaluated on the device with ``sb_eval``."""
    if HAVE_PYTENSOR:  # pragma: no cover
        itypes = [pt.dvector, pt.dvector, pt.dmatrix, pt.dvector]
        otypes = [pt.dmatrix]

    def perform(self, node, inputs, outputs):
        params, params_fixed, y, tvals = inputs
        self._set_params(params, params_fixed)
        problem = self._solver._problem
        p = problem.flat_params(self._solver._user_data)
        out = np.empty((len(tvals), problem.n_states))
        self._solver._engine.eval(0, np.ascontiguousarray(tvals, dtype=np.float64),
                                  np.ascontiguousarray(y, dtype=np.float64), p, None, out,
                                  params_shared=True)
        if not np.isfinite(out).all():
            raise ValueError('Bad ode rhs return code: 1')
        outputs[0][0] = out


def initial_sensitivities(problem) -> np.ndarray:
    """``sens0[n_deriv, n_states]``: zero, except that a derivative parameter living under
    ``__initial_values`` is the initial value of the state with the same path, so its
    sensitivity starts as the matching unit vector (reference :211-230)."""
    subset = problem.params_subset
    state_slices = problem.state_subset.flat_slices
    sens0 = np.zeros((subset.n_subset, problem.n_states))
    row = 0
    for path in subset.subset_paths:
        n_items = int(np.prod(subset.flat_shapes[path], dtype=int))
        if path and path[0] == '__initial_values':
            start = state_slices[tuple(path[1:])].start
            for i in range(n_items):
                sens0[row + i, start + i] = 1.0
        row += n_items
    return sens0


class SolveODE(_SolverOp):
    """Forward solve with forward sensitivities (reference :186-263): outputs the solution
    ``[n_t, n_s]`` and ``d solution / d params`` ``[n_t, n_deriv, n_s]``."""
    if HAVE_PYTENSOR:  # pragma: no cover
        itypes = [pt.dvector, pt.dvector, pt.dvector, pt.dscalar, pt.dvector]
        otypes = [pt.dmatrix, pt.dtensor3]

    def __init__(self, solver):
        super().__init__(solver)
        self._sens0 = initial_sensitivities(solver._problem)

    def perform(self, node, inputs, outputs):
        y0, params, params_fixed, t0, tvals = inputs
        y_out, sens_out = self._solver.make_output_buffers(tvals)
        self._set_params(params, params_fixed)
        try:
            self._solver.solve(float(t0), tvals, np.asarray(y0, dtype=np.float64), y_out,
                               sens0=self._sens0, sens_out=sens_out)
        except SolverError:
            y_out[...] = np.nan
            sens_out[...] = np.nan
        outputs[0][0] = y_out
        outputs[1][0] = sens_out

    def grad(self, inputs, g):  # pragma: no cover - needs pytensor
        g, g_grad = g
        _, params, params_fixed, t0, tvals = inputs
        assert str(g_grad) == '<DisconnectedType>'
        solution, sens = self(*inputs)
        return [
            pt.zeros_like(inputs[0]),
            pt.sum(g[:, None, :] * sens, (0, -1)),
            grad_not_implemented(self, 2, params_fixed),
            grad_not_implemented(self, 3, t0),
            (EvalRhs(self._solver)(params, params_fixed, solution, tvals) * g).sum(-1),
        ]


class SolveODEAdjoint(_SolverOp):
    """Forward solve; its gradient is :class:`SolveODEAdjointBackward` (reference :266-308)."""
    if HAVE_PYTENSOR:  # pragma: no cover
        itypes = [pt.dvector, pt.dvector, pt.dvector, pt.dscalar, pt.dvector]
        otypes = [pt.dmatrix]

    def perform(self, node, inputs, outputs):
        y0, params, params_fixed, t0, tvals = inputs
        y_out, _, _ = self._solver.make_output_buffers(tvals)
        self._set_params(params, params_fixed)
        try:
            self._solver.solve_forward(float(t0), tvals, np.asarray(y0, dtype=np.float64), y_out)
        except SolverError:
            y_out[:] = np.nan
        outputs[0][0] = y_out

    def grad(self, inputs, g):  # pragma: no cover - needs pytensor
        g, = g
        y0, params, params_fixed, t0, tvals = inputs
        solution = self(*inputs)
        lamda, gradient = SolveODEAdjointBackward(self._solver)(y0, params, params_fixed, g, t0, tvals)
        return [
            -lamda,
            gradient,
            grad_not_implemented(self, 2, params_fixed),
            grad_not_implemented(self, 3, t0),
            (EvalRhs(self._solver)(params, params_fixed, solution, tvals) * g).sum(-1),
        ]


class SolveODEAdjointBackward(_SolverOp):
    """(lamda(t0), dL/dparams) from the cotangents ``g[n_t, n_s]`` (reference :311-344).  Where
    the reference re-runs ``solve_forward`` and then ``solve_backward`` as two loops of CVODES
    calls, this is ONE library call (``sb_solve_adjoint``)."""
    if HAVE_PYTENSOR:  # pragma: no cover
        itypes = [pt.dvector, pt.dvector, pt.dvector, pt.dmatrix, pt.dscalar, pt.dvector]
        otypes = [pt.dvector, pt.dvector]

    def perform(self, node, inputs, outputs):
        y0, params, params_fixed, grads, t0, tvals = inputs
        self._set_params(params, params_fixed)
        tvals = np.asarray(tvals, dtype=np.float64)
        y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(1, -1)
        grads = np.ascontiguousarray(grads, dtype=np.float64)[None]
        _, grad_out, lamda_out, status = self._solver.solve_adjoint_batch(
            float(t0), tvals, y0, None, grads)
        if status[0] != 0:                       # as_pytensor.py:339-341
            grad_out[:] = np.nan
            lamda_out[:] = np.nan
        outputs[0][0] = lamda_out[0]
        outputs[1][0] = grad_out[0]


class SolveODEAdjointBatch(_SolverOp):
    """Batched forward solve: ``y0[B, n_s]``, ``params[B, n_deriv]`` (derivative parameters per
    draw), ``params_fixed[n_fixed]`` shared, ``t0``, ``tvals`` -> ``y[B, n_t, n_s]``.  Failed draws
    are NaN rows, as the batch-1 Op does for its single draw."""
    if HAVE_PYTENSOR:  # pragma: no cover
        itypes = [pt.dmatrix, pt.dmatrix, pt.dvector, pt.dscalar, pt.dvector]
        otypes = [pt.dtensor3]

    def _full_params(self, params, params_fixed) -> np.ndarray:
        subset = self._solver._problem.params_subset
        B = params.shape[0]
        full = np.empty((B, subset.n_items))
        full[:, subset.subset_flat_index] = params
        full[:, subset.remainder_flat_index] = np.asarray(params_fixed)[None, :]
        return full

    def perform(self, node, inputs, outputs):
        y0, params, params_fixed, t0, tvals = inputs
        y, _ = self._solver.solve_forward_batch(
            float(t0), np.asarray(tvals, dtype=np.float64),
            np.ascontiguousarray(y0, dtype=np.float64), self._full_params(params, params_fixed))
        outputs[0][0] = y

    def grad(self, inputs, g):  # pragma: no cover - needs pytensor
        g, = g
        y0, params, params_fixed, t0, tvals = inputs
        lamda, gradient = SolveODEAdjointBackwardBatch(self._solver)(
            y0, params, params_fixed, g, t0, tvals)
        return [
            -lamda,
            gradient,
            grad_not_implemented(self, 2, params_fixed),
            grad_not_implemented(self, 3, t0),
            grad_not_implemented(self, 4, tvals),
        ]


class SolveODEAdjointBackwardBatch(SolveODEAdjointBatch):
    """(lamda(t0)[B, n_s], dL/dparams[B, n_deriv]) for a batch of draws and cotangents
    ``g[B, n_t, n_s]``."""
    if HAVE_PYTENSOR:  # pragma: no cover
        itypes = [pt.dmatrix, pt.dmatrix, pt.dvector, pt.dtensor3, pt.dscalar, pt.dvector]
        otypes = [pt.dmatrix, pt.dmatrix]

    def perform(self, node, inputs, outputs):
        y0, params, params_fixed, grads, t0, tvals = inputs
        _, grad_out, lamda_out, status = self._solver.solve_adjoint_batch(
            float(t0), np.asarray(tvals, dtype=np.float64),
            np.ascontiguousarray(y0, dtype=np.float64), self._full_params(params, params_fixed),
            np.ascontiguousarray(grads, dtype=np.float64))
        outputs[0][0] = lamda_out
        outputs[1][0] = grad_out

    def grad(self, inputs, g):  # pragma: no cover
        raise NotImplementedError('second derivatives are not available')
