"""``SympyProblem``: an ODE right-hand side written once with sympy symbols.

Same constructor, attributes and error behaviour as the reference's
``sunode/symode/problem.py`` (class at :24-158, ``_make_dydt`` at :160-230): the user function
``rhs(t, y, p)`` is called once with attribute trees of sympy symbols (states ``positive=True``,
parameters ``real=True``, :78-79) and must return a (nested) dict with one entry per state.
From it the Jacobian (:142), the parameter derivative for the *derivative* parameters (:144),
the adjoint right-hand side ``-lamda^T J`` (:147) and the quadrature integrand
``lamda^T df/dp`` (:148) are derived symbolically.

The back half differs by design: instead of numba functions with SUNDIALS signatures the
problem yields one :class:`~sunode_b200.symode.codegen.GeneratedSource` whose CUDA flavour is
compiled into the sm_100a integrator kernels and whose C flavour backs the Python-callable
``make_rhs()/make_jac_dense()/...`` evaluators (same call signatures as the reference's, e.g.
``rhs(out, t, y, user_data) -> int`` with 1 meaning "non-finite output", :262-270).
"""
from __future__ import annotations

from itertools import product
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np
import sympy as sym

from .. import basic, dtypesubset
from ..problem import Problem
from . import codegen
from .hostfuncs import HostFunctions

Path = Tuple[str, ...]
Shape = Tuple[int, ...]


def _scalar_item(item: np.ndarray) -> Any:
    if hasattr(item, 'shape') and item.shape == ():
        return item.item()
    return item


class SympyProblem(Problem):
    def __init__(
        self,
        params: Dict[str, Any],
        states: Dict[str, Any],
        rhs_sympy: Callable[[sym.Symbol, Any, Any], Dict[str, Any]],
        derivative_params: List[Path],
        coords: Optional[Dict[str, Any]] = None,
        simplify: Optional[Callable[[sym.Expr], sym.Expr]] = None,
    ):
        derivative_params = [tuple(p) for p in derivative_params]
        self.params_subset = dtypesubset.DTypeSubset(
            params, derivative_params, fixed_dtype=basic.data_dtype, coords=coords)
        self.coords = self.params_subset.coords
        self.params_dtype = self.params_subset.dtype
        self.state_subset = dtypesubset.DTypeSubset(
            states, [], fixed_dtype=basic.data_dtype, coords=self.coords)
        self.state_dtype = self.state_subset.dtype
        self._rhs_sympy_func = rhs_sympy
        self._simplify_func = simplify
        self._simplify = np.vectorize(simplify if simplify is not None else (lambda x: x),
                                      otypes=[object])

        unknown = set(derivative_params) - set(self.params_subset.paths)
        if unknown:
            raise ValueError('Unknown derivative parameters: %s' % sorted(unknown))
        self._check_subset_dtype(self.params_subset.subset_dtype)

        # ---- symbols, in flat (declaration) order -------------------------------------
        self._sym_time = sym.Symbol('time', real=True)

        def make_symbols(shapes: Dict[Path, Shape], **assume: Any) -> Dict[Path, np.ndarray]:
            return {path: sym.symarray('_'.join(path), shape, **assume)
                    for path, shape in shapes.items()}

        state_syms = make_symbols(self.state_subset.flat_shapes, positive=True)
        param_syms = make_symbols(self.params_subset.flat_shapes, real=True)

        # name -> access path, as in the reference (:81-95); kept for introspection/debugging
        self._varmap: Dict[str, Tuple[Any, ...]] = {}
        for kind, table in (('state', state_syms), ('params', param_syms)):
            for path, arr in table.items():
                for idxs in product(*[range(n) for n in arr.shape]):
                    entry: Tuple[Any, ...] = (kind,) + path
                    if idxs != ():
                        entry = entry + (idxs,)
                    self._varmap[arr[idxs].name] = entry

        subset_paths = set(self.params_subset.subset_paths)
        deriv_syms = [arr.ravel() for path, arr in param_syms.items() if path in subset_paths]
        fixed_syms = [arr.ravel() for path, arr in param_syms.items() if path not in subset_paths]
        empty = np.zeros((0,), dtype=object)
        self._sym_deriv_paramsvec = np.concatenate(deriv_syms) if deriv_syms else empty
        self._sym_fixed_paramsvec = np.concatenate(fixed_syms) if fixed_syms else empty
        all_params = [arr.ravel() for arr in param_syms.values()]
        self._sym_paramsvec = np.concatenate(all_params) if all_params else empty
        self._sym_statevec = np.concatenate([arr.ravel() for arr in state_syms.values()])

        self._sym_params = self.params_subset.as_dataclass(
            'Params', self._sym_deriv_paramsvec, self._sym_fixed_paramsvec, item_map=_scalar_item)
        self._sym_states = self.state_subset.as_dataclass(
            'State', [], self._sym_statevec, item_map=_scalar_item)

        # ---- the user's right-hand side and everything derived from it ----------------
        dydt = self._make_dydt()
        self._sym_dydt = np.array(dydt).ravel()

        n_s, n_d = self.n_states, self.n_params
        self._sym_sens = sym.symarray('sens', (n_d, n_s))
        self._sym_lamda = sym.symarray('lamda', n_s)
        for idxs in product(*[range(n) for n in self._sym_lamda.shape]):
            self._varmap[self._sym_lamda[idxs].name] = ('lamda', idxs)
        for idxs in product(*[range(n) for n in self._sym_sens.shape]):
            self._varmap[self._sym_sens[idxs].name] = ('sens', idxs)

        self._sym_dydt_jac = np.array(dydt.jacobian(list(self._sym_statevec))).reshape(n_s, n_s)
        if n_d:
            self._sym_dydp = np.array(
                dydt.jacobian(list(self._sym_deriv_paramsvec))).reshape(n_s, n_d)
        else:
            self._sym_dydp = np.zeros((n_s, 0), dtype=object)
        self._sym_dlamdadt = -self._sym_lamda @ self._sym_dydt_jac
        self._sym_quad_rhs = self._sym_lamda @ self._sym_dydp

        self.user_data_dtype = np.dtype([
            ('params', self.params_subset.dtype),
            ('tmp_nstates_nstates', np.float64, (n_s, n_s)),
            ('tmp_nparams_nstates', np.float64, (n_d, n_s)),
            ('tmp2_nparams_nstates', np.float64, (n_d, n_s)),
            ('error_states', self.state_dtype),
            ('error_rhs', np.float64, (n_s,)),
            ('error_jac', np.float64, (n_s, n_s)),
        ])

        self._generated: Optional[codegen.GeneratedSource] = None
        self._host: Optional[HostFunctions] = None

    # ------------------------------------------------------------------ construction helpers
    @staticmethod
    def _check_subset_dtype(dtype: np.dtype, path: Optional[str] = None) -> None:
        if dtype.fields is None:
            if dtype.base != basic.data_dtype:
                raise ValueError('Derivative param %s has incorrect dtype %s. Should be %s'
                                 % (path, dtype.base, basic.data_dtype))
            return
        for name, (sub, *_rest) in dtype.fields.items():
            SympyProblem._check_subset_dtype(sub, name if path is None else path + '.' + name)

    def _make_dydt(self) -> sym.Matrix:
        rhs = self._rhs_sympy_func(self._sym_time, self._sym_states, self._sym_params)
        if not isinstance(rhs, dict):
            raise ValueError('The right-hand-side function must return a dict of states.')
        self._state_leaf_dims = {
            path: names for path, (_, names) in
            dtypesubset.as_flattened(self.state_subset.dims).items()}
        rhs = self._clone_rhs(rhs)

        flat: List[Any] = []
        for path in self.state_subset.paths:
            node = rhs
            for name in path[:-1]:
                if not isinstance(node, dict) or name not in node:
                    raise ValueError('No right-hand-side for state %s' % '.'.join(path))
                node = node[name]
            if not isinstance(node, dict) or path[-1] not in node:
                raise ValueError('No right-hand-side for state %s' % '.'.join(path))
            value = node.pop(path[-1])
            shape = self.state_subset.flat_shapes[path]
            dims = tuple(self._state_leaf_dims[path])
            flat.extend(self._flatten_value('.'.join(path), value, shape, dims))

        leftover = dtypesubset.as_flattened(rhs)
        if leftover:
            raise ValueError('Unknown state variables: %s' % ['.'.join(p) for p in leftover])
        return sym.Matrix(flat) if flat else sym.Matrix(0, 1, [])

    def _clone_rhs(self, rhs: Dict[str, Any], prefix: Path = ()) -> Dict[str, Any]:
        """Copy the dict structure down to (but excluding) the state leaves."""
        out: Dict[str, Any] = {}
        for key, val in rhs.items():
            path = prefix + (key,)
            is_branch = any(len(p) > len(path) and p[:len(path)] == path
                            for p in self.state_subset.paths)
            if isinstance(val, dict) and is_branch:
                out[key] = self._clone_rhs(val, path)
            else:
                out[key] = val
        return out

    def _flatten_value(self, name: str, value: Any, shape: Shape, dims: Tuple[str, ...]) -> List[Any]:
        """Accept sympy/numpy arrays, nested lists, or dicts keyed by coordinate labels
        (reference :165-206)."""
        total = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if hasattr(value, 'shape') and not isinstance(value, sym.Expr):
            if tuple(value.shape) != tuple(shape):
                raise ValueError('Invalid shape for right-hand-side state %s. It is %s but we '
                                 'expected %s.' % (name, tuple(value.shape), shape))
            if hasattr(value, 'dims') and tuple(value.dims) != tuple(dims):
                raise ValueError('Invalid dims for right-hand-side state %s.' % name)
            if isinstance(value, sym.NDimArray):
                return list(value.reshape(total)) if total else []
            if hasattr(value, 'data') and hasattr(value, 'dims'):  # xarray.DataArray
                return list(np.asarray(value.data, dtype=object).ravel())
            return list(np.asarray(value, dtype=object).reshape((total,)))
        if isinstance(value, (list, tuple)):
            if not shape or len(value) != shape[0]:
                raise ValueError('Invalid shape for right-hand-side state %s.' % name)
            out: List[Any] = []
            for item in value:
                out.extend(self._flatten_value(name, item, shape[1:], dims[1:]))
            return out
        if isinstance(value, dict):
            if not shape or len(value) != shape[0]:
                raise ValueError('Invalid shape for right-hand-side state %s.' % name)
            out = []
            for label in self.coords[dims[0]]:
                out.extend(self._flatten_value(name, value[label], shape[1:], dims[1:]))
            return out
        if shape == ():
            return [value]
        raise ValueError('Unknown righ-hand-side for state %s.' % name)

    # ------------------------------------------------------------------ generated code
    @property
    def generated(self) -> codegen.GeneratedSource:
        """CUDA / C source of all problem functions (generated once, lazily)."""
        if self._generated is None:
            deriv_index = self.params_subset.subset_flat_index
            simp = self._simplify
            self._generated = codegen.generate(
                time=self._sym_time,
                states=list(self._sym_statevec),
                params=list(self._sym_paramsvec),
                lamda=list(self._sym_lamda),
                sens=self._sym_sens,
                deriv_index=[int(i) for i in deriv_index],
                dydt=list(simp(np.array(self._sym_dydt, dtype=object))),
                jac=simp(self._sym_dydt_jac),
                dydp=simp(self._sym_dydp) if self._sym_dydp.size else self._sym_dydp,
                dlamdadt=list(simp(np.array(self._sym_dlamdadt, dtype=object))),
                quad_rhs=list(simp(np.array(self._sym_quad_rhs, dtype=object)))
                if self._sym_quad_rhs.size else [],
            )
        return self._generated

    @property
    def host_functions(self) -> HostFunctions:
        if self._host is None:
            self._host = HostFunctions(self.generated)
        return self._host

    # pickling: sympy objects pickle fine, the ctypes handle does not
    def __getstate__(self):
        _ = self.generated          # the generated source travels; the user's callables need not
        state = dict(self.__dict__)
        state['_host'] = None
        state['_simplify'] = None
        state['_rhs_sympy_func'] = None
        state['_simplify_func'] = None
        # attribute trees of symbols: dynamically created dataclasses, only needed while the
        # user's rhs is being traced in __init__
        state['_sym_params'] = None
        state['_sym_states'] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        f = self._simplify_func
        self._simplify = np.vectorize(f if f is not None else (lambda x: x), otypes=[object])

    # ------------------------------------------------------------------ python-callable evaluators
    def _params_of(self, user_data) -> np.ndarray:
        if self.params_dtype.itemsize == 0:
            return np.zeros(0)
        return np.array(user_data['params']).reshape(1).view(np.float64)

    def _state_of(self, y) -> np.ndarray:
        y = np.asarray(y)
        if y.dtype == self.state_dtype:
            y = y.reshape(1).view(np.float64)
        return np.ascontiguousarray(y, dtype=np.float64).ravel()

    def make_rhs(self, *, debug=False):
        """``rhs(out, t, y, user_data) -> int`` (reference :251-282)."""
        host = self.host_functions
        if debug:
            print(self.generated.c)

        def rhs(out, t, y, user_data):
            flag = host.rhs(float(t), self._state_of(y), self._params_of(user_data), out)
            if flag:
                user_data['error_rhs'][...] = out
                user_data['error_states'] = np.asarray(self._state_of(y)).view(self.state_dtype)[0]
            return flag
        return rhs

    def make_jac_dense(self, *, debug=False):
        """``jac(out, t, y, fy, user_data) -> int``; ``out[i, j] = df_i/dy_j`` (:342-371)."""
        host = self.host_functions

        def jac_dense(out, t, y, fy, user_data):
            tmp = np.empty(self.n_states * self.n_states)
            flag = host.jac(float(t), self._state_of(y), self._params_of(user_data), tmp)
            out[...] = tmp.reshape(self.n_states, self.n_states).T  # column-major -> [i, j]
            if flag:
                user_data['error_jac'][...] = out
            return flag
        return jac_dense

    def make_adjoint_rhs(self, *, debug=False):
        """``adj(out, t, y, lamda, user_data) -> int`` computing ``-J^T lamda`` (:284-311)."""
        host = self.host_functions

        def adjoint(out, t, y, lamda, user_data):
            return host.adj_rhs(float(t), self._state_of(y),
                                np.ascontiguousarray(lamda, dtype=np.float64),
                                self._params_of(user_data), out)
        return adjoint

    def make_adjoint_quad_rhs(self, *, debug=False):
        """``quad(out, t, y, lamda, user_data) -> int``: ``lamda^T df/dp`` (:313-340)."""
        host = self.host_functions

        def quad_rhs(out, t, y, lamda, user_data):
            return host.quad_rhs(float(t), self._state_of(y),
                                 np.ascontiguousarray(lamda, dtype=np.float64),
                                 self._params_of(user_data), out)
        return quad_rhs

    def make_adjoint_jac_dense(self, *, debug=False):
        """``jacB(out, t, y, yB, fyB, user_data) -> int``; ``out = -J^T`` (:406-433)."""
        host = self.host_functions

        def jac_dense(out, t, y, yB, fyB, user_data):
            tmp = np.empty(self.n_states * self.n_states)
            flag = host.adj_jac(float(t), self._state_of(y), self._params_of(user_data), tmp)
            out[...] = tmp.reshape(self.n_states, self.n_states).T
            return flag
        return jac_dense

    def make_rhs_jac_prod(self, *, debug=False):
        """``jac_prod(out, v, t, y, fy, user_data) -> int`` computing ``J v`` (:373-403)."""
        jac = self.make_jac_dense()

        def jac_prod(out, v, t, y, fy, user_data):
            J = np.empty((self.n_states, self.n_states))
            flag = jac(J, t, y, fy, user_data)
            out[...] = J @ np.asarray(v, dtype=np.float64)
            return int(flag or not np.isfinite(out).all())
        return jac_prod

    def make_adjoint_jac_prod(self, *, debug=False):
        """``jac_prod(out, vB, t, y, yB, fyB, user_data)`` computing ``-J^T vB`` (:435-465)."""
        jac = self.make_jac_dense()

        def jac_prod(out, vB, t, y, yB, fyB, user_data):
            J = np.empty((self.n_states, self.n_states))
            flag = jac(J, t, y, None, user_data)
            out[...] = -(J.T @ np.asarray(vB, dtype=np.float64))
            return int(flag or not np.isfinite(out).all())
        return jac_prod

    def make_sensitivity_rhs(self, *, debug=False):
        """``sens(out, t, y, yS, user_data) -> int`` with ``out[k] = J yS[k] + df/dp_k``
        (:557-583); ``yS``/``out`` have shape ``(n_params, n_states)``."""
        host = self.host_functions

        def sens_rhs(out, t, y, yS, user_data):
            tmp = np.empty(self.n_params * self.n_states)
            flag = host.sens_rhs(float(t), self._state_of(y),
                                 np.ascontiguousarray(yS, dtype=np.float64).ravel(),
                                 self._params_of(user_data), tmp)
            out[...] = tmp.reshape(self.n_params, self.n_states)
            return flag
        return sens_rhs
