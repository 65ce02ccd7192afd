"""Problem protocol shared by the solvers.

Counterpart of the reference's ``sunode/problem.py``: the parts that describe *data* are kept
(``n_states``/``n_params`` at problem.py:92-98, ``make_user_data`` :51-52, the parameter
scatter/extract helpers :54-90, ``flat_solution_as_dict`` :147-154, ``solution_to_xarray``
:100-145).  The ``make_sundials_*`` trampolines (:156-494) have no counterpart: there are no
function pointers on the device — the generated ``__device__`` functions are compiled into the
integrator kernels (see ``symode/codegen.py`` and ``csrc/sb_kernels.cuh``).
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from .dtypesubset import DTypeSubset, as_nested


class Problem:
    params_dtype: np.dtype
    params_subset: DTypeSubset
    state_dtype: np.dtype
    state_subset: DTypeSubset
    user_data_dtype: np.dtype
    coords: Dict[str, Any]

    # ------------------------------------------------------------------ sizes
    @property
    def n_states(self) -> int:
        return self.state_subset.n_items

    @property
    def n_params(self) -> int:
        """Number of *derivative* parameters (the reference's meaning, problem.py:96-98)."""
        return self.params_subset.n_subset

    @property
    def n_params_total(self) -> int:
        """Number of all scalar parameters (length of the flat parameter vector)."""
        return self.params_subset.n_items

    # ------------------------------------------------------------------ user data
    def make_user_data(self) -> np.ndarray:
        return np.zeros((), dtype=self.user_data_dtype).view(np.recarray)

    def update_params(self, user_data: np.ndarray, params: np.ndarray) -> None:
        user_data.params.fill(params)

    def update_subset_params(self, user_data: np.ndarray, params: np.ndarray) -> None:
        view = user_data.params.view(self.params_subset.subset_view_dtype)
        view.fill(params)

    def update_remaining_params(self, user_data: np.ndarray, params: np.ndarray) -> None:
        view = user_data.params.view(self.params_subset.remainder.subset_view_dtype)
        view.fill(params)

    def extract_params(self, user_data: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.full((1,), np.nan, dtype=self.params_dtype)[0]
        out.fill(user_data.params)
        return out

    def flat_params(self, user_data: np.ndarray) -> np.ndarray:
        """All parameters as one contiguous float64 vector, in declaration order."""
        if self.params_dtype.itemsize == 0:
            return np.zeros(0, dtype=np.float64)
        return np.array(user_data.params).reshape(1).view(np.float64).copy()

    # ------------------------------------------------------------------ presentation
    def flat_solution_as_dict(self, solution: Any) -> Dict[str, Any]:
        """Split ``solution[..., n_states]`` into the nested state dict (problem.py:147-154)."""
        slices = self.state_subset.flat_slices
        shapes = self.state_subset.flat_shapes
        views = {}
        for path in self.state_subset.paths:
            views[path] = solution[:, slices[path]].reshape((-1,) + shapes[path])
        return as_nested(views)

    def solution_to_xarray(self, tvals, solution, user_data, sensitivity=None,
                           *, unstack_state=True, unstack_params=True):
        """xarray export (problem.py:100-145).  xarray is imported lazily; it is not part of
        the compute path and is absent from the build image."""
        import xarray as xr

        assert sensitivity is None, 'TODO'
        solution = np.ascontiguousarray(solution).view(self.state_dtype)[..., 0]
        params = self.extract_params(user_data)

        def leaves(array, dims, prefix):
            out = {}
            for name in array.dtype.names:
                if array[name].dtype.fields is None:
                    out['_'.join(prefix + [name])] = (tuple(dims[name][1]), array[name])
                else:
                    out.update(leaves(array[name], dims[name], prefix + [name]))
            return out

        data = xr.Dataset(coords=self.coords)
        data['time'] = ('time', tvals)
        if unstack_state:
            for name, (dims, vals) in leaves(solution, self.state_subset.dims, ['solution']).items():
                if name in data:
                    raise ValueError(f"Variable {name} is not unique.")
                data[name] = (('time',) + dims, vals)
        else:
            data['solution'] = ('time', solution)
        if unstack_params:
            for name, (dims, vals) in leaves(params, self.params_subset.dims, ['parameters']).items():
                if name in data:
                    raise ValueError(f"Variable {name} is not unique.")
                data[name] = (dims, vals)
        else:
            data['parameters'] = params
        return data
