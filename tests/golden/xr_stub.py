"""A recording stand-in for ``xarray`` (TEST INFRASTRUCTURE): ``Dataset`` keeps what is assigned to
it -- name -> (dims, values) -- so that the xarray export of the reference
(sunode/problem.py:100-145) and ours can be compared without the real package, which is not in
the image.  Used by tests/golden/make_xarray_golden.py (run on the reference) and
tests/test_xarray_export.py (run on sunode_b200)."""
import sys
import types

import numpy as np


class DataArray:            # only used in isinstance checks by the reference
    pass


class Dataset:
    def __init__(self, coords=None):
        self.coords = dict(coords or {})
        self.vars = {}

    def __contains__(self, name):
        return name in self.vars or name in self.coords

    def __setitem__(self, name, value):
        if isinstance(value, tuple):
            dims, vals = value
        else:                                   # data['parameters'] = params (a structured scalar)
            dims, vals = (), value
        if isinstance(dims, str):
            dims = (dims,)
        self.vars[name] = (tuple(dims), np.asarray(vals))

    def summary(self):
        """JSON-able description: coords, and per variable dims / dtype / shape / flat float values."""
        out = {'coords': {k: [str(x) for x in np.asarray(v).ravel()] for k, v in self.coords.items()}, 'vars': {}}
        for name, (dims, vals) in self.vars.items():
            flat = np.ascontiguousarray(vals)
            if flat.dtype.fields is not None:
                flat = flat.reshape(-1).view(np.float64) if flat.dtype.itemsize else np.zeros(0)
            out['vars'][name] = {'dims': list(dims), 'shape': list(vals.shape), 'dtype': str(vals.dtype),
                                 'values': [float(x) for x in np.asarray(flat, dtype=np.float64).ravel()]}
        return out


def install():
    mod = types.ModuleType('xarray')
    mod.Dataset, mod.DataArray = Dataset, DataArray
    sys.modules['xarray'] = mod
    return mod


def cases():
    """(name, params, states, rhs, derivative_params, coords)"""
    def nested(t, y, p):
        return {'a': p.c.d * y.a, 'b': {'c': [3. * y.b.c[1], 4. * y.a[0]]}}

    def named(t, y, p):
        return {'x': -p.k * y.x, 'total': y.x[0] + y.x[1] + y.x[2]}
    return [
        ('nested', {'c': {'d': 3}, 'f': 4}, {'a': 3, 'b': {'c': 2}}, nested, [('c', 'd')], None),
        ('named', {'k': 'species', 'unused': ()}, {'x': 'species', 'total': ()}, named, [('k',)],
         {'species': ['fox', 'hare', 'lynx']}),
    ]


def inputs(n_states, n_params_total, seed=7):
    rng = np.random.default_rng(seed)
    tvals = np.linspace(0, 1, 4)
    return tvals, rng.standard_normal((len(tvals), n_states)), rng.standard_normal(n_params_total)
