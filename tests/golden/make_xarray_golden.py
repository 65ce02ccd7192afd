"""Generates tests/golden/xarray_golden.json by running the REFERENCE's xarray export
(/root/reference/sunode/problem.py:100-145, reached through Solver.as_xarray, solver.py:428-433)
on a recording stand-in for xarray (tests/golden/xr_stub.py).

    python tests/golden/make_xarray_golden.py     # needs /root/reference; not run by the tests
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.golden import xr_stub                              # noqa: E402
xr_stub.install()
from tests.golden.make_codegen_golden import load_reference   # noqa: E402


def main():
    RefProblem = load_reference()
    out = {}
    for name, params, states, rhs, deriv, coords in xr_stub.cases():
        prob = RefProblem(params, states, rhs, deriv, coords=coords)
        n_all = prob.params_dtype.itemsize // 8
        tvals, sol, p = xr_stub.inputs(prob.n_states, n_all)
        ud = prob.make_user_data()
        ud.params = p.view(prob.params_dtype)[0]
        for us in (True, False):
            for up in (True, False):
                ds = prob.solution_to_xarray(tvals, sol.copy(), ud, unstack_state=us, unstack_params=up)
                out['%s/%d%d' % (name, us, up)] = ds.summary()
    with open(os.path.join(HERE, 'xarray_golden.json'), 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print('wrote xarray_golden.json', sorted(out))


if __name__ == '__main__':
    main()
