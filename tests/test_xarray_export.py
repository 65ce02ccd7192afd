"""``Solver.as_xarray`` / ``Problem.solution_to_xarray`` against what the reference's own export
(sunode/problem.py:100-145) produces on the same inputs.  xarray is not in the image: both sides
run on a recording stand-in (tests/golden/xr_stub.py) that keeps name -> (dims, values); the
golden was recorded by running the reference (tests/golden/make_xarray_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest

from sunode_b200 import SympyProblem
from tests.golden import xr_stub

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, 'golden', 'xarray_golden.json')) as fh:
    GOLD = json.load(fh)


@pytest.fixture
def stub_xarray():
    saved = sys.modules.get('xarray')
    xr_stub.install()
    yield
    if saved is None:
        sys.modules.pop('xarray', None)
    else:
        sys.modules['xarray'] = saved


@pytest.mark.parametrize('case', [c[0] for c in xr_stub.cases()])
def test_xarray_export_matches_reference(stub_xarray, case):
    name, params, states, rhs, deriv, coords = next(c for c in xr_stub.cases() if c[0] == case)
    prob = SympyProblem(params, states, rhs, deriv, coords=coords)
    tvals, sol, p = xr_stub.inputs(prob.n_states, prob.n_params_total)
    ud = prob.make_user_data()
    ud.params = p.view(prob.params_dtype)[0]
    for us in (True, False):
        for up in (True, False):
            ds = prob.solution_to_xarray(tvals, sol.copy(), ud, unstack_state=us, unstack_params=up)
            got, want = ds.summary(), GOLD['%s/%d%d' % (name, us, up)]
            assert got['coords'] == want['coords']
            assert sorted(got['vars']) == sorted(want['vars'])
            for var, w in want['vars'].items():
                g = got['vars'][var]
                assert g['dims'] == w['dims'] and g['shape'] == w['shape'], var
                np.testing.assert_array_equal(g['values'], w['values'])
