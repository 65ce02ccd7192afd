"""Shared constants, return codes and error helpers.

Counterpart of the pieces of the reference's ``sunode/basic.py`` that survive without SUNDIALS:
``data_dtype``/``index_dtype`` (basic.py:40-43), the ``ERRORS`` code->name map (basic.py:49-55)
and ``check`` (basic.py:84-91).  The numeric codes are the CVODES ones the reference exposes
(include/cvodes/16_cvodes.h:45-106) so that per-instance ``status`` values written by the CUDA
kernels read the same as the flags the reference raises on.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

data_dtype = np.dtype(np.float64)
index_dtype = np.dtype(np.int64)

CV_SUCCESS = 0
CV_TSTOP_RETURN = 1
CV_ROOT_RETURN = 2
CV_WARNING = 99
CV_TOO_MUCH_WORK = -1
CV_TOO_MUCH_ACC = -2
CV_ERR_FAILURE = -3
CV_CONV_FAILURE = -4
CV_LINIT_FAIL = -5
CV_LSETUP_FAIL = -6
CV_LSOLVE_FAIL = -7
CV_RHSFUNC_FAIL = -8
CV_FIRST_RHSFUNC_ERR = -9
CV_REPTD_RHSFUNC_ERR = -10
CV_UNREC_RHSFUNC_ERR = -11
CV_CONSTR_FAIL = -15
CV_MEM_FAIL = -20
CV_MEM_NULL = -21
CV_ILL_INPUT = -22
CV_NO_MALLOC = -23
CV_BAD_K = -24
CV_BAD_T = -25
CV_BAD_DKY = -26
CV_TOO_CLOSE = -27
CV_QRHSFUNC_FAIL = -31
CV_FIRST_QRHSFUNC_ERR = -32
CV_REPTD_QRHSFUNC_ERR = -33
CV_UNREC_QRHSFUNC_ERR = -34
CV_NO_ADJ = -101
CV_NO_FWD = -102
CV_NO_BCK = -103
CV_BAD_TB0 = -104
CV_REIFWD_FAIL = -105
CV_FWD_FAIL = -106
CV_GETY_BADT = -107

ERRORS: Dict[int, str] = {
    value: name for name, value in list(globals().items())
    if name.startswith('CV_') and isinstance(value, int)
}


class SolverError(RuntimeError):
    """Raised by the batch-1 solve calls when an instance fails (reference solver.py:21)."""


def check(retcode: int) -> None:
    """Raise on a non-zero library return code, in the reference's wording (basic.py:84-91)."""
    if isinstance(retcode, (int, np.integer)) and retcode != 0:
        name = ERRORS.get(int(retcode), 'UNKNOWN')
        raise ValueError('Bad return code from sundials: %s (%s)' % (name, int(retcode)))
