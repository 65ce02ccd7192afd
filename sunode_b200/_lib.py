"""ctypes binding of ``libsunode_b200.so`` (``include/sunode_b200.h``).

This module takes the place of the reference's ``sunode/basic.py:29-30`` (``lib``/``ffi`` of the
cffi extension ``_sundials_cvodes``).  The library is looked up in-tree; if it is missing it is
built once from ``csrc/`` (host C++ only, see ``_build.py``).  There is no fallback: if neither
works the import fails, and every call that needs a device raises :class:`DeviceError` when no
CUDA driver / GPU is present.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

from . import _build

SB_MEM_HOST = 0
SB_MEM_DEVICE = 1
SB_OK = 0
SB_ERR_CUDA = -1001
SB_ERR_NVRTC = -1002
SB_ERR_ARG = -1003
SB_ERR_STATE = -1004
SB_STATS_PER_INSTANCE = 8

_DP = ctypes.POINTER(ctypes.c_double)
_IP = ctypes.POINTER(ctypes.c_int32)
_VP = ctypes.c_void_p


class LibraryError(RuntimeError):
    """A call into libsunode_b200 failed (message from ``sb_last_error``)."""


class DeviceError(LibraryError):
    """No usable CUDA device / driver for a call that needs one."""


_lib: Optional[ctypes.CDLL] = None


def _declare(lib: ctypes.CDLL) -> None:
    c_int, c_i64, c_double, c_size = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t
    sig = {
        'sb_version': (c_int, []),
        'sb_last_error': (ctypes.c_char_p, []),
        'sb_device_count': (c_int, [ctypes.POINTER(c_int)]),
        'sb_nvrtc_version': (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
        'sb_compile': (c_int, [ctypes.c_char_p, ctypes.c_char_p, c_int, c_int,
                               ctypes.POINTER(_VP), ctypes.POINTER(c_size),
                               ctypes.POINTER(ctypes.c_void_p)]),
        'sb_free': (None, [_VP]),
        'sb_problem_create': (c_int, [ctypes.POINTER(_VP), c_int, c_int, c_int, _VP, c_size, c_int]),
        'sb_problem_destroy': (c_int, [_VP]),
        'sb_set_tolerances': (c_int, [_VP, c_double, _DP, c_int]),
        'sb_set_tolerances_b': (c_int, [_VP, c_double, c_double]),
        'sb_set_sens_scaling': (c_int, [_VP, ctypes.POINTER(c_double), c_int]),
        'sb_set_quad_tolerances_b': (c_int, [_VP, c_double, c_double]),
        'sb_set_max_num_steps': (c_int, [_VP, c_int, c_int]),
        'sb_set_max_num_steps_b': (c_int, [_VP, c_int, c_int]),
        'sb_set_history_capacity': (c_int, [_VP, c_int]),
        'sb_set_backward_trace': (c_int, [_VP, _VP, _VP]),
        'sb_solve_forward': (c_int, [_VP, c_i64, c_double, _VP, c_int, _VP, _VP, _VP, _VP, _VP,
                                     c_int, c_int, _VP]),
        'sb_solve_forward_sens': (c_int, [_VP, c_i64, c_double, _VP, c_int, _VP, _VP, _VP, c_int,
                                          _VP, _VP, _VP, _VP, c_int, _VP]),
        'sb_solve_backward': (c_int, [_VP, c_i64, c_double, c_double, _VP, c_int, _VP, _VP, c_int,
                                      _VP, _VP, _VP, _VP, c_int, _VP]),
        'sb_solve_adjoint': (c_int, [_VP, c_i64, c_double, _VP, c_int, _VP, _VP, _VP, c_int,
                                     _VP, _VP, _VP, _VP, _VP, _VP, c_int, _VP]),
        'sb_set_workspace_limit': (c_int, [_VP, c_size]),
        'sb_last_chunks': (c_int, [_VP]),
        'sb_forward_fail_index': (c_int, [_VP, c_i64, _VP]),
        'sb_eval': (c_int, [_VP, c_int, c_i64, _VP, _VP, _VP, c_int, _VP, _VP, c_int, _VP]),
        'sb_synchronize': (c_int, [_VP]),
        'sb_last_kernel_ms': (c_int, [_VP] + [ctypes.POINTER(ctypes.c_float)] * 3),
        'sb_launch_count': (c_i64, [_VP]),
        'sb_kernel_info': (c_int, [_VP] + [ctypes.POINTER(c_int)] * 6),
        'sb_host_alloc': (c_int, [ctypes.POINTER(_VP), c_size]),
        'sb_host_free': (c_int, [_VP]),
    }
    for name, (restype, argtypes) in sig.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes


EXPORTS = (
    'sb_version', 'sb_last_error', 'sb_device_count', 'sb_nvrtc_version', 'sb_compile', 'sb_free',
    'sb_problem_create', 'sb_problem_destroy', 'sb_set_tolerances', 'sb_set_tolerances_b',
    'sb_set_sens_scaling', 'sb_set_quad_tolerances_b', 'sb_set_max_num_steps', 'sb_set_max_num_steps_b',
    'sb_set_history_capacity', 'sb_set_backward_trace', 'sb_solve_forward', 'sb_solve_forward_sens', 'sb_solve_backward',
    'sb_solve_adjoint', 'sb_set_workspace_limit', 'sb_last_chunks', 'sb_forward_fail_index',
    'sb_eval', 'sb_synchronize', 'sb_last_kernel_ms', 'sb_launch_count', 'sb_kernel_info',
    'sb_host_alloc', 'sb_host_free',
)


def lib() -> ctypes.CDLL:
    """The loaded library (built on first use if the in-tree .so is absent or stale)."""
    global _lib
    if _lib is None:
        path = _build.LIB_PATH
        if _build.needs_build():
            try:
                path = _build.build_library()
            except Exception as err:  # noqa: BLE001
                if not os.path.exists(path):
                    raise ImportError(
                        'libsunode_b200.so is missing and could not be built: %s' % err) from err
                # a stale library embeds OLD kernel sources: say so instead of running them silently
                import warnings
                warnings.warn('libsunode_b200.so is older than its sources and the rebuild failed '
                              '(%s); the stale library is used' % err, RuntimeWarning, stacklevel=2)
        handle = ctypes.CDLL(path)
        _declare(handle)
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().sb_last_error()
    return msg.decode('utf-8', 'replace') if msg else ''


def check(code: int) -> None:
    """Raise on a library-level error (NOT per-instance integrator flags)."""
    if code == SB_OK:
        return
    msg = last_error()
    if code == SB_ERR_CUDA:
        raise DeviceError('sunode_b200: %s' % msg)
    raise LibraryError('sunode_b200 error %d: %s' % (code, msg))


def nvrtc_version() -> str:
    major, minor = ctypes.c_int(0), ctypes.c_int(0)
    check(lib().sb_nvrtc_version(ctypes.byref(major), ctypes.byref(minor)))
    return '%d.%d' % (major.value, minor.value)


def device_count() -> int:
    n = ctypes.c_int(0)
    check(lib().sb_device_count(ctypes.byref(n)))
    return n.value
