"""The C-ABI library: loads without a GPU, exports every symbol the header declares, compiles
kernels for sm_100a without a GPU, and fails loudly (no CPU fallback) when a device is needed."""
import ctypes
import os
import re

import numpy as np
import pytest

from sunode_b200 import _build, _engine, _lib, examples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, 'include', 'sunode_b200.h')) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sb_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_lib.EXPORTS) == declared           # the ctypes table covers the whole header
    assert lib.sb_version() >= 100


def test_header_is_plain_c():
    """No torch / C++ types in the boundary: the header must compile as C99."""
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'hdr.c')
        with open(src, 'w') as fh:
            fh.write('#include "sunode_b200.h"\nint main(void) { return sb_version() < 0; }\n')
        subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-c', src, '-I',
                        os.path.join(ROOT, 'include'), '-o', os.path.join(tmp, 'hdr.o')], check=True)


def test_compile_for_sm100a_without_gpu(tmp_path, monkeypatch):
    # a scratch cache: this compile must not replace the cubin build() put into the tree
    monkeypatch.setenv('SUNODE_B200_CACHE', str(tmp_path))
    gen = examples.lotka_volterra().generated
    cubin, path = _engine.compile_cubin(gen, use_cache=False)
    assert os.path.dirname(path) == str(tmp_path)
    assert cubin[:4] == b'\x7fELF' and os.path.exists(path)
    import subprocess
    out = subprocess.run(['cuobjdump', '-elf', path], capture_output=True, text=True).stdout
    assert 'sm_100' in out or 'SM100' in out.upper() or 'EF_CUDA_SM100' in out.upper()
    for kernel in ('sb_forward', 'sb_backward', 'sb_tables', 'sb_eval'):
        assert kernel.encode() in cubin


def test_compile_error_is_reported():
    lib = _lib.lib()
    cubin, size, log = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_void_p()
    code = lib.sb_compile(b'#define SB_NS 1\nthis is not CUDA\n', b'sm_100a', 32, 1,
                          ctypes.byref(cubin), ctypes.byref(size), ctypes.byref(log))
    assert code == _lib.SB_ERR_NVRTC and not cubin.value
    assert b'error' in ctypes.string_at(log.value)
    lib.sb_free(log)
    assert 'nvrtcCompileProgram' in _lib.last_error()
    code = lib.sb_compile(b'', b'sm_100a', 33, 1, ctypes.byref(cubin), ctypes.byref(size), None)
    assert code == _lib.SB_ERR_ARG


def test_argument_validation_without_device():
    lib = _lib.lib()
    assert lib.sb_set_tolerances(None, 1e-8, None, 1) == _lib.SB_ERR_ARG
    assert lib.sb_problem_destroy(None) == _lib.SB_OK
    assert lib.sb_launch_count(None) == 0
    handle = ctypes.c_void_p()
    assert lib.sb_problem_create(ctypes.byref(handle), 0, 0, 0, None, 0, 0) == _lib.SB_ERR_ARG


def _no_gpu():
    try:
        return _lib.device_count() == 0
    except _lib.DeviceError:
        return True


@pytest.mark.skipif(not _no_gpu(), reason='needs a machine WITHOUT a GPU')
def test_no_silent_cpu_fallback():
    """Without a device the product path must raise, never compute on the CPU."""
    from sunode_b200.solver import AdjointSolver, Solver
    prob = examples.lotka_volterra()
    with pytest.raises(_lib.DeviceError):
        Solver(prob)
    with pytest.raises(_lib.DeviceError):
        AdjointSolver(prob)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under sunode_b200/ may reference it."""
    pkg = os.path.join(ROOT, 'sunode_b200')
    for dirpath, _, files in os.walk(pkg):
        if '_cache' in dirpath:
            continue
        for name in files:
            if name.endswith(('.py', '.cpp', '.cuh', '.h')):
                with open(os.path.join(dirpath, name)) as fh:
                    text = fh.read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), name
                assert 'liboracle' not in text and 'cvodes_port' not in text, name


def test_embedded_sources_are_current():
    _build.write_embedded()
    with open(os.path.join(_build.CSRC, 'sb_embedded.inc')) as fh:
        inc = fh.read()
    with open(os.path.join(_build.CSRC, 'sb_bdf.cuh')) as fh:
        assert fh.read()[:2000] in inc
