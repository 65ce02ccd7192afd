#!/usr/bin/env python
"""Run `-m gpu` tests WITHOUT a GPU: the C-ABI library on the stand-in driver of
tests/emu/fake_cuda.cpp, kernels executed by the host emulation of the device sources.

    python tools/dry_run_gpu_tests.py tests/test_zz_options_gpu.py [-k expr ...]

Checks the host-side plumbing and the tests' thresholds before GPU minutes are spent.  Tests that
need torch CUDA tensors or lane groups as such cannot run here."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tmp = tempfile.mkdtemp(prefix='sb_fake_')
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([cxx, '-O1', '-std=c++17', '-fPIC', '-shared', '-I', '/usr/local/cuda/include',
                    os.path.join(ROOT, 'tests', 'emu', 'fake_cuda.cpp'), '-o',
                    os.path.join(tmp, 'libcuda.so.1'), '-ldl'], check=True)
    env = dict(os.environ)
    env['LD_LIBRARY_PATH'] = tmp + os.pathsep + env.get('LD_LIBRARY_PATH', '')
    env['SUNODE_B200_CACHE'] = os.path.join(tmp, 'cache')
    env['PYTHONPATH'] = ROOT + os.pathsep + env.get('PYTHONPATH', '')
    args = sys.argv[1:] or ['tests/test_zz_options_gpu.py']
    cmd = [sys.executable, '-m', 'pytest', '-m', 'gpu', '-q', '-p', 'tests.emu.dryrun_plugin'] + args
    sys.exit(subprocess.run(cmd, env=env, cwd=ROOT).returncode)


if __name__ == '__main__':
    main()
