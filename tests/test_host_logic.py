"""CPU tier: the HOST half of the C ABI (csrc/sb_api.cpp -- staging buffers, argument blocks,
kernel selection, history strides, status routing) and the Python layer above it, run without a
GPU on a stand-in driver (tests/emu/fake_cuda.cpp) whose cuLaunchKernel executes the host
emulation of the same kernels.  Results must equal the emulated kernels called directly, bit for
bit.  The GPU tier runs the same calls on the real driver."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_driver_env(tmp_path):
    lib = tmp_path / 'libcuda.so.1'
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([cxx, '-O1', '-std=c++17', '-fPIC', '-shared', '-I', '/usr/local/cuda/include',
                    os.path.join(ROOT, 'tests', 'emu', 'fake_cuda.cpp'), '-o', str(lib), '-ldl'],
                   check=True)
    env = dict(os.environ)
    env['LD_LIBRARY_PATH'] = str(tmp_path) + os.pathsep + env.get('LD_LIBRARY_PATH', '')
    env['SUNODE_B200_CACHE'] = str(tmp_path / 'cache')      # keep the tree's cubins untouched
    env['PYTHONPATH'] = ROOT + os.pathsep + env.get('PYTHONPATH', '')
    env.pop('SUNODE_B200_DEFINES', None)
    return env


def test_c_abi_host_logic_on_a_fake_driver(tmp_path):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'emu', 'host_logic_child.py')],
                          env=_fake_driver_env(tmp_path), capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and 'ALL OK' in proc.stdout, proc.stdout + proc.stderr


def test_smoke_entry_point_on_a_fake_driver(tmp_path):
    """``__graft_entry__.smoke()`` as the driver calls it, with the kernels emulated: its calls,
    shapes and envelopes are checked before a GPU box runs it."""
    code = ('from tests.emu import dryrun_plugin; dryrun_plugin.pytest_configure(None); '
            'import __graft_entry__ as g; g.smoke()')
    proc = subprocess.run([sys.executable, '-c', code], env=_fake_driver_env(tmp_path), cwd=ROOT,
                          capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and 'smoke ok' in proc.stdout, proc.stdout + proc.stderr
