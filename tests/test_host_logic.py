"""CPU tier: the HOST half of the C ABI (csrc/sb_api.cpp -- staging buffers, argument blocks,
kernel selection, history strides, status routing) and the Python layer above it, run without a
GPU on a stand-in driver (tests/emu/fake_cuda.cpp) whose cuLaunchKernel executes the host
emulation of the same kernels.  Results must equal the emulated kernels called directly, bit for
bit.  The GPU tier runs the same calls on the real driver."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_host_logic_on_a_fake_driver(tmp_path):
    cuda_inc = '/usr/local/cuda/include'
    lib = tmp_path / 'libcuda.so.1'
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.run([cxx, '-O1', '-std=c++17', '-fPIC', '-shared', '-I', cuda_inc,
                    os.path.join(ROOT, 'tests', 'emu', 'fake_cuda.cpp'), '-o', str(lib), '-ldl'],
                   check=True)
    env = dict(os.environ)
    env['LD_LIBRARY_PATH'] = str(tmp_path) + os.pathsep + env.get('LD_LIBRARY_PATH', '')
    env['SUNODE_B200_CACHE'] = str(tmp_path / 'cache')      # keep the tree's cubins untouched
    env.pop('SUNODE_B200_DEFINES', None)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'emu', 'host_logic_child.py')],
                          env=env, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and 'ALL OK' in proc.stdout, proc.stdout + proc.stderr
