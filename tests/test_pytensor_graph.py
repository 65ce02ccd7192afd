"""CPU tier: the graph half of wrappers/as_pytensor.py (solve_ivp, make_node, grad) EXECUTED --
on a miniature stand-in for PyTensor (tests/emu/mini_pytensor; the real package is not in the
image) and the stand-in CUDA driver.  The scenario is the reference's sunode/test_pytensor.py."""
import os
import subprocess
import sys

from tests.test_host_logic import _fake_driver_env

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_pytensor_scenario_on_stand_ins(tmp_path):
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'emu', 'pytensor_graph_child.py')],
                          env=_fake_driver_env(tmp_path), capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0 and 'ALL OK' in proc.stdout, proc.stdout + proc.stderr
