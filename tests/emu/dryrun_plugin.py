"""pytest plugin for tools/dry_run_gpu_tests.py -- TEST INFRASTRUCTURE ONLY.

Runs `-m gpu` tests on the stand-in driver (fake_cuda.cpp): every cubin request is answered by
building the host emulation of the same kernels (same problem, same build options) and pointing
the fake driver at it.  What this checks before a GPU box is spent: that the tests' calls go
through the C ABI, and that their thresholds hold for the emulated arithmetic."""
import os
import tempfile

_work = tempfile.mkdtemp(prefix='sb_dryrun_')


def pytest_configure(config):
    from sunode_b200 import _engine
    from tests.emu.emu import Emulator

    class _Gen:
        def __init__(self, gen):
            self.generated = gen

    def compile_cubin(gen, *, defines=(), **kw):
        env = [d.strip() for d in os.environ.get('SUNODE_B200_DEFINES', '').split(',') if d.strip()]
        # lane groups are not emulated behind the fake driver (it reports one lane per instance)
        alld = tuple(d for d in tuple(defines) + tuple(env) if not d.startswith('SB_NO_GROUP')
                     and not d.startswith('SB_GROUP'))
        emu = Emulator(_Gen(gen), _work, defines=alld)
        os.environ['SB_FAKE_EMU_LIB'] = emu.lib._name
        return b'\x7fELF' + bytes(60), '<emulated>'

    _engine.compile_cubin = compile_cubin
