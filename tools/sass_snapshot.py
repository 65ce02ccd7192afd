#!/usr/bin/env python
"""Per-kernel SASS snapshots of the default builds of the benchmark problems (no GPU needed).

    python tools/sass_snapshot.py save  DIR     # before touching csrc/
    python tools/sass_snapshot.py check DIR     # after: which kernels changed?

Used to prove that an addition behind a build option (SB_HERMITE, SB_CONSTRAINTS, SB_FUND ...)
leaves the kernels of the measured path byte-identical, so that numbers in profiles/ stay valid.
"""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunode_b200 import _build, _engine, examples   # noqa: E402


def kernels_of(problem):
    _, path = _engine.compile_cubin(problem.generated)
    sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True, check=True).stdout
    sass = '\n'.join(l for l in sass.splitlines() if not re.match(r'^\s*/\*[0-9a-f]*\*/\s*$', l)) + '\n'
    parts = re.split(r'\n\s*Function : (\w+)\n', sass)
    return {parts[i]: parts[i + 1] for i in range(1, len(parts) - 1, 2)}


def main():
    mode, out = sys.argv[1], sys.argv[2]
    _build.build_library()
    os.makedirs(out, exist_ok=True)
    changed = 0
    for make in examples.problem_list():
        for name, text in kernels_of(make()).items():
            path = os.path.join(out, '%s.%s.sass' % (make.__name__, name))
            if mode == 'save':
                with open(path, 'w') as fh:
                    fh.write(text)
            else:
                with open(path) as fh:
                    same = fh.read() == text
                changed += not same
                print('%-16s %-18s %s' % (make.__name__, name, 'identical' if same else 'CHANGED'))
    if mode != 'save':
        print('%d kernel(s) changed' % changed)


if __name__ == '__main__':
    main()
