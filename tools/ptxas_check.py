"""nvcc -Xptxas -v of one problem's translation unit (registers / spills / SASS size per kernel).

    python tools/ptxas_check.py [lotka_volterra|robertson|seir] [-DNAME=VALUE ...]
"""
import os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sunode_b200 import _build, _engine, examples

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('-') else 'lotka_volterra'
defs = [a for a in sys.argv[1:] if a.startswith('-D')]
prob = getattr(examples, name)()
tmp = os.path.join(os.path.dirname(_build.CSRC), '..', 'build')
os.makedirs(tmp, exist_ok=True)
src = os.path.join(tmp, name + '_check.cu')
with open(src, 'w') as fh:
    fh.write(prob.generated.cuda + '\n#include "sb_kernels.cuh"\n')
cub = os.path.join(tmp, name + '_check.cubin')
cmd = ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xptxas', '-v',
       '-DSB_BLOCK=%d' % _engine.DEFAULT_BLOCK, *defs, '-I', _build.CSRC, '-cubin', '-o', cub, src]
p = subprocess.run(cmd, capture_output=True, text=True)
if p.returncode:
    print(p.stdout + p.stderr); sys.exit(1)
cur = None
for l in p.stderr.splitlines():
    m = re.search(r"Compiling entry function '(\w+)'", l)
    if m: cur = m.group(1)
    m = re.search(r'Used (\d+) registers', l)
    if m: print('%-18s %s' % (cur, l.strip().replace('ptxas info    : ', '')))
    if 'spill' in l and 'bytes stack' in l: print('%-18s %s' % (cur, l.strip().replace('ptxas info    : ', '')))
sass = subprocess.run(['cuobjdump', '-sass', cub], capture_output=True, text=True).stdout
cur = None; n = {}
for l in sass.splitlines():
    m = re.search(r'Function : (\w+)', l)
    if m: cur = m.group(1); n[cur] = 0
    elif cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): n[cur] += 1
print('SASS instructions:', n)
