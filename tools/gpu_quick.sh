#!/bin/bash
# build + one short device-resident bench of a workload on the GPU box (A/B runs during kernel work)
# usage: tools/gpu_quick.sh [workload] [extra bench args]
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -3
W=${1:-lv_adj}; shift || true
/usr/local/graft/bin/gpurun --timeout 900 -- "python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $* | python -c \"import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('RESULT', d['value'], d['roofline']['kernel_ms_all'], d['config']['failed_instances'], d['roofline']['registers'])\"" 2>&1 | grep -E "RESULT|status=|Error|error" | head
