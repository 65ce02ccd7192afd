#!/bin/bash
# A/B of kernel build variants on the GPU box: tools/ab_defines.sh <workload> <bench args...> -- <defines1> <defines2> ...
# ("-" = the default build); prints value / kernel times per variant.  Run under gpurun.
W=$1; shift
ARGS=()
while [ "$1" != "--" ] && [ $# -gt 0 ]; do ARGS+=("$1"); shift; done
shift
for D in "$@"; do
  if [ "$D" = "-" ]; then unset SUNODE_B200_DEFINES; else export SUNODE_B200_DEFINES="$D"; fi
  python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary "${ARGS[@]}" 2> gpurun_out/ab_err.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
    print('AB', '$D', '%.4e' % d['value'], {k: round(v,3) for k,v in r['kernel_ms_all'].items()}, r['registers'], d['config']['failed_instances'], r['mean_steps'])
except Exception as e:
    print('AB', '$D', 'FAILED', e); print(open('gpurun_out/ab_err.log').read()[-1500:])
"
done
