#!/usr/bin/env python
"""Print the headline and the secondary entries of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def show(e):
    e2e = e.get('e2e') or {}
    print('%-14s %-11s N=%d value %.4e  e2e %s  pageable %s  cpu %s  ms/step %.3f  kernels %s  fail %s  clocks %s %s' % (
        e['config']['workload'], e['config']['backward_schedule'], e['n_gpus'], e['value'],
        '%.4e' % e2e['value'] if e2e else None,
        '%.4e' % e2e['pageable']['value'] if e2e.get('pageable') else None,
        '%.3e/%d' % (e['cpu_baseline']['value'], e['cpu_baseline']['cores']) if 'cpu_baseline' in e else None,
        e['ms_per_step'], {k: round(v, 3) for k, v in e['roofline']['kernel_ms_all'].items()},
        e['config']['failed_instances'], e['clocks']['sm_mhz'], e['clocks']['reasons']))


show(d)
for e in d.get('secondary', []):
    show(e)
