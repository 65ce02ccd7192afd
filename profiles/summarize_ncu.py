#!/usr/bin/env python
"""Summarise an ncu report: headline raw metrics + stall-reason mix + hottest SASS segments.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__thread_inst_executed_pred_on_per_inst_executed.ratio',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_op_read.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'smsp__average_warp_latency_per_inst_issued.ratio', 'smsp__warps_eligible.avg.per_cycle_active',
    'sm__cycles_elapsed.max',
]


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu_csv(rep, 'raw')
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')]
        print('== kernel %s  (ncu --set full --clock-control none; replayed, cold caches)' % name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('%-70s %-14s %s' % (k, units[i], vals[i]))
    rows = ncu_csv(rep, 'source')
    hdr, data = rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = {hdr[i]: 0 for i in cols}
    ie, it, isrc = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('Source')
    n_inst = warp_inst = thr_inst = 0
    for r in data:
        for i in cols:
            try:
                tot[hdr[i]] += int(r[i])
            except ValueError:
                pass
        try:
            warp_inst += int(r[ie]); thr_inst += int(r[it]); n_inst += 1
        except ValueError:
            pass
    s = sum(tot.values()) or 1
    print('\n== warp stall samples (all SASS instructions of the kernel)')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print('%-28s %9d %5.1f%%' % (k, v, 100.0 * v / s))
    print('\nSASS instructions: %d   warp-level executed: %d   avg active threads / instruction: %.2f'
          % (n_inst, warp_inst, thr_inst / max(warp_inst, 1)))
    mix = {}
    for r in data:
        op = r[isrc].split()
        if not op:
            continue
        name = op[1] if op[0].startswith('@') and len(op) > 1 else op[0]
        name = name.split('.')[0]
        try:
            mix[name] = mix.get(name, 0) + int(r[ie])
        except ValueError:
            pass
    print('\n== executed instruction mix (warp-level, top 15)')
    for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:15]:
        print('%-10s %12d %5.1f%%' % (k, v, 100.0 * v / max(warp_inst, 1)))


if __name__ == '__main__':
    main()
