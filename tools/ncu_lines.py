#!/usr/bin/env python
"""Per-source-line view of an ncu report (kernels compiled with -lineinfo).

    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top]

For every source line: warp-level instructions executed, average active threads, idle-lane
instruction slots (32*warp - thread), stall samples; plus the same aggregated over the functions of
sb_bdf.cuh / sb_kernels.cuh (by line range).
"""
import csv, io, os, re, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None
hdr = None
lines = {}   # (file, line) -> [warp, thread, samples, n_sass]
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        fname = os.path.basename(r[1]); continue
    if len(r) == 2:
        continue
    if r and r[0] == 'Line No':
        hdr = r
        ie = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed')
        ismp = hdr.index('# Samples')
        continue
    if hdr is None or not r:
        continue
    if r[0] != '':          # a source line summary row
        cur = (fname, int(r[0]))
        lines.setdefault(cur, [0, 0, 0, 0])
        continue
    try:
        w, t, s = int(r[ie]), int(r[it]), int(r[ismp])
    except ValueError:
        continue
    e = lines[cur]; e[0] += w; e[1] += t; e[2] += s; e[3] += 1

tw = sum(v[0] for v in lines.values()); tt = sum(v[1] for v in lines.values()); ts = sum(v[2] for v in lines.values())
print('total warp inst %d, avg active %.2f, samples %d, sass %d' % (tw, tt / tw, ts, sum(v[3] for v in lines.values())))

# function ranges from the sources
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'sunode_b200', 'csrc')
funcs = {}
for f in ('sb_bdf.cuh', 'sb_kernels.cuh'):
    starts = []
    for n, l in enumerate(open(os.path.join(root, f)), 1):
        m = re.match(r'\s*(?:SB_ROOT_FN|SB_EXACT_FN|__device__ __forceinline__|extern "C" __global__)[^(]*?(\w+)\s*\(', l)
        if m and not l.strip().startswith('//'):
            starts.append((n, m.group(1)))
        m2 = re.match(r'^sb_(\w+)\(const __grid', l)
        if m2:
            starts.append((n, 'sb_' + m2.group(1)))
    funcs[f] = starts

def func_of(f, n):
    best = '?'
    for s, name in funcs.get(f, []):
        if s <= n: best = name
        else: break
    return best

agg = {}
for (f, n), v in lines.items():
    k = (f, func_of(f, n)) if f in funcs else (f, '-')
    a = agg.setdefault(k, [0, 0, 0, 0])
    for i in range(4): a[i] += v[i]
print('\n%-34s %8s %6s %7s %7s %7s %5s' % ('function', 'warpinst', '%', 'active', 'idle%', 'smp%', 'sass'))
idle_tot = 32 * tw - tt
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] == 0: continue
    print('%-34s %8.1fM %5.1f%% %7.2f %6.1f%% %6.1f%% %5d' % (k[0][:14] + ':' + k[1], v[0] / 1e6, 100 * v[0] / tw, v[1] / v[0],
          100 * (32 * v[0] - v[1]) / idle_tot, 100 * v[2] / ts, v[3]))
print('\n%-24s %8s %6s %7s %7s %5s' % ('line', 'warpinst', '%', 'active', 'smp%', 'sass'))
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-24s %8.1fM %5.1f%% %7.2f %6.1f%% %5d' % ('%s:%d' % k, v[0] / 1e6, 100 * v[0] / tw, v[1] / max(v[0], 1), 100 * v[2] / ts, v[3]))
