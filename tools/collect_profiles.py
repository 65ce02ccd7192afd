#!/usr/bin/env python
"""Turn what tools/profile_round.sh left in gpurun_out/ into the tracked files under profiles/:
bench records, the launch list with kernel shares, per-kernel ncu summaries, ncu_traffic.json."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'

for f in ('bench_default', 'bench_lv_adj', 'bench_lv_fwd', 'bench_robertson_adj', 'bench_seir_adj_32768',
          'bench_lv_adj_reference_arm'):
    if os.path.exists('%s/%s.json' % (O, f)):
        shutil.copy('%s/%s.json' % (O, f), '%s/%s_%s.json' % (P, tag, f))

rows = [r for r in csv.reader(open(O + '/launches_bench_lv_adj.csv')) if r and not r[0].startswith('==')]
hdr = rows[0]
ik, ib, ig, iv, iid = (hdr.index(k) for k in ('Kernel Name', 'Block Size', 'Grid Size', 'Metric Value', 'ID'))
tot, out = {}, []
for r in rows[1:]:
    try:
        ns = float(r[iv].replace(',', ''))
    except (ValueError, IndexError):
        continue
    name = r[ik].split('(')[0]
    tot[name] = tot.get(name, 0) + ns
    out.append((r[iid], name, r[ib], r[ig], int(ns)))
s = sum(tot.values())
share = ', '.join('%s=%.1f%%' % (k.split('<')[0], 100 * v / s) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:5])
with open('%s/%s_launches_bench_lv_adj.csv' % (P, tag), 'w') as fh:
    fh.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline\n')
    fh.write('# per-launch times are serialised / cold-cache: compare SHARES.  share of all launches: %s\n' % share)
    fh.write('# (sb_backward and sb_backward_flat are launched back to back; the build whose interval schedule does not suit the batch returns at once)\n')
    fh.write('id,kernel,block,grid,gpu__time_duration.sum [ns]\n')
    for r in out:
        fh.write('%s,%s,"%s","%s",%d\n' % r)
print(share)


def dram(rep):
    rows = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)))
    h, u, v = rows[0], rows[1], rows[2]
    def get(k):
        i = h.index(k)
        return float(v[i].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u[i]]
    return int(get('dram__bytes_read.sum') + get('dram__bytes_write.sum'))


reps = {'lv_adj:65536:sb_backward': 'ncu_sb_backward_lv', 'lv_fwd:65536:sb_forward': 'ncu_sb_forward_lv',
        'seir_adj:32768:sb_backward': 'ncu_sb_backward_seir',
        'robertson_adj:16384:sb_backward_flat': 'ncu_sb_backward_flat_robertson',
        'lv_adj:65536:sb_backward_fund': 'ncu_sb_backward_fund_lv', 'lv_adj:65536:sb_tables': 'ncu_sb_tables_lv',
        'seir_adj:32768:sb_forward': 'ncu_sb_forward_seir'}
t = {'_comment': 'dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from an `ncu --set full` capture '
                 '(profiles/%s_ncu_*.txt; caches flushed before every replay pass, dirty lines still in L2 at the end '
                 'of the launch are not counted), keyed by workload:batch:kernel; bench.py copies the matching entry '
                 'into roofline.traffic' % tag}
for key, rep in reps.items():
    path = '%s/%s.ncu-rep' % (O, rep)
    if not os.path.exists(path):
        continue
    t[key] = dram(path)
    with open('%s/%s_%s.txt' % (P, tag, rep), 'w') as fh:
        fh.write(subprocess.run([sys.executable, P + '/summarize_ncu.py', path], capture_output=True, text=True).stdout)
        fh.write('\n== per source function (tools/ncu_lines.py; inlined frames are counted under every frame they belong to)\n')
        lines = subprocess.run([sys.executable, ROOT + '/tools/ncu_lines.py', path, '25'], capture_output=True, text=True).stdout
        fh.write('\n'.join(lines.splitlines()[:75]) + '\n')
json.dump(t, open(P + '/ncu_traffic.json', 'w'), indent=1)
print(t)
