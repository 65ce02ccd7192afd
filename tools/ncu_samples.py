#!/usr/bin/env python
"""Source lines of an ncu report ranked by stall SAMPLES (where the warps wait), with the top
stall reasons of each line.    python tools/ncu_samples.py x.ncu-rep [top]"""
import csv, io, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = hdr = None; L = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = os.path.basename(r[1]); continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No':
        hdr = r; ie = hdr.index('Instructions Executed'); it = hdr.index('Thread Instructions Executed'); ism = hdr.index('# Samples')
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr is None or not r: continue
    if r[0] != '': cur = (fname, int(r[0])); continue
    try: w, t, s = int(r[ie]), int(r[it]), int(r[ism])
    except ValueError: continue
    a = L.setdefault(cur, [0, 0, 0, {}]); a[0] += w; a[1] += t; a[2] += s
    for i, h in stall:
        try: a[3][h] = a[3].get(h, 0) + int(r[i])
        except ValueError: pass
ts = sum(v[2] for v in L.values())
for k, v in sorted(L.items(), key=lambda kv: -kv[1][2])[:top]:
    why = ', '.join('%s %d%%' % (h[6:], 100 * n / max(sum(v[3].values()), 1)) for h, n in sorted(v[3].items(), key=lambda kv: -kv[1])[:3] if n)
    print('%-24s smp %5.2f%%  warpinst %7.1fM  active %5.2f  %s' % ('%s:%d' % k, 100 * v[2] / ts, v[0] / 1e6, v[1] / max(v[0], 1), why))
