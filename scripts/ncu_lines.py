#!/usr/bin/env python
"""Aggregate `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per CUDA source line."""
import csv, sys
def num(x):
    try: return int(x)
    except Exception: return 0
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; cur = None; agg = {}; src = {}; fname = ''
for r in rows:
    if r and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != '':
        cur = (fname, r[0]); src[cur] = r[1]
        agg.setdefault(cur, [0, 0])
        agg[cur][0] += num(r[iS]); agg[cur][1] += num(r[iI])
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("samples", tot, "instr", toti)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print("%6d %5.1f%% %9d  %s:%s  %s" % (v[0], 100.0 * v[0] / max(tot, 1), v[1], k[0], k[1], src[k].strip()[:110]))
