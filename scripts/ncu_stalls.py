#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: stall-reason totals and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
out = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r; idx = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr): continue
    out.append(r)
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {h: 0 for h in st}; ns = 0; ni = 0
for r in out:
    for h in st: tot[h] += int(r[idx[h]] or 0)
    ns += int(r[idx['# Samples']] or 0); ni += int(r[idx['Instructions Executed']] or 0)
print("samples", ns, "warp-instructions", ni, "sass lines", len(out))
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print("  %-24s %8d %5.1f%%" % (h, v, 100.0 * v / ns))
for r in sorted(out, key=lambda r: -int(r[idx['# Samples']] or 0))[:n]:
    print(r[idx['# Samples']].rjust(7), r[idx['Instructions Executed']].rjust(9), r[idx['Source']][:100])
