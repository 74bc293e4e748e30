#!/usr/bin/env python
"""Top stall locations of one kernel launch from an ncu source-page CSV.
usage: ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv; tools/ncu_src_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][1][:150])
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break  # next launch
    if len(r) >= len(hdr) - 1 and r[0].startswith('0x'):
        data.append(r + [''] * (len(hdr) - len(r)))
isrc, isamp, iexe = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
agg = {}
for r in data:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:n]:
    s = {hdr[i]: int(r[i] or 0) for i in stalls if int(r[i] or 0) > 0}
    s = dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])
    print(r[isamp].rjust(7), r[iexe].rjust(9), r[isrc][:90].ljust(90), s)
