#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu -i rep --page source --csv` output.  usage: stall_hot.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    s = int(r[ix["# Samples"]] or 0)
    tot += s
    data.append((s, r))
print("total samples", tot)
agg = {k: 0 for k in stalls}
for s, r in data:
    for k in stalls:
        agg[k] += int(r[ix[k]] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for n, (s, r) in enumerate(data):
    r.append(n)
for s, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = sorted(((int(r[ix[k]] or 0), k) for k in stalls), reverse=True)[:3]
    print(f"{s:6d} {100*s/tot:5.1f}%  #{r[-1]:5d} {r[ix['Source']][:70]:70s} {top}")
