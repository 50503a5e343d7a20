#!/usr/bin/env python
"""Executed warp instructions by SASS opcode from `ncu --page source --csv`.  usage: op_mix.py src.csv [points]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
pts = float(sys.argv[2]) if len(sys.argv) > 2 else None
c = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    n = int(r[ix["Instructions Executed"]] or 0)
    src = r[ix["Source"]].split()
    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
    op = op.split(".")[0]
    c[op] += n; tot += n
print("total warp instructions", tot, "" if not pts else f"= {tot*32/pts:.1f} thread instr per point")
for op, n in c.most_common(30):
    print(f"{op:12s} {n:12d} {100*n/tot:5.1f}%" + ("" if not pts else f"  {n*32/pts:6.2f}/pt"))
