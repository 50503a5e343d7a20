#!/usr/bin/env python
"""Instruction mix (executed warp instructions by SASS opcode) from `ncu --page source --csv` of one kernel."""
import csv, sys, collections, subprocess, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sys.argv[2:], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, smp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    op = r[isrc].strip().split()
    if not op: continue
    m = op[1] if op[0].startswith("@") else op[0]
    m = m.rstrip(";")
    base = m.split(".")[0]
    key = base if base not in ("LDG","STG","LDS","STS","LD","ST","RED","ATOMG","BAR","MUFU") else m
    n = int(r[iex]); mix[key] += n; tot += n; smp[key] += int(r[ismp])
print("total warp instr", tot)
for k, v in mix.most_common(40):
    print(f"{k:28s} {v:12d} {100*v/tot:6.2f}%   samples {smp[k]}")
