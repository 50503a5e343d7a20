"""Opcode histogram per kernel of the shipped library (cuobjdump -sass libtsd_b200/libtsdgpu.so): the evidence of which
Blackwell units every kernel uses (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP / UTMALDG = TMA copies,
SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed FP32).  Usage: python profiles/tools/sass_ops.py > profiles/r02_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "libtsd_b200", "libtsdgpu.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "USETMAXREG", "BAR",
       "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "HMMA", "LDS", "STS", "LDG", "STG", "LDC", "LDL", "STL", "SHFL", "RED", "ATOM", "MUFU")
kern, counts = None, {}
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        op = m.group(1)
        counts[kern][op] += 1
print(f"# opcode histogram of {os.path.relpath(so, ROOT)} (static instruction counts per kernel; only opcodes that identify a unit)")
for k in sorted(counts):
    c = counts[k]
    tot = sum(c.values())
    sel = {op: c[op] for op in KEY if c.get(op)}
    name = re.sub(r"\(.*", "", k)
    print(f"{name}: total {tot}  " + "  ".join(f"{op} {n}" for op, n in sel.items()))
