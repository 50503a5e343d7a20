"""Prints the per-phase clock trace written by TSDGPU_OLS_PROF (ols16k.cu: CTA 0, 6 consecutive blocks, 16 math warps,
14 time stamps).  Usage: python profiles/tools/ols_trace.py trace.bin"""
import sys
import numpy as np
a = np.fromfile(sys.argv[1], np.uint32).reshape(6, 16, 14).astype(np.int64)
names = ["top", "P1 loads", "P1 fft32", "P1 sts", "row wait", "P2 in+fft16+tw", "P2 T1", "P2 fft32 H ifft32", "P2 T2", "P2 tw+ifft16+sts",
         "bar2", "P3 lds", "P3 ifft32+stg"]
t0 = a[1, :, 0].min()
for b in range(1, 5):
    print(f"--- block {b}: period {a[b + 1, :, 0].min() - a[b, :, 0].min()} cycles")
    d = np.diff(a[b, :, :13], axis=1)
    print("warp " + " ".join(f"{n[:9]:>9s}" for n in names[1:]) + "    start      end")
    for w in range(16):
        print(f"{w:4d} " + " ".join(f"{int(x):9d}" for x in d[w]) + f" {a[b, w, 0] - t0:8d} {a[b, w, 12] - t0:8d}")
    print("mean " + " ".join(f"{int(x):9d}" for x in d.mean(0)))
