"""Real-valued data through the tensor-core FIR (fir_tc2_kernel<true, true>) against the FP32 FMA kernel: 2048 ch x 1 Mi f32,
127 taps, 64 Ki step() blocks (the byte volume of BASELINE config 3).  Usage: python profiles/fir_real_quick.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
h = F.design_rif_fen(127, "lp", 0.1)
nchan, n, blk = 2048, 1 << 20, 1 << 16
x = torch.randn((nchan, n), dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
for tc in ("1", "0"):
    os.environ["TSDGPU_FIR_TC"] = tc
    f = F.filtre_rif(h, np.float32, nchan)
    def step():
        for b in range(n // blk):
            f.step(x[:, b * blk:(b + 1) * blk], out=y[:, b * blk:(b + 1) * blk])
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"TSDGPU_FIR_TC={tc}: {ms:.3f} ms per pass, {nchan * n / ms / 1e6:.1f} G real samples/s, {nchan * n * 8 / ms / 1e6:.0f} GB/s of {6554.6:.0f}")
