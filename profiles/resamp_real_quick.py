"""filtre_itrp on real-valued data (the reference's own acceptance-test type): rate of the current path.
Usage: python profiles/resamp_real_quick.py"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
lut = oracle.port().itrp_sinc_lut(64, 256, 0.4)
nchan, n = 512, 1 << 20
for dt, tdt in ((np.float32, torch.float32), (np.complex64, torch.complex64)):
    x = torch.randn((nchan, n), dtype=tdt, device="cuda")
    f = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), nchan, dt)
    for _ in range(2): y = f.step(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): y = f.step(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{np.dtype(dt).name}: {ms:.2f} ms, {nchan * n / ms / 1e6:.1f} G input samples/s", flush=True)
