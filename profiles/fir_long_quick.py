"""Long direct FIRs (cf32 data): overlap-save kernel vs FP32 FMA kernel.  Usage: python profiles/fir_long_quick.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
nchan, n = 256, 1 << 20
x = torch.randn((nchan, n), dtype=torch.complex64, device="cuda")
y = torch.empty_like(x)
for K in (128, 512, 2048):
    h = F.design_rif_fen(K - 1 if K % 2 == 0 else K, "lp", 0.1)
    for ols in ("1", "0"):
        os.environ["TSDGPU_FIR_OLS"] = ols
        f = F.filtre_rif(h, np.complex64, nchan)
        for _ in range(2): f.step(x, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f.step(x, out=y)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"K={len(h)} TSDGPU_FIR_OLS={ols}: {ms:.2f} ms, {nchan * n / ms / 1e6:.1f} Gsamples/s", flush=True)
# complex taps (cf32 x cf32): no tensor-core kernel, 8 K flop per sample in the direct form
for K in (33, 63, 127):
    rng = np.random.default_rng(K)
    hc = ((rng.standard_normal(K) + 1j * rng.standard_normal(K)) / np.sqrt(K)).astype(np.complex64)
    for ols in ("1", "0"):
        os.environ["TSDGPU_FIR_OLS"] = ols
        f = F.filtre_rif(hc, np.complex64, nchan)
        for _ in range(2): f.step(x, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f.step(x, out=y)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"complex taps K={K} TSDGPU_FIR_OLS={ols}: {ms:.2f} ms, {nchan * n / ms / 1e6:.1f} Gsamples/s", flush=True)
# real-valued data, long filters: channel pairs on the overlap-save kernel vs the FMA kernel
xr = torch.randn((2 * nchan, n), dtype=torch.float32, device="cuda")
yr = torch.empty_like(xr)
for K in (255, 511):
    h = F.design_rif_fen(K, "lp", 0.1)
    for ols in ("1", "0"):
        os.environ["TSDGPU_FIR_OLS"] = ols
        f = F.filtre_rif(h, np.float32, 2 * nchan)
        for _ in range(2): f.step(xr, out=yr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f.step(xr, out=yr)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"real data K={K} TSDGPU_FIR_OLS={ols}: {ms:.2f} ms, {2 * nchan * n / ms / 1e6:.1f} G real samples/s", flush=True)
