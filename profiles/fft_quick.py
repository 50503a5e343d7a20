"""Quick GPU check of the 65536-point plan (fft64k_pipe.cu): error against numpy's double FFT for a few batch sizes and
in-place use, then device-resident timing of fft + ifft round trips.  Usage: python profiles/fft_quick.py [time batch]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import libtsd_b200
from libtsd_b200 import fourier as Fo

libtsd_b200.init(0)
N = 65536

def check(batch, inplace=False, seed=0):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((batch, N)) + 1j * rng.standard_normal((batch, N))).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    plan = Fo.tfrplan_creation(N, batch=batch)
    X = plan.step(xd, True, out=xd if inplace else None)
    Xh = X.cpu().numpy()
    idx = sorted(set([0, batch - 1, batch // 2]))
    err = 0.0
    for c in idx:
        ref = np.fft.fft(x[c].astype(np.complex128)) / np.sqrt(N)
        err = max(err, np.abs(Xh[c] - ref).max() / np.sqrt(np.mean(np.abs(ref) ** 2)))
    back = plan.step(X, False).cpu().numpy()
    err2 = max(np.abs(back[c] - x[c]).max() / np.sqrt(np.mean(np.abs(x[c]) ** 2)) for c in idx)
    print(f"batch={batch} inplace={inplace} fwd err/rms={err:.3e}  round trip err/rms={err2:.3e}", flush=True)

if len(sys.argv) <= 1 or sys.argv[1] != "time":
    for b in (1, 3, 50, 300):
        check(b)
    check(97, inplace=True)
else:
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    x = torch.randn((batch, N), dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    plan = Fo.tfrplan_creation(N, batch=batch)
    for _ in range(2):
        plan.step(x, True, out=y); plan.step(y, False, out=x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        plan.step(x, True, out=y); plan.step(y, False, out=x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"batch={batch}: {ms:.3f} ms per round trip -> {batch * N / ms / 1e6:.1f} G round trips/s = {batch * N * 32 / ms / 1e6:.0f} GB/s "
          f"({batch * N * 32 / ms / 1e6 / 6554.6:.3f} of measured HBM)", flush=True)
