"""Rates of the secondary paths (not BASELINE configs): stock resample() chains outside [0.5, 2), exact-delay interpolators,
FFT plans of other sizes, the windowed block filter, arbitrary-H block filter.  Usage: python profiles/secondary_paths_quick.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F, fourier as Fo
libtsd_b200.init(0)

def timeit(name, fn, samples, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:55s} {ms:9.3f} ms  {samples / ms / 1e6:8.1f} Gsamples/s", flush=True)

nchan, n = 256, 1 << 20
x = torch.randn((nchan, n), dtype=torch.complex64, device="cuda")
for ratio in (0.3, 0.1, 3.0, 7.3):
    f = F.filtre_reechan(ratio, nchan)
    timeit(f"filtre_reechan ratio {ratio}", lambda: f.step(x), nchan * n)
for nm, it in (("lineaire", F.itrp_lineaire()), ("lagrange(3)", F.itrp_lagrange(3))):
    f = F.filtre_itrp(147 / 160, it, nchan)
    timeit(f"filtre_itrp {nm}", lambda: f.step(x), nchan * n)
for N in (256, 1024, 4096, 16384, 32768, 131072, 1000, 12345):
    b = max(1, (nchan * n) // N)
    xx = torch.randn((b, N), dtype=torch.complex64, device="cuda")
    yy = torch.empty_like(xx)
    plan = Fo.tfrplan_creation(N, batch=b)
    timeit(f"fft plan N={N} batch={b}", lambda: plan.step(xx, True, out=yy), b * N)
# windowed (Hann, 50 % overlap) block filter, Ne = 2048
cfg = Fo.FiltreFFTConfig(2048, 0, avec_fenetrage=True)
flt, N = Fo.filtre_fft(cfg, nchan)
timeit(f"filtre_fft avec_fenetrage Ne=2048 N={N}", lambda: flt.step(x), nchan * n)
# arbitrary H (not the transform of few taps), N = 65536
rng = np.random.default_rng(0)
H = (rng.standard_normal(65536) + 1j * rng.standard_normal(65536)).astype(np.complex64)
flt2, N2 = Fo.filtre_fft(Fo.FiltreFFTConfig(61441, 4095, H=H), 64)
x2 = torch.randn((64, 1 << 22), dtype=torch.complex64, device="cuda")
timeit(f"filtre_fft arbitrary H N={N2}", lambda: flt2.step(x2), 64 * (1 << 22))
H3 = (rng.standard_normal(1024) + 1j * rng.standard_normal(1024)).astype(np.complex64)
flt3, N3 = Fo.filtre_fft(Fo.FiltreFFTConfig(512, 512, H=H3), nchan)
timeit(f"filtre_fft arbitrary H N={N3}", lambda: flt3.step(x), nchan * n)
