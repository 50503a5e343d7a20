"""Quick GPU check of the single-SM overlap-save kernel (ols16k.cu): error against a float64 FIR on the host for a few
shapes, then device-resident timing.  Usage: python profiles/ols_quick.py [nchan_for_timing] [log2 n]"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import libtsd_b200
from libtsd_b200 import fourier as Fo, filtrage as F

libtsd_b200.init(0)

def check(K, Ne, nchan, n, chunks=None, seed=0):
    rng = np.random.default_rng(seed)
    h = rng.standard_normal(K).astype(np.float32)
    N = Fo.prochaine_puissance_de_2(Ne + K)
    H = Fo.ola_make_H(h, N)
    flt, _ = Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, K, H=H, fir_len=K), nchan)
    x = (rng.standard_normal((nchan, n)) + 1j * rng.standard_normal((nchan, n))).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    outs = []
    pos = 0
    for c in (chunks or [n]):
        outs.append(flt.step(xd[:, pos:pos + c]).clone())
        pos += c
    y = torch.cat(outs, dim=1).cpu().numpy()
    D = Ne - K
    nout = y.shape[1]
    err = 0.0
    for c in range(min(nchan, 3)):
        ref = np.convolve(x[c].astype(np.complex128), h.astype(np.float64))[:n]
        ref = np.concatenate([np.zeros(D), ref])[:nout]
        err = max(err, np.abs(y[c] - ref).max() / np.sqrt(np.mean(np.abs(ref) ** 2)))
    print(f"K={K} Ne={Ne} nchan={nchan} n={n} chunks={chunks} out={nout} err/rms={err:.3e}", flush=True)
    return err

if len(sys.argv) <= 1 or sys.argv[1] != "time":
    check(127, 512, 1, 20000)
    check(4095, 61441, 2, 400000)
    check(4095, 61441, 3, 400001, chunks=[65536, 65537, 100000, 3, 169325])
    check(8000, 123072, 2, 600000)
    check(31, 2000, 150, 50000)
else:
    nchan = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    n = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 22)
    h = F.design_rif_fen(4095, "lp", 0.1)
    H = Fo.ola_make_H(h, 65536)
    flt, _ = Fo.filtre_fft(Fo.FiltreFFTConfig(61441, 4095, H=H, fir_len=4095), nchan)
    x = torch.randn((nchan, n), dtype=torch.complex64, device="cuda")
    y = torch.empty((nchan, n + 61441), dtype=torch.complex64, device="cuda")
    for _ in range(2):
        flt.step(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        flt.step(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"nchan={nchan} n={n}: {ms:.3f} ms/step, {nchan * n / ms / 1e6:.1f} Gsamples/s", flush=True)
