"""Resampler step time for a [512][1 Mi] input that is contiguous vs a slice of a [512][8 Mi] buffer (channel stride 64 MiB)."""
import sys, torch
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
lut = oracle.port().itrp_sinc_lut(64, 256, 0.4)
nchan, n = 512, 1 << 20
def run(x, tag):
    f = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), nchan)
    cap = int(n * 147 / 160) + 64
    y = torch.empty((nchan, cap), dtype=torch.complex64, device="cuda")
    for _ in range(3): f.step(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    libtsd_b200._lib.lib().tsdgpu_timing_reset() if hasattr(libtsd_b200._lib.lib(), "tsdgpu_timing_reset") else None
    e0.record()
    for _ in range(5): f.step(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    print(tag, "step %.3f ms" % (e0.elapsed_time(e1) / 5))
xa = torch.empty((nchan, n), dtype=torch.complex64, device="cuda"); torch.view_as_real(xa).normal_()
run(xa, "contiguous [512][1Mi]      ")
xb = torch.empty((nchan, 8 * n), dtype=torch.complex64, device="cuda"); torch.view_as_real(xb).normal_()
run(xb[:, :n], "slice of [512][8Mi]        ")
run(xb[:, 3 * n: 4 * n], "slice 3 of [512][8Mi]      ")
yb = torch.empty((nchan, 8 * n), dtype=torch.complex64, device="cuda")
