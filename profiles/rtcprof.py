import sys, ctypes, numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
lut = oracle.port().itrp_sinc_lut(64, 256, 0.4)
nchan, n = 512, 1 << 20
f = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), nchan)
x = torch.empty((nchan, n), dtype=torch.complex64, device="cuda"); torch.view_as_real(x).normal_()
for _ in range(3): y = f.step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y = f.step(x)
e1.record()
torch.cuda.synchronize()
print("step of %d ch x %d: %.3f ms" % (nchan, n, e0.elapsed_time(e1) / 5))
L = ctypes.CDLL(libtsd_b200._lib.SO_PATH)
L.tsdgpu_debug_rtcprof_dump(b"gpurun_out/rtcprof.bin")
raw = np.fromfile("gpurun_out/rtcprof.bin", dtype=np.int64)
a = raw[: 1024 * 32 * 4].reshape(1024, 32, 4)
life = raw[1024 * 32 * 4:].reshape(8192, 4)
nz = a[:, 24, 3] > 0   # CTAs whose MMA warp ran (pair mode: the leaders)
a = a[nz]
print("CTAs traced", len(a))
names = {4: "conv0", 8: "gen0", 23: "gen15", 24: "mma"}
for w, nm in names.items():
    m = a[:, w].mean(axis=0)
    print(f"{nm:6s} waitA {m[0]:9.0f}  waitB {m[1]:9.0f}  work {m[2]:9.0f}  total {m[3]:9.0f}")

c = a[:, 31]
print("per CTA (thread 0): prologue %.0f  roles %.0f  final sync %.0f cycles" % tuple(c[:, :3].mean(axis=0)))
ends = np.sort(c[:, 3])
print("end-time spread of the traced CTAs: %.1f us; first-wave CTAs end after %.1f us" % ((ends[-1] - ends[0]) / 1e3, 0.0))

L = life[life[:, 1] > 0]
t0 = L[:, 0].min()
st, en = (L[:, 0] - t0) / 1e3, (L[:, 1] - t0) / 1e3
print("CTAs %d, SMs used %d, kernel span %.1f us, mean life %.1f us, sum of lives / span = %.1f CTAs resident" % (len(L), len(set(L[:, 2])), en.max(), (en - st).mean(), (en - st).sum() / en.max()))
for q in range(0, 100, 10):
    lo, hi = np.percentile(st, q), np.percentile(st, q + 10)
    m = (st >= lo) & (st <= hi)
    print("  CTAs started in [%6.0f, %6.0f] us: mean life %.1f us" % (lo, hi, (en - st)[m].mean()))
per_sm = {}
for s_, e_, sm, r in zip(st, en, L[:, 2], L[:, 3]):
    per_sm.setdefault(int(sm), []).append((s_, e_))
gaps = []
for sm, v in per_sm.items():
    v.sort()
    gaps += [b[0] - a_[1] for a_, b in zip(v[:-1], v[1:])]
print("gap between consecutive CTAs on an SM: mean %.1f us, median %.1f us, max %.1f us; CTAs per SM %.1f" % (np.mean(gaps), np.median(gaps), np.max(gaps), len(L) / len(per_sm)))
