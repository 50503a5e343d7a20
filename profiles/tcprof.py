import sys, ctypes, numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
h = oracle.port().design_rif_fen(127, "lp", 0.1)
nchan, n = 1024, 65536
f = F.filtre_rif(h, np.complex64, nchan)
x = torch.empty((nchan, n), dtype=torch.complex64, device="cuda"); torch.view_as_real(x).normal_()
y = torch.empty_like(x)
for _ in range(3): f.step(x, out=y)
torch.cuda.synchronize()
L = ctypes.CDLL(libtsd_b200._lib.SO_PATH)
L.tsdgpu_debug_tcprof_dump(b"gpurun_out/tcprof.bin")
a = np.fromfile("gpurun_out/tcprof.bin", dtype=np.int64).reshape(1024, 16, 4)
a = a[:37]   # blockIdx.x range for this launch (37 spans)
names = ["epi0","epi1","epi2","epi3","prodA0","prodA1","prodA2","prodA3","prodB0","prodB1","prodB2","prodB3","mma"]
for w in range(13):
    m = a[:36, w].mean(axis=0)
    print(f"{names[w]:7s} waitA {m[0]:9.0f}  waitB {m[1]:9.0f}  work {m[2]:9.0f}  total {m[3]:9.0f}")

b = np.fromfile("gpurun_out/tcprof.bin", dtype=np.int64).reshape(1024, 16, 4)[:592, 14]
gt = (b[:, 2] - b[:, 2].min()) / 1e3
dur = b[:, 1]
print("main cycles: mean %.0f min %d max %d" % (dur.mean(), dur.min(), dur.max()))
order = np.argsort(gt)
print("end time (us) quantiles:", np.percentile(gt, [0, 10, 25, 50, 75, 90, 100]).round(1))
for w in range(4):
    sel = order[w * 148:(w + 1) * 148]
    print("wave-ish", w, "end %.1f..%.1f us, main cycles mean %.0f" % (gt[sel].min(), gt[sel].max(), dur[sel].mean()))
sm = b[:, 3]
print("CTAs per SM: min", np.bincount(sm.astype(int)).min(), "max", np.bincount(sm.astype(int)).max())
