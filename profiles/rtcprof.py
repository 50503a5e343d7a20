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
L = ctypes.CDLL(libtsd_b200._lib.SO_PATH)
L.tsdgpu_debug_rtcprof_dump(b"gpurun_out/rtcprof.bin")
a = np.fromfile("gpurun_out/rtcprof.bin", dtype=np.int64).reshape(1024, 32, 4)
nz = a[:, 24, 3] > 0   # CTAs whose MMA warp ran (pair mode: the leaders)
a = a[nz]
print("CTAs traced", len(a))
names = {4: "conv0", 8: "gen0", 23: "gen15", 24: "mma"}
for w, nm in names.items():
    m = a[:, w].mean(axis=0)
    print(f"{nm:6s} waitA {m[0]:9.0f}  waitB {m[1]:9.0f}  work {m[2]:9.0f}  total {m[3]:9.0f}")
