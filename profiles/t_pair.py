import sys, numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
P = oracle.port()
lut = P.itrp_sinc_lut(64, 256, 0.4)
nchan, n = 128, 40000
rng = np.random.default_rng(1)
x = (rng.standard_normal((nchan, n)) + 1j * rng.standard_normal((nchan, n))).astype(np.complex64)
g = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), nchan)
y = g.step(x)
ref = P.itrp(147.0 / 160.0, lut, 256).step(x[77])
print("shape", y.shape, len(ref), "err", np.abs(y[77] - ref).max())
ref = P.itrp(147.0 / 160.0, lut, 256).step(x[3])
print("err ch3", np.abs(y[3] - ref).max())
