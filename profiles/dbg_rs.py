import sys, numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
lut = oracle.port().itrp_sinc_lut(64, 256, 0.4)
nchan, n = 64, 1 << 22
g = torch.Generator(device="cuda"); g.manual_seed(5)
x = torch.empty((nchan, n), dtype=torch.complex64, device="cuda"); torch.view_as_real(x).normal_(generator=g)
ys = []
for rep in range(3):
    f = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), nchan)
    ys.append(f.step(x).clone())
torch.cuda.synchronize()
for rep in (1, 2):
    d = (ys[rep] != ys[0])
    print("rep", rep, "differing samples", int(d.sum()))
    if d.any():
        idx = d.nonzero()
        ch = idx[:, 0].unique()
        cols = idx[:, 1]
        print("  channels", ch[:10].tolist(), "n", len(ch), "cols min/max", int(cols.min()), int(cols.max()), "first cols", cols[:12].tolist())
        c0 = int(cols[0]); r0 = int(idx[0, 0])
        print("  values", ys[0][r0, c0].item(), ys[rep][r0, c0].item())
        print("  distinct tiles", (cols // 128).unique()[:20].tolist())
