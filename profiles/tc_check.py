"""Quick check of the tensor-core FIR path against the CPU oracle (run under gpurun with TSDGPU_FIR_TC=1)."""
import sys, numpy as np
sys.path.insert(0, ".")
import libtsd_b200, oracle
from libtsd_b200 import filtrage as F
libtsd_b200.init(0)
O = oracle.ref() if oracle.have_ref() else oracle.port()
rng = np.random.default_rng(1)
def cn(*s): return (rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64)
for K, nchan, n in ((127, 3, 1000), (127, 70, 4096), (31, 2, 500), (1, 1, 300), (100, 64, 65536), (127, 130, 20001)):
    h = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    f = F.filtre_rif(h, np.complex64, nchan)
    refs = [O.fir(1, h) for _ in range(min(nchan, 4))]
    errs = []
    for blk in (n, 130, 7):
        x = cn(nchan, blk)
        y = f.step(x)
        for c, r in enumerate(refs):
            yr = r.step(x[c])
            errs.append(float(np.max(np.abs(y[c] - yr)) / np.sqrt(np.mean(np.abs(x) ** 2))))
    print(f"K={K} nchan={nchan} n={n}: max rel err {max(errs):.3e}", flush=True)
