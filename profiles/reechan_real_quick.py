import sys
import numpy as np, torch
sys.path.insert(0, ".")
import libtsd_b200
from libtsd_b200 import filtrage as F, detection as D
libtsd_b200.init(0)
def timeit(name, fn, samples, reps=3):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:50s} {ms:9.3f} ms  {samples / ms / 1e6:8.1f} Gsamples/s", flush=True)
nchan, n = 256, 1 << 20
xr = torch.randn((nchan, n), dtype=torch.float32, device="cuda")
for ratio in (0.3, 147 / 160, 3.0):
    f = F.filtre_reechan(ratio, nchan, np.float32)
    timeit(f"filtre_reechan<float> ratio {ratio:.3f}", lambda: f.step(xr), nchan * n)
