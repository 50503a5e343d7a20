"""Two (or more) ranks: symmetric-memory buffers (torch.distributed._symmetric_memory, CUDA backend = cuMem + peer mapping),
peer copies by the copy engines, and the block filter storing straight into a PEER's buffer.  Launch:
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/multi/p2p_gather_test.py"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import torch.distributed._symmetric_memory as symm_mem
import libtsd_b200
from libtsd_b200 import fourier as Fo, filtrage as F
libtsd_b200.init(lr)

n = 1 << 22
cpr = 64
t = symm_mem.empty((world, cpr, n + 61441), dtype=torch.complex64, device=f"cuda:{lr}")
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
peers = [hdl.get_buffer(r, t.shape, t.dtype) for r in range(world)]
print(rank, "rendezvous ok; multicast:", getattr(hdl, "has_multicast_support", None), flush=True)
src = torch.full((cpr, n + 61441), float(rank + 1), dtype=torch.complex64, device="cuda")
hdl.barrier()
# 1. copy-engine pushes of my shard into slot [rank] of every peer's buffer
side = torch.cuda.Stream()
for rep in range(3):
    torch.cuda.synchronize(); hdl.barrier()
    t0 = time.perf_counter()
    with torch.cuda.stream(side):
        for r in range(world):
            if r != rank:
                peers[r][rank].copy_(src, non_blocking=True)
    side.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(f"push of {src.numel() * 8 / 1e9:.2f} GB to {world - 1} peer(s): {dt * 1e3:.2f} ms = {(world - 1) * src.numel() * 8 / dt / 1e9:.0f} GB/s sent", flush=True)
hdl.barrier()
ok = all(bool((t[r] == float(r + 1)).all()) for r in range(world) if r != rank)
print(rank, "peer pushes landed:", ok, flush=True)
# 2. the block filter storing straight into the peer's memory (kernel stores over NVLink)
h = F.design_rif_fen(4095, "lp", 0.1)
H = Fo.ola_make_H(h, 65536)
Ne = 61441
x = torch.empty((cpr, n), dtype=torch.complex64, device="cuda"); torch.view_as_real(x).normal_()
n_out = Ne * (n // Ne)
loc = torch.empty((cpr, n + 61441), dtype=torch.complex64, device="cuda")
f1, _ = Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, 4095, H=H, fir_len=4095), cpr)
f2, _ = Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, 4095, H=H, fir_len=4095), cpr)
y_loc = f1.step(x, out=loc)
dst = peers[(rank + 1) % world][rank]
hdl.barrier()
y_rem = f2.step(x, out=dst)
torch.cuda.synchronize(); hdl.barrier()
# what the neighbour wrote into MY buffer must equal what it computed locally: compare through a gather of checksums
mine = t[(rank - 1) % world][:, :n_out]
cs_remote = torch.view_as_real(mine).double().sum().reshape(1)
cs_local = torch.view_as_real(loc[:, :n_out]).double().sum().reshape(1)
allr = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
alll = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
dist.all_gather(allr, cs_remote); dist.all_gather(alll, cs_local)
if rank == 0:
    for r in range(world):
        print("checksum written by rank", (r - 1) % world, "into rank", r, ":", float(allr[r]), "local:", float(alll[(r - 1) % world]), flush=True)
for name, out in (("local", loc), ("peer", dst)):
    f = Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, 4095, H=H, fir_len=4095), cpr)[0]
    for _ in range(2): f.step(x, out=out)
    torch.cuda.synchronize(); hdl.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f.step(x, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(rank, f"filter with {name} output: {ms:.3f} ms = {cpr * n / ms / 1e6:.1f} Gsamples/s per GPU", flush=True)
    hdl.barrier()
dist.destroy_process_group()
