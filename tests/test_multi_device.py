"""One process, two devices (tsdgpu_init_devices / tsdgpu_set_device / tsdgpu_gather): objects created on the second device
give the bit-identical results of the first one (shared-memory opt-ins, constant tables and tensor maps are per device),
and the gather collects the shards on one device.  Skipped on a single-GPU box."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_second_device_in_one_process_and_gather(cpu_oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import libtsd_b200
    from libtsd_b200 import _lib, filtrage as F, fourier as Fo
    L = _lib.lib()
    devs = (C.c_int * 2)(0, 1)
    _lib.check(L.tsdgpu_init_devices(devs, 2))
    rng = np.random.default_rng(2)
    nchan, n = 130, 40000
    xh = (rng.standard_normal((nchan, n)) + 1j * rng.standard_normal((nchan, n))).astype(np.complex64)
    h127, h511 = cpu_oracle.design_rif_fen(127, "lp", 0.1), cpu_oracle.design_rif_fen(511, "lp", 0.1)
    lut = cpu_oracle.itrp_sinc_lut(64, 256, 0.4)
    res, keep = {}, {}
    for d in (0, 1):
        _lib.check(L.tsdgpu_set_device(d))
        torch.cuda.set_device(d)
        with torch.cuda.stream(torch.cuda.Stream(device=d)):
            x = torch.from_numpy(xh).cuda(d)
            y_fir = F.filtre_rif(h127, np.complex64, nchan).step(x).clone()            # tensor-core kernel, tensor maps
            y_long = F.filtre_rif(h511, np.complex64, nchan).step(x).clone()           # overlap-save kernel, delay 0
            y_rs = F.filtre_itrp(147 / 160, F.InterpolateurLUT(lut), nchan).step(x).clone()   # tcgen05 resampler, CTA pairs
            plan = Fo.tfrplan_creation(4096, batch=8)
            X = plan.step(x[:8, :4096].contiguous(), True).clone()
            libtsd_b200.synchronize()
            torch.cuda.synchronize(d)
        keep[d] = (y_fir, y_long, y_rs, X)
        res[d] = tuple(t.cpu() for t in keep[d])
    for a, b in zip(res[0], res[1]):
        assert a.shape == b.shape and torch.equal(a, b)
    # one channel against the oracle (the values are not just equal, they are right)
    ref = cpu_oracle.fir(1, h127).step(xh[5])
    assert np.max(np.abs(res[1][0][5].numpy() - ref)) <= 1e-5 * np.sqrt(np.mean(np.abs(xh) ** 2))
    # gather: shard 0 = first half of the channels from device 0, shard 1 = second half from device 1, onto device 0
    half = nchan // 2
    s0, s1 = keep[0][0][:half].contiguous(), keep[1][0][half:].contiguous()
    dst = torch.zeros((nchan, n), dtype=torch.complex64, device="cuda:0")
    offs = (C.c_longlong * 2)(0, s0.numel() * 8)
    srcs = (C.c_void_p * 2)(s0.data_ptr(), s1.data_ptr())
    sdev = (C.c_int * 2)(0, 1)
    nbytes = (C.c_longlong * 2)(s0.numel() * 8, s1.numel() * 8)
    _lib.check(L.tsdgpu_gather(C.c_void_p(dst.data_ptr()), 0, offs, srcs, sdev, nbytes, 2))
    assert torch.equal(dst.cpu(), res[0][0])
    _lib.check(L.tsdgpu_set_device(0))
    torch.cuda.set_device(0)
