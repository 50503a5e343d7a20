"""BASELINE-size runs on the GPU checked through size-independent properties (the CPU oracle would need
minutes to hours at these sizes): a channel subset against the oracle, linearity, round trips, block
partition independence, counts.  Device-resident buffers, like bench.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def env():
    import torch
    import libtsd_b200
    libtsd_b200.init(0)
    return torch, libtsd_b200


def randc(torch, nchan, n, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    x = torch.empty((nchan, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).normal_(generator=g)
    return x


def rms(t):
    return float(t.abs().pow(2).mean().sqrt())


def test_fft_config2_full(env, cpu_oracle):
    """4096 x 65536 forward + inverse: unitary, round trip, and channels 0 / 2047 / 4095 against the oracle."""
    torch, tsd = env
    from libtsd_b200 import fourier as Fo
    x = randc(torch, 4096, 65536, 0x7D5D0002)
    plan = Fo.tfrplan_creation(65536, batch=4096)
    X = plan.step(x, True)
    x2 = plan.step(X, False)
    tsd.synchronize()
    assert abs(rms(X) / rms(x) - 1) < 1e-5                       # unitary (fourier.cc:119-120)
    assert rms(x2 - x) / rms(x) <= 5e-6                          # test-fourier.cc:287-312 bar
    ref = cpu_oracle.fft(65536)
    for c in (0, 2047, 4095):
        Xr = ref.step(x[c].cpu().numpy(), True)
        assert np.max(np.abs(X[c].cpu().numpy() - Xr)) / rms(x) <= TOL
    del X, x2


def test_fir_config3_full(env, cpu_oracle):
    """1024 ch x 1 Mi, 127 taps, 16 step() calls of 64 Ki: equals the one-shot call (state carried bit-exactly),
    and channels 0 / 1023 equal the oracle."""
    torch, tsd = env
    from libtsd_b200 import filtrage as F
    h = cpu_oracle.design_rif_fen(127, "lp", 0.1)
    nchan, n, blk = 1024, 1 << 20, 65536
    x = randc(torch, nchan, n, 0x7D5D0003)
    y = torch.empty_like(x)
    f = F.filtre_rif(h, np.complex64, nchan)
    for b in range(n // blk):
        f.step(x[:, b * blk:(b + 1) * blk], out=y[:, b * blk:(b + 1) * blk])
    y1 = F.filtre_rif(h, np.complex64, nchan).step(x)
    tsd.synchronize()
    assert f.index == n % 127
    assert torch.equal(y, y1)                                   # block partition independence, bit for bit
    for c in (0, 1023):
        yr = cpu_oracle.fir(1, h).step(x[c].cpu().numpy())
        assert np.max(np.abs(y[c].cpu().numpy() - yr)) / rms(x) <= TOL


def test_ola_config4_channels(env, cpu_oracle):
    """filtre_fft K = 4095, Ne = 61441 on 32 channels x 16 Mi (one GPU's share of config 4 on 8 GPUs):
    output count, delay Ne - K, linearity, and channel 31 against the oracle on its first 2 Mi samples."""
    torch, tsd = env
    from libtsd_b200 import fourier as Fo
    K, Ne, nchan, n = 4095, 61441, 32, 1 << 24
    h = cpu_oracle.design_rif_fen(K, "lp", 0.1)
    H = cpu_oracle.ola_make_H(h, 65536)
    cfg = Fo.FiltreFFTConfig(Ne, K, H=H, fir_len=K)
    x1, x2 = randc(torch, nchan, n, 0x7D5D0004), randc(torch, nchan, n, 99)
    f, N = Fo.filtre_fft(cfg, nchan)
    y1 = f.step(x1)
    y2 = Fo.filtre_fft(cfg, nchan)[0].step(x2)
    y12 = Fo.filtre_fft(cfg, nchan)[0].step(x1 + 2 * x2)
    tsd.synchronize()
    assert N == 65536 and y1.shape == (nchan, 273 * Ne) and f.residual == n - 273 * Ne == 3823
    assert float(y1[:, : Ne - K].abs().max()) <= TOL * rms(x1)                                  # delay Ne - K
    assert float((y12 - (y1 + 2 * y2)).abs().max()) / rms(x1) <= 3 * TOL                        # linearity
    m = 2 * 1024 * 1024
    r = cpu_oracle.ola(Ne, K, H)
    yr = r.step(x1[31, :m].cpu().numpy())
    assert np.max(np.abs(y1[31, : len(yr)].cpu().numpy() - yr)) / rms(x1) <= TOL
    # generic overlap-add form gives the same stream
    yg = Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, K, H=H, fir_len=0), nchan)[0].step(x1)
    tsd.synchronize()
    assert float((yg - y1).abs().max()) / rms(x1) <= TOL


def test_resampler_config5_channels(env, port, cpu_oracle):
    """147/160, 64 taps x 257 phases on 64 channels x 8 Mi: output count of SURVEY §8 a13 (7 707 034), same result
    when fed in 64 Ki blocks, channel 63 against the oracle on its first 1 Mi samples."""
    torch, tsd = env
    from libtsd_b200 import filtrage as F
    lut = cpu_oracle.itrp_sinc_lut(64, 256, 0.4)
    nchan, n = 64, 1 << 23
    x = randc(torch, nchan, n, 0x7D5D0005)
    f = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), nchan)
    y = f.step(x)
    tsd.synchronize()
    assert y.shape == (nchan, 7707034)
    g = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), nchan)
    parts = [g.step(x[:, i:i + 65536 * 16]) for i in range(0, n, 65536 * 16)]
    tsd.synchronize()
    assert torch.equal(torch.cat(parts, dim=1), y)
    m = 1 << 20
    yr = port.itrp(147.0 / 160.0, lut, 256).step(x[63, :m].cpu().numpy())
    assert np.max(np.abs(y[63, : len(yr)].cpu().numpy() - yr)) / rms(x) <= TOL


def test_pageable_host_buffers_match_device_path(env):
    """Host entry points with ordinary (pageable) numpy buffers go through the pinned bounce buffers and the copy-thread
    pool (runtime.cu: stage_in / stage_out) over several pipeline chunks; pinned buffers take the direct copies.  Both must
    give what the device-resident call gives: bit for bit where the arithmetic does not depend on how the call is cut into
    chunks (direct FIR, FFT plan), to rounding (<= 1e-5 of RMS, lengths exact) where the kernel's internal block grid moves
    with the chunk boundaries (block filter, tensor-core resampler)."""
    torch, tsd = env
    from libtsd_b200 import filtrage as F, fourier as Fo
    rng = np.random.default_rng(5)
    nchan, n = 4, 1 << 22          # 128 MiB of input: three pipeline chunks of 48 MiB
    x = (rng.standard_normal((nchan, n), dtype=np.float32) + 1j * rng.standard_normal((nchan, n), dtype=np.float32)).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    xp = torch.from_numpy(x).pin_memory()
    h = F.design_rif_fen(127, "lp", 0.1)
    # direct FIR: pageable, pinned, device; and in place on pageable memory
    yd = F.filtre_rif(h, np.complex64, nchan).step(xd).cpu().numpy()
    assert np.array_equal(F.filtre_rif(h, np.complex64, nchan).step(x), yd)
    yp = torch.empty((nchan, n), dtype=torch.complex64).pin_memory()
    F.filtre_rif(h, np.complex64, nchan).step(xp.numpy(), out=yp.numpy())
    assert np.array_equal(yp.numpy(), yd)
    xi = x.copy()
    F.filtre_rif(h, np.complex64, nchan).step(xi, out=xi)
    assert np.array_equal(xi, yd)
    # block filter (ragged output length per chunk) and resampler (fewer outputs than inputs)
    h2 = F.design_rif_fen(1023, "lp", 0.2)
    cfg = Fo.FiltreFFTConfig(15361, 1023, H=Fo.ola_make_H(h2, 16384), fir_len=1023)
    od = Fo.filtre_fft(cfg, nchan)[0].step(xd).cpu().numpy()
    oh = Fo.filtre_fft(cfg, nchan)[0].step(x)
    assert oh.shape == od.shape and np.max(np.abs(oh - od)) <= TOL * np.sqrt(np.mean(np.abs(od) ** 2))
    rd = F.filtre_reechan(147.0 / 160.0, nchan).step(xd).cpu().numpy()
    rh = F.filtre_reechan(147.0 / 160.0, nchan).step(x)
    assert rh.shape == rd.shape and np.max(np.abs(rh - rd)) <= TOL * np.sqrt(np.mean(np.abs(rd) ** 2))
    # FFT plan, groups of transforms
    xf = x.reshape(nchan * 64, 65536)
    Xd = Fo.tfrplan_creation(65536, batch=nchan * 64).step(torch.from_numpy(xf).cuda(), True).cpu().numpy()
    assert np.array_equal(Fo.tfrplan_creation(65536, batch=nchan * 64).step(xf, True), Xd)
