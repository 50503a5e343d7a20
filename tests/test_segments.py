"""Halo split of a long single stream: planner and gather on the CPU (world_size-2 gloo), and on the
GPU every path split into segments that start from the halo (tsdgpu_*_set_state / set_history) against the one-shot
call: output lengths identical, samples within the parity bar."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from libtsd_b200.segments import plan_segments

TOL = 1e-5


def test_plan_segments():
    for n in (0, 1, 5, 1000, 123457):
        for nseg in (1, 2, 3, 8):
            for align in (1, 61, 4096):
                sp = plan_segments(n, nseg, align)
                assert len(sp) == nseg and sp[0][0] == 0 and sp[-1][1] == n
                for (a0, a1), (b0, _) in zip(sp, sp[1:]):
                    assert a0 <= a1 == b0
                    assert a1 % align == 0 or a1 == n


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from libtsd_b200.segments import gather_segments
    P = oracle.port()
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    h = P.design_rif_fen(31, "lp", 0.25)
    s, e = plan_segments(n, world)[rank]
    # the segment's samples as a halo-started object produces them: y[s:e] of the stream filter
    halo = max(0, s - 30)
    y_local = P.fir(1, h).step(x[halo:e])[s - halo:][None]
    y = gather_segments(torch.from_numpy(y_local)).numpy()
    if rank == 0:
        y_ref = P.fir(1, h).step(x)[None]
        q.put((y.shape == y_ref.shape, float(np.max(np.abs(y - y_ref)))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_segment_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1001, q)) for r in range(2)]
    for p in procs:
        p.start()
    same_shape, err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same_shape and err == 0.0      # K-1 samples of halo reproduce the stream filter exactly


# ------------------------------------------------------------------------------------------------ GPU
def _cn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(1e-30, np.sqrt(np.mean(np.abs(b) ** 2))))


@pytest.fixture(scope="module")
def tsd():
    import libtsd_b200
    libtsd_b200.init(0)
    return libtsd_b200


@pytest.mark.gpu
@pytest.mark.parametrize("nseg", [2, 5])
def test_split_equals_one_shot_every_path(tsd, cpu_oracle, nseg):
    from libtsd_b200 import filtrage as F, fourier as Fo, segments as S
    rng = np.random.default_rng(nseg)
    nchan, n = 2, 300007
    x = _cn(rng, nchan, n)
    # direct FIR
    h = cpu_oracle.design_rif_fen(127, "lp", 0.1)
    mk = lambda: F.filtre_rif(h, np.complex64, nchan)          # noqa: E731
    y1, ys = mk().step(x), S.run_split(mk, S.start_fir, x, nseg)
    assert ys.shape == y1.shape and _rel(ys, y1) <= TOL
    # FFT-domain filter, BASELINE config-4 shape and the reference's default shape
    for K, Ne in ((4095, 61441), (127, 512)):
        hk = cpu_oracle.design_rif_fen(K, "lp", 0.1)
        N = cpu_oracle.p2(Ne + K)
        H = Fo.ola_make_H(hk, N)
        mk = lambda: Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, K, H=H, fir_len=K), nchan)[0]     # noqa: E731
        y1 = mk().step(x)
        for align in (1, Ne):
            ys = S.run_split(mk, S.start_ola, x, nseg, align)
            assert ys.shape == y1.shape                      # same number of emitted blocks: bit-exact bookkeeping
            assert _rel(ys, y1) <= TOL
    # LUT resampler: the segment start phase comes from the host schedule
    for ratio in (147 / 160, 1.37):
        mk = lambda: F.filtre_itrp(ratio, F.itrp_sinc(F.InterpolateurSincConfig(64, 256, 0.4, "hn")), nchan)    # noqa: E731
        y1, ys = mk().step(x), S.run_split(mk, S.start_itrp, x, nseg)
        assert ys.shape == y1.shape and _rel(ys, y1) <= TOL
    # polyphase stages
    h15 = cpu_oracle.design_rif_fen(15, "lp", 0.25)
    for mk in (lambda: F.filtre_rif_decim(h15, 3, np.complex64, nchan), lambda: F.filtre_rif_demi_bande(h15, np.complex64, nchan),
               lambda: F.filtre_rif_ups(h15, 2, np.complex64, nchan)):
        y1, ys = mk().step(x), S.run_split(mk, S.start_polyphase, x, nseg)
        assert ys.shape == y1.shape and _rel(ys, y1) <= TOL


@pytest.mark.gpu
def test_long_single_stream_split(tsd, cpu_oracle):
    """One stream of 32 Mi samples, device-resident, K = 4095 / Ne = 61441: 8 halo-started segments == one-shot, and the
    first block against the oracle."""
    from libtsd_b200 import fourier as Fo, segments as S
    n = 1 << 25
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.empty((1, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).normal_(generator=g)
    K, Ne = 4095, 61441
    h = cpu_oracle.design_rif_fen(K, "lp", 0.1)
    H = Fo.ola_make_H(h, 65536)
    mk = lambda: Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, K, H=H, fir_len=K), 1)[0]        # noqa: E731
    y1 = mk().step(x)
    ys = S.run_split(mk, S.start_ola, x, 8, Ne)
    tsd.synchronize()
    assert ys.shape == y1.shape == (1, Ne * (n // Ne))
    rms = float(y1.abs().pow(2).mean().sqrt())
    assert float((ys - y1).abs().max()) / rms <= TOL
    ref = cpu_oracle.ola(Ne, K, cpu_oracle.ola_make_H(h, 65536)).step(x[0, : 3 * Ne].cpu().numpy())
    assert np.max(np.abs(ys[0, : len(ref)].cpu().numpy() - ref)) / rms <= TOL
