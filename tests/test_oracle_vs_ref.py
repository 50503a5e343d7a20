"""Pins the C restatement (oracle/tsd_oracle.c) on the reference's OWN code compiled in place
(oracle/_ref/libtsdref.so).  Streaming paths must be bit-exact; skipped where the reference build
is absent (it needs /root/reference at build time)."""
import numpy as np
import pytest


def cn(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def test_p2_and_cost_model(port, ref):
    for i in list(range(1, 3000)) + [61441 + 4095, 65535, 65536, 65537, 1 << 20, (1 << 20) + 1, (1 << 21) + 1]:
        assert port.p2(i) == ref.p2(i)
    for M in (3, 127, 512, 2560, 4095):
        assert port.ola_complexite_optimise(M) == ref.ola_complexite_optimise(M)
    assert ref.ola_complexite_optimise(4095)[1:] == (65536, 4094, 61442)


@pytest.mark.parametrize("n,fc", [(31, 0.25), (127, 0.1), (4095, 0.1), (15, 0.25), (64, 0.3)])
def test_design_rif_fen(port, ref, n, fc):
    assert np.array_equal(port.design_rif_fen(n, "lp", fc), ref.design_rif_fen(n, "lp", fc))


def test_readme_spot_values(ref):
    h = ref.design_rif_fen(31, "lp", 0.25)   # BASELINE.md §2 spot values
    assert abs(h[0] - (-5.44303075e-05)) < 1e-11 and abs(h[15] - 0.499930501) < 1e-8 and abs(h.sum() - 1) < 1e-6


@pytest.mark.parametrize("K,P,fc", [(64, 256, 0.4), (15, 256, 0.4), (127, 256, 0.5), (16, 100, 0.3)])
def test_sinc_lut(port, ref, K, P, fc):
    assert np.array_equal(port.itrp_sinc_lut(K, P, fc), ref.itrp_sinc_lut(K, P, fc))


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("K", [1, 2, 31, 127])
def test_fir_bit_exact(port, ref, kind, K):
    rng = np.random.default_rng(K * 3 + kind)
    taps = cn(rng, K) if kind == 2 else rng.standard_normal(K).astype(np.float32)
    a, b = port.fir(kind, taps), ref.fir(kind, taps)
    for n in (1, 7, 100, 127, 128, 1000):
        x = rng.standard_normal(n).astype(np.float32) if kind == 0 else cn(rng, n)
        assert np.array_equal(a.step(x), b.step(x))


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 128, 1024, 4096, 65536])
def test_fft_bit_exact(port, ref, n):
    rng = np.random.default_rng(n)
    x = cn(rng, n)
    a, b = port.fft(n), ref.fft(n)
    X1, X2 = a.step(x), b.step(x)
    assert np.array_equal(X1, X2)
    assert np.array_equal(a.step(X1, False), b.step(X2, False))


@pytest.mark.parametrize("Ne,nz,K", [(1500, 127, 127), (0, 127, 127), (61441, 4095, 4095), (512, 0, 0)])
def test_ola_bit_exact(port, ref, Ne, nz, K):
    rng = np.random.default_rng(Ne + nz)
    N = ref.p2((Ne or 512) + nz)
    H = ref.ola_make_H(ref.design_rif_fen(K, "lp", 0.1), N) if K else None
    a, b = port.ola(Ne, nz, H), ref.ola(Ne, nz, H)
    assert a.N == b.N == N
    for n in (100, 1000, a.Ne, 5000, 3, 0, 70000):
        x = cn(rng, n)
        ya, yb = a.step(x), b.step(x)
        assert len(ya) == len(yb) and np.array_equal(ya, yb)


@pytest.mark.parametrize("Ne,nz,useH", [(512, 0, False), (512, 512, True), (100, 28, True), (1000, 24, True),
                                        (101, 27, True), (33, 31, False), (61440, 4096, True)])
def test_ola_fenetre_bit_exact(port, ref, Ne, nz, useH):
    """Hann-window 50 % overlap mode (fourier.cc:884-930), incl. odd Ne and the svg/x2 aliasing of the reference."""
    rng = np.random.default_rng(Ne * 7 + nz)
    N = ref.p2(Ne + nz)
    H = cn(rng, N) if useH else None
    assert np.array_equal(port.fenetre("hn", Ne, False), ref.fenetre("hn", Ne, False))
    a, b = port.ola(Ne, nz, H, True), ref.ola(Ne, nz, H, True)
    lens = []
    for n in (Ne, Ne, 37, 3 * Ne + 5, 1, Ne - 1, 2 * Ne):
        x = cn(rng, n)
        ya, yb = a.step(x), b.step(x, cap=8 * Ne + n)
        lens.append(len(yb))
        assert len(ya) == len(yb) and np.array_equal(ya, yb)
    assert lens[0] == 0 and lens[1] == Ne     # the first block of the stream emits nothing (:900-903)


def test_ola_make_H_close(port, ref):
    h = ref.design_rif_fen(4095, "lp", 0.1)
    Ha, Hb = port.ola_make_H(h, 65536), ref.ola_make_H(h, 65536)
    assert np.max(np.abs(Ha - Hb)) / np.max(np.abs(Hb)) < 1e-6   # complex plan vs the reference's rfft route


@pytest.mark.parametrize("ratio,K,fc", [(147 / 160, 64, 0.4), (1.5, 127, 0.5), (0.5, 15, 0.25), (1.9999, 15, 0.4),
                                        (np.pi / 2, 31, 0.4), (1.0, 15, 0.4)])
def test_itrp_bit_exact(port, ref, ratio, K, fc):
    rng = np.random.default_rng(K)
    lut = ref.itrp_sinc_lut(K, 256, fc)
    a, b = port.itrp(ratio, lut, 256), ref.itrp(ratio, K, 256, fc)
    for n in (1, 10, 1000, 65536, 0, 777):
        x = cn(rng, n)
        ya, yb = a.step(x), b.step(x)
        assert len(ya) == len(yb) and np.array_equal(ya, yb)


# ref driver numbering (ref_driver.cc:171-183): 0 demi-bande, 1 ups, 2 decim; port / C ABI: 0 ups, 1 demi-bande, 2 decim
_REF_KIND = {0: 1, 1: 0, 2: 2}


@pytest.mark.parametrize("kind,K,R", [(0, 15, 2), (0, 15, 3), (0, 16, 4), (0, 7, 5), (1, 15, 2), (1, 17, 2), (1, 31, 2), (1, 8, 2),
                                      (2, 15, 2), (2, 31, 3), (2, 5, 7), (2, 1, 2)])
def test_polyphase_bit_exact(port, ref, kind, K, R):
    """FiltreRIFUps / FiltreRIFDemiBande / FiltreRIFDecim (polyphase.cc): output counts per call and samples."""
    rng = np.random.default_rng(100 * kind + K + R)
    taps = rng.standard_normal(K).astype(np.float32) if K not in (15, 31) else ref.design_rif_fen(K, "lp", 0.25)
    a, b = port.polyphase(kind, taps, R), ref.polyphase(_REF_KIND[kind], taps, R)
    for n in (1, 2, 3, 10, 0, 1000, 7, 4096, 1):
        x = cn(rng, n)
        ya, yb = a.step(x), b.step(x, cap=n * max(R, 1) + 16)
        assert len(ya) == len(yb) and np.array_equal(ya, yb)


def test_reechan_chain_matches_reference(port, ref):
    """filtre_reechan<cfloat>(ratio) (ra.cc:104-177) rebuilt from the port's stages equals the reference object,
    per call, for ratios that need half-band / x2 stages."""
    rng = np.random.default_rng(5)
    for ratio in (0.1, 0.25, 0.3, 3.0, 4.0, 7.3, 2.0, 0.5, 1.0):
        nd, nu, post, fcut, use = port.reechan_plan(ratio)
        coefs = ref.design_rif_fen(15, "lp", 0.25)
        stages = [port.polyphase(1, coefs) for _ in range(nd)] + [port.polyphase(0, coefs, 2) for _ in range(nu)]
        itrp = port.itrp(post, ref.itrp_sinc_lut(15, 256, fcut), 256) if use else None
        r = ref.reechan(ratio)
        for n in (1000, 37, 4096, 1, 513):
            x = cn(rng, n)
            y = x
            if ratio != 1:
                for st in stages:
                    y = st.step(y)
                if itrp is not None:
                    y = itrp.step(y)
            yr = r.step(x, cap=int(n * max(ratio, 1) * 2 + 64))
            assert len(y) == len(yr) and np.array_equal(y, yr), ratio


def test_reference_quirks(ref):
    """filtre_rif_fft<cfloat> drops the imaginary part; delay Ne - K."""
    rng = np.random.default_rng(0)
    h = ref.design_rif_fen(127, "lp", 0.1)
    x = cn(rng, 8192)
    y = ref.rif_fft(1, h).step(x)
    assert np.max(np.abs(y.imag)) == 0
    yd = np.convolve(x.astype(np.complex128), h.astype(np.float64))[:8192]
    d = 512 - 127
    assert np.max(np.abs(y[d:].real - yd[: 8192 - d].real)) < 2e-6
    with pytest.raises(ValueError):
        ref.rif_fft(1, np.ones(600, np.float32))       # undefined behaviour in the reference for K > 512
    assert list(ref.tampon_trace(512, [100, 1000, 5000, 2092])) == [512] * 16   # test-tsd.cc:479-497
