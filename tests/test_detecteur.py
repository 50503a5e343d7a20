"""Correlation detector (reference src/fourier/detection.cc, tests/test-detecteur.cc pattern: a motif buried in
noise at known positions): GPU score signal and detections against the reference object, block by block."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tsd():
    import libtsd_b200
    libtsd_b200.init(0)
    return libtsd_b200


def _stream(rng, n, motif, positions, gains, noise):
    x = (noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    for p, g in zip(positions, gains):
        x[p: p + len(motif)] += np.complex64(g) * motif
    return x


def _compare(tsd, ref, motif, Ne, x, blocks, seuil=0.5):
    from libtsd_b200 import detection as D
    got = []
    g = D.detecteur_creation(D.DetecteurConfig(Ne=Ne, motif=motif, seuil=seuil, gere_detection=got.append))
    r = ref.detecteur(motif, Ne=Ne, seuil=seuil)
    pos, ndet = 0, 0
    for nb in blocks:
        n = nb * g.Ne
        xb = x[pos: pos + n]
        pos += n
        s_ref, d_ref = r.step(xb)
        s = g.step(xb)
        assert s.shape == s_ref.shape
        assert np.max(np.abs(s - s_ref)) <= 2e-5                     # scores live in [0, 1]
        d = g.detections
        assert len(d) == len(d_ref), (len(d), len(d_ref))
        for a, b in zip(d, d_ref):
            assert a.position == int(b["position"])
            assert abs(a.position_prec - b["position_prec"]) <= 2e-3
            assert abs(a.score - b["score"]) <= 2e-5
            assert abs(a.gain - b["gain"]) <= 1e-3 * max(1.0, abs(b["gain"]))
            assert abs(np.angle(np.exp(1j * (a.theta - b["theta"])))) <= 1e-3
            assert abs(a.sigma_noise - b["sigma_noise"]) <= 1e-3 * max(1e-3, b["sigma_noise"]) + 1e-6
            assert abs(a.SNR_dB - b["SNR_dB"]) <= 0.02
        ndet += len(d)
    assert len(got) == ndet                                         # the callback saw every detection
    return ndet, g


def test_detecteur_short_motif(tsd, ref):
    rng = np.random.default_rng(1)
    M, Ne = 128, 1024
    motif = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
    # inside a block, across a block border, and peaks landing on the first / last sample of a call (border cases of
    # detection.cc:352-384): the peak of a motif starting at p sits at output p + Ne
    positions = [700, 1024 - 60, 3 * 1024 - 1 - 1024 + 0, 5 * 1024 - 1024, 9000]
    gains = [0.8 * np.exp(0.3j), 1.5 * np.exp(-1.1j), 0.5, 1.0j, 2.0]
    x = _stream(rng, 16 * Ne, motif, positions, gains, 0.05)
    n, g = _compare(tsd, ref, motif, Ne, x, [1, 1, 1, 2, 1, 3, 1, 4, 2])
    assert n == len(positions) and (g.Ne, g.N, g.M) == (1024, 2048, 128)


def test_detecteur_auto_block_and_config4_size(tsd, ref):
    """M = 4095 with the reference's own choice of block (ola_complexité_optimise: Ne = 61442, N = 65536)."""
    rng = np.random.default_rng(2)
    M = 4095
    motif = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
    Ne = 61442
    x = _stream(rng, 4 * Ne, motif, [30000, 100000, 2 * Ne - 2000], [1.0, 0.3 * np.exp(2j), 0.7], 0.2)
    n, g = _compare(tsd, ref, motif, 0, x, [1, 2, 1])
    assert n == 3 and (g.Ne, g.N) == (Ne, 65536)


def test_detecteur_long_motif_generic_path(tsd, ref):
    """M - 1 > 8192: the correlator falls back to the N-point block filter with the gains conj(fft(motif)) as data."""
    rng = np.random.default_rng(3)
    M, Ne = 9000, 23769                                   # N = 32768
    motif = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
    x = _stream(rng, 4 * Ne, motif, [5000, 40000], [1.0, 0.6j], 0.3)
    n, g = _compare(tsd, ref, motif, Ne, x, [2, 2])
    assert n == 2 and g.N == 32768


def test_detecteur_errors(tsd):
    from libtsd_b200 import detection as D
    with pytest.raises(tsd.TsdGpuError):
        D.detecteur_creation(D.DetecteurConfig(Ne=1000, motif=np.zeros(64, np.complex64)))        # null motif
    g = D.detecteur_creation(D.DetecteurConfig(Ne=1024, motif=np.ones(64, np.complex64)))
    with pytest.raises(tsd.TsdGpuError, match="multiple of Ne"):
        g.step(np.zeros(1000, np.complex64))
