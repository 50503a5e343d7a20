"""The C restatement against the committed golden vectors (tests/golden/reference_vectors.npz, produced by
the reference's own code through tests/golden/make_golden.py).  Runs everywhere, needs no GPU and no
/root/reference."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
STRIDE = 61


def cn(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def test_integers(port):
    assert [port.p2(int(v)) for v in G["p2_in"]] == list(G["p2_out"])
    c, nf, nz, ne = port.ola_complexite_optimise(4095)
    assert [nf, nz, ne] == list(G["ola_opt_4095"]) and np.float32(c) == G["ola_opt_4095_C"][0]
    w, blocks = 0, []
    for n in (100, 1000, 5000, 2092):   # TamponNv2 (tsd.cc:332-370) bookkeeping
        tot = w + n
        blocks += [512] * (tot // 512)
        w = tot % 512
    assert blocks == list(G["tampon_512"])


def test_design(port):
    assert np.array_equal(port.design_rif_fen(31, "lp", 0.25), G["readme_h"])
    assert np.array_equal(port.design_rif_fen(127, "lp", 0.1), G["h127"])
    assert np.array_equal(port.design_rif_fen(4095, "lp", 0.1), G["h4095"])
    assert np.array_equal(port.itrp_sinc_lut(64, 256, 0.4), G["lut64"])
    assert np.array_equal(port.itrp_sinc_lut(15, 256, 0.4), G["lut15"])


def test_readme_example(port):
    assert np.array_equal(port.fir(0, G["readme_h"]).step(G["readme_x"]), G["readme_y"])
    assert abs(G["readme_y"][0] - (-5.30200559e-05)) > 0  # value depends on the RNG; see make_golden.py


def test_fir(port):
    f = port.fir(1, G["h127"])
    x, i, ys = G["fir_x"], 0, []
    for n in G["fir_blocks"]:
        ys.append(f.step(x[i:i + n]))
        i += n
    assert np.array_equal(np.concatenate(ys), G["fir_y"])
    assert np.array_equal(port.fir(2, G["firc_taps"]).step(x[:500]), G["firc_y"])


def test_fft(port):
    for n in (8, 1024):
        p = port.fft(n)
        assert np.array_equal(p.step(G[f"fft{n}_x"], True), G[f"fft{n}_X"])
        assert np.array_equal(p.step(G[f"fft{n}_x"], False), G[f"fft{n}_xi"])
    x = cn(np.random.default_rng(int(G["fft65536_seed"][0])), 65536)
    p = port.fft(65536)
    X = p.step(x, True)
    assert np.array_equal(X[::STRIDE], G["fft65536_X_sub"])
    assert np.array_equal(p.step(X, False)[::STRIDE], G["fft65536_rt_sub"])


def test_ola_small(port):
    o = port.ola(0, 127, G["ola_small_H"])
    x, i, ys, lens = G["ola_small_x"], 0, [], []
    for n in G["ola_small_chunks"]:
        y = o.step(x[i:i + n])
        i += n
        ys.append(y)
        lens.append(len(y))
    assert lens == list(G["ola_small_lens"])
    assert np.array_equal(np.concatenate(ys), G["ola_small_y"])


def test_ola_fenetre(port):
    assert np.array_equal(port.fenetre("hn", 512, False), G["ola_fen_w"])
    o = port.ola(512, 512, G["ola_fen_H"], True)
    x, i, ys, lens = G["ola_fen_x"], 0, [], []
    for n in G["ola_fen_chunks"]:
        y = o.step(x[i:i + n])
        i += n
        ys.append(y)
        lens.append(len(y))
    assert lens == list(G["ola_fen_lens"])
    assert np.array_equal(np.concatenate(ys), G["ola_fen_y"])


def test_ola_big(port):
    x = cn(np.random.default_rng(int(G["ola_big_seed"][0])), 200000)
    H = port.ola_make_H(G["h4095"], 65536)
    assert np.max(np.abs(H[::STRIDE] - G["ola_big_H_sub"])) / np.max(np.abs(H)) < 1e-6
    o = port.ola(61441, 4095, H)
    ys, lens = [], []
    for i in range(0, 200000, 65536):
        y = o.step(x[i:i + 65536])
        lens.append(len(y))
        ys.append(y)
    assert lens == list(G["ola_big_lens"])
    y = np.concatenate(ys)[::STRIDE]
    # H comes from the port's complex plan here (the reference used its rfft route): rounding-level difference
    assert np.max(np.abs(y - G["ola_big_y_sub"])) < 5e-6


def test_resampler(port):
    r = port.itrp(147.0 / 160.0, G["lut64"], 256)
    x, i, ys = G["rs_x"], 0, []
    for n in G["rs_blocks"]:
        ys.append(r.step(x[i:i + n]))
        i += n
    assert [len(v) for v in ys] == list(G["rs_lens"])
    assert np.array_equal(np.concatenate(ys), G["rs_y"])
    nd, nu, post, fcut, use = port.reechan_plan(147.0 / 160.0)
    assert (nd, nu, use) == (0, 0, True) and np.float32(fcut) == np.float32(0.4)
    assert np.array_equal(port.itrp(post, G["lut15"], 256).step(x), G["rs15_y"])
    r = port.itrp(147.0 / 160.0, G["lut64"], 256)
    z = np.zeros(65536, np.complex64)
    assert [len(r.step(z)) for _ in range(6)] == list(G["rs_counts_64k"]) == [60212, 60211, 60211, 60211, 60211, 60212]


def test_polyphase_and_chains(port):
    """polyphase.cc stages and full filtre_reechan chains against vectors produced by the reference build."""
    x, blocks = G["poly_x"], list(G["poly_blocks"])
    for name, kind, rr in (("ups2", 0, 2), ("demi", 1, 2), ("decim3", 2, 3)):   # port kinds: 0 ups, 1 demi-bande, 2 decim
        f = port.polyphase(kind, G["h15"], rr)
        ys, i = [], 0
        for n in blocks:
            ys.append(f.step(x[i:i + n]))
            i += n
        assert [len(v) for v in ys] == list(G[f"poly_{name}_lens"])
        assert np.array_equal(np.concatenate(ys), G[f"poly_{name}_y"])
    for name, ratio in (("r0p1", 0.1), ("r7p3", 7.3)):
        nd, nu, post, fcut, use = port.reechan_plan(ratio)
        stages = [port.polyphase(1, G["h15"]) for _ in range(nd)] + [port.polyphase(0, G["h15"], 2) for _ in range(nu)]
        itrp = port.itrp(post, port.itrp_sinc_lut(15, 256, fcut), 256) if use else None
        ys, i = [], 0
        for n in blocks:
            y = x[i:i + n]
            i += n
            for st in stages:
                y = st.step(y)
            ys.append(itrp.step(y) if itrp is not None else y)
        assert [len(v) for v in ys] == list(G[f"reechan_{name}_lens"])
        assert np.array_equal(np.concatenate(ys), G[f"reechan_{name}_y"])
