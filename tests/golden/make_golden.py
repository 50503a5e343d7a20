"""Generates tests/golden/*.npz by running the REFERENCE's own code (oracle/_ref/libtsdref.so, built in
place from /root/reference by `make -C oracle ref`).  Run in the authoring container only:

    python tests/golden/make_golden.py

Inputs come from numpy's PCG64 (`default_rng(seed)`, stable across numpy versions) and are stored next
to the outputs, so the fixtures are self-contained on the GPU box where /root/reference does not exist.
Large cases (N = 65536) store every STRIDE-th output sample instead of the full vector.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

STRIDE = 61


def cn(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def main():
    R = oracle.ref()
    out = {}

    # README example (config 1): design_rif_fen(31,"lp",0.25); 500 float samples
    h31 = R.design_rif_fen(31, "lp", 0.25)
    rng = np.random.default_rng(0x7D5D0001)
    x = (np.cos(2 * np.pi * 0.01 * np.arange(500)) + 0.1 * rng.standard_normal(500)).astype(np.float32)
    out["readme_h"] = h31
    out["readme_x"] = x
    out["readme_y"] = R.filtrer(h31, x)

    # taps / LUT of the benchmark configurations
    out["h127"] = R.design_rif_fen(127, "lp", 0.1)
    out["h4095"] = R.design_rif_fen(4095, "lp", 0.1)
    out["lut64"] = R.itrp_sinc_lut(64, 256, 0.4)
    out["lut15"] = R.itrp_sinc_lut(15, 256, 0.4)

    # streaming FIR, 127 taps, cf32, blocks 1000/24/3000
    rng = np.random.default_rng(0x7D5D0003)
    x = cn(rng, 4024)
    f = R.fir(1, out["h127"])
    out["fir_x"] = x
    out["fir_blocks"] = np.array([1000, 24, 3000], np.int32)
    out["fir_y"] = np.concatenate([f.step(x[:1000]), f.step(x[1000:1024]), f.step(x[1024:])])
    # complex taps
    tc = cn(rng, 33)
    out["firc_taps"] = tc
    out["firc_y"] = R.fir(2, tc).step(x[:500])

    # FFT plans
    for n in (8, 1024):
        rng = np.random.default_rng(n)
        x = cn(rng, n)
        p = R.fft(n)
        out[f"fft{n}_x"] = x
        out[f"fft{n}_X"] = p.step(x, True)
        out[f"fft{n}_xi"] = p.step(x, False)
    rng = np.random.default_rng(0x7D5D0002)
    x = cn(rng, 65536)
    p = R.fft(65536)
    X = p.step(x, True)
    out["fft65536_seed"] = np.array([0x7D5D0002], np.int64)
    out["fft65536_X_sub"] = X[::STRIDE]
    out["fft65536_rt_sub"] = p.step(X, False)[::STRIDE]

    # OLA small: Ne = 512 (default), K = 127, chunks 100/1000/512/5000
    rng = np.random.default_rng(77)
    x = cn(rng, 6612)
    H = R.ola_make_H(out["h127"], 1024)
    o = R.ola(0, 127, H)
    chunks = [100, 1000, 512, 5000]
    ys, lens, i = [], [], 0
    for c in chunks:
        y = o.step(x[i:i + c])
        i += c
        ys.append(y)
        lens.append(len(y))
    out["ola_small_x"] = x
    out["ola_small_H"] = H
    out["ola_small_chunks"] = np.array(chunks, np.int32)
    out["ola_small_lens"] = np.array(lens, np.int32)
    out["ola_small_y"] = np.concatenate(ys)

    # periodogramme_tfd (fourier.cc:1451-1481): N = 64 (N2 = 64) and N = 100 (N2 = 128, zero-padded frames)
    rng = np.random.default_rng(80)
    xp = cn(rng, 1000)
    out["pg_x"] = xp
    out["pg_w64"] = R.fenetre("hn", 64, False)
    out["pg_w100"] = R.fenetre("hn", 100, False)
    out["pg_M64"] = R.periodogramme_tfd(xp, 64)
    out["pg_M100"] = R.periodogramme_tfd(xp, 100)

    # rt_spectrum / Spectrum (fourier.cc:1162-1343): plain averaging, and sub-blocks + sweep + masks on a non power-of-two Nf
    rng = np.random.default_rng(81)
    for key, (BS, nmeans, nsubs, sweep, step, mbf, mhf, fen) in {"a": (512, 3, 1, False, 0, 0, 0, 1), "b": (1200, 2, 3, True, 250, 4, 7, 1)}.items():
        sp = R.spectrum(BS, nmeans, nsubs, sweep, step, mbf, mhf, fen)
        xs = cn(rng, 2 * nmeans * BS)
        ys = [sp.step(xs[i * BS:(i + 1) * BS]) for i in range(2 * nmeans)]
        out["sp_cfg_" + key] = np.array([BS, nmeans, nsubs, int(sweep), step, mbf, mhf, fen], np.int32)
        out["sp_x_" + key] = xs
        out["sp_lens_" + key] = np.array([len(y) for y in ys], np.int32)
        out["sp_y_" + key] = np.concatenate(ys)

    # rééchan_freq (fourier.cc:1391-1419): real and complex input, up and down
    rng = np.random.default_rng(79)
    xr = rng.standard_normal(1000).astype(np.float32)
    xc = cn(rng, 777)
    out["rfq_xr"] = xr
    out["rfq_xc"] = xc
    out["rfq_loms"] = np.array([1.5, 0.7, 2.0, 0.37], np.float32)
    for i, lom in enumerate(out["rfq_loms"]):
        out[f"rfq_yr{i}"] = R.reechan_freq(xr, float(lom))
        out[f"rfq_yc{i}"] = R.reechan_freq(xc, float(lom))

    # OLA, Hann-window 50 % overlap mode (fourier.cc:884-930): Ne = 512, N = 1024, random spectral gain
    rng = np.random.default_rng(78)
    x = cn(rng, 6000)
    H = cn(rng, 1024)
    o = R.ola(512, 512, H, True)
    chunks = [512, 100, 1000, 512, 3888]
    ys, lens, i = [], [], 0
    for c in chunks:
        y = o.step(x[i:i + c], cap=8192)
        i += c
        ys.append(y)
        lens.append(len(y))
    out["ola_fen_x"] = x
    out["ola_fen_H"] = H
    out["ola_fen_w"] = R.fenetre("hn", 512, False)
    out["ola_fen_chunks"] = np.array(chunks, np.int32)
    out["ola_fen_lens"] = np.array(lens, np.int32)
    out["ola_fen_y"] = np.concatenate(ys)

    # OLA config-4 shape: K = 4095, Ne = 61441, N = 65536, 200000 samples, sub-sampled output
    rng = np.random.default_rng(0x7D5D0004)
    x = cn(rng, 200000)
    H = R.ola_make_H(out["h4095"], 65536)
    o = R.ola(61441, 4095, H)
    lens = []
    ys = []
    for i in range(0, 200000, 65536):
        y = o.step(x[i:i + 65536])
        lens.append(len(y))
        ys.append(y)
    y = np.concatenate(ys)
    out["ola_big_seed"] = np.array([0x7D5D0004], np.int64)
    out["ola_big_H_sub"] = H[::STRIDE]
    out["ola_big_lens"] = np.array(lens, np.int32)
    out["ola_big_y_sub"] = y[::STRIDE]

    # resampler 147/160, sinc 64 x 257, blocks 2000/1/777
    rng = np.random.default_rng(0x7D5D0005)
    x = cn(rng, 2778)
    r = R.itrp(147.0 / 160.0, 64, 256, 0.4)
    ys = [r.step(x[:2000]), r.step(x[2000:2001]), r.step(x[2001:])]
    out["rs_x"] = x
    out["rs_blocks"] = np.array([2000, 1, 777], np.int32)
    out["rs_lens"] = np.array([len(v) for v in ys], np.int32)
    out["rs_y"] = np.concatenate(ys)
    # stock resample() (15 taps)
    out["rs15_y"] = R.reechan(147.0 / 160.0).step(x)
    # output counts of six 64 Ki blocks (60212,60211,60211,60211,60211,60212)
    r = R.itrp(147.0 / 160.0, 64, 256, 0.4)
    z = np.zeros(65536, np.complex64)
    out["rs_counts_64k"] = np.array([len(r.step(z)) for _ in range(6)], np.int32)

    # polyphase stages (polyphase.cc) on design_rif_fen(15, "lp", 0.25): x2 interpolator, half-band, decimator by 3,
    # blocks 1000/1/777 (ref driver kinds: 0 demi-bande, 1 ups, 2 decim)
    h15 = R.design_rif_fen(15, "lp", 0.25)
    out["h15"] = h15
    x = cn(np.random.default_rng(0x7D5D0006), 1778)
    out["poly_x"] = x
    out["poly_blocks"] = np.array([1000, 1, 777], np.int32)
    for name, kind, rr in (("ups2", 1, 2), ("demi", 0, 2), ("decim3", 2, 3)):
        f = R.polyphase(kind, h15, rr)
        ys = [f.step(x[:1000], cap=4096), f.step(x[1000:1001], cap=64), f.step(x[1001:], cap=4096)]
        out[f"poly_{name}_lens"] = np.array([len(v) for v in ys], np.int32)
        out[f"poly_{name}_y"] = np.concatenate(ys)
    # full resample() chains: ratio 0.1 (3 half-bands + interpolator 0.8) and 7.3 (2 x2 stages + interpolator 1.825)
    for name, ratio in (("r0p1", 0.1), ("r7p3", 7.3)):
        f = R.reechan(ratio)
        ys = [f.step(x[:1000], cap=16384), f.step(x[1000:1001], cap=64), f.step(x[1001:], cap=16384)]
        out[f"reechan_{name}_lens"] = np.array([len(v) for v in ys], np.int32)
        out[f"reechan_{name}_y"] = np.concatenate(ys)
    # rfft (RTFRPlan, fourier.cc:280-355)
    xr = np.random.default_rng(0x7D5D0007).standard_normal(1024).astype(np.float32)
    out["rfft_x"] = xr
    out["rfft_X"] = R.rfft(xr)

    # FFT plans for n not a power of two: even split (fourier.cc:438-462) and chirp-z (fourier.cc:237-255)
    for n in (12, 15, 1000, 12345):
        xx = cn(np.random.default_rng(1000 + n), n)
        pl = R.fft(n)
        out[f"fftnp{n}_x"] = xx
        out[f"fftnp{n}_X"] = pl.step(xx, True)
        out[f"fftnp{n}_xi"] = pl.step(xx, False)

    # integer bookkeeping
    out["p2_in"] = np.array([1, 2, 3, 5, 127, 512, 639, 1024, 1025, 65535, 65536, 65537, 61441 + 4095, 1 << 20], np.int32)
    out["p2_out"] = np.array([R.p2(int(v)) for v in out["p2_in"]], np.int32)
    c, nf, nz, ne = R.ola_complexite_optimise(4095)
    out["ola_opt_4095"] = np.array([nf, nz, ne], np.int32)
    out["ola_opt_4095_C"] = np.array([c], np.float32)
    out["tampon_512"] = R.tampon_trace(512, [100, 1000, 5000, 2092])

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
