"""The reference's own acceptance test of its rate adapters (tests/test-ra.cc:11-146), restated as a checker:
a 2 kHz sine sampled at 100 kHz (1000 samples) goes through the adapter; the output must be a pure sine at
f2 / (ratio * fe): length within 1 % of ratio * n, >= 80 % of the windowed spectrum's energy in the three bins around
the expected frequency, strongest spur <= -50 dB, peak-to-peak amplitude within 10 %."""
import numpy as np


def hann_periodic(n):
    t = np.arange(n, dtype=np.float64) / n - 0.5
    return 0.5 + 0.5 * np.cos(2 * np.pi * t)


def sine_input(n=1000, fe=100e3, f2=2e3):
    t = np.arange(n, dtype=np.float32) / np.float32(fe)
    return np.sin(t * np.float32(2 * np.pi * f2)).astype(np.float32)


def verifie_sinus(x, f):
    """test-ra.cc:11-56 -> (score, max spur in dB)."""
    x = np.asarray(x, np.float64)
    n = len(x)
    X = np.abs(np.fft.fft(x * hann_periodic(n))[: n // 2]) ** 2
    idx = int(f * n)
    ef = X[idx] + (X[idx - 1] if idx > 0 else 0.0) + (X[idx + 1] if idx + 1 < n // 2 else 0.0)
    score = ef / X.sum()
    if idx >= 10:
        X[idx - 10: idx + 10] = 0
    return score, 10 * np.log10(X.max() / ef)


def check_adapter(step, ratio, max_spur_db=-50.0, fe=100e3, f2=2e3):
    """step: callable real float32 [n] -> real output.  Returns the measured figures; raises AssertionError like
    test_ra_unit (test-ra.cc:58-146)."""
    x = sine_input(1000, fe, f2)
    y = np.asarray(step(x))
    s_in, _ = verifie_sinus(x, f2 / fe)
    assert s_in >= 0.8
    score, spur = verifie_sinus(y, f2 / (ratio * fe))
    assert score >= 0.8, f"not a pure sine: score {score}"
    assert 100.0 * abs(len(y) - ratio * len(x)) / len(x) < 1, "output length"
    amp1, amp2 = x.max() - x.min(), y.max() - y.min()
    assert 100 * (amp1 - amp2) / amp1 < 10, "amplitude"
    assert spur <= max_spur_db, f"spur {spur:.1f} dB"
    return score, spur
