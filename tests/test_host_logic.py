"""Host-side mirror (design helpers, stage planner, bookkeeping) against the oracle — no GPU."""
import numpy as np
import pytest

from libtsd_b200 import filtrage as F
from libtsd_b200 import fourier as Fo


@pytest.mark.parametrize("n,fc", [(31, 0.25), (127, 0.1), (4095, 0.1), (15, 0.25), (64, 0.3)])
def test_design_rif_fen(port, n, fc):
    assert np.max(np.abs(F.design_rif_fen(n, "lp", fc) - port.design_rif_fen(n, "lp", fc))) <= 1e-7


@pytest.mark.parametrize("K,fc", [(64, 0.4), (15, 0.4), (127, 0.5)])
def test_itrp_sinc_lut(port, K, fc):
    it = F.itrp_sinc(F.InterpolateurSincConfig(K, 256, fc, "hn"))
    assert it.lut.shape == (257, K)
    assert np.max(np.abs(it.lut - port.itrp_sinc_lut(K, 256, fc))) <= 2e-7
    assert np.array_equal(it.coefs(0.5), it.lut[128])


def test_p2_and_make_H(port):
    assert Fo.prochaine_puissance_de_2(61441 + 4095) == 65536
    h = port.design_rif_fen(127, "lp", 0.1)
    assert np.max(np.abs(Fo.ola_make_H(h, 1024) - port.ola_make_H(h, 1024))) < 1e-5
