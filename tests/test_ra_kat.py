"""Known-answer test of the reference for its rate adapters, as the reference writes it (tests/test-ra.cc:148-217): REAL
float data through filtre_itrp<float> with itrp_cspline<float>() and itrp_sinc<float>({127, 256, 0.5, "hn"}),
filtre_reechan<float>, and the <float,float> half-band, xR and /R polyphase stages; sine purity, length and amplitude.
CPU: the reference build itself through the restated checker (tests/ra_kat.py) - pins the checker.  GPU: the CUDA path
through the same checker, same types."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ra_kat  # noqa: E402

RATIOS = (1.0, 1.5, 0.5, 2.0, 1.2, float(np.float32(np.pi)))   # test-ra.cc:215


@pytest.fixture(scope="module")
def ref():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return oracle.ref()


def _cstep(f, **kw):
    return lambda x: np.real(f.step(x.astype(np.complex64), **kw))


def _rstep(f, **kw):
    return lambda x: f.step(x.astype(np.float32), **kw)


def test_reference_passes_its_own_kat(ref):
    for ratio in RATIOS:
        ra_kat.check_adapter(_rstep(ref.itrp2(ratio, "cspline", cplx=False), cap=9000), ratio)      # test-ra.cc:152-155
        ra_kat.check_adapter(_rstep(ref.itrp2(ratio, "sinc", cplx=False, K=127, nphases=256, fcut=0.5), cap=9000), ratio)
        ra_kat.check_adapter(_rstep(ref.reechan(ratio, cplx=False), cap=9000), ratio)               # test-ra.cc:162-164
        ra_kat.check_adapter(_cstep(ref.reechan(ratio), cap=9000), ratio)
        ra_kat.check_adapter(_cstep(ref.itrp(ratio, 127, 256, 0.5), cap=9000), ratio)
    h = ref.design_rif_fen(15, "lp", 0.25)
    ra_kat.check_adapter(_cstep(ref.polyphase(0, h), cap=9000), 0.5)          # ref driver kinds: 0 half-band, 1 xR, 2 /R
    ra_kat.check_adapter(_cstep(ref.polyphase(1, h, 2), cap=9000), 2.0)
    for R in (2, 3, 4, 5, 8):
        ra_kat.check_adapter(_cstep(ref.polyphase(2, ref.design_rif_fen(15, "lp", 0.5 / R), R), cap=9000), 1.0 / R)


@pytest.mark.gpu
def test_gpu_rate_adapters_sine_purity():
    import libtsd_b200
    from libtsd_b200 import filtrage as F
    libtsd_b200.init(0)
    f32 = np.float32
    for ratio in RATIOS:
        # as written in test-ra.cc:148-165: T = float, cspline / sinc(127, 256, 0.5) / filtre_reechan
        ra_kat.check_adapter(_rstep(F.filtre_itrp(ratio, F.itrp_cspline(), T=f32)), ratio)
        ra_kat.check_adapter(_rstep(F.filtre_itrp(ratio, F.itrp_sinc(F.InterpolateurSincConfig(127, 256, 0.5, "hn")), T=f32)), ratio)
        ra_kat.check_adapter(_rstep(F.filtre_reechan(ratio, T=f32)), ratio)
        # the cfloat instantiations (the fast kernels)
        ra_kat.check_adapter(_cstep(F.filtre_reechan(ratio)), ratio)
        ra_kat.check_adapter(_cstep(F.filtre_itrp(ratio, F.itrp_sinc(F.InterpolateurSincConfig(127, 256, 0.5, "hn")))), ratio)
    h = F.design_rif_fen(15, "lp", 0.25)
    ra_kat.check_adapter(_rstep(F.filtre_rif_demi_bande(h, f32)), 0.5)          # test-ra.cc:169-175 <float,float>
    ra_kat.check_adapter(_rstep(F.filtre_rif_ups(h, 2, f32)), 2.0)
    for R in (2, 3, 4, 5, 8):
        ra_kat.check_adapter(_rstep(F.filtre_rif_decim(F.design_rif_fen(15, "lp", 0.5 / R), R, f32)), 1.0 / R)
