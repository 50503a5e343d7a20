"""The C-ABI library loads without a GPU and exports every symbol include/tsdgpu.h declares;
host-only entry points work; device entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tsdgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsdgpu_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header():
    from libtsd_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tsdgpu.h but not exported"
    assert sorted(_lib.EXPORTS) == names, "python binding table out of sync with the header"


def test_host_only_entries(port):
    from libtsd_b200 import _lib
    L = _lib.lib()
    for i in (1, 2, 3, 100, 65536, 65537, 61441 + 4095):
        assert L.tsdgpu_p2(i) == port.p2(i)
    # cost model of the block filter (fourier.cc:708-735)
    from libtsd_b200 import fourier as Fo
    for M in (3, 127, 512, 2560, 4095):
        assert Fo.ola_complexite_optimise(M) == port.ola_complexite_optimise(M)
    assert Fo.ola_complexite_optimise(4095)[1:] == (65536, 4094, 61442)
    # resampler schedule == the oracle's recurrence, for several ratios and block partitions
    for ratio in (147 / 160, 1.5, 0.5, 1.9999, 3.7, 0.3, 1.0, 0.999999, 0.50001, 0.75, 2 / 3, 0.6180339, 0.97, 0.51, 44100 / 48000):
        phase_g, phase_o = ctypes.c_float(0), 0.0
        for n in (1, 10, 1000, 65536, 777, 300001):
            cap = int(np.ceil(np.float32(ratio) * n) + 10)
            a, b = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            no = ctypes.c_longlong()
            rc = L.tsdgpu_resamp_schedule(ctypes.byref(phase_g), ctypes.c_float(ratio), 256, n,
                                          a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(no))
            assert rc == 0
            ia, ib, phase_o = port.itrp_schedule(phase_o, ratio, 256, n)
            assert no.value == len(ia)
            assert np.array_equal(a[: no.value], ia) and np.array_equal(b[: no.value], ib)
            assert np.float32(phase_g.value) == np.float32(phase_o)


def test_resampler_schedule_integer_runs(port):
    """Ratios in [8/9, 1) take the integer-exact run form of the phase recurrence inside the library
    (resamp.cu:resamp_schedule_runs): it must reproduce the reference's float32 loop bit for bit, for any chunking of
    the stream (carried phase) and any number of phases."""
    from libtsd_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(0)
    ratios = [147 / 160, 0.999999, 0.97, 44100 / 48000, 8 / 9, 0.8889, 0.9000001, 0.99999994] + list(rng.uniform(0.8889, 0.99999, 12))
    for nph in (256, 100, 7):
        for ratio in ratios:
            phase_g, phase_o = ctypes.c_float(0), 0.0
            for n in [1, 7, 1000, 65536, 777, 300001, 12, 13, 11, 5] + [int(v) for v in rng.integers(1, 5000, 6)]:
                cap = int(np.ceil(np.float32(ratio) * n) + 32)
                a, b = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
                no = ctypes.c_longlong()
                rc = L.tsdgpu_resamp_schedule(ctypes.byref(phase_g), ctypes.c_float(ratio), nph, n,
                                              a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(no))
                assert rc == 0
                ia, ib, phase_o = port.itrp_schedule(phase_o, ratio, nph, n)
                assert no.value == len(ia)
                assert np.array_equal(a[: no.value], ia) and np.array_equal(b[: no.value], ib)
                assert np.float32(phase_g.value) == np.float32(phase_o)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import libtsd_b200
    from libtsd_b200 import filtrage as F
    with pytest.raises(libtsd_b200.TsdGpuError, match="no CUDA device|no CPU fallback"):
        F.filtre_rif(np.ones(3, np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "libtsd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "oracle/" not in txt and "tsdo_" not in txt, f
