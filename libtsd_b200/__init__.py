"""libtsd_b200 — B200-native (sm_100a) implementation of libtsd's filtering hot path.

Scope: direct-form FIR (`filtre_rif`), batched complex FFT plans (`FFTPlan`),
FFT-domain block filtering (`filtre_fft` / OLA) and the LUT resampler (`filtre_itrp`), behind the
reference's own step() interface.  The compute lives in libtsdgpu.so (C ABI: include/tsdgpu.h);
this package is the thin host-side mirror of the reference's names.  No CPU fallback.
"""
from . import _lib, detection, filtrage, fourier, segments  # noqa: F401
from ._lib import TsdGpuError, init, launch_count, synchronize, use_torch_stream  # noqa: F401

__all__ = ["filtrage", "fourier", "detection", "segments", "TsdGpuError", "init", "synchronize", "launch_count", "use_torch_stream"]
