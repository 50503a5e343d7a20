"""ctypes binding of libtsdgpu.so (the C ABI declared in include/tsdgpu.h).

The product path is the CUDA library; there is no CPU fallback.  If the shared object is
missing or no B200 is usable, every call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# TSDGPU_LIB selects another build of the same library (kernel A/B experiments under profiles/); default = the in-tree build
SO_PATH = os.environ.get("TSDGPU_LIB") or os.path.join(HERE, "libtsdgpu.so")
HOST, DEVICE = 0, 1

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class TsdGpuError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the reference raises `échec`, commun.hpp:152-163)."""


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], stdout=out)
    return SO_PATH


_SIGS = {
    "tsdgpu_init": (_i, [_i]),
    "tsdgpu_init_devices": (_i, [C.POINTER(_i), _i]),
    "tsdgpu_set_device": (_i, [_i]),
    "tsdgpu_current_device": (_i, []),
    "tsdgpu_shutdown": (_i, []),
    "tsdgpu_gather": (_i, [_vp, _i, C.POINTER(_ll), C.POINTER(_vp), C.POINTER(_i), C.POINTER(_ll), _i]),
    "tsdgpu_set_stream": (_i, [_vp]),
    "tsdgpu_synchronize": (_i, []),
    "tsdgpu_last_error": (C.c_char_p, []),
    "tsdgpu_launch_count": (_ll, [_i]),
    "tsdgpu_p2": (_i, [_i]),
    "tsdgpu_ola_complexite": (_i, [_i, _i, C.POINTER(_f), C.POINTER(_i), C.POINTER(_i)]),
    "tsdgpu_ola_complexite_optimise": (_i, [_i, C.POINTER(_f), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "tsdgpu_timing_enable": (_i, [_i]),
    "tsdgpu_timing_read": (_i, [C.POINTER(C.c_double), C.POINTER(_ll)]),
    "tsdgpu_fir_create": (_i, [_i, _vp, _i, _i, C.POINTER(_vp)]),
    "tsdgpu_fir_step": (_i, [_vp, _vp, _ll, _i, _vp, _ll, _i]),
    "tsdgpu_fir_get_state": (_i, [_vp, _vp, C.POINTER(_i)]),
    "tsdgpu_fir_set_state": (_i, [_vp, _vp, _i]),
    "tsdgpu_fir_set_history": (_i, [_vp, _vp, _ll]),
    "tsdgpu_fir_destroy": (_i, [_vp]),
    "tsdgpu_fft_plan": (_i, [_i, _i, C.POINTER(_vp)]),
    "tsdgpu_fft_exec": (_i, [_vp, _vp, _ll, _vp, _ll, _i, _i]),
    "tsdgpu_fft_destroy": (_i, [_vp]),
    "tsdgpu_ola_create": (_i, [_i, _i, _vp, _i, _i, C.POINTER(_vp)]),
    "tsdgpu_ola_create_fen": (_i, [_i, _i, _vp, _vp, _i, C.POINTER(_vp)]),
    "tsdgpu_ola_create_cb": (_i, [_i, _i, _vp, _vp, _vp, _i, C.POINTER(_vp)]),
    "tsdgpu_periodogramme_tfd": (_i, [_vp, C.c_longlong, _i, _i, _i, _vp, _vp, C.c_longlong, C.POINTER(_i), C.POINTER(_i), _i]),
    "tsdgpu_spectrum_create": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _i, C.POINTER(_vp)]),
    "tsdgpu_spectrum_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "tsdgpu_spectrum_step": (_i, [_vp, _vp, C.c_longlong, _i, _vp, C.c_longlong, C.POINTER(_i), _i]),
    "tsdgpu_spectrum_destroy": (_i, [_vp]),
    "tsdgpu_ola_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "tsdgpu_ola_out_count": (_ll, [_vp, _i]),
    "tsdgpu_ola_step": (_i, [_vp, _vp, _ll, _i, _vp, _ll, C.POINTER(_ll), _i]),
    "tsdgpu_ola_state_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "tsdgpu_ola_get_state": (_i, [_vp, C.POINTER(_i), C.POINTER(_ll), _vp, _vp, _vp]),
    "tsdgpu_ola_set_state": (_i, [_vp, _i, _ll, _vp, _vp, _vp]),
    "tsdgpu_ola_destroy": (_i, [_vp]),
    "tsdgpu_detect_create": (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    "tsdgpu_detect_dims": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_f)]),
    "tsdgpu_detect_step": (_i, [_vp, _vp, _ll, _i, _vp, _ll, _vp, _ll, _i]),
    "tsdgpu_detect_destroy": (_i, [_vp]),
    "tsdgpu_resamp_create": (_i, [_f, _vp, _i, _i, _i, C.POINTER(_vp)]),
    "tsdgpu_resamp_create_ex": (_i, [_f, _vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "tsdgpu_resamp_create_exact": (_i, [_f, _i, _i, _i, _i, C.POINTER(_vp)]),
    "tsdgpu_resamp_out_count": (_ll, [_vp, _i]),
    "tsdgpu_resamp_phase": (_f, [_vp]),
    "tsdgpu_resamp_step": (_i, [_vp, _vp, _ll, _i, _vp, _ll, _ll, C.POINTER(_ll), _i]),
    "tsdgpu_resamp_get_state": (_i, [_vp, C.POINTER(_f), _vp]),
    "tsdgpu_resamp_set_state": (_i, [_vp, _f, _vp]),
    "tsdgpu_resamp_destroy": (_i, [_vp]),
    "tsdgpu_poly_create": (_i, [_i, _vp, _i, _i, _i, _i, C.POINTER(_vp)]),
    "tsdgpu_poly_out_count": (_ll, [_vp, _i]),
    "tsdgpu_poly_state": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "tsdgpu_poly_step": (_i, [_vp, _vp, _ll, _i, _vp, _ll, C.POINTER(_ll), _i]),
    "tsdgpu_poly_get_state": (_i, [_vp, C.POINTER(_ll), C.POINTER(_i), _vp]),
    "tsdgpu_poly_set_state": (_i, [_vp, _ll, _i, _vp]),
    "tsdgpu_poly_destroy": (_i, [_vp]),
    "tsdgpu_resamp_schedule": (_i, [C.POINTER(_f), _f, _i, _i, _vp, _vp, _ll, C.POINTER(_ll)]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise TsdGpuError(
                f"{SO_PATH} is missing: build it with libtsd_b200._lib.build() / __graft_entry__.build(). "
                "There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise TsdGpuError((lib().tsdgpu_last_error() or b"?").decode("utf-8", "replace"))


def init(device: int = 0) -> None:
    check(lib().tsdgpu_init(device))


def use_torch_stream() -> None:
    """Enqueue the library's launches on torch's current CUDA stream."""
    import torch
    h = torch.cuda.current_stream().cuda_stream
    if not h:
        raise TsdGpuError("use_torch_stream: torch's current stream is the legacy default stream (handle 0), which the "
                          "C ABI reads as 'library stream'; make a torch.cuda.Stream() current first")
    check(lib().tsdgpu_set_stream(_vp(h)))


def bind_torch_stream() -> None:
    """Device-resident (torch) buffers: run the library on torch's CURRENT stream so that the call is ordered after the
    kernels that produced its inputs and before the consumers of its outputs (torch's legacy default stream has handle 0,
    which the C ABI reads as "library stream"; cudaStreamLegacy = 1 names it explicitly)."""
    import torch
    h = torch.cuda.current_stream().cuda_stream or 1
    check(lib().tsdgpu_set_stream(_vp(h)))


def synchronize() -> None:
    check(lib().tsdgpu_synchronize())


def launch_count(reset: bool = False) -> int:
    return int(lib().tsdgpu_launch_count(1 if reset else 0))


def timing_enable(on: bool = True) -> None:
    check(lib().tsdgpu_timing_enable(1 if on else 0))


def timing_read():
    """(total milliseconds, launches) of the dominant kernels since the last read."""
    ms, n = C.c_double(), _ll()
    check(lib().tsdgpu_timing_read(C.byref(ms), C.byref(n)))
    return ms.value, int(n.value)
