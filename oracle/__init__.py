"""TEST INFRASTRUCTURE ONLY — ctypes access to the two CPU oracles.

* ``port``  : oracle/libtsd_oracle.so, the plain-C restatement (oracle/tsd_oracle.c).
* ``ref``   : oracle/_ref/libtsdref.so, the reference's own sources compiled in place from
              /root/reference (oracle/Makefile).  Present only when it was built in the
              authoring container; the .so travels to the GPU box with the snapshot.

Nothing under ``libtsd_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` do.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libtsd_oracle.so")
# TSDREF_LIB selects another build of the same reference sources (bench.py: the AVX2 timing build, oracle/Makefile `native`)
REF_SO = os.environ.get("TSDREF_LIB") or os.path.join(HERE, "_ref", "libtsdref.so")

_vp = C.c_void_p
_f = C.c_float
_i = C.c_int


def build(verbose: bool = False) -> None:
    """Compile the C restatement and, when /root/reference is mounted, the reference build."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "-j8", "all"], stdout=out)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def _c64(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.complex64)


def _f32(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float32)


# --------------------------------------------------------------------------- port
class _Port:
    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        L = C.CDLL(PORT_SO)
        for name in ("tsdo_fir_new", "tsdo_fft_new", "tsdo_ola_new", "tsdo_ola_new2", "tsdo_itrp_new", "tsdo_poly_new"):
            getattr(L, name).restype = _vp
        L.tsdo_itrp_phase.restype = _f
        self.L = L

    # -- integers / design
    def p2(self, i: int) -> int:
        return self.L.tsdo_p2(_i(i))

    def ola_complexite_optimise(self, M: int):
        c, nf, nz, ne = _f(), _i(), _i(), _i()
        self.L.tsdo_ola_complexite_optimise(_i(M), C.byref(c), C.byref(nf), C.byref(nz), C.byref(ne))
        return c.value, nf.value, nz.value, ne.value

    def design_rif_fen(self, n: int, typ: str, fc: float, win: str = "hn") -> np.ndarray:
        assert typ in ("lp", "pb")
        h = np.zeros(n, np.float32)
        rc = self.L.tsdo_design_rif_fen_lp(_i(n), _f(fc), win.encode(), _ptr(h))
        if rc:
            raise RuntimeError("tsdo_design_rif_fen_lp failed")
        return h

    def itrp_sinc_lut(self, K: int, nphases: int, fcut: float, win: str = "hn") -> np.ndarray:
        lut = np.zeros((nphases + 1, K), np.float32)
        self.L.tsdo_itrp_sinc_lut(_i(K), _i(nphases), _f(fcut), win.encode(), _ptr(lut))
        return lut

    def reechan_plan(self, ratio: float):
        nd, nu, post, fcut, use = _i(), _i(), _f(), _f(), _i()
        self.L.tsdo_reechan_plan(_f(ratio), C.byref(nd), C.byref(nu), C.byref(post), C.byref(fcut), C.byref(use))
        return nd.value, nu.value, post.value, fcut.value, bool(use.value)

    def ola_make_H(self, h, N: int) -> np.ndarray:
        h = _f32(h)
        H = np.zeros(N, np.complex64)
        if self.L.tsdo_ola_make_H(_ptr(h), _i(len(h)), _i(N), _ptr(H)):
            raise RuntimeError("tsdo_ola_make_H failed")
        return H

    # -- streaming objects
    def fir(self, kind: int, taps):
        return _PortFir(self.L, kind, taps)

    def fft(self, n: int):
        return _PortFft(self.L, n)

    def ola(self, Ne: int, nb_zeros_min: int, H=None, avec_fenetrage: bool = False):
        return _PortOla(self.L, Ne, nb_zeros_min, H, avec_fenetrage)

    def fenetre(self, nom: str, n: int, sym: bool = True) -> np.ndarray:
        w = np.empty(n, np.float32)
        if self.L.tsdo_fenetre(nom.encode(), _i(n), _i(1 if sym else 0), _ptr(w)):
            raise ValueError("window not available: " + nom)
        return w

    def itrp(self, ratio: float, lut: np.ndarray, nphases: int):
        return _PortItrp(self.L, ratio, lut, nphases)

    def polyphase(self, kind: int, taps, R: int = 2, cplx: bool = True):
        """kind 0: filtre_rif_ups(c, R), 1: filtre_rif_demi_bande(c), 2: filtre_rif_decim(c, R) (polyphase.cc)."""
        return _PortPoly(self.L, kind, taps, R, cplx)

    def itrp_schedule(self, phase: float, ratio: float, nphases: int, n: int):
        cap = int(np.ceil(np.float32(ratio) * n) + 10)
        a = np.zeros(cap, np.int32)
        b = np.zeros(cap, np.int32)
        ph = _f(phase)
        no = _i()
        rc = self.L.tsdo_itrp_schedule(C.byref(ph), _f(ratio), _i(nphases), _i(n), _ptr(a), _ptr(b), _i(cap), C.byref(no))
        if rc:
            raise RuntimeError("tsdo_itrp_schedule overflow")
        return a[: no.value].copy(), b[: no.value].copy(), ph.value


class _PortFir:
    def __init__(self, L, kind, taps):
        self.L, self.kind = L, kind
        t = _c64(taps) if kind == 2 else _f32(taps)
        self.K = len(t)
        self.h = _vp(L.tsdo_fir_new(_i(kind), _ptr(t), _i(self.K)))
        if not self.h:
            raise RuntimeError("tsdo_fir_new failed (K must be > 0)")

    def step(self, x):
        x = _f32(x) if self.kind == 0 else _c64(x)
        y = np.empty_like(x)
        self.L.tsdo_fir_step(self.h, _ptr(x), _i(len(x)), _ptr(y))
        return y

    @property
    def index(self):
        return self.L.tsdo_fir_index(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdo_fir_free(self.h)


class _PortFft:
    def __init__(self, L, n):
        self.L, self.n = L, n
        self.h = _vp(L.tsdo_fft_new(_i(n)))
        if not self.h:
            raise RuntimeError("tsdo_fft_new: n must be a power of two")

    def step(self, x, forward=True):
        x = _c64(x)
        assert len(x) == self.n
        y = np.empty_like(x)
        self.L.tsdo_fft_step(self.h, _ptr(x), _i(1 if forward else 0), _ptr(y))
        return y

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdo_fft_free(self.h)


class _PortOla:
    def __init__(self, L, Ne, nzmin, H, avec_fenetrage=False):
        self.L = L
        self.fen = bool(avec_fenetrage)
        Hc = None if H is None else _c64(H)
        self.h = _vp(L.tsdo_ola_new2(_i(Ne), _i(nzmin), None if Hc is None else _ptr(Hc), _i(1 if self.fen else 0)))
        if not self.h:
            raise RuntimeError("tsdo_ola_new failed")
        a, b, c, d = _i(), _i(), _i(), _i()
        L.tsdo_ola_dims(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        self.Ne, self.N, self.Nz = a.value, b.value, c.value
        if Hc is not None and len(Hc) != self.N:
            raise ValueError("H must have N bins")

    @property
    def residual(self):
        a, b, c, d = _i(), _i(), _i(), _i()
        self.L.tsdo_ola_dims(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return d.value

    def step(self, x):
        x = _c64(x)
        cap = self.Ne * ((self.residual + len(x)) // self.Ne)
        y = np.empty(cap, np.complex64)
        no = _i()
        if self.L.tsdo_ola_step(self.h, _ptr(x), _i(len(x)), _ptr(y), C.byref(no)):
            raise RuntimeError("OLA: N_zeros > Ne (the reference runs out of its buffers here)")
        assert no.value == cap or self.fen   # windowed mode: the first block of the stream emits nothing
        return y[: no.value]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdo_ola_free(self.h)


class _PortPoly:
    def __init__(self, L, kind, taps, R, cplx):
        self.L, self.cplx = L, bool(cplx)
        t = _f32(taps)
        self.h = _vp(L.tsdo_poly_new(_i(kind), _ptr(t), _i(len(t)), _i(R), _i(1 if cplx else 0)))
        if not self.h:
            raise ValueError("tsdo_poly_new: bad arguments")

    @property
    def index(self):
        return self.L.tsdo_poly_index(self.h)

    @property
    def cnt(self):
        return self.L.tsdo_poly_cnt(self.h)

    def step(self, x):
        x = _c64(x) if self.cplx else _f32(x)
        y = np.empty(self.L.tsdo_poly_out_count(self.h, _i(len(x))), x.dtype)
        self.L.tsdo_poly_step(self.h, _ptr(x), _i(len(x)), _ptr(y))
        return y

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdo_poly_free(self.h)


class _PortItrp:
    def __init__(self, L, ratio, lut, nphases):
        self.L = L
        lut = _f32(lut)
        assert lut.shape[0] == nphases + 1
        self.K = lut.shape[1]
        self.ratio = ratio
        self.h = _vp(L.tsdo_itrp_new(_f(ratio), _ptr(lut), _i(self.K), _i(nphases)))

    @property
    def phase(self):
        return self.L.tsdo_itrp_phase(self.h)

    def step(self, x):
        x = _c64(x)
        if len(x) == 0:
            return np.zeros(0, np.complex64)
        cap = self.L.tsdo_itrp_capacity(self.h, _i(len(x)))
        y = np.empty(cap, np.complex64)
        no = _i()
        rc = self.L.tsdo_itrp_step(self.h, _ptr(x), _i(len(x)), _ptr(y), _i(cap), C.byref(no))
        if rc:
            raise RuntimeError(f"tsdo_itrp_step failed ({rc})")
        return y[: no.value].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdo_itrp_free(self.h)


# --------------------------------------------------------------------------- ref
class _RefFilter:
    def __init__(self, L, handle, err):
        if not handle:
            raise RuntimeError("reference: " + err())
        self.L, self.h, self._err = L, _vp(handle), err
        self.cplx = bool(L.tsdref_filter_is_complex(self.h))

    def step(self, x, cap=None):
        x = _c64(x) if self.cplx else _f32(x)
        if cap is None:
            cap = 4 * len(x) + 1024
        y = np.empty(cap, x.dtype)
        no = _i()
        if self.L.tsdref_filter_step(self.h, _ptr(x), _i(len(x)), _ptr(y), _i(cap), C.byref(no)):
            raise RuntimeError("reference: " + self._err())
        return y[: no.value].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdref_filter_free(self.h)


class _RefSpectrum:
    """rt_spectrum(SpectrumConfig) (fourier.hpp:909-952, fourier.cc:1162-1343); fenetre: 0 = none, 1 = Hann, 3 = Hamming."""

    def __init__(self, L, err, BS, nmeans, nsubs, sweep_active, sweep_step, masque_bf, masque_hf, fenetre):
        self.L, self._err = L, err
        nf, ns = _i(), _i()
        self.h = _vp(L.tsdref_spectrum_new(_i(BS), _i(nmeans), _i(nsubs), _i(1 if sweep_active else 0), _i(sweep_step), _i(masque_bf),
                                           _i(masque_hf), _i(fenetre), C.byref(nf), C.byref(ns)))
        if not self.h:
            raise RuntimeError("reference: " + err())
        self.BS, self.Nf, self.Ns = BS, nf.value, ns.value

    def step(self, x):
        x = _c64(x)
        y = np.empty(self.Ns + 8, np.float32)
        no = _i()
        if self.L.tsdref_spectrum_step(self.h, _ptr(x), _i(len(x)), _ptr(y), _i(len(y)), C.byref(no)):
            raise RuntimeError("reference: " + self._err())
        return y[: no.value].copy()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdref_spectrum_free(self.h)


class _RefDetect:
    FIELDS = ("position", "position_prec", "score", "gain", "theta", "SNR_dB", "sigma_noise")

    def __init__(self, L, motif, Ne, seuil, mode, err):
        m = _c64(motif)
        self.L, self._err = L, err
        self.h = _vp(L.tsdref_detect_new(_ptr(m), _i(len(m)), _i(Ne), _f(seuil), _i(1 if mode == "rif" else 0)))
        if not self.h:
            raise RuntimeError("reference: " + err())

    def step(self, x):
        """-> (score[n] float32, [dict per detection])"""
        x = _c64(x)
        score = np.empty(len(x), np.float32)
        dets = np.empty((4096, 7), np.float32)
        nd = _i()
        if self.L.tsdref_detect_step(self.h, _ptr(x), _i(len(x)), _ptr(score), _ptr(dets), _i(4096), C.byref(nd)):
            raise RuntimeError("reference: " + self._err())
        return score, [dict(zip(self.FIELDS, map(float, dets[i]))) for i in range(nd.value)]

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdref_detect_free(self.h)


class _RefPlan:
    def __init__(self, L, n, err):
        self.L, self._err = L, err
        self.h = _vp(L.tsdref_fftplan_new(_i(n), _i(1)))
        if not self.h:
            raise RuntimeError("reference: " + err())

    def step(self, x, forward=True):
        x = _c64(x)
        y = np.empty_like(x)
        if self.L.tsdref_fftplan_step(self.h, _ptr(x), _i(len(x)), _i(1 if forward else 0), _ptr(y)):
            raise RuntimeError("reference: " + self._err())
        return y

    def __del__(self):
        if getattr(self, "h", None):
            self.L.tsdref_fftplan_free(self.h)


class _Ref:
    """The reference's own code (see oracle/ref_shim/ref_driver.cc for the symbols)."""

    def __init__(self):
        L = C.CDLL(REF_SO)
        L.tsdref_last_error.restype = C.c_char_p
        for name in ("tsdref_fir_new", "tsdref_rif_fft_new", "tsdref_ola_new", "tsdref_ola_new2", "tsdref_itrp_new",
                     "tsdref_reechan_new", "tsdref_fftplan_new", "tsdref_polyphase_new", "tsdref_itrp_new2",
                     "tsdref_reechan_new_f32", "tsdref_detect_new", "tsdref_spectrum_new"):
            getattr(L, name).restype = _vp
        self.L = L

    def _err(self) -> str:
        return (self.L.tsdref_last_error() or b"").decode("utf-8", "replace")

    def p2(self, i):
        return self.L.tsdref_p2(_i(i))

    def ola_complexite_optimise(self, M):
        c, nf, nz, ne = _f(), _i(), _i(), _i()
        if self.L.tsdref_ola_complexite_optimise(_i(M), C.byref(c), C.byref(nf), C.byref(nz), C.byref(ne)):
            raise RuntimeError(self._err())
        return c.value, nf.value, nz.value, ne.value

    def design_rif_fen(self, n, typ, fc, win="hn"):
        h = np.zeros(n, np.float32)
        if self.L.tsdref_design_rif_fen(_i(n), typ.encode(), _f(fc), win.encode(), _ptr(h)):
            raise RuntimeError(self._err())
        return h

    def itrp_sinc_lut(self, K, nphases, fcut, win="hn"):
        lut = np.zeros((nphases + 1, K), np.float32)
        if self.L.tsdref_itrp_sinc_lut(_i(K), _i(nphases), _f(fcut), win.encode(), _ptr(lut)):
            raise RuntimeError(self._err())
        return lut

    def fir(self, kind, taps):
        t = _c64(taps) if kind == 2 else _f32(taps)
        return _RefFilter(self.L, self.L.tsdref_fir_new(_i(kind), _ptr(t), _i(len(t))), self._err)

    def rif_fft(self, kind, taps):
        t = _f32(taps)
        if len(t) > 512:
            # fourier.cc:954-960 fixes Ne = 512; with K > 512 svg.tail(N_zeros) starts before the
            # buffer (tableau.cc:520 does not catch a negative start) -> memory corruption.
            raise ValueError("reference filtre_rif_fft is undefined behaviour for K > 512")
        return _RefFilter(self.L, self.L.tsdref_rif_fft_new(_i(kind), _ptr(t), _i(len(t))), self._err)

    def fenetre(self, nom, n, sym=True):
        w = np.empty(n, np.float32)
        if self.L.tsdref_fenetre(nom.encode(), _i(n), _i(1 if sym else 0), _ptr(w)):
            raise RuntimeError(self._err())
        return w

    def ola(self, Ne, nb_zeros_min, H=None, avec_fenetrage=False):
        N = self.p2((Ne if Ne > 0 else 512) + nb_zeros_min)
        if N - (Ne if Ne > 0 else 512) > (Ne if Ne > 0 else 512):
            raise ValueError("reference OLA is undefined behaviour when N_zeros > Ne")
        no = _i()
        if H is None:
            h = self.L.tsdref_ola_new2(_i(Ne), _i(nb_zeros_min), None, _i(N), C.byref(no), _i(1 if avec_fenetrage else 0))
        else:
            Hc = _c64(H)
            assert len(Hc) == N
            h = self.L.tsdref_ola_new2(_i(Ne), _i(nb_zeros_min), _ptr(Hc), _i(N), C.byref(no), _i(1 if avec_fenetrage else 0))
        f = _RefFilter(self.L, h, self._err)
        f.N = no.value
        return f

    def spectrum(self, BS=1024, nmeans=10, nsubs=1, sweep_active=False, sweep_step=1024, masque_bf=0, masque_hf=0, fenetre=1):
        return _RefSpectrum(self.L, self._err, BS, nmeans, nsubs, sweep_active, sweep_step, masque_bf, masque_hf, fenetre)

    def periodogramme_tfd(self, x, N):
        """periodogramme_tfd(x, N) (fourier.cc:1451-1481) -> [frames, N2/2] float32 (dB)."""
        xx = _c64(x)
        cap = 4 * len(xx) + 4 * N + 64
        out = np.empty(cap, np.float32)
        r, c = _i(), _i()
        if self.L.tsdref_periodogramme_tfd(_ptr(xx), _i(len(xx)), _i(N), _ptr(out), _i(cap), C.byref(r), C.byref(c)):
            raise RuntimeError(self._err())
        return out[: r.value * c.value].reshape(r.value, c.value).copy()

    def reechan_freq(self, x, lom):
        """rééchan_freq<T>(x, lom) (fourier.cc:1391-1419), T = float or cfloat after the dtype of x."""
        cplx = np.iscomplexobj(x)
        xx = _c64(x) if cplx else _f32(x)
        cap = int(len(xx) * max(lom, 1.0)) + 8
        y = np.empty(cap, xx.dtype)
        no = _i()
        if self.L.tsdref_reechan_freq(_i(1 if cplx else 0), _ptr(xx), _i(len(xx)), _f(lom), _ptr(y), _i(cap), C.byref(no)):
            raise RuntimeError(self._err())
        return y[: no.value].copy()

    def ola_make_H(self, h, N):
        h = _f32(h)
        H = np.zeros(N, np.complex64)
        if self.L.tsdref_ola_make_H(_ptr(h), _i(len(h)), _i(N), _ptr(H)):
            raise RuntimeError(self._err())
        return H

    def itrp(self, ratio, K, nphases, fcut):
        return _RefFilter(self.L, self.L.tsdref_itrp_new(_f(ratio), _i(K), _i(nphases), _f(fcut)), self._err)

    def reechan(self, ratio, cplx=True):
        if not cplx:
            return _RefFilter(self.L, self.L.tsdref_reechan_new_f32(_f(ratio)), self._err)
        return _RefFilter(self.L, self.L.tsdref_reechan_new(_f(ratio)), self._err)

    def itrp2(self, ratio, kind, cplx=True, K=0, nphases=256, fcut=0.5, degree=0):
        """filtre_itrp<T>(ratio, itrp): kind = "sinc" | "cspline" | "lineaire" | "lagrange"; T = cfloat or float."""
        k = {"sinc": 0, "cspline": 1, "lineaire": 2, "lagrange": 3}[kind]
        return _RefFilter(self.L, self.L.tsdref_itrp_new2(_f(ratio), _i(k), _i(1 if cplx else 0), _i(K), _i(nphases), _f(fcut),
                                                          _i(degree)), self._err)

    def detecteur(self, motif, Ne=0, seuil=0.5, mode="ola"):
        """détecteur_création({Ne, motif, seuil, mode}) (detection.cc:511-514)."""
        return _RefDetect(self.L, motif, Ne, seuil, mode, self._err)

    def cspline_lut(self, n=256, c=0.0):
        lut = np.empty((n + 1, 4), np.float32)
        if self.L.tsdref_cspline_lut(_i(n), _f(c), _ptr(lut)):
            raise RuntimeError(self._err())
        return lut

    def polyphase(self, kind, taps, R=2):
        t = _f32(taps)
        return _RefFilter(self.L, self.L.tsdref_polyphase_new(_i(kind), _ptr(t), _i(len(t)), _i(R)), self._err)

    def fft(self, n):
        return _RefPlan(self.L, n, self._err)

    def rfft(self, x):
        x = _f32(x)
        y = np.zeros(len(x), np.complex64)
        if self.L.tsdref_rfft(_ptr(x), _i(len(x)), _ptr(y)):
            raise RuntimeError(self._err())
        return y

    def filtrer(self, taps, x):
        t, x = _f32(taps), _f32(x)
        y = np.empty_like(x)
        if self.L.tsdref_filtrer_f32(_ptr(t), _i(len(t)), _ptr(x), _i(len(x)), _ptr(y)):
            raise RuntimeError(self._err())
        return y

    def tampon_trace(self, N, chunks):
        ch = np.ascontiguousarray(chunks, np.int32)
        cap = int(ch.sum() // max(N, 1) + 8)
        em = np.zeros(cap, np.int32)
        no = _i()
        if self.L.tsdref_tampon_trace(_i(N), _ptr(ch), _i(len(ch)), _ptr(em), _i(cap), C.byref(no)):
            raise RuntimeError(self._err())
        return em[: no.value].copy()


_port = None
_ref = None


def port() -> _Port:
    global _port
    if _port is None:
        _port = _Port()
    return _port


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref() -> _Ref:
    global _ref
    if _ref is None:
        if not have_ref():
            raise RuntimeError("oracle/_ref/libtsdref.so not built (needs /root/reference; run make -C oracle ref)")
        _ref = _Ref()
    return _ref
