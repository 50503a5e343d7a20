"""Host-side mirror of tsd::filtrage for the GPU hot path (names and argument meaning follow the
reference: core/include/tsd/filtrage.hpp; English aliases follow core/include/dsp/filter.hpp).

Every object keeps the reference's per-object semantics (one object == one stream, state carried
across step() calls); the only new surface is ``nchan``: a batch of independent channels that share
the same configuration, laid out [nchan, n].  Arrays may be numpy (host memory: copied in and out)
or torch CUDA tensors (device memory: no copies, asynchronous on the library stream).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._buf import Batch, empty_like_batch, restore_shape
from ._lib import TsdGpuError, check, lib

_vp = C.c_void_p


class FiltreGen:
    """tsd::FiltreGen<Te,Ts> (tsd.hpp:626-657): ``step(x) -> y``; the callee sizes ``y``."""

    nchan = 1

    def step(self, x):  # pragma: no cover - abstract
        raise NotImplementedError

    # English skin (dsp::FilterGen)
    def __call__(self, x):
        return self.step(x)


# ----------------------------------------------------------------------------- design helpers
def _sinc(T, f):
    """tsd::sinc(T, f) (divers.cc:6-12), float32."""
    T = np.float32(T)
    f = np.asarray(f, np.float32)
    a = np.float32(np.pi) * T * f
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.sin(a, dtype=np.float32) / (np.float32(np.pi) * f)
    return np.where(np.abs(a) < np.float32(1e-7), T, r).astype(np.float32)


def _linspace(a, b, n):
    """tsd::linspace (tsd.hpp:916-931): double step, float32 storage."""
    a = np.float32(a)
    b = np.float32(b)
    x = np.empty(n, np.float32)
    if n > 0:
        x[0] = a
    if n > 1:
        step = (float(b) - float(a)) / (n - 1)
        x[1:] = (float(a) + step * np.arange(1, n)).astype(np.float32)
    return x


def fenetre(type_: str, n: int, symetrique: bool = True) -> np.ndarray:
    """tsd::filtrage::fenêtre for "re"/"hn"/"hm" (fenetres.cc:17-59,125-128,204-232)."""
    if type_ in ("", "re", "aucune", "none"):
        return np.ones(n, np.float32)
    if type_ in ("hn", "hann"):
        a = np.float32(0.5)
    elif type_ in ("hm", "hamming"):
        a = np.float32(0.54)
    else:
        raise TsdGpuError(f"fenêtre: type '{type_}' non supporté")
    tmin = -(n // 2)
    if n % 2 == 0:
        tmax = n // 2 if symetrique else (n - 1) // 2
    else:
        tmax = n // 2 if symetrique else np.float32(n // 2) - (np.float32(n) - 1) / np.float32(n)
    t = _linspace(np.float32(tmin) / np.float32(n), np.float32(tmax) / np.float32(n), n)
    return (a + (np.float32(1) - a) * np.cos(np.float32(2 * np.pi) * t, dtype=np.float32)).astype(np.float32)


def design_rif_fen(n: int, type_: str, fc: float, fen: str = "hn") -> np.ndarray:
    """tsd::filtrage::design_rif_fen (rif-fen.cc:30-107), low-pass only ("lp"/"pb")."""
    if type_ not in ("lp", "pb"):
        raise TsdGpuError("design_rif_fen: seul le type 'lp' est disponible dans cette version")
    c = n // 2 if n % 2 else (n - 1) // 2
    h = _sinc(np.float32(2) * np.float32(fc), (np.arange(n) - c).astype(np.float32)) * fenetre(fen, n, True)
    h = h.astype(np.float32)
    if type_ == "lp":
        h = (h / np.float32(np.sum(h, dtype=np.float64))).astype(np.float32)
    return h


design_fir_wnd = design_rif_fen


# ----------------------------------------------------------------------------- FIR
class FiltreRIF(FiltreGen):
    """GPU counterpart of FiltreRIF<T,Tc> (filtre-rt.cc:53-109), created by :func:`filtre_rif`."""

    def __init__(self, coefs, T=np.complex64, nchan: int = 1):
        coefs = np.asarray(coefs)
        taps_complex = np.iscomplexobj(coefs)
        T = np.dtype(T).type
        if T not in (np.float32, np.complex64):
            raise TsdGpuError("filtre_rif: T doit être float32 ou complex64")
        if taps_complex and T is np.float32:
            raise TsdGpuError("filtre_rif: coefficients complexes avec des données réelles")
        self.kind = 0 if T is np.float32 else (2 if taps_complex else 1)
        self.dtype = T
        self.coefs = np.ascontiguousarray(coefs, np.complex64 if taps_complex else np.float32)
        self.K = int(self.coefs.shape[0])
        self.nchan = int(nchan)
        h = _vp()
        check(lib().tsdgpu_fir_create(self.kind, self.coefs.ctypes.data_as(_vp), self.K, self.nchan, C.byref(h)))
        self._h = h

    def step(self, x, out=None):
        b = Batch(x, self.dtype, self.nchan)
        if out is None:
            y = empty_like_batch(b, self.dtype, b.n)
            yb = Batch(y, self.dtype, self.nchan, "y")
        else:
            yb = Batch(out, self.dtype, self.nchan, "y")
            y = yb.arr
            if yb.n < b.n or yb.mem != b.mem:
                raise TsdGpuError("filtre_rif.step: tampon de sortie incompatible")
        check(lib().tsdgpu_fir_step(self._h, b.ptr, b.stride, b.n, yb.ptr, yb.stride, b.mem))
        return restore_shape(y[:, : b.n], b.ndim)

    @property
    def index(self) -> int:
        """Ring index of the reference object: (samples so far) mod K (filtre-rt.cc:89)."""
        i = C.c_int()
        check(lib().tsdgpu_fir_get_state(self._h, None, C.byref(i)))
        return i.value

    def get_state(self):
        fen = np.zeros((self.nchan, self.K), self.dtype)
        i = C.c_int()
        check(lib().tsdgpu_fir_get_state(self._h, fen.ctypes.data_as(_vp), C.byref(i)))
        return fen, i.value

    def set_state(self, fen, index: int):
        fen = np.ascontiguousarray(fen, self.dtype).reshape(self.nchan, self.K)
        check(lib().tsdgpu_fir_set_state(self._h, fen.ctypes.data_as(_vp), int(index)))

    def set_history(self, hist, samples_so_far: int):
        """Start in the middle of a stream: ``hist[nchan, K-1]`` = the K-1 samples before the first one to come (oldest
        first), ``samples_so_far`` = their position in the stream (halo split, libtsd_b200.segments)."""
        hist = np.ascontiguousarray(hist, self.dtype).reshape(self.nchan, max(self.K - 1, 0))
        check(lib().tsdgpu_fir_set_history(self._h, hist.ctypes.data_as(_vp) if hist.size else None, int(samples_so_far)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().tsdgpu_fir_destroy(h)
            except Exception:
                pass


def filtre_rif(coefs, T=np.complex64, nchan: int = 1) -> FiltreRIF:
    """sptr<FiltreGen<T>> filtre_rif<Tc,T>(coefs) (filtrage.hpp:1367-1368, filtre-rt.cc:171-175)."""
    return FiltreRIF(coefs, T, nchan)


filter_fir = filtre_rif


def filtrer(h, x):
    """tsd::filtrage::filtrer(Design, x) for FIR designs (filtrage.hpp:1684-1711): one-shot filter."""
    x_is_t = hasattr(x, "is_cuda")
    cplx = (x.is_complex() if x_is_t else np.iscomplexobj(x))
    nchan = 1 if x.ndim == 1 else x.shape[0]
    return filtre_rif(h, np.complex64 if cplx else np.float32, nchan).step(x)


filter = filtrer  # noqa: A001  (dsp::filter, dsp/filter.hpp:1662-1666)


def convol(h, x):
    """tsd::filtrage::convol(h, x) (filtrage.hpp:1774-1780): filtre_rif<Tc,T>(h)->step(x), as many outputs as inputs."""
    return filtrer(h, x)


def filtfilt(h, x):
    """tsd::filtrage::filtfilt(h, x) (filtrage.hpp:1761-1765): filtrer(h, filtrer(h, x).reverse()).reverse()."""
    rev = (lambda a: a.flip(-1)) if hasattr(x, "is_cuda") else (lambda a: np.ascontiguousarray(a[..., ::-1]))
    return rev(filtrer(h, rev(filtrer(h, x))))


# ----------------------------------------------------------------------------- interpolators
@dataclass
class InterpolateurSincConfig:
    """filtrage.hpp:1914-1927"""
    ncoefs: int = 31
    nphases: int = 256
    fcut: float = 0.5
    fenetre: str = "hn"


class InterpolateurSinc:
    """Windowed-sinc interpolator LUT (itrp.cc:10-55).  ``lut[p]`` = coefficients for delay p/nphases."""

    def __init__(self, config: InterpolateurSincConfig):
        if not (0 <= config.fcut <= 0.5):
            raise TsdGpuError("interpolateur sinc : fréquence normalisée attendue (entre 0 et 0.5)")
        self.config = config
        K, P = config.ncoefs, config.nphases
        self.K = K
        self.delais = 0.5 * K
        self.nom = f"sinc - ncoefs={K}, nphases={P}, fcut={config.fcut}, fen={config.fenetre}"
        lin = _linspace(-(K // 2), (K - 1) // 2, K)
        scale = np.float32(2 * np.pi / K)
        i = np.arange(K)
        lut = np.empty((P + 1, K), np.float32)
        for p in range(P + 1):
            tau = np.float32((1.0 * p) / P)
            h = _sinc(np.float32(2) * np.float32(config.fcut), (i - K // 2).astype(np.float32) - tau)
            if config.fenetre == "hn":
                t = ((lin - tau) * scale).astype(np.float32)
                h = h * (np.float32(0.5) + np.float32(0.5) * np.cos(t, dtype=np.float32))
            lut[p] = h
        self.lut = lut

    def coefs(self, tau: float) -> np.ndarray:
        idx = int(np.float32(tau) * np.float32(self.config.nphases))
        return self.lut[idx]


def itrp_sinc(config: InterpolateurSincConfig = InterpolateurSincConfig()) -> InterpolateurSinc:
    return InterpolateurSinc(config)


class InterpolateurLUT:
    """Any InterpolateurRIF given directly by its coefficient table [nphases+1, K]."""

    def __init__(self, lut):
        self.lut = np.ascontiguousarray(lut, np.float32)
        self.K = int(self.lut.shape[1])


class InterpolateurCSpline:
    """itrp_cspline<T>() (itrp.cc:59-80): Catmull-Rom family, K = 4, LUT of n + 1 = 257 delays built by
    cspline_calc_lut / cspline_filtre / cspline_calc (itrp.cc:292-320) with the reference's float32 operations."""

    def __init__(self, n: int = 256, c: float = 0.0):
        f = np.float32
        self.K, self.delais, self.nom = 4, 1.5, "cspline"
        lut = np.empty((n + 1, 4), np.float32)
        one_c = f(1) - f(c)
        for p in range(n + 1):
            t = f(p) / f(n)
            tm1 = t - f(1)
            h0 = (f(1) + f(2) * t) * tm1 * tm1
            h1 = t * tm1 * tm1
            h2 = t * t * (f(3) - f(2) * t)
            h3 = t * t * tm1
            lut[p] = (-one_c * h1 / f(2), h0 - one_c * h3 / f(2), h2 + one_c * h1 / f(2), one_c * h3 / f(2))
        self.lut = lut


class InterpolateurLineaire:
    """itrp_lineaire<T>() (itrp.cc:82-95): coefficients {1 - tau, tau} at the EXACT delay (no LUT)."""
    K, delais, nom, exact_kind, degre = 2, 0.5, "linéaire", 1, 1


class InterpolateurLagrange:
    """itrp_lagrange<T>(d) (itrp.cc:97-127): Lagrange polynomial of degree d through d + 1 samples, evaluated at the
    EXACT delay (no LUT)."""
    exact_kind = 2

    def __init__(self, degre: int):
        self.degre = int(degre)
        self.K, self.delais, self.nom = self.degre + 1, 0.5 * self.degre, f"Lagrange degré {self.degre}"


def itrp_cspline() -> InterpolateurCSpline:
    return InterpolateurCSpline()


def itrp_lineaire() -> InterpolateurLineaire:
    return InterpolateurLineaire()


def itrp_lagrange(degre: int) -> InterpolateurLagrange:
    return InterpolateurLagrange(degre)


class AdaptationRythmeSimple(FiltreGen):
    """GPU counterpart of AdaptationRythmeSimple (ra.cc:13-79), created by :func:`filtre_itrp`."""

    def __init__(self, ratio: float, itrp, nchan: int = 1, T=np.complex64):
        T = np.dtype(T).type
        if T not in (np.float32, np.complex64):
            raise TsdGpuError("filtre_itrp: T doit être float32 ou complex64")
        self.ratio = float(np.float32(ratio))
        self.itrp = itrp
        self.nchan = int(nchan)
        self.dtype = T
        cplx = 1 if T is np.complex64 else 0
        h = _vp()
        if getattr(itrp, "exact_kind", 0):
            # linear / Lagrange: coefficients evaluated at the exact float32 phase on the device
            self.K, self.nphases = int(itrp.K), 0
            check(lib().tsdgpu_resamp_create_exact(C.c_float(self.ratio), int(itrp.exact_kind), int(itrp.degre), cplx,
                                                   self.nchan, C.byref(h)))
        else:
            lut = np.ascontiguousarray(itrp.lut, np.float32)
            self.nphases = lut.shape[0] - 1
            self.K = lut.shape[1]
            check(lib().tsdgpu_resamp_create_ex(C.c_float(self.ratio), lut.ctypes.data_as(_vp), self.K, self.nphases, cplx,
                                                self.nchan, C.byref(h)))
        self._h = h

    @property
    def phase(self) -> float:
        return float(lib().tsdgpu_resamp_phase(self._h))

    def out_count(self, n: int) -> int:
        return int(lib().tsdgpu_resamp_out_count(self._h, int(n)))

    def get_state(self):
        """(phase, hist[nchan, K-1]) — ra.cc:16-22."""
        ph = C.c_float()
        hist = np.zeros((self.nchan, max(self.K - 1, 0)), self.dtype)
        check(lib().tsdgpu_resamp_get_state(self._h, C.byref(ph), hist.ctypes.data_as(_vp) if hist.size else None))
        return float(ph.value), hist

    def set_state(self, phase: float, hist):
        hist = np.ascontiguousarray(hist, self.dtype).reshape(self.nchan, max(self.K - 1, 0))
        check(lib().tsdgpu_resamp_set_state(self._h, C.c_float(phase), hist.ctypes.data_as(_vp) if hist.size else None))

    def step(self, x, out=None):
        """``out`` (optional, same memory space as ``x``) needs room for ceil(n * ratio) + 16 samples per channel; the
        result is the exact-length view of the buffer the samples were written to (no extra host copy)."""
        b = Batch(x, self.dtype, self.nchan)
        if b.n == 0:
            return restore_shape(empty_like_batch(b, self.dtype, 0), b.ndim)
        # the exact count comes out of the phase recurrence, which the library runs once inside the call: size the
        # buffer by the bound ceil(n * ratio) + 16 and return the exact-length view (a separate out_count() query would
        # run the whole recurrence a second time on the host)
        cap = int(np.ceil(b.n * max(self.ratio, 0.0))) + 16
        if out is None:
            y = empty_like_batch(b, self.dtype, cap)
            yb = Batch(y, self.dtype, self.nchan, "y")
        else:
            yb = Batch(out, self.dtype, self.nchan, "y")
            y = yb.arr
            if yb.mem != b.mem:
                raise TsdGpuError("filtre_itrp.step: tampon de sortie incompatible")
            cap = yb.n   # the library reports TSDGPU_ERR_SIZE itself if the exact count does not fit
        no = C.c_longlong()
        check(lib().tsdgpu_resamp_step(self._h, b.ptr, b.stride, b.n, yb.ptr, max(yb.stride, 1), cap, C.byref(no), b.mem))
        return restore_shape(y[:, : no.value], b.ndim)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().tsdgpu_resamp_destroy(h)
            except Exception:
                pass


def filtre_itrp(ratio: float, itrp, nchan: int = 1, T=np.complex64) -> AdaptationRythmeSimple:
    """sptr<FiltreGen<T>> filtre_itrp<T>(ratio, itrp), T = cfloat or float (filtrage.hpp:2039, ra.cc:185-195); ``itrp`` from
    itrp_sinc / itrp_cspline (LUT) or itrp_lineaire / itrp_lagrange (exact delay)."""
    return AdaptationRythmeSimple(ratio, itrp, nchan, T)


filter_itrp = filtre_itrp


# ----------------------------------------------------------------------------- polyphase stages
POLY_UPS, POLY_DEMI_BANDE, POLY_DECIM = 0, 1, 2


class FiltrePolyphase(FiltreGen):
    """GPU counterpart of FiltreRIFUps / FiltreRIFDemiBande / FiltreRIFDecim (polyphase.cc:54-341)."""

    def __init__(self, kind: int, coefs, R: int = 2, T=np.complex64, nchan: int = 1):
        T = np.dtype(T).type
        if T not in (np.float32, np.complex64):
            raise TsdGpuError("filtre polyphase: T doit être float32 ou complex64")
        self.kind, self.dtype, self.nchan = int(kind), T, int(nchan)
        self.R = 2 if kind == POLY_DEMI_BANDE else int(R)
        self.coefs = np.ascontiguousarray(coefs, np.float32)
        h = _vp()
        check(lib().tsdgpu_poly_create(self.kind, self.coefs.ctypes.data_as(_vp), int(self.coefs.shape[0]), self.R,
                                       1 if T is np.complex64 else 0, self.nchan, C.byref(h)))
        self._h = h

    def out_count(self, n: int) -> int:
        return int(lib().tsdgpu_poly_out_count(self._h, int(n)))

    @property
    def state(self):
        """(ring index, decimation counter) of the reference object (`index`, `odd` / `cnt`)."""
        i, c = C.c_int(), C.c_int()
        check(lib().tsdgpu_poly_state(self._h, C.byref(i), C.byref(c)))
        return i.value, c.value

    def get_state(self):
        """(samples so far, decimation counter, hist[nchan, L-1])."""
        tot, cnt = C.c_longlong(), C.c_int()
        check(lib().tsdgpu_poly_get_state(self._h, C.byref(tot), C.byref(cnt), None))
        L = self.hist_len
        hist = np.zeros((self.nchan, L), self.dtype)
        check(lib().tsdgpu_poly_get_state(self._h, None, None, hist.ctypes.data_as(_vp) if hist.size else None))
        return tot.value, cnt.value, hist

    def set_state(self, total: int, cnt: int, hist):
        hist = np.ascontiguousarray(hist, self.dtype).reshape(self.nchan, self.hist_len)
        check(lib().tsdgpu_poly_set_state(self._h, int(total), int(cnt), hist.ctypes.data_as(_vp) if hist.size else None))

    @property
    def hist_len(self) -> int:
        """Delay line of the reference object minus one: K (half-band, decimator) or ceil(K / R) (interpolator)."""
        K = int(self.coefs.shape[0])
        L = (K + self.R - 1) // self.R if self.kind == POLY_UPS else K
        return max(L - 1, 0)

    def step(self, x):
        b = Batch(x, self.dtype, self.nchan)
        cnt = self.out_count(b.n)
        y = empty_like_batch(b, self.dtype, cnt)
        if b.n == 0:
            return restore_shape(y, b.ndim)
        yb = Batch(y, self.dtype, self.nchan, "y")
        no = C.c_longlong()
        check(lib().tsdgpu_poly_step(self._h, b.ptr, b.stride, b.n, yb.ptr if cnt else None, max(yb.stride, 1),
                                     C.byref(no), b.mem))
        assert no.value == cnt
        return restore_shape(y, b.ndim)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().tsdgpu_poly_destroy(h)
            except Exception:
                pass


def filtre_rif_ups(coefs, R: int, T=np.complex64, nchan: int = 1) -> FiltrePolyphase:
    """filtre_rif_ups<float,T>(c, R) (polyphase.cc:246-341,356-360): xR polyphase interpolator."""
    return FiltrePolyphase(POLY_UPS, coefs, R, T, nchan)


def filtre_rif_demi_bande(coefs, T=np.complex64, nchan: int = 1) -> FiltrePolyphase:
    """filtre_rif_demi_bande<float,T>(c) (polyphase.cc:54-149,350-354): half-band decimator by 2."""
    return FiltrePolyphase(POLY_DEMI_BANDE, coefs, 2, T, nchan)


def filtre_rif_decim(coefs, R: int, T=np.complex64, nchan: int = 1) -> FiltrePolyphase:
    """filtre_rif_decim<float,T>(c, R) (polyphase.cc:156-239,344-348): FIR followed by decimation by R."""
    return FiltrePolyphase(POLY_DECIM, coefs, R, T, nchan)


def filtre_rif_ups_delais(nc: int, R: int) -> float:
    """filtre_rif_ups_délais (polyphase.cc:363-369)."""
    pad = (R - (nc % R)) if (nc % R) else 0
    return (nc - 1) / 2.0 + pad


filter_fir_ups, filter_fir_half_band, filter_fir_decim = filtre_rif_ups, filtre_rif_demi_bande, filtre_rif_decim


class AdaptationRythmeArbitraire(FiltreGen):
    """filtre_reechan<cfloat>(ratio) (ra.cc:84-183): half-band decimators while the factor is < 0.5, x2 polyphase
    interpolators while it is >= 2 (both on design_rif_fen(15, "lp", 0.25, "hn")), then the arbitrary-ratio LUT
    interpolator itrp_sinc({15, 256, min(0.4, f/2), "hn"}) unless |f - 1| < 1e-6."""

    def __init__(self, ratio: float, nchan: int = 1, T=np.complex64):
        T = np.dtype(T).type
        self.dtype = T
        r = np.float32(ratio)
        if r <= 0 or np.isinf(r) or r >= 1e9:   # ra.cc:108-112: logged, ratio forced to 1
            r = np.float32(1)
        self.ratio = float(r)
        self.nchan = int(nchan)
        f = r
        self.nb_decimateurs = 0
        self.nb_surechantillonneurs = 0
        while f < 0.5:
            self.nb_decimateurs += 1
            f = np.float32(f * 2)
        while f >= 2:
            self.nb_surechantillonneurs += 1
            f = np.float32(f / 2)
        self.facteur_post_interpolation = float(f)
        coefs = design_rif_fen(15, "lp", 0.25, "hn")                      # ra.cc:135
        self.decimateurs = [filtre_rif_demi_bande(coefs, T, nchan) for _ in range(self.nb_decimateurs)]
        self.surechantillonneurs = [filtre_rif_ups(coefs, 2, T, nchan) for _ in range(self.nb_surechantillonneurs)]
        fcut = min(np.float32(0.4), np.float32(f / 2))
        self.interpolateur = filtre_itrp(float(f), itrp_sinc(InterpolateurSincConfig(15, 256, float(fcut), "hn")), nchan, T)

    def step(self, x, out=None):
        """``out`` (optional) is handed to the final interpolator (see AdaptationRythmeSimple.step); it is ignored when
        the chain ends with a polyphase stage."""
        if self.ratio == 1:                        # ra.cc:162-163
            return x.clone() if hasattr(x, "clone") else np.array(x, self.dtype)
        y = x
        for d in self.decimateurs:                 # ra.cc:166-167
            y = d.step(y)
        for s in self.surechantillonneurs:         # ra.cc:169-170
            y = s.step(y)
        if abs(np.float32(self.facteur_post_interpolation) - 1) < 1e-6:   # ra.cc:173-174
            if y is x:
                return x.clone() if hasattr(x, "clone") else np.array(x, self.dtype)
            return y
        return self.interpolateur.step(y, out=out)


def filtre_reechan(ratio: float, nchan: int = 1, T=np.complex64) -> AdaptationRythmeArbitraire:
    """filtre_reechan<T>(ratio), T = cfloat or float (ra.cc:180-183,190-191)."""
    return AdaptationRythmeArbitraire(ratio, nchan, T)


filter_resample = filtre_reechan


def reechan(x, r: float):
    """tsd::rééchan(x, r) (tsd.hpp:700-705) / dsp::resample (dsp/dsp.hpp:499-503)."""
    nchan = 1 if x.ndim == 1 else x.shape[0]
    cplx = x.is_complex() if hasattr(x, "is_complex") else np.iscomplexobj(x)
    return filtre_reechan(r, nchan, np.complex64 if cplx else np.float32).step(x)


resample = reechan
