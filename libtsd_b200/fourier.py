"""Host-side mirror of tsd::fourier for the GPU hot path (reference: core/include/tsd/fourier.hpp;
English aliases: core/include/dsp/fourier.hpp)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._buf import Batch, empty_like_batch, restore_shape
from ._lib import TsdGpuError, check, lib
from .filtrage import FiltreGen

_vp = C.c_void_p


def prochaine_puissance_de_2(i: int) -> int:
    """tsd.cc:287-291 (float log, evaluated by the library with the reference's expression)."""
    return int(lib().tsdgpu_p2(int(i)))


def ola_complexite(M: int, Ne: int):
    """ola_complexité(M, Ne) -> (C, Nf, Nz) (fourier.cc:708-713): FLOP per input sample of the block filter."""
    c, nf, nz = C.c_float(), C.c_int(), C.c_int()
    check(lib().tsdgpu_ola_complexite(int(M), int(Ne), C.byref(c), C.byref(nf), C.byref(nz)))
    return c.value, nf.value, nz.value


def ola_complexite_optimise(M: int):
    """ola_complexité_optimise(M) -> (C, Nf, Nz, Ne) (fourier.cc:715-735): block length 2^k - (M-1) of least cost."""
    c, nf, nz, ne = C.c_float(), C.c_int(), C.c_int(), C.c_int()
    check(lib().tsdgpu_ola_complexite_optimise(int(M), C.byref(c), C.byref(nf), C.byref(nz), C.byref(ne)))
    return c.value, nf.value, nz.value, ne.value


ola_complexity, ola_complexity_optimize = ola_complexite, ola_complexite_optimise


class FFTPlan:
    """tsd::fourier::FFTPlan (fourier.hpp:19-32) backed by the GPU plan.

    Like TFRPlanDefaut the plan is always unitary (``normalize`` is accepted and ignored,
    fourier.cc:119-120,362) and re-plans itself when the input length changes (fourier.cc:416-417).
    ``batch`` transforms are laid out [batch, n].
    """

    def __init__(self, n: int = -1, avant: bool = True, normalize: bool = True, batch: int = 1):
        self.n = -1
        self.avant = avant
        self.batch = int(batch)
        self._h = None
        if n >= 0:
            self.configure(n, avant, normalize)

    def configure(self, n: int, avant: bool = True, normalize: bool = True):
        self._free()
        self.n = int(n)
        self.avant = avant
        h = _vp()
        check(lib().tsdgpu_fft_plan(self.n, self.batch, C.byref(h)))
        self._h = h

    def step(self, x, avant: Optional[bool] = None, out=None):
        if avant is None:
            avant = self.avant
        nb = 1 if x.ndim == 1 else x.shape[0]
        n = x.shape[-1]
        if n <= 0:
            raise TsdGpuError("Echec assertion : x.rows() > 0.")   # fourier.cc:414
        if n != self.n or nb != self.batch:
            self.batch = nb
            self.configure(n, self.avant)
        b = Batch(x, np.complex64, self.batch)
        if out is None:
            y = empty_like_batch(b, np.complex64, n)
        else:
            y = out if out.ndim == 2 else out[None]
        yb = Batch(y, np.complex64, self.batch, "y")
        check(lib().tsdgpu_fft_exec(self._h, b.ptr, b.stride, yb.ptr, yb.stride, 1 if avant else 0, b.mem))
        return restore_shape(yb.arr, x.ndim)

    def _free(self):
        if self._h:
            try:
                lib().tsdgpu_fft_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def __del__(self):
        self._free()


def tfrplan_creation(n: int = -1, avant: bool = True, normaliser: bool = True, batch: int = 1) -> FFTPlan:
    """sptr<FFTPlan> tfrplan_création(n, avant, normaliser) (fourier.hpp:69, fourier.cc:475-481)."""
    return FFTPlan(n, avant, normaliser, batch)


fftplan_new = tfrplan_creation


def fft(x):
    """tsd::fourier::fft for complex input (fourier.hpp:163-170): unitary DFT."""
    return FFTPlan().step(x, True)


def ifft(X):
    """tsd::fourier::ifft (fourier.hpp:199-205): unitary inverse DFT."""
    return FFTPlan().step(X, False)


def _tfr_rotation(n: int) -> np.ndarray:
    """tfr_rotation<float>(n) (fourier.cc:32-46): cdouble recurrence r *= w0, stored as cfloat."""
    w0 = np.exp(-2j * np.pi / n)
    out = np.empty(n, np.complex64)
    r = 1.0 + 0.0j
    for i in range(n):
        out[i] = r
        r *= w0
    return out


def rfft(x):
    """tsd::fourier::rfft (fourier.hpp:116-122) = RTFRPlan<float>::step (fourier.cc:280-355): n real samples ->
    n complex bins of the unitary DFT.  Even n: one n/2-point complex transform (GPU plan) of the packed signal, the
    reference's post-twiddle, then csym_forçage (fourier.hpp:264-282); odd n: plain complex transform.
    Set-up path (it builds H for filtre_rif_fft, fourier.cc:962-965): host numpy around the GPU plan."""
    x = np.ascontiguousarray(x, np.float32)
    n = x.shape[0]
    if n == 0:
        raise TsdGpuError("Echec assertion : x.rows() > 0.")
    if n & 1:
        return FFTPlan().step(x.astype(np.complex64), True)
    h = n // 2
    Xt = FFTPlan().step((x[0::2] + 1j * x[1::2]).astype(np.complex64), True)
    rot = _tfr_rotation(n)
    i = np.arange(h + 1)
    X1 = np.where(i == h, Xt[0], Xt[np.minimum(i, h - 1)])
    X2 = np.where(i > 0, Xt[(h - i) % h], Xt[0])
    j2, r2 = np.complex64(1j * 0.5 / np.sqrt(2)), np.complex64(0.5 / np.sqrt(2))
    y = np.zeros(n, np.complex64)
    y[: h + 1] = (r2 * (X1 + np.conj(X2)) - j2 * (X1 - np.conj(X2)) * rot[: h + 1]).astype(np.complex64)
    y[0] = y[0].real
    y[h] = y[h].real
    if h > 1:
        y[n - (h - 1):] = np.conj(y[1:h][::-1])
    return y


def periodogramme_tfd(x, N: int, fenetre=None):
    """Tabf periodogramme_tfd(x, N) (fourier.hpp:967, fourier.cc:1451-1481): 10 log10(|X|^2 + 1e-20) of the first N2/2
    bins of every frame of the windowed filtre_fft object (two frames per block of N samples, Hann window, N2 = p2(N)).
    x: [n] -> [frames, N2/2]; [nchan, n] -> [nchan, frames, N2/2] (numpy = host, torch.cuda = device)."""
    nchan = 1 if x.ndim == 1 else int(x.shape[0])
    b = Batch(x, np.complex64, nchan)
    N = int(N)
    if fenetre is None:
        from .filtrage import fenetre as _fen
        fenetre = _fen("hn", N, False)                       # fourier.cc:796
    w = np.ascontiguousarray(fenetre, np.float32)
    if w.shape != (N,):
        raise TsdGpuError(f"periodogramme_tfd: la fenêtre doit comporter N = {N} points")
    frames, bins = 2 * (b.n // N), prochaine_puissance_de_2(N) // 2
    if b.torch:
        import torch
        out = torch.empty((nchan, frames, bins), dtype=torch.float32, device=b.arr.device)
        optr = out.data_ptr()
    else:
        out = np.empty((nchan, frames, bins), np.float32)
        optr = out.ctypes.data
    nf, nb = C.c_int(), C.c_int()
    check(lib().tsdgpu_periodogramme_tfd(b.ptr, b.stride, b.n, nchan, N, w.ctypes.data_as(_vp), _vp(optr),
                                         max(frames * bins, 1), C.byref(nf), C.byref(nb), b.mem))
    assert (nf.value, nb.value) == (frames, bins)
    return out[0] if b.ndim == 1 else out


@dataclass
class SpectrumConfig:
    """tsd::fourier::SpectrumConfig (fourier.hpp:909-947)."""
    BS: int = 1024                # dimension des blocs d'entrée
    nmeans: int = 10              # nombre de spectres moyennés
    nsubs: int = 1                # sous-blocs par bloc (balayage fréquentiel / multi-threading)
    sweep_active: bool = False
    sweep_step: int = 1024
    sweep_masque_bf: int = 0
    sweep_masque_hf: int = 0
    fenetre: str = "hn"           # Fenetre::HANN by default; "re" / "hm"

    def Nf(self) -> int:
        return self.BS // self.nsubs

    def Ns(self) -> int:
        return self.Nf() + (self.nsubs - 1) * self.sweep_step if self.sweep_active else self.Nf()


class Spectrum:
    """rt_spectrum(config) (fourier.hpp:952; Spectrum, fourier.cc:1162-1343): averaged power spectrum in dB.  step(x) takes
    exactly BS samples per channel and returns Ns values per channel when the block completes `nmeans` blocks, else an
    empty array (the reference resizes y to 0).  ``nchan`` independent channels share the configuration."""

    def __init__(self, config: SpectrumConfig, nchan: int = 1, fenetre=None):
        from .filtrage import fenetre as _fen
        self.config, self.nchan = config, int(nchan)
        Nf = config.Nf()
        if fenetre is None:
            f = _fen(config.fenetre, Nf, False).astype(np.float32)
            # f = sqrt(Nf / abs2(f).somme()) * f  (fourier.cc:1203): float32 throughout
            f = (np.float32(np.sqrt(np.float32(Nf) / np.sum(f * f, dtype=np.float32))) * f).astype(np.float32)
        else:
            f = np.ascontiguousarray(fenetre, np.float32)
        if f.shape != (Nf,):
            raise TsdGpuError(f"rt_spectrum: la fenêtre doit comporter Nf = {Nf} points")
        self.fenetre = f
        self._h = _vp()
        check(lib().tsdgpu_spectrum_create(int(config.BS), int(config.nmeans), int(config.nsubs), 1 if config.sweep_active else 0,
                                           int(config.sweep_step), int(config.sweep_masque_bf), int(config.sweep_masque_hf),
                                           f.ctypes.data_as(_vp), self.nchan, C.byref(self._h)))
        nf, ns = C.c_int(), C.c_int()
        check(lib().tsdgpu_spectrum_dims(self._h, C.byref(nf), C.byref(ns)))
        self.Nf, self.Ns = nf.value, ns.value
        self._cnt = 0

    def step(self, x):
        b = Batch(x, np.complex64, self.nchan)
        last = self._cnt + 1 == self.config.nmeans
        if b.torch:
            import torch
            out = torch.empty((self.nchan, self.Ns if last else 0), dtype=torch.float32, device=b.arr.device)
            optr = out.data_ptr() if last else None
        else:
            out = np.empty((self.nchan, self.Ns if last else 0), np.float32)
            optr = out.ctypes.data if last else None
        no = C.c_int()
        check(lib().tsdgpu_spectrum_step(self._h, b.ptr, b.stride, b.n, _vp(optr), max(self.Ns, 1), C.byref(no), b.mem))
        self._cnt = 0 if last else self._cnt + 1
        assert no.value == (self.Ns if last else 0)
        return out[0] if b.ndim == 1 else out

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().tsdgpu_spectrum_destroy(h)
            except Exception:
                pass


def rt_spectrum(config: SpectrumConfig, nchan: int = 1, fenetre=None) -> Spectrum:
    return Spectrum(config, nchan, fenetre)


def reechan_freq(x, lom: float):
    """rééchan_freq<T>(x, lom) (fourier.hpp:143, fourier.cc:1391-1419): delay-free resampling of a whole signal by
    zero-padding (lom > 1) or truncating (lom < 1) its spectrum.  n2 = round(n * lom); the two transforms run on GPU
    plans of n and n2 points (any size).  Like the reference, the result is the REAL part of the inverse transform,
    also for complex input (``return real(x)``, :1408,1416)."""
    x = np.asarray(x)
    cplx = np.iscomplexobj(x)
    x = np.ascontiguousarray(x, np.complex64 if cplx else np.float32)
    if lom == 1:
        return x.copy()
    n = x.shape[0]
    n2 = int(np.floor(np.float32(n) * np.float32(lom) + np.float32(0.5)))      # (entier) round(n * lom), float32 product
    X = fft(x) if cplx else rfft(x)
    X2 = np.zeros(n2, np.complex64)
    h = n // 2 if lom > 1 else n2 // 2
    X2[:h] = X[:h]
    X2[n2 - h:] = X[n - h:]
    y = ifft(X2) * np.float32(np.sqrt(np.float32(lom)))
    return y.real.astype(np.complex64) if cplx else np.ascontiguousarray(y.real)


resample_freq = reechan_freq


@dataclass
class FiltreFFTConfig:
    """fourier.hpp:305-320.  The reference's ``traitement_freq`` callback is a host std::function; the
    GPU path takes the callback of FiltreFFTRIF, "X *= H" (fourier.cc:956-959), as data:

    H        : N complex gains (None = identity callback)
    fir_len  : K > 0 declares H = fft([0^(N-K), h]) * sqrt(N) (fourier.cc:962-965) and lets the
               library use its overlap-save form; 0 keeps the generic overlap-add.
    """
    dim_blocs_temporel: int = 0
    nb_zeros_min: int = 0
    avec_fenetrage: bool = False
    H: Optional[np.ndarray] = None
    fir_len: int = 0
    fenetre: Optional[np.ndarray] = None   # windowed mode: the Ne window values as data (default: Hann, periodic)
    # the reference's own field: arbitrary host callback ``f(X)`` on the N bins of every block (numpy complex64 view,
    # modified in place; ``f(X, chan)`` if it takes two arguments).  Runs FFT -> host -> callback -> device -> IFFT.
    traitement_freq: Optional[object] = None


FFTFilterConfig = FiltreFFTConfig


class OLA(FiltreGen):
    """GPU counterpart of OLA<cfloat> (fourier.cc:737-932): plain mode and, with ``avec_fenetrage``, the
    Hann-window 50 % overlap mode (:884-930; window = fenêtre("hn", Ne, non), or ``config.fenetre`` as data)."""

    def __init__(self, config: FiltreFFTConfig, nchan: int = 1):
        self.config = config
        self.nchan = int(nchan)
        Ne = config.dim_blocs_temporel if config.dim_blocs_temporel > 0 else 512
        N = prochaine_puissance_de_2(Ne + config.nb_zeros_min)
        Hp = None
        if config.H is not None:
            H = np.ascontiguousarray(config.H, np.complex64)
            if H.shape != (N,):
                raise TsdGpuError(f"filtre_fft: H doit comporter N = {N} points")
            Hp = H.ctypes.data_as(_vp)
            self._H = H
        h = _vp()
        self._cb = None
        if config.traitement_freq is not None:
            if config.H is not None:
                raise TsdGpuError("filtre_fft: H et traitement_freq sont exclusifs")
            import inspect
            fn = config.traitement_freq
            two = len(inspect.signature(fn).parameters) >= 2

            def _tramp(user, chan, X, n):
                a = np.ctypeslib.as_array(C.cast(X, C.POINTER(C.c_float)), shape=(2 * n,)).view(np.complex64)
                fn(a, chan) if two else fn(a)
            self._cb = C.CFUNCTYPE(None, _vp, C.c_int, _vp, C.c_int)(_tramp)
            wp = None
            if config.avec_fenetrage:
                w = config.fenetre
                if w is None:
                    from .filtrage import fenetre
                    w = fenetre("hn", Ne, False)
                w = np.ascontiguousarray(w, np.float32)
                self._w = w
                wp = w.ctypes.data_as(_vp)
            check(lib().tsdgpu_ola_create_cb(int(config.dim_blocs_temporel), int(config.nb_zeros_min), self._cb, None, wp,
                                             self.nchan, C.byref(h)))
        elif config.avec_fenetrage:
            w = config.fenetre
            if w is None:
                from .filtrage import fenetre
                w = fenetre("hn", Ne, False)                       # fourier.cc:796
            w = np.ascontiguousarray(w, np.float32)
            if w.shape != (Ne,):
                raise TsdGpuError(f"filtre_fft: la fenêtre doit comporter Ne = {Ne} points")
            check(lib().tsdgpu_ola_create_fen(int(config.dim_blocs_temporel), int(config.nb_zeros_min), Hp,
                                              w.ctypes.data_as(_vp), self.nchan, C.byref(h)))
        else:
            check(lib().tsdgpu_ola_create(int(config.dim_blocs_temporel), int(config.nb_zeros_min), Hp,
                                          int(config.fir_len), self.nchan, C.byref(h)))
        self._h = h
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(lib().tsdgpu_ola_dims(h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.Ne, self.N, self.N_zeros = a.value, b.value, c.value

    @property
    def residual(self) -> int:
        d = C.c_int()
        check(lib().tsdgpu_ola_dims(self._h, None, None, None, C.byref(d)))
        return d.value

    def out_count(self, n: int) -> int:
        return int(lib().tsdgpu_ola_out_count(self._h, int(n)))

    def state_dims(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(lib().tsdgpu_ola_state_dims(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def get_state(self):
        """dict(residual, blocks_done, carry[nchan, carry_len], svg, last) — see tsdgpu_ola_get_state."""
        cl, sl, ll = self.state_dims()
        carry = np.zeros((self.nchan, cl), np.complex64)
        svg = np.zeros((self.nchan, sl), np.complex64) if sl else None
        last = np.zeros((self.nchan, ll), np.complex64) if ll else None
        r, bd = C.c_int(), C.c_longlong()
        check(lib().tsdgpu_ola_get_state(self._h, C.byref(r), C.byref(bd), carry.ctypes.data_as(_vp),
                                         svg.ctypes.data_as(_vp) if sl else None, last.ctypes.data_as(_vp) if ll else None))
        return dict(residual=r.value, blocks_done=bd.value, carry=carry, svg=svg, last=last)

    def set_state(self, residual: int, blocks_done: int, carry, svg=None, last=None):
        cl, sl, ll = self.state_dims()
        carry = np.ascontiguousarray(carry, np.complex64).reshape(self.nchan, cl)
        svg = np.ascontiguousarray(svg, np.complex64).reshape(self.nchan, sl) if (svg is not None and sl) else None
        last = np.ascontiguousarray(last, np.complex64).reshape(self.nchan, ll) if (last is not None and ll) else None
        check(lib().tsdgpu_ola_set_state(self._h, int(residual), int(blocks_done), carry.ctypes.data_as(_vp),
                                         svg.ctypes.data_as(_vp) if svg is not None else None,
                                         last.ctypes.data_as(_vp) if last is not None else None))

    def step(self, x, out=None):
        b = Batch(x, np.complex64, self.nchan)
        cnt = self.out_count(b.n)
        if out is None:
            y = empty_like_batch(b, np.complex64, cnt)
        else:
            y = out if out.ndim == 2 else out[None]
        if b.n == 0:
            return restore_shape(y[:, :0], b.ndim)
        yb = Batch(y, np.complex64, self.nchan, "y")
        no = C.c_longlong()
        check(lib().tsdgpu_ola_step(self._h, b.ptr, b.stride, b.n, yb.ptr if cnt else None, max(yb.stride, 1),
                                    C.byref(no), b.mem))
        assert no.value == cnt
        return restore_shape(yb.arr[:, :cnt], b.ndim)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().tsdgpu_ola_destroy(h)
            except Exception:
                pass


def filtre_fft(config: FiltreFFTConfig, nchan: int = 1):
    """tuple<sptr<Filtre<cfloat,cfloat,FiltreFFTConfig>>, entier> filtre_fft(config)
    (fourier.hpp:370, fourier.cc:935-940): returns (filter, N)."""
    f = OLA(config, nchan)
    return f, f.N


filter_fft = filtre_fft


def ola_make_H(h, N: int) -> np.ndarray:
    """H of the FiltreFFTRIF convention (fourier.cc:962-965): h2.tail(K) = h; H = fft(h2) * sqrt(N).
    Set-up helper evaluated in float64 on the host."""
    h = np.asarray(h, np.float64)
    h2 = np.zeros(N, np.float64)
    h2[N - len(h):] = h
    return np.fft.fft(h2).astype(np.complex64)


class FiltreFFTRIF(FiltreGen):
    """filtre_rif_fft<T>(h) (fourier.cc:946-990).  ``Ne`` defaults to the reference's 512; the
    reference is only defined for K <= 512 (with more taps N_zeros > Ne and it indexes before its
    buffer, fourier.cc:870), so longer filters must pass a larger ``Ne`` explicitly.
    ``compat_real_output`` reproduces the reference's ``real(...)`` of the output (fourier.cc:976)."""

    def __init__(self, h, Ne: int = 0, nchan: int = 1, compat_real_output: bool = False):
        h = np.ascontiguousarray(h, np.float32)
        K = len(h)
        ne = Ne if Ne > 0 else 512
        N = prochaine_puissance_de_2(ne + K)
        cfg = FiltreFFTConfig(dim_blocs_temporel=Ne, nb_zeros_min=K, H=ola_make_H(h, N), fir_len=K)
        self.ola = OLA(cfg, nchan)
        self.nchan = nchan
        self.compat_real_output = compat_real_output

    def step(self, x):
        y = self.ola.step(x)
        if self.compat_real_output:
            y = y.real.astype(np.complex64) if isinstance(y, np.ndarray) else (y.real + 0j)
        return y


def filtre_rif_fft(h, Ne: int = 0, nchan: int = 1, compat_real_output: bool = False) -> FiltreFFTRIF:
    return FiltreFFTRIF(h, Ne, nchan, compat_real_output)


filter_fir_fft = filtre_rif_fft
