"""Correlation detector — GPU counterpart of détecteur_création(config) / Detecteur::step
(reference core/include/tsd/fourier.hpp:546-679, core/src/fourier/detection.cc:68-516, MODE_OLA).

The hot part — correlation with the motif through the block filter, energy of the same M samples, normalisation — is the
C ABI's tsdgpu_detect_step (libtsd_b200/csrc/ola.cu).  What follows in the reference is sparse, serial host logic on the
score signal (erosion over M samples, quadratic interpolation of the peak, gain / phase / noise estimate per detection,
detection.cc:262-500); it is restated here line by line, including its quirks (a peak on the last sample of a call is
finished at the beginning of the next one; an aborted border case drops the remaining candidates of that call)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from ._lib import lib, check, TsdGpuError, HOST

_vp = C.c_void_p
f32 = np.float32


@dataclass
class Detection:
    """fourier.hpp:546-573"""
    position: int = 0
    position_prec: float = 0.0
    score: float = 0.0
    gain: float = 0.0
    theta: float = 0.0
    SNR_dB: float = 0.0
    sigma_noise: float = 0.0


@dataclass
class DetecteurConfig:
    """fourier.hpp:577-601 (mode: only MODE_OLA is routed to the GPU; the FIR mode is `filtre_rif` + the same host logic)."""
    Ne: int = 0
    motif: Optional[np.ndarray] = None
    seuil: float = 0.5
    gere_detection: Optional[Callable[[Detection], None]] = None


def _delais(x: np.ndarray, tau: float) -> np.ndarray:
    """délais<cfloat>(x, tau) (fourier.cc:607-626,686-698): integer delays shift, fractional ones rotate the spectrum of the
    signal zero-padded to twice its length."""
    if np.floor(tau) == tau:
        t = int(tau)
        if t == 0:
            return x
        y = np.zeros_like(x)
        if t > 0:
            y[t:] = x[: len(x) - t]
        else:
            y[: len(x) + t] = x[-t:]
        return y
    n = 2 * len(x)
    x2 = np.zeros(n, np.complex64)
    x2[n // 4: n // 4 + n // 2] = x
    X = (np.fft.fft(x2) / np.sqrt(n)).astype(np.complex64)
    i = np.arange(n, dtype=np.float32)
    ang = (f32(-2) * f32(np.pi) * i * f32(tau) / f32(n) + f32(np.pi) * f32(tau)).astype(np.float32)
    rot = np.exp(1j * ang.astype(np.float64)).astype(np.complex64)
    X = X * np.fft.fftshift(rot)
    return (np.fft.ifft(X) * np.sqrt(n)).astype(np.complex64)[n // 4: n // 4 + n // 2]


class Detecteur:
    """One stream (the reference object is single-channel).  step(x) -> score (float32, len(x)); the detections of the call
    are in ``self.detections`` and are handed to ``config.gere_detection`` like the reference does."""

    def __init__(self, config: DetecteurConfig):
        if config.motif is None:
            raise TsdGpuError("détecteur : motif manquant")
        self.config = config
        self.motif0 = np.ascontiguousarray(config.motif, np.complex64)
        h = _vp()
        check(lib().tsdgpu_detect_create(self.motif0.ctypes.data_as(_vp), len(self.motif0), int(config.Ne), 1, C.byref(h)))
        self._h = h
        a, b, c, d, e = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_float()
        check(lib().tsdgpu_detect_dims(h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)))
        self.Ne, self.N, self.M, self.delais_corr = a.value, b.value, c.value, d.value
        self.norme_motif = f32(e.value)
        self.lar = np.zeros(self.delais_corr + 1, np.complex64)      # LigneARetardExt (detection.cc:27-64,193)
        self.pic_final_a_traiter = False
        self.pic_final = Detection()
        self.lc = self.lc0 = np.complex64(0)
        self.alc = self.alc0 = f32(0)
        self.itr = 0
        self.dernier_n = 0
        self.detections: List[Detection] = []
        self.corr = None

    def _lar_step(self, x):
        K, n = len(self.lar), len(x)
        if n >= K:
            self.lar = x[-K:].copy()
        else:
            self.lar[: K - n] = self.lar[n:].copy()
            self.lar[K - n:] = x

    def step(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex64)
        n = len(x)
        score = np.empty(n, np.float32)
        corr = np.empty(n, np.complex64)
        check(lib().tsdgpu_detect_step(self._h, x.ctypes.data_as(_vp), n, n, score.ctypes.data_as(_vp), n, corr.ctypes.data_as(_vp),
                                       n, HOST))
        self.corr = corr
        self.detections = []
        self._pics(x, score, corr)
        return score

    # detection.cc:262-500
    def _pics(self, x, y, corr):
        cfg, M, n, N = self.config, self.M, len(y), self.N
        y2 = np.zeros(n, np.float32)
        for i in range(0, n, M):
            seg = y[i: i + M]
            idx = int(np.argmax(seg))
            y2[i + idx] = seg[idx]
        lst = np.nonzero(y2 > f32(cfg.seuil))[0].tolist()
        lst2 = []
        if self.pic_final_a_traiter:
            self.pic_final_a_traiter = False
            lst2.append(-1)
        for idx in lst:
            if not any((y[i2] > y[idx]) and (abs(idx - i2) < M) for i2 in lst):
                lst2.append(idx)
        for idx in lst2:
            det = Detection()
            if idx == -1:
                det = Detection(**vars(self.pic_final))
                det.position -= self.dernier_n
                det.position_prec = float(f32(det.position_prec) - f32(self.dernier_n))
            else:
                det.score = float(y[idx])
                det.position = idx - self.delais_corr
                det.theta = float(np.angle(corr[idx]))
                det.gain = float(f32(abs(corr[idx])) / (self.norme_motif / np.sqrt(f32(N))))
                det.position_prec = float(det.position)
            if idx == -1:
                ac0, c0, ac1, c1, ac2, c2 = self.alc0, self.lc0, self.alc, self.lc, y[0], corr[0]
                if ac1 < ac2:
                    break                                  # `goto suite`: leaves the loop
            elif idx == 0:
                ac0, c0, ac1, c1, ac2, c2 = self.alc, self.lc, y[0], corr[0], y[1], corr[1]
                if ac1 < ac0:
                    break
            elif idx == n - 1:
                self.pic_final_a_traiter = True
                self.pic_final = det
                break
            else:
                ac0, c0, ac1, c1, ac2, c2 = y[idx - 1], corr[idx - 1], y[idx], corr[idx], y[idx + 1], corr[idx + 1]
            ac0, ac1, ac2 = f32(ac0), f32(ac1), f32(ac2)
            delta = (ac2 - ac0) / (f32(2) * (f32(2) * ac1 - ac2 - ac0))      # qint_loc
            if delta < -0.5 or delta > 0.5:
                delta = f32(min(max(delta, f32(-0.5)), f32(0.5)))
            det.position_prec = float(f32(det.position_prec) + delta)
            r = np.sqrt(f32(N)) / self.norme_motif
            g2 = (np.complex64(c1) - (np.complex64(c0) - np.complex64(c2)) * np.complex64(delta) * np.complex64(0.25)) * np.complex64(r)
            det.gain = float(abs(g2))
            det.theta = float(np.angle(g2))
            recu_theo = (self.motif0 * np.complex64(det.gain * np.exp(1j * det.theta))).astype(np.complex64)
            recu_theo = _delais(recu_theo, float(delta))
            ident = idx - self.delais_corr
            if ident >= 0:
                cur = M
            elif -M < ident < 0:
                cur = M + ident
            else:
                cur = 0
            avant = M - cur
            recu = np.zeros(M, np.complex64)
            if cur:
                recu[M - cur:] = x[max(ident, 0): max(ident, 0) + cur]
            if avant < M:
                if avant:
                    recu[:avant] = self.lar[len(self.lar) - avant:]
            else:
                recu[:avant] = self.lar[ident + len(self.lar): ident + len(self.lar) + avant]
            bruit = recu - recu_theo
            var_bruit = f32(np.mean(np.abs(bruit[1: M - 1]).astype(np.float32) ** 2))
            var_signal = f32((f32(det.gain) * self.norme_motif) ** 2 / f32(M))
            det.sigma_noise = float(np.sqrt(var_bruit))
            det.SNR_dB = float(10 * np.log10(var_signal / var_bruit)) if var_bruit > 0 else float("inf")
            self.detections.append(det)
            if cfg.gere_detection:
                cfg.gere_detection(det)
        self._lar_step(x)
        self.alc0, self.lc0 = y[n - 2], corr[n - 2]
        self.alc, self.lc = y[n - 1], corr[n - 1]
        self.itr += 1
        self.dernier_n = n

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib().tsdgpu_detect_destroy(h)
            except Exception:
                pass


def detecteur_creation(config: DetecteurConfig) -> Detecteur:
    """sptr<Detecteur> détecteur_création(config) (fourier.hpp:679, detection.cc:511-514)."""
    return Detecteur(config)


detector_new = detecteur_creation
