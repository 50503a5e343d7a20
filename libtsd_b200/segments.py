"""Halo split of ONE long stream into time segments (SURVEY §5 "long single stream", §8e row 2).

A segment that starts at stream sample S needs, besides its own samples, only the filter's memory at S:
  FIR        the K-1 samples before S                                   (filtre-rt.cc:56-58)
  filtre_fft the re-blocking position (S mod Ne, S div Ne) and the samples still inside the block filter's window —
             `carry_len` of them — before S; FIR-derived gains only (overlap-save form: no partial sums to hand over)
                                                                         (fourier.cc:813-833, tsd.cc:332-370)
  filtre_itrp the float32 phase at input S — from the data-independent schedule (ra.cc:58-73) — and the K-1 samples
             before S
  polyphase  (S, the decimation counter at S, the L-1 samples before S)  (polyphase.cc:100-105,213-218)
All of it is READ from the source buffer (the halo), nothing is exchanged between segments; every segment is then an
ordinary object stepping its own samples, on any GPU, and the concatenation of the segment outputs IS the one-shot
output (same lengths, bit for bit in the bookkeeping).  The only communication is the final gather of the outputs
(`gather_segments`, any torch.distributed backend)."""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from ._lib import lib, check
from .shard import channel_shard


def plan_segments(n: int, nseg: int, align: int = 1) -> List[Tuple[int, int]]:
    """[start, end) of ``nseg`` contiguous time segments covering [0, n); starts are multiples of ``align`` (use Ne for
    filtre_fft to keep every segment on whole blocks — not required for correctness, it only avoids re-blocking
    residuals inside a segment)."""
    if n < 0 or nseg < 1 or align < 1:
        raise ValueError("plan_segments: invalid argument")
    out, prev = [], 0
    for s in range(nseg):
        _, cnt = channel_shard(n, s, nseg)
        end = prev + cnt
        if s + 1 < nseg:
            end = min(n, ((end + align // 2) // align) * align)
        else:
            end = n
        end = max(end, prev)
        out.append((prev, end))
        prev = end
    return out


def _np_dtype(x):
    if isinstance(x, np.ndarray):
        return x.dtype
    return np.complex64 if x.is_complex() else np.float32


def _halo(x, start: int, length: int) -> np.ndarray:
    """The ``length`` samples before ``start`` (zeros before the stream start), oldest first; x is [nchan, n] (numpy, or a
    torch tensor in host or device memory: only the halo itself is brought to the host)."""
    h = np.zeros((x.shape[0], length), _np_dtype(x))
    a = max(0, start - length)
    if start > a:
        piece = x[:, a:start]
        h[:, length - (start - a):] = piece if isinstance(piece, np.ndarray) else piece.cpu().numpy()
    return h


def resamp_phase_at(ratio: float, start: int) -> float:
    """Phase of the reference's recurrence (ra.cc:58-73) after ``start`` input samples, from phase 0."""
    ph, no = C.c_float(0.0), C.c_longlong()
    done = 0
    while done < start:                       # the C entry takes an int count
        step = min(start - done, 1 << 30)
        check(lib().tsdgpu_resamp_schedule(C.byref(ph), C.c_float(np.float32(ratio)), 256, int(step), None, None, 0, C.byref(no)))
        done += step
    return float(ph.value)


def start_fir(flt, x: np.ndarray, start: int) -> None:
    """Puts a fresh filtre_rif object in the state it has after x[:, :start]."""
    flt.set_history(_halo(x, start, flt.K - 1), start)


def start_ola(flt, x: np.ndarray, start: int) -> None:
    """Puts a fresh filtre_fft object (FIR-derived gains, fir_len > 0) in the state it has after x[:, :start]."""
    cl, sl, ll = flt.state_dims()
    if sl or ll:
        raise ValueError("halo split needs the overlap-save form (fir_len > 0): partial sums cannot be read from the input")
    flt.set_state(start % flt.Ne, start // flt.Ne, _halo(x, start, cl))


def start_itrp(flt, x: np.ndarray, start: int) -> None:
    flt.set_state(resamp_phase_at(flt.ratio, start), _halo(x, start, flt.K - 1))


def start_polyphase(flt, x: np.ndarray, start: int) -> None:
    from .filtrage import POLY_UPS
    cnt = 0 if flt.kind == POLY_UPS else start % flt.R      # cnt <- (n + cnt) % R per call (polyphase.cc:100-105,213-218)
    flt.set_state(start, cnt, _halo(x, start, flt.hist_len))


def run_segment(make, starter, x: np.ndarray, span: Tuple[int, int]):
    """One segment: fresh object from ``make()``, start state from the halo, one step over x[:, start:end]."""
    s, e = span
    flt = make()
    if s > 0:
        starter(flt, x, s)
    seg = x[:, s:e]
    return flt.step(np.ascontiguousarray(seg) if isinstance(seg, np.ndarray) else seg)


def run_split(make, starter, x, nseg: int, align: int = 1):
    """All segments one after the other on this GPU (what each rank does for its own span in a multi-GPU run)."""
    parts = [run_segment(make, starter, x, sp) for sp in plan_segments(x.shape[1], nseg, align)]
    if isinstance(parts[0], np.ndarray):
        return np.concatenate(parts, axis=1)
    import torch
    return torch.cat(parts, dim=1)


def gather_segments(y_local, group=None):
    """Final gather (the only collective of the path): every rank contributes its segment's output [nchan, n_r] (n_r may
    differ), every rank receives the concatenation in rank order.  Works with nccl and gloo."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([y_local.shape[1]], dtype=torch.int64, device=y_local.device)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(v.item()) for v in ns]
    pad = torch.zeros((y_local.shape[0], max(ns)), dtype=y_local.dtype, device=y_local.device)
    pad[:, : y_local.shape[1]] = y_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:, :k] for o, k in zip(out, ns)], dim=1)
