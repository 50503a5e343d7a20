#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native libtsd filtering hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ola|fft|fir|resample] [--impl reference]

Metric (BASELINE.json): Gsamples/s of cf32 input samples filtered.  Default workload = BASELINE
config 4, the FFT-domain filter `filtre_fft` (K = 4095 taps, Ne = 61441, N = 65536) on 256 channels x
16 Mi cf32 samples per GPU; the other workloads are BASELINE configs 2, 3 and 5.  A "step" is one
step() over the whole per-GPU batch.  Multi-GPU: channels are independent, every rank owns its own
256 channels (weak scaling, no data-path collective); the timed region is bracketed by a barrier and a
device synchronize, and the reported time is the max over ranks.

The JSON line carries, besides the driver contract: `roofline` (HBM-bound: algorithmic bytes of the
dominant kernel / its CUDA-event duration / measured HBM copy bandwidth), `cpu_baseline` (the reference's
own CPU code, oracle/_ref, on all host cores over a bounded sample), `e2e` (same metric through the C ABI
with pinned HOST buffers, copies inside the timed region) and `clocks` (nvidia-smi during the run).

`--impl reference` times the reference's CPU implementation (oracle/_ref, else the C port) on the same
workload definition with all host threads; under torchrun only rank 0 runs it.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (BASELINE config text, algorithmic HBM bytes per input sample)
    "ola": ("overlap-save filtre_fft: 256 ch x 16Mi cf32 per GPU, 4095-tap filter, Ne=61441, N=65536", 16.0),
    "fft": ("batched fft/ifft: 4096 ch x 65536-pt cf32, forward+inverse round trip", 32.0),
    "fir": ("direct FIR: 1024 ch x 1Mi cf32, 127-tap low-pass, filtre_rif step() in 64Ki blocks", 16.0),
    "resample": ("polyphase/LUT resampler 147/160 on 512 ch x 8Mi cf32, sinc LUT 64 taps x 257 phases", 8.0 + 8.0 * 147 / 160),
    # secondary line of SURVEY 8(d) config 5: the stock resample() chain (15-tap interpolator only at this ratio)
    "reechan": ("stock resample() / filtre_reechan 147/160 on 512 ch x 8Mi cf32 (15-tap sinc interpolator x 257 phases)", 8.0 + 8.0 * 147 / 160),
    # the reference's DEFAULT block-filter shape: filtre_rif_fft(h) = Ne 512, N 1024 (fourier.cc:946-990)
    "rif_fft": ("filtre_rif_fft default shape: 1024 ch x 1Mi cf32, 255-tap low-pass, Ne=512, N=1024", 16.0),
    # a long direct-form filter through the reference's filtre_rif interface (filtre-rt.cc:53-109): 511 taps
    "fir_long": ("direct FIR, long filter: 256 ch x 1Mi cf32, 511-tap low-pass, filtre_rif one step() per channel", 16.0),
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_workload_setup(workload):
    """Builds the per-thread job of the CPU reference arm; returns (make_job, samples_per_job, kind, sample_text)."""
    import oracle
    have_ref = oracle.have_ref()
    O = oracle.ref() if have_ref else oracle.port()
    kind = "reference" if have_ref else "port"
    rng = np.random.default_rng(0xC0FFEE)

    def cn(n):
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)

    if workload == "ola":
        h = O.design_rif_fen(4095, "lp", 0.1)
        H = O.ola_make_H(h, 65536)
        n = 32 * 61441
        x = cn(n)

        def make_job():
            f = O.ola(61441, 4095, H)
            return lambda: f.step(x)
        return make_job, n, kind, f"1 channel x {n} cf32 samples (32 Ne-blocks) per host thread, filtre_fft K=4095 Ne=61441"
    if workload == "fft":
        x = cn(65536)
        reps = 64

        def make_job():
            p = O.fft(65536)

            def job():
                for _ in range(reps):
                    p.step(p.step(x, True), False)
            return job
        return make_job, reps * 65536, kind, f"{reps} x 65536-pt fwd+inv round trips per host thread"
    if workload == "fir":
        h = O.design_rif_fen(127, "lp", 0.1)
        blocks = 16
        x = cn(65536)

        def make_job():
            f = O.fir(1, h)

            def job():
                for _ in range(blocks):
                    f.step(x)
            return job
        return make_job, blocks * 65536, kind, f"1 channel x {blocks} step() of 65536 cf32 per host thread, 127 taps"
    if workload == "resample":
        n = 1 << 20
        x = cn(n)
        if have_ref:
            def make_job():
                f = O.itrp(147.0 / 160.0, 64, 256, 0.4)
                return lambda: f.step(x)
        else:
            lut = O.itrp_sinc_lut(64, 256, 0.4)

            def make_job():
                f = O.itrp(147.0 / 160.0, lut, 256)
                return lambda: f.step(x)
        return make_job, n, kind, f"1 channel x {n} cf32 samples per host thread, filtre_itrp 147/160, sinc 64x257"
    if workload == "fir_long":
        h = O.design_rif_fen(511, "lp", 0.1)
        x = cn(1 << 18)

        def make_job():
            f = O.fir(1, h)
            return lambda: f.step(x)
        return make_job, 1 << 18, kind, "1 channel x 262144 cf32 per host thread, 511 taps"
    if workload == "rif_fft":
        h = O.design_rif_fen(255, "lp", 0.1)
        n = 1 << 20
        x = cn(n)
        if have_ref:
            def make_job():
                f = O.rif_fft(1, h)
                return lambda: f.step(x)
        else:
            H = O.ola_make_H(h, 1024)

            def make_job():
                f = O.ola(512, 255, H)
                return lambda: f.step(x)
        return make_job, n, kind, f"1 channel x {n} cf32 samples per host thread, filtre_rif_fft 255 taps (Ne=512, N=1024)"
    if workload == "reechan":
        n = 1 << 20
        x = cn(n)
        if have_ref:
            def make_job():
                f = O.reechan(147.0 / 160.0)
                return lambda: f.step(x)
        else:
            lut = O.itrp_sinc_lut(15, 256, 0.4)

            def make_job():
                f = O.itrp(147.0 / 160.0, lut, 256)
                return lambda: f.step(x)
        return make_job, n, kind, f"1 channel x {n} cf32 samples per host thread, filtre_reechan 147/160"
    raise SystemExit(f"unknown workload {workload}")


def cpu_run(workload, steps, warmup, threads=None):
    """Each step: every host thread runs one job on its own filter object (objects are per-channel and
    not thread-safe in the reference).  Returns (Gsamples/s, ms_per_step, threads, kind, sample_text)."""
    threads = threads or (os.cpu_count() or 1)
    make_job, n_job, kind, text = cpu_workload_setup(workload)
    jobs = [None] * threads

    def build(i):
        jobs[i] = make_job()   # constructed inside its worker thread (avoids false sharing)

    def run_all(fn):
        ts = [threading.Thread(target=fn, args=(i,)) for i in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    run_all(build)
    for _ in range(warmup):
        run_all(lambda i: jobs[i]())
    t0 = time.perf_counter()
    for _ in range(steps):
        run_all(lambda i: jobs[i]())
    dt = time.perf_counter() - t0
    total = float(n_job) * threads * steps
    return total / dt / 1e9, dt / steps * 1e3, threads, kind, text


# --------------------------------------------------------------------------------------- GPU arm
class _Setup:
    """Set-up tables (taps, H, LUT) of the GPU arms, from the package's own host-side mirror of the reference's design
    functions — oracle/ is touched by the cpu_baseline / --impl reference legs only."""

    @staticmethod
    def design_rif_fen(n, type_, fc):
        from libtsd_b200 import filtrage as F
        return F.design_rif_fen(n, type_, fc)

    @staticmethod
    def ola_make_H(h, N):
        from libtsd_b200 import fourier as Fo
        return Fo.ola_make_H(h, N)

    @staticmethod
    def itrp_sinc_lut(K, nphases, fcut):
        from libtsd_b200 import filtrage as F
        return F.itrp_sinc(F.InterpolateurSincConfig(K, nphases, fcut, "hn")).lut


class GpuWorkload:
    """Builds the device-resident synthetic batch and the filter objects of one workload."""

    def __init__(self, name, scale=1.0):
        import torch
        from libtsd_b200 import filtrage as F, fourier as Fo
        self.name = name
        self.torch = torch
        O = _Setup
        g = torch.Generator(device="cuda")
        g.manual_seed({"ola": 0x7D5D0004, "fft": 0x7D5D0002, "fir": 0x7D5D0003, "resample": 0x7D5D0005, "reechan": 0x7D5D0005, "rif_fft": 0x7D5D0006, "fir_long": 0x7D5D0007}[name])

        pad = int(os.environ.get("TSDGPU_BENCH_PAD", "0"))   # experiment: channel stride n + pad instead of n

        def randc(nchan, n):
            x = torch.empty((nchan, n + pad), dtype=torch.complex64, device="cuda")
            torch.view_as_real(x).normal_(generator=g)
            return x[:, :n]

        def emptyc(nchan, n):
            return torch.empty((nchan, n + pad), dtype=torch.complex64, device="cuda")[:, :n]

        if name == "ola":
            self.nchan, self.n = max(1, int(256 * scale)), 1 << 24
            h = O.design_rif_fen(4095, "lp", 0.1)
            self.H = O.ola_make_H(h, 65536)
            self.flt, N = Fo.filtre_fft(Fo.FiltreFFTConfig(61441, 4095, H=self.H, fir_len=4095), self.nchan)
            assert N == 65536
            self.x = randc(self.nchan, self.n)
            self.y = emptyc(self.nchan, self.flt.Ne * ((self.n + self.flt.Ne - 1) // self.flt.Ne))
            self.samples_per_step = self.nchan * self.n
            self.step = lambda: self.flt.step(self.x, out=self.y)
        elif name == "fft":
            self.nchan, self.n = max(1, int(4096 * scale)), 65536
            self.plan = Fo.tfrplan_creation(65536, batch=self.nchan)
            self.x = randc(self.nchan, self.n)
            self.X = emptyc(self.nchan, self.n)
            self.y = emptyc(self.nchan, self.n)
            self.samples_per_step = self.nchan * self.n

            def step():
                self.plan.step(self.x, True, out=self.X)
                self.plan.step(self.X, False, out=self.y)
            self.step = step
        elif name == "fir":
            self.nchan, self.n, self.blk = max(1, int(1024 * scale)), 1 << 20, 65536
            h = O.design_rif_fen(127, "lp", 0.1)
            self.flt = F.filtre_rif(h, np.complex64, self.nchan)
            self.x = randc(self.nchan, self.n)
            self.y = emptyc(self.nchan, self.n)
            self.samples_per_step = self.nchan * self.n

            def step():
                for b in range(self.n // self.blk):
                    s = slice(b * self.blk, (b + 1) * self.blk)
                    self.flt.step(self.x[:, s], out=self.y[:, s])
            self.step = step
        elif name == "resample":
            self.nchan, self.n = max(1, int(512 * scale)), 1 << 23
            lut = O.itrp_sinc_lut(64, 256, 0.4)
            self.flt = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(lut), self.nchan)
            self.x = randc(self.nchan, self.n)
            self.samples_per_step = self.nchan * self.n
            self.step = lambda: self.flt.step(self.x)
        elif name == "rif_fft":
            self.nchan, self.n = max(1, int(1024 * scale)), 1 << 20
            h = O.design_rif_fen(255, "lp", 0.1)
            self.flt = Fo.filtre_rif_fft(h, nchan=self.nchan).ola
            assert self.flt.N == 1024 and self.flt.Ne == 512
            self.x = randc(self.nchan, self.n)
            self.y = emptyc(self.nchan, self.n)
            self.samples_per_step = self.nchan * self.n
            self.step = lambda: self.flt.step(self.x, out=self.y)
        elif name == "fir_long":
            self.nchan, self.n = max(1, int(256 * scale)), 1 << 20
            h = O.design_rif_fen(511, "lp", 0.1)
            self.flt = F.filtre_rif(h, np.complex64, self.nchan)
            self.x = randc(self.nchan, self.n)
            self.y = emptyc(self.nchan, self.n)
            self.samples_per_step = self.nchan * self.n
            self.step = lambda: self.flt.step(self.x, out=self.y)
        elif name == "reechan":
            self.nchan, self.n = max(1, int(512 * scale)), 1 << 23
            self.flt = F.filtre_reechan(147.0 / 160.0, self.nchan)
            self.x = randc(self.nchan, self.n)
            self.samples_per_step = self.nchan * self.n
            self.step = lambda: self.flt.step(self.x)
        else:
            raise SystemExit(f"unknown workload {name}")


KERNELS = {
    "ola": "ols16k_kernel<1> (single-SM overlap-save, 16384-point transforms, TMEM constants, TMA bulk prefetch)",
    "fft": "fft64k_pipe_kernel (TMA-fed persistent four-step pipeline: 3-D tensor-map column tiles, 6-slot shared-memory ring)",
    "fir": "fir_tc2_kernel (tcgen05 3xTF32 Toeplitz GEMM; persistent CTAs, tensor-map loads and stores)",
    "resample": "resamp_tc_kernel (tcgen05 3xTF32 banded filter-bank GEMM, CTA pairs, tensor-map loads and stores)",
    "reechan": "resamp_tc_kernel (tcgen05 3xTF32 banded filter-bank GEMM, CTA pairs, tensor-map loads and stores)",
    "rif_fft": "ols16k_kernel<1> (same single-SM overlap-save kernel: the device's transform size is independent of Ne / N)",
    "fir_long": "ols16k_kernel<1> (delay 0, FIR history as carry: filtre_rif with >= 128 taps on cf32 data)",
}


def make_config(workload, world):
    """The SAME dict in both arms (the driver compares them)."""
    text, _ = WORKLOADS[workload]
    return {"workload": text, "per_gpu_batch": "fixed (weak scaling)", "l2": "inputs larger than L2 (no flush needed)",
            "parallelism": f"channel-sharded x{world}, no collective"}


def pin_affinity(local_rank):
    """Best effort: run this rank on the CPUs nvidia-smi lists as local to its GPU (NUMA), so that the pinned staging buffers
    and the copy threads of different ranks do not share one socket's memory controllers."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        hdr = None
        for ln in out.splitlines():
            cols = [c.strip() for c in ln.split("\t") if c.strip()]
            if hdr is None and any("CPU Affinity" in c for c in cols):
                hdr = cols
                continue
            if hdr and cols and cols[0] == f"GPU{local_rank}":
                idx = [i for i, c in enumerate(hdr) if "CPU Affinity" in c][0] + 1   # data rows carry the row label first
                spec = cols[idx] if idx < len(cols) else ""
                cpus = set()
                for part in spec.split(","):
                    if "-" in part:
                        a, b = part.split("-")
                        cpus.update(range(int(a), int(b) + 1))
                    elif part.strip().isdigit():
                        cpus.add(int(part))
                cpus &= os.sched_getaffinity(0)
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    return sorted(cpus)
    except Exception:
        pass
    return None


def load_traffic(workload):
    """DRAM bytes of ONE launch of the dominant kernel at the benched size, from the committed ncu capture of this same
    command (profiles/r02_traffic_<workload>.json, written by profiles/tools/traffic_capture.sh)."""
    f = os.path.join(ROOT, "profiles", f"r02_traffic_{workload}.json")
    if os.path.exists(f):
        with open(f) as fh:
            return json.load(fh)
    return None


def measure_workload(name, steps, warmup, scale, stream, barrier, world, sample_clocks_on=None):
    """Device-resident timing of one workload: returns (dict for the JSON line, samples per step)."""
    import torch
    import libtsd_b200
    wl = GpuWorkload(name, scale)
    torch.cuda.synchronize()
    for _ in range(warmup):
        wl.step()
    barrier()
    sampler = ClockSampler(sample_clocks_on) if sample_clocks_on is not None else None
    if sampler:
        sampler.start()
    libtsd_b200.launch_count(reset=True)
    libtsd_b200._lib.timing_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(steps):
        wl.step()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    kern_ms, kern_launches = libtsd_b200._lib.timing_read()
    libtsd_b200._lib.timing_enable(False)
    launches = libtsd_b200.launch_count()
    clocks = sampler.stop() if sampler else None
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    samples_per_step = wl.samples_per_step
    wl.step = None   # the step closure refers back to the workload: break the cycle so that the buffers go now
    del wl
    gc.collect()
    torch.cuda.empty_cache()
    _, bps = WORKLOADS[name]
    peaks, peak_kind = load_peaks()
    peak = float(peaks["hbm_gbs"])
    achieved = bps * samples_per_step * steps / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    roofline = {"bound": "hbm", "kernel": KERNELS[name], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_kind,
                "algorithmic_bytes_per_sample": bps, "kernel_ms_per_step": kern_ms / steps, "kernel_launches": kern_launches,
                "algorithmic_bytes_per_launch": (bps * samples_per_step * steps / kern_launches) if kern_launches else None}
    tr = load_traffic(name)
    if tr and kern_launches:
        roofline["traffic"] = tr.get("dram_bytes_per_launch")
        roofline["traffic_source"] = tr.get("source")
    res = {"value": float(samples_per_step) * world * steps / (elapsed_ms * 1e-3) / 1e9, "unit": "Gsamples/s",
           "ms_per_step": elapsed_ms / steps, "steps": steps, "roofline": roofline, "gpu_launches": launches, "clocks": clocks,
           "config": make_config(name, world)}
    return res, samples_per_step


def e2e_measure(name, steps, warmup, barrier=None, world=1, pageable=False):
    """Same metric through the C ABI with HOST buffers (H2D + D2H inside the timed region).  pageable = False: pinned
    buffers (cudaHostAlloc through torch); True: plain malloc'd numpy arrays, what a caller that owns ordinary Veccf
    storage hands over (the library stages them itself)."""
    import torch
    from libtsd_b200 import filtrage as F, fourier as Fo
    O = _Setup
    rng = np.random.default_rng(1)

    def hostbuf(nchan, n):
        if pageable:
            a = np.empty((nchan, n), np.complex64)
            t = None
        else:
            t = torch.empty((nchan, n), dtype=torch.complex64).pin_memory()
            a = t.numpy()
        row_r = rng.standard_normal(n, dtype=np.float32)
        row_i = rng.standard_normal(n, dtype=np.float32)
        a.real[...] = row_r[None]
        a.imag[...] = row_i[None]
        return t, a

    if name == "ola":
        # the BASELINE batch itself at N = 1 (256 channels x 16 Mi); a share of it per rank when several GPUs pull from
        # the same host memory; a bounded subset for the pageable variant
        nchan = 32 if pageable else (256 if world == 1 else max(32, 256 // world))
        n = 1 << 24
        H = O.ola_make_H(O.design_rif_fen(4095, "lp", 0.1), 65536)
        flt, _ = Fo.filtre_fft(Fo.FiltreFFTConfig(61441, 4095, H=H, fir_len=4095), nchan)
        while True:
            try:
                tx, x = hostbuf(nchan, n)
                ty, y = hostbuf(nchan, 61441 * (n // 61441 + 1))
                break
            except RuntimeError:
                if nchan <= 8:
                    raise
                nchan //= 2
                flt, _ = Fo.filtre_fft(Fo.FiltreFFTConfig(61441, 4095, H=H, fir_len=4095), nchan)
        step = lambda: flt.step(x, out=y)   # noqa: E731
        out_per_step = 61441 * (n // 61441)
    elif name == "fft":
        nchan, n = 256, 65536
        plan = Fo.tfrplan_creation(65536, batch=nchan)
        tx, x = hostbuf(nchan, n)
        ty, y = hostbuf(nchan, n)

        def step():
            plan.step(x, True, out=y)
            plan.step(y, False, out=y)
        out_per_step = 2 * n
        n = 2 * n   # two host round trips per step
    elif name == "fir":
        nchan, n = 64, 1 << 20
        flt = F.filtre_rif(O.design_rif_fen(127, "lp", 0.1), np.complex64, nchan)
        tx, x = hostbuf(nchan, n)
        ty, y = hostbuf(nchan, n)
        step = lambda: flt.step(x, out=y)   # noqa: E731
        out_per_step = n
    elif name == "fir_long":
        nchan, n = 64, 1 << 20
        flt = F.filtre_rif(O.design_rif_fen(511, "lp", 0.1), np.complex64, nchan)
        tx, x = hostbuf(nchan, n)
        ty, y = hostbuf(nchan, n)
        step = lambda: flt.step(x, out=y)   # noqa: E731
        out_per_step = n
    elif name == "rif_fft":
        nchan, n = 256, 1 << 20
        flt = Fo.filtre_rif_fft(O.design_rif_fen(255, "lp", 0.1), nchan=nchan).ola
        tx, x = hostbuf(nchan, n)
        ty, y = hostbuf(nchan, n)
        step = lambda: flt.step(x, out=y)   # noqa: E731
        out_per_step = n
    elif name == "reechan":
        nchan, n = 128, 1 << 20
        flt = F.filtre_reechan(147.0 / 160.0, nchan)
        tx, x = hostbuf(nchan, n)
        ty, y = hostbuf(nchan, int(n * 147 / 160) + 32)
        step = lambda: flt.step(x, out=y)   # noqa: E731
        out_per_step = int(n * 147 / 160)
    else:
        nchan, n = 128, 1 << 20
        flt = F.filtre_itrp(147.0 / 160.0, F.InterpolateurLUT(O.itrp_sinc_lut(64, 256, 0.4)), nchan)
        tx, x = hostbuf(nchan, n)
        ty, y = hostbuf(nchan, int(n * 147 / 160) + 32)
        step = lambda: flt.step(x, out=y)   # noqa: E731
        out_per_step = int(n * 147 / 160)
    for _ in range(warmup):
        step()
    if barrier is not None:
        barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()   # host-memory calls return when y is valid on the host
    dt = (time.perf_counter() - t0) / steps
    # the link ceiling of the same buffers: H2D of x and D2H of y at the same time, no kernel (pinned buffers only)
    ceiling = None
    if not pageable:
        dx = torch.empty((nchan, x.shape[1]), dtype=torch.complex64, device="cuda")
        dy = torch.empty((nchan, y.shape[1]), dtype=torch.complex64, device="cuda")   # whole rows: one contiguous copy each way
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        if barrier is not None:
            barrier()
        t1 = time.perf_counter()
        with torch.cuda.stream(s_in):
            dx.copy_(tx, non_blocking=True)
        with torch.cuda.stream(s_out):
            ty.copy_(dy, non_blocking=True)
        torch.cuda.synchronize()
        ceiling = time.perf_counter() - t1
        del dx, dy
    samples = nchan * (n if name != "fft" else n // 2)
    res = {"samples_per_step": samples, "seconds_per_step": dt, "unit": "Gsamples/s", "h2d_bytes_per_step": int(nchan * n * 8),
           "d2h_bytes_per_step": int(nchan * out_per_step * 8), "host_memory": "pageable (malloc)" if pageable else "pinned",
           "sample": f"{nchan} channels x {n if name != 'fft' else n // 2} cf32 in {'pageable' if pageable else 'pinned'} host memory per step and per GPU"}
    if ceiling:
        res["copy_only_seconds"] = ceiling
    return res


def strong_and_gathered(steps, warmup, stream, barrier, world, rank):
    """BASELINE config 4 as north_star states it: 256 channels sharded over the N GPUs (256/N each), no collective during
    the compute ("strong"), then the same with the final NCCL all-gather of the output shards, issued per channel group on
    a side stream so that the transfer of group g overlaps the filtering of group g+1 ("gathered")."""
    import torch
    import torch.distributed as dist
    from libtsd_b200 import fourier as Fo
    total_ch, n, Ne = 256, 1 << 24, 61441
    cpr = total_ch // world
    H = _Setup.ola_make_H(_Setup.design_rif_fen(4095, "lp", 0.1), 65536)
    G = 4 if cpr % 4 == 0 else 1
    cg = cpr // G
    n_out = Ne * (n // Ne + 1)   # the objects carry their re-blocking residual from step to step: a step emits 273 or 274 blocks
    g = torch.Generator(device="cuda")
    g.manual_seed(0x7D5D0004 + rank)
    x = torch.empty((cpr, n), dtype=torch.complex64, device="cuda")
    torch.view_as_real(x).normal_(generator=g)
    y = torch.empty((G, cg, n_out), dtype=torch.complex64, device="cuda")
    flts = [Fo.filtre_fft(Fo.FiltreFFTConfig(Ne, 4095, H=H, fir_len=4095), cg)[0] for _ in range(G)]

    def compute(gi):
        flts[gi].step(x[gi * cg:(gi + 1) * cg], out=y[gi])

    def timed(body):
        for _ in range(warmup):
            body()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            body()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    ms_strong = timed(lambda: [compute(gi) for gi in range(G)])
    samples = float(total_ch) * n
    recv = (world - 1) / world * total_ch * n_out * 8.0          # bytes every GPU receives per step
    del y
    torch.cuda.empty_cache()

    # (a) gather by the COPY ENGINES over peer-mapped (symmetric) memory: every rank's kernel stores its shard straight into
    # its slot of the local gathered buffer, and a side stream pushes the slot into the same place of every peer's buffer
    # (cudaMemcpyAsync peer-to-peer: DMA over NVLink, no SM taken from the persistent filter kernel) while the next channel
    # group is being filtered.  Measured on 2 GPUs (profiles/multi/p2p_gather_test.py): copy engines 771 GB/s per direction;
    # the filter kernel storing directly into peer memory only 498 GB/s (62 Gsamples/s: remote stores stall the math warps),
    # NCCL's all-gather kernel next to the persistent kernel 390 GB/s.
    gath_ce = None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        full = symm_mem.empty((G, world, cg, n_out), dtype=torch.complex64, device=f"cuda:{torch.cuda.current_device()}")
        hdl = symm_mem.rendezvous(full, dist.group.WORLD)
        peers = [hdl.get_buffer(r, full.shape, full.dtype) for r in range(world)]
        side = torch.cuda.Stream()   # ONE side stream: a stream per peer measured slower at N = 8 (63 vs 74 Gsamples/s)
        pushed = [torch.cuda.Event() for _ in range(G)]
        first = [True]

        def gathered_ce():
            cur = torch.cuda.current_stream()
            for gi in range(G):
                if not first[0]:
                    cur.wait_event(pushed[gi])                 # last step's push of this slot has read it
                flts[gi].step(x[gi * cg:(gi + 1) * cg], out=full[gi, rank])
                ev = torch.cuda.Event()
                ev.record(cur)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    for d in range(1, world):
                        r = (rank + d) % world                 # every rank starts with another peer: no hot spot
                        peers[r][gi, rank].copy_(full[gi, rank], non_blocking=True)
                    pushed[gi].record(side)
            first[0] = False
            cur.wait_stream(side)                              # the step ends when my pushes have landed

        hdl.barrier()
        ms_ce = timed(gathered_ce)
        hdl.barrier()
        # every slot of my buffer must hold the shard its owner computed: compare checksums of slot [0][r] with the owner's
        cs = torch.stack([torch.view_as_real(full[0, r, 0, :Ne * (n // Ne)]).double().sum() for r in range(world)])
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        same = all(bool(torch.equal(allcs[0], c)) for c in allcs)
        gath_ce = {"value": samples / (ms_ce * 1e-3) / 1e9, "unit": "Gsamples/s", "ms_per_step": ms_ce, "channels_per_gpu": cpr,
                   "collective": f"none: the kernel stores its shard into the local gathered buffer, copy engines push it to the {world - 1} peer(s) "
                                 f"over NVLink per group of {cg} channels (symmetric memory), overlapped with the next group's filtering",
                   "link_gbs_per_gpu_received": recv / (ms_ce * 1e-3) / 1e9, "link_peak_gbs": 770.0,
                   "link_peak_source": "peer-copy figure of /opt/skills/guides/B200_PROFILING.md",
                   "link_frac": recv / (ms_ce * 1e-3) / 1e9 / 770.0, "layout": "[group][rank][channel][sample]",
                   "all_ranks_hold_identical_data": same}
        del peers, hdl, full
    except Exception as e:   # symmetric memory unavailable (no peer access): the NCCL figure below stands alone
        gath_ce = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    gc.collect()
    torch.cuda.empty_cache()

    # (b) the same with NCCL's all-gather (its kernel has to share the SMs with the persistent filter kernel)
    y = torch.empty((G, cg, n_out), dtype=torch.complex64, device="cuda")
    full = torch.empty((G, world, cg, n_out), dtype=torch.complex64, device="cuda")   # gathered layout [group][rank][channel][sample]

    def gathered():
        works = []
        for gi in range(G):
            compute(gi)
            works.append(dist.all_gather_into_tensor(full[gi], y[gi], async_op=True))   # NCCL's stream waits for the step
        for w in works:
            w.wait()
    ms_gath = timed(gathered)
    del x, y, full, flts
    gc.collect()
    torch.cuda.empty_cache()
    nccl = {"value": samples / (ms_gath * 1e-3) / 1e9, "unit": "Gsamples/s", "ms_per_step": ms_gath, "channels_per_gpu": cpr,
            "collective": f"ncclAllGather per group of {cg} channels, overlapped with the next group's filtering",
            "link_gbs_per_gpu_received": recv / (ms_gath * 1e-3) / 1e9, "link_peak_gbs": 770.0,
            "link_peak_source": "peer-copy figure of /opt/skills/guides/B200_PROFILING.md",
            "link_frac": recv / (ms_gath * 1e-3) / 1e9 / 770.0,
            "layout": "[group][rank][channel][sample]"}
    if gath_ce and "value" in gath_ce and gath_ce["value"] >= nccl["value"]:
        gath = dict(gath_ce)
        gath["nccl_all_gather"] = nccl
    else:
        gath = dict(nccl)
        gath["copy_engine_gather"] = gath_ce
    return ({"value": samples / (ms_strong * 1e-3) / 1e9, "unit": "Gsamples/s", "ms_per_step": ms_strong, "channels_per_gpu": cpr,
             "scaling": "strong", "collective": "none"}, gath)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="ola", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="channel-count multiplier (debug only; 1.0 = BASELINE size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (no `workloads`, `strong`, `gathered` blocks)")
    ap.add_argument("--gather-only", action="store_true", help="debug: skip the `workloads` block (headline + strong / gathered only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    text, bytes_per_sample = WORKLOADS[args.workload]

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return
        val, ms, threads, kind, sample = cpu_run(args.workload, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": "Gsamples/s filtered (cf32)", "value": val, "unit": "Gsamples/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": make_config(args.workload, args.gpus),
                "cpu_baseline": {"value": val, "unit": "Gsamples/s", "cores": threads, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = pin_affinity(local_rank) if world > 1 else None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import libtsd_b200
    libtsd_b200.init(local_rank)
    # every launch of the library and the timing events go to ONE explicit (non-default) stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    libtsd_b200.use_torch_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    head, samples_per_step = measure_workload(args.workload, args.steps, args.warmup, args.scale, stream, barrier, world,
                                              sample_clocks_on=local_rank if rank == 0 else None)

    # the other BASELINE configs, same measurement, in the same run (fewer steps: they only need a stable mean)
    extra = {}
    if not args.no_extra and not args.gather_only:
        for name in ("fft", "fir", "resample", "reechan", "rif_fft", "fir_long", "ola"):
            if name == args.workload:
                continue
            r, _ = measure_workload(name, min(args.steps, 6), 3, args.scale, stream, barrier, world,
                                    sample_clocks_on=local_rank if rank == 0 else None)
            r.pop("config")
            extra[name] = r

    strong = gath = None
    if world > 1 and not args.no_extra and args.workload == "ola" and 256 % world == 0:
        strong, gath = strong_and_gathered(min(args.steps, 6), 3, stream, barrier, world, rank)

    # end to end through the C ABI with host buffers: every rank drives its own GPU, same metric
    e2e = e2e_page = None
    if not args.no_e2e:
        torch.cuda.empty_cache()

        def run_e2e(pageable):
            m = e2e_measure(args.workload, 2, 1, barrier if world > 1 else None, world, pageable)
            sec = m.pop("seconds_per_step")
            smp = m.pop("samples_per_step")
            cop = m.pop("copy_only_seconds", None)
            if world > 1:
                t = torch.tensor([sec, cop or 0.0], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec, cop = float(t[0].item()), (float(t[1].item()) or None)
            d = dict(value=smp * world / sec / 1e9, **m)
            if cop:
                # what the host link alone allows for these buffers (both directions at once, all ranks at the same time)
                d["link_ceiling"] = {"value": smp * world / cop / 1e9, "unit": "Gsamples/s",
                                     "gbs_each_way_per_gpu": m["h2d_bytes_per_step"] / cop / 1e9,
                                     "how": "H2D of the step's input and D2H of its output issued together, no kernel"}
            return d
        e2e = run_e2e(False)
        if not args.no_extra:
            e2e_page = run_e2e(True)
            if rank == 0:
                e2e["pageable"] = {k: e2e_page[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "sample", "host_memory")}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            v, ms, threads, kind, sample = cpu_run(args.workload, 2, 1)
            cpu = {"value": v, "unit": "Gsamples/s", "cores": threads, "kind": kind, "sample": sample,
                   "flags": "-O3 -march=x86-64 (the reference's release flags, core/std-makefile-defs:171)"}
            nat = cpu_run_native(args.workload)
            if nat:
                cpu["value_avx2_build"] = nat
        line = {"metric": "Gsamples/s filtered (cf32)", "value": head["value"], "unit": "Gsamples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": head["config"], "roofline": head["roofline"], "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": head["gpu_launches"], "clocks": head["clocks"]}
        if extra:
            line["workloads"] = extra
        if strong:
            line["strong"] = strong
            line["gathered"] = gath
        if affinity:
            line["cpu_affinity_rank0"] = f"{len(affinity)} CPUs local to GPU {local_rank}"
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_run_native(workload):
    """SURVEY 8(d) also asks for the CPU figure of an -march=native-class build (timing only: FMA contraction makes it
    differ from the parity build).  The GPU box's CPU is not the build container's, so the portable stand-in is
    x86-64-v3 (AVX2 + FMA), built by `make -C oracle native` into oracle/_ref/libtsdref_v3.so."""
    so = os.path.join(ROOT, "oracle", "_ref", "libtsdref_v3.so")
    if not os.path.exists(so):
        return None
    try:
        env = dict(os.environ, TSDREF_LIB=so)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "2",
                              "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env).stdout.strip().splitlines()
        return {"value": json.loads(out[-1])["value"], "unit": "Gsamples/s", "flags": "-O3 -march=x86-64-v3 (AVX2 + FMA)"}
    except Exception:
        return None


if __name__ == "__main__":
    main()
