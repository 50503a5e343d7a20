#!/usr/bin/env python
"""gpurun_out/{launches_resample_tc{1,0}.csv, prof_resample_tc{1,0}.ncu-rep} of profiles/run_profiles_resample.sh -> tracked summaries:
   profiles/r01_resample_launches.csv (both kernels), profiles/r01_resample_ncu.txt, profiles/traffic_resample.json"""
import collections, csv, io, json, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
SAMPLES = 512 * (1 << 20)   # input samples of one launch (a step() at BASELINE size is 8 launches of 512 channels x 1 Mi)
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}
with open(os.path.join(ROOT, "profiles", "r01_resample_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none\n")
    f.write("# one step() call of filtre_itrp 147/160 (sinc LUT 64 x 257) on 512 channels x 8 Mi cf32 (profiles/run_profiles_resample.sh)\n")
    f.write("variant,id,kernel,grid,block,duration_ns,dram_read_bytes,dram_write_bytes\n")
    traffic = {}
    for tc in (1, 0):
        rows = [r for r in csv.reader(open(os.path.join(OUT, f"launches_resample_tc{tc}.csv"))) if len(r) > 10 and r[0].isdigit()]
        L = collections.OrderedDict()
        for r in rows:
            d = L.setdefault(int(r[0]), {"kernel": r[4].split("(")[0].replace("void ", "").replace("tsdgpu::", ""), "grid": r[8], "block": r[7]})
            d[r[12]] = float(r[14].replace(",", "")) * UNIT.get(r[13], 1)
        agg = collections.OrderedDict()
        for i, d in L.items():
            t, rd, wr = d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
            f.write(f"{'tensor' if tc else 'fma'},{i},{d['kernel']},\"{d['grid']}\",\"{d['block']}\",{t:.0f},{rd:.0f},{wr:.0f}\n")
            a = agg.setdefault(d["kernel"], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
        tot = sum(a[1] for a in agg.values())
        for k, a in agg.items():
            f.write(f"# {'tensor' if tc else 'fma'}: {k}: {a[0]} launches, mean {a[1]/a[0]/1e3:.1f} us, {100*a[1]/tot:.1f} % of device time, "
                    f"DRAM {(a[2]+a[3])/a[0]/1e6:.1f} MB per launch\n")
        main = [a for k, a in agg.items() if "hist" not in k][0]
        traffic[tc] = (main[2] + main[3]) / main[0]
json.dump({"kernel": "resamp_tc_kernel", "dram_bytes_per_launch": traffic[1], "samples_in_profiled_launch": SAMPLES,
           "dram_bytes_per_sample": traffic[1] / SAMPLES, "fma_kernel_dram_bytes_per_sample": traffic[0] / SAMPLES,
           "note": "mean over the 8 launches of one step() at the full BASELINE size (each 512 channels x 1 Mi in, x147/160 out), caches not flushed between launches; algorithmic bytes per input sample = 8 + 8*147/160 = 15.35"},
          open(os.path.join(ROOT, "profiles", "traffic_resample.json"), "w"))
print("DRAM B/sample tensor %.2f fma %.2f" % (traffic[1] / SAMPLES, traffic[0] / SAMPLES))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second"]
with open(os.path.join(ROOT, "profiles", "r01_resample_ncu.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none, one launch (512 channels x 1 Mi-input slice, sinc LUT 64 x 257, ratio 147/160) of each resampler kernel\n")
    for tc in (1, 0):
        out = subprocess.run(["ncu", "-i", os.path.join(OUT, f"prof_resample_tc{tc}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(out)))
        hdr, units, vals = rr[0], rr[1], rr[2]
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        f.write(f"== {d.get('Kernel Name', '?')[:90]}\n")
        for k in hdr:
            if k in KEYS or k.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed") or k.endswith("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"):
                f.write(f"{k:100s} {u[k]:16s} {d[k]}\n")
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("per_issue_active.ratio"):
                try:
                    if float(d[k]) >= 0.3: f.write(f"{k:100s} {u[k]:16s} {d[k]}\n")
                except ValueError:
                    pass
    sass = subprocess.run("cuobjdump -sass %s | grep -c UTCHMMA" % os.path.join(ROOT, "libtsd_b200", "libtsdgpu.so"), shell=True, capture_output=True, text=True).stdout.strip()
    f.write(f"# SASS evidence: {sass} UTCHMMA (tcgen05.mma) instructions, LDTM/STTM (tcgen05.ld/st) in fir_tc_kernel and resamp_tc_kernel (cuobjdump -sass libtsdgpu.so)\n")
