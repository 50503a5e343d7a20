#!/usr/bin/env python
"""Turns gpurun_out/{launches_*.csv, prof_*.ncu-rep} into the tracked round summaries under profiles/:
   r01_<workload>_launches.csv  (kernel, duration ns per launch)      r01_<workload>_ncu.txt (key counters)
   traffic_<workload>.json      (DRAM bytes per launch of the dominant kernel, read by bench.py)"""
import csv, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
STALLS = "smsp__average_warps_issue_stalled_"
for w in ("ola", "fft", "fir", "resample"):
    lc = os.path.join(OUT, f"launches_{w}.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 10 and r[0].isdigit()]
        agg = collections.OrderedDict()
        with open(os.path.join(ROOT, "profiles", f"r01_{w}_launches.csv"), "w") as f:
            f.write("id,kernel,grid,block,duration_ns\n")
            for r in rows:
                name = r[4].split("(")[0].replace("void ", "")
                f.write(f"{r[0]},{name},\"{r[8]}\",\"{r[7]}\",{r[-1]}\n")
                agg.setdefault(name, [0, 0.0])
                agg[name][0] += 1
                agg[name][1] += float(r[-1])
            tot = sum(v[1] for k, v in agg.items() if "tsdgpu" in k) or 1.0
            f.write("# share of the library's device time by kernel (torch fill kernels excluded)\n")
            for k, v in agg.items():
                if "tsdgpu" in k:
                    f.write(f"# {k}: {v[0]} launches, {v[1]/1e3:.1f} us total, {100*v[1]/tot:.1f} %\n")
    rep = os.path.join(OUT, f"prof_{w}.ncu-rep")
    if os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        with open(os.path.join(ROOT, "profiles", f"r01_{w}_ncu.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, one launch of {d.get('Kernel Name','?')[:100]}\n")
            f.write(f"# command: see profiles/run_profiles.sh (workload {w}, reduced channel count so that ncu's replays stay short)\n")
            for k in KEYS:
                if k in d: f.write(f"{k:90s} {u[k]:16s} {d[k]}\n")
            for k in hdr:
                if k.startswith(STALLS) and k.endswith("per_issue_active.ratio"): f.write(f"{k:90s} {u[k]:16s} {d[k]}\n")
        rd = float(d["dram__bytes_read.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
        # input samples handled by the profiled launch (sizes of profiles/run_profiles.sh)
        samples = {"ola": 4 * (1 << 24), "fft": 256 * 65536, "fir": 64 * 65536, "resample": 64 * (1 << 20)}[w]
        json.dump({"kernel": d.get("Kernel Name", "?")[:80], "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                   "samples_in_profiled_launch": samples, "dram_bytes_per_sample": (rd + wr) / samples,
                   "note": "one ncu --set full capture at the reduced size of profiles/run_profiles.sh (per launch); small launches "
                           "leave part of the output dirty in L2, so dram_write can undercount"},
                  open(os.path.join(ROOT, "profiles", f"traffic_{w}.json"), "w"))
print("done")
