#!/usr/bin/env python
"""gpurun_out/{launches_ola.csv, prof_ola.ncu-rep} of profiles/run_profiles_ola_staged.sh -> tracked summaries:
   profiles/r01_ola_launches.csv   every launch of one step(): kernel, grid, device time, DRAM bytes (+ shares)
   profiles/r01_ola_ncu.txt        key counters of one launch of each stage kernel (ncu --set full)
   profiles/traffic_ola.json       DRAM bytes per input sample of the whole step (read by bench.py)"""
import collections, csv, io, json, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
SAMPLES = 4 * (1 << 24)   # --scale 0.0157 -> 4 channels x 16 Mi
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}

rows = [r for r in csv.reader(open(os.path.join(OUT, "launches_ola.csv"))) if len(r) > 10 and r[0].isdigit()]
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(int(r[0]), {"kernel": r[4].split("(")[0].replace("void ", "").replace("tsdgpu::", ""), "grid": r[8], "block": r[7]})
    d[r[12]] = float(r[14].replace(",", "")) * UNIT.get(r[13], 1)
agg = collections.OrderedDict()
with open(os.path.join(ROOT, "profiles", "r01_ola_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none\n")
    f.write("# one step() of filtre_fft (K=4095, Ne=61441, N=65536) on 4 channels x 16 Mi samples: 30 chunks x 3 stage kernels + carry update\n")
    f.write("id,kernel,grid,block,duration_ns,dram_read_bytes,dram_write_bytes\n")
    for i, d in L.items():
        t, rd, wr = d.get("gpu__time_duration.sum", 0), d.get("dram__bytes_read.sum", 0), d.get("dram__bytes_write.sum", 0)
        f.write(f"{i},{d['kernel']},\"{d['grid']}\",\"{d['block']}\",{t:.0f},{rd:.0f},{wr:.0f}\n")
        a = agg.setdefault(d["kernel"], [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr
    tot = sum(a[1] for a in agg.values())
    f.write("# share of the step's device time by kernel (launches serialised by ncu: compare shares, not absolutes)\n")
    for k, a in agg.items():
        f.write(f"# {k}: {a[0]} launches, {a[1]/1e3:.1f} us, {100*a[1]/tot:.1f} %, DRAM read {a[2]/1e6:.1f} MB, write {a[3]/1e6:.1f} MB\n")
rd = sum(a[2] for a in agg.values()); wr = sum(a[3] for a in agg.values())
json.dump({"kernel": "ola64k_stage<0|1|2> (all launches of one step)", "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
           "samples_in_profiled_launch": SAMPLES, "dram_bytes_per_sample": (rd + wr) / SAMPLES,
           "note": "sum over the 90 stage launches + carry update of one step() at 4 channels x 16 Mi (profiles/run_profiles_ola_staged.sh), "
                   "caches not flushed between launches; the last chunks' output can still sit dirty in L2, so dram_write can undercount"},
          open(os.path.join(ROOT, "profiles", "traffic_ola.json"), "w"))
print("DRAM bytes/sample", (rd + wr) / SAMPLES, {k: round(100 * a[1] / tot, 1) for k, a in agg.items()})

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
out = subprocess.run(["ncu", "-i", os.path.join(OUT, "prof_ola.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(out)))
hdr, units = rr[0], rr[1]
with open(os.path.join(ROOT, "profiles", "r01_ola_ncu.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --cache-control none, one launch of each stage kernel of the staged filtre_fft\n")
    f.write("# (37 blocks x 16 tiles = 592 CTAs per launch, profiled alone: in a real step four such launches overlap); see run_profiles_ola_staged.sh\n")
    for vals in rr[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        f.write(f"== {d.get('Kernel Name', '?')[:90]}\n")
        for k in KEYS:
            if k in d: f.write(f"{k:90s} {u[k]:16s} {d[k]}\n")
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("per_issue_active.ratio"):
                try:
                    if float(d[k]) >= 0.3: f.write(f"{k:90s} {u[k]:16s} {d[k]}\n")
                except ValueError:
                    pass
