#!/usr/bin/env python
"""Turns gpurun_out/{r02_launches_*.csv, r02_prof_*.ncu-rep} (profiles/run_profiles_r02.sh) into the tracked round-2
summaries:  profiles/r02_<workload>_launches.csv (kernel, duration per launch, share of the step) and
profiles/r02_<workload>_ncu.txt (key counters, stall reasons per issue, hottest SASS lines).  Runs here (no GPU)."""
import collections, csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
STALLS = "smsp__average_warps_issue_stalled_"
SCALE = {"ola": 0.125, "fft": 0.0625, "fir": 0.25, "resample": 0.125}
for w in sys.argv[1:] or ("ola", "fft", "fir", "resample"):
    lc = os.path.join(OUT, f"r02_launches_{w}.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 10 and r[0].isdigit()]
        agg = collections.OrderedDict()
        with open(os.path.join(ROOT, "profiles", f"r02_{w}_launches.csv"), "w") as f:
            f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --workload {w} --scale {SCALE[w]} "
                    "--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extra  (3 warm-up steps + 1 timed step; cold-cache, serialised)\n")
            f.write("id,kernel,grid,block,duration_ns\n")
            for r in rows:
                name = r[4].split("(")[0].replace("void ", "")
                f.write(f"{r[0]},{name},\"{r[8]}\",\"{r[7]}\",{r[-1]}\n")
                agg.setdefault(name, [0, 0.0])
                agg[name][0] += 1
                agg[name][1] += float(r[-1])
            ours = lambda k: not k.startswith("at::") and "elementwise" not in k and "distribution" not in k
            tot = sum(v[1] for k, v in agg.items() if ours(k)) or 1.0
            f.write("# share of the library's device time by kernel (torch's own fill/random kernels excluded)\n")
            for k, v in agg.items():
                if ours(k):
                    f.write(f"# {k}: {v[0]} launches, {v[1]/1e3:.1f} us total, {100*v[1]/tot:.1f} %\n")
    rep = os.path.join(OUT, f"r02_prof_{w}.ncu-rep")
    if os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        with open(os.path.join(ROOT, "profiles", f"r02_{w}_ncu.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on, one launch of {d.get('Kernel Name','?')[:100]}\n")
            f.write(f"# command: profiles/run_profiles_r02.sh (workload {w}, --scale {SCALE[w]}: reduced channel count so ncu's replays stay short)\n")
            for k in KEYS:
                if k in d: f.write(f"{k:90s} {u[k]:16s} {d[k]}\n")
            for k in hdr:
                if k.startswith(STALLS) and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
                    f.write(f"{k:90s} {u[k]:16s} {d[k]}\n")
            src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
            srows = list(csv.reader(io.StringIO(src)))
            h = next((i for i, r in enumerate(srows) if "Source" in r and "# Samples" in r), None)
            if h is not None:
                sh = srows[h]; ix = {c: i for i, c in enumerate(sh)}
                data = [r for r in srows[h + 1:] if len(r) == len(sh)]
                stalls = [c for c in sh if c.startswith("stall_") and "Not Issued" not in c]
                tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
                by = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
                f.write(f"# warp-state samples over the SASS page: {tot} samples, {len(data)} instructions\n")
                f.write("# by reason (%): " + ", ".join(f"{k[6:]} {100*v/tot:.1f}" for k, v in sorted(by.items(), key=lambda kv: -kv[1]) if 100 * v / tot >= 0.5) + "\n")
                ops = collections.Counter()
                for r in data:
                    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
                    if op.startswith("@"): op = r[ix["Source"]].split()[1]
                    ops[op.split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
                tw = sum(ops.values()) or 1
                f.write("# executed warp instructions by opcode (%): " + ", ".join(f"{k} {100*v/tw:.1f}" for k, v in ops.most_common(14)) + "\n")
                f.write("# hottest SASS instructions (share of samples, dominant reason)\n")
                for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:12]:
                    s = int(r[ix["# Samples"]] or 0)
                    why = max(stalls, key=lambda k: int(r[ix[k]] or 0))
                    f.write(f"#   {100*s/tot:5.2f}% {why[6:]:20s} {r[ix['Source']][:90]}\n")
print("done")
