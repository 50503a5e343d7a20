"""gpurun_out/traffic_<workload>.csv (ncu --csv, profiles/tools/traffic_capture.sh) -> profiles/r02_traffic_<workload>.json.
Usage: python profiles/tools/traffic_json.py ola 68719476736 [kernels per timed launch]
(bench.py times one exec of the FFT plan as ONE launch although the staged schedule issues 256 stage kernels for it: the
third argument sums that many captured kernels into one launch.)"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
wl, alg = sys.argv[1], float(sys.argv[2])
group = int(sys.argv[3]) if len(sys.argv) > 3 else 1
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f"traffic_{wl}.csv"))) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
per = {}
for r in rows:
    per.setdefault(r[ix["ID"]], {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
launches = []
for k, d in per.items():
    rd = d["dram__bytes_read.sum"][0] * scale[d["dram__bytes_read.sum"][1]]
    wr = d["dram__bytes_write.sum"][0] * scale[d["dram__bytes_write.sum"][1]]
    launches.append({"kernel": d["kernel"], "dram_read": rd, "dram_write": wr, "time": d["gpu__time_duration.sum"]})
n = len(launches) / group
tot = sum(l["dram_read"] + l["dram_write"] for l in launches) / n
out = {"workload": wl, "dram_bytes_per_launch": tot, "dram_read_per_launch": sum(l["dram_read"] for l in launches) / n,
       "dram_write_per_launch": sum(l["dram_write"] for l in launches) / n, "algorithmic_bytes_per_launch": alg,
       "ratio_to_algorithmic": tot / alg, "launches_captured": n, "kernels_per_launch": group, "kernel": launches[0]["kernel"],
       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum of bench.py --workload {wl} at the full BASELINE size (profiles/tools/traffic_capture.sh)"}
json.dump(out, open(os.path.join(ROOT, "profiles", f"r02_traffic_{wl}.json"), "w"), indent=1)
print(json.dumps(out))
