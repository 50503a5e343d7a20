import sys
sys.path.insert(0, ".")
sys.argv = ["bench.py"]
import bench, torch
import libtsd_b200
libtsd_b200.init(0)
m = bench.e2e_measure("ola", 2, 1, None, 1, True)
print("pageable", m["samples_per_step"] / m["seconds_per_step"] / 1e9, "Gsamples/s")
