"""bench.py contract, CPU side: the reference arm (`--impl reference`) prints ONE JSON line with the keys the driver reads,
its `config` dict is the one the GPU arm prints for the same workload (`same_config`), and it never touches the GPU library."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_ref():
    sys.path.insert(0, ROOT)
    import oracle
    return oracle.have_ref()


@pytest.mark.parametrize("workload", ["ola", "fir"])
def test_reference_arm_line(workload):
    if not _have_ref():
        pytest.skip("oracle/_ref/libtsdref.so not built (needs /root/reference)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Gsamples/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.make_config(workload, 1)
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"] or "Gsamples" in d["metric"]
