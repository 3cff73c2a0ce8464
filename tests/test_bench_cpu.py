"""CPU-side checks of bench.py: the reference arm's JSON line (contract keys, same workload string as the product arm) and
the shard-plan / schedule-free parts that do not need a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["metric"] == "J/K Fock builds/sec" and line["unit"] == "builds/s" and line["higher_is_better"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    # the same workload string as the product arm builds for N = 494 (the driver compares them)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"]["workload"] == bench.workload_string(494)
    # measured, not extrapolated, step time; the per-build figure says that it is scaled
    assert line["ms_per_step"] > 0 and "scaled" in line["cpu_baseline"]["sample"]
