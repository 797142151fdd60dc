"""bench.py's JSON contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`: the reference's own CPU kernels from oracle/_ref, or the oracle port) prints ONE
JSON line on stdout carrying the base contract's keys, the same metric / unit / config as the CUDA arm,
and the cpu_baseline / e2e objects of the tier contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--tile", "32"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "particle-pushes/s per full PIC step" and d["unit"] == "particle-pushes/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"] and d["sample"] == d["cpu_baseline"]["sample"]
    assert "32^3 cells" in d["sample"] and "Juettner" in d["sample"]            # the line states what the CPU arm actually ran
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same workload description the CUDA arm prints
    import bench
    import argparse
    args = argparse.Namespace(cells=512, tile=32, ppc=16)       # --tile 32 keeps this CPU test short; the bench default is 64
    assert d["config"] == bench.workload_config(args, 1)
    assert "workload" in d["config"] and "model" not in d["config"]
