"""bench.py's reference arm on the CPU (the only arm that runs without a GPU): one small step of the oracle port, the JSON line
contract of the tier (impl, metric / unit / higher_is_better of BASELINE.json's metric, cpu_baseline describing the run, e2e with
zero copies), and that the arm honours --steps / --warmup."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-batch", "32"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["unit"] == "molecules/s" and d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert "molecules" in d["metric"] and isinstance(base.get("metric", ""), str)
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "32 pairs" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
