"""Reads bench.py's JSON line from stdin and prints the headline numbers."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("n", d.get("n_gpus"), "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]),
      "roof", round((d.get("roofline") or {}).get("frac", 0), 3), "launches", d.get("gpu_launches"))
