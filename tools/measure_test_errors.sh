#!/bin/bash
# Runs the GPU suite with a vanishing gradient tolerance and lists, per test, the largest measured error (from the assertion messages).
mkdir -p gpurun_out
MOLCLR_TEST_RTOL_GRAD=1e-12 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_measure.log 2>&1
python - <<'PY'
import re
txt = open("gpurun_out/pytest_measure.log").read()
blocks = re.split(r"\n_{5,} (test_[^\n]+?) _{5,}\n", txt)
for name, body in zip(blocks[1::2], blocks[2::2]):
    errs = [float(x) for x in re.findall(r"\('[\w\.]+', ([0-9.e+-]+)\)", body)]
    m = re.search(r"assert ([0-9.e+-]+) < 1e-12", body)
    if m:
        errs.append(float(m.group(1)))
    pairs = re.findall(r"\('([\w\.]+)', ([0-9.e+-]+)\)", body)
    worst = max(pairs, key=lambda kv: float(kv[1])) if pairs else None
    print(f"{name:90s} max {max(errs) if errs else float('nan'):.3e}  {worst[0] if worst else ''}")
PY
