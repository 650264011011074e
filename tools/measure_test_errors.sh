#!/bin/bash
# Runs the GPU suite with a vanishing gradient tolerance and lists, per test, the largest measured error (from the assertion messages).
mkdir -p gpurun_out
MOLCLR_TEST_RTOL_GRAD=1e-12 timeout 1200 python -m pytest tests -m gpu -q -vv > gpurun_out/pytest_measure.log 2>&1
python - <<'PY'
import re
txt = open("gpurun_out/pytest_measure.log").read()
blocks = re.split(r"\n_{5,} (test_[^\n]+?) _{5,}\n", txt)
for name, body in zip(blocks[1::2], blocks[2::2]):
    def fl(x):
        try:
            return float(x)
        except ValueError:
            return None
    pairs = [(k, fl(v)) for k, v in re.findall(r"\('([\w\.]+)', ([0-9.e+-]+)\)", body)]
    pairs = [(k, v) for k, v in pairs if v is not None]
    extra = [v for v in (fl(x) for x in re.findall(r"where ([0-9.e+-]+) = rel_err", body)) if v is not None]
    allv = [v for _, v in pairs] + extra
    worst = max(pairs, key=lambda kv: kv[1]) if pairs else ("", 0.0)
    print(f"{name[:95]:95s} max {max(allv) if allv else float('nan'):.3e}  {worst[0]}")
PY
