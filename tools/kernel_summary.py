#!/usr/bin/env python
"""Summarises an `ncu --metrics <list> --csv` capture of one step (tools/gpu_r2_ncu.sh): per kernel instance the launch count and the
mean of every metric.    python tools/kernel_summary.py gpurun_out/kernels_X.csv [--md]"""
import collections
import csv
import re
import sys

path, md = sys.argv[1], "--md" in sys.argv
rows = list(csv.reader(open(path, errors="replace")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
ki, mi, vi, ui, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("molclr::", "")[:70]
    unit = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    if r[mi].startswith(("dram__bytes", "lts__t_bytes")):
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
    per.setdefault(name, collections.defaultdict(list))[r[mi]].append(v)
cols = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"), ("lts__t_bytes.sum", "L2 MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU smem %"), ("launch__registers_per_thread", "regs")]
mean = lambda xs: sum(xs) / len(xs) if xs else float("nan")
items = sorted(per.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"]))
total = sum(sum(v["gpu__time_duration.sum"]) for _, v in items)
if md:
    print(f"Source: `{path}` (one steady-state step, B = 4096 pairs; ncu replays kernels serialised with cold caches: compare shares and ratios).\n")
    print("| kernel | launches | total us | share | " + " | ".join(c for _, c in cols) + " |")
    print("|---|---:|---:|---:|" + "---:|" * len(cols))
for n, m in items:
    t = m["gpu__time_duration.sum"]
    vals = [mean(m[k]) for k, _ in cols]
    if md:
        print(f"| `{n}` | {len(t)} | {sum(t):.0f} | {100 * sum(t) / total:.1f}% | " + " | ".join(f"{v:.1f}" for v in vals) + " |")
    else:
        print(f"{n:70s} n={len(t):3d} tot={sum(t):8.0f}us " + " ".join(f"{c}={v:.1f}" for (_, c), v in zip(cols, vals)))
print(f"\ntotal {total / 1e3:.3f} ms" if not md else f"\nTotal {total / 1e3:.3f} ms summed device time.")
