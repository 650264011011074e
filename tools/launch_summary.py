#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total and share.

    python tools/launch_summary.py gpurun_out/launches_X.csv [--md] > profiles/launches_X.md
"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i
            break
    else:
        raise SystemExit("no header in " + path)
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    out = []
    for r in rows[start + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        out.append((r[ki], v * scale))
    return out


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("molclr::", "")
    return name[:90]


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    launches = load(path)
    agg = collections.OrderedDict()
    for n, us in launches:
        a = agg.setdefault(short(n), [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(v[1] for v in agg.values())
    items = sorted(agg.items(), key=lambda kv: -kv[1][1])
    if md:
        print(f"Source: `{path}` — {len(launches)} launches, {total / 1e3:.3f} ms summed device time "
              "(ncu per-launch times are cold-cache and serialised: compare SHARES).\n")
        print("| kernel | launches | total us | us/launch | share |")
        print("|---|---:|---:|---:|---:|")
        for n, (c, us) in items:
            print(f"| `{n}` | {c} | {us:.1f} | {us / c:.1f} | {100 * us / total:.1f}% |")
    else:
        for n, (c, us) in items:
            print(f"{us:10.1f}us {c:4d} {us / c:8.1f}us/launch {100 * us / total:5.1f}%  {n}")
        print(f"total {total / 1e3:.3f} ms over {len(launches)} launches")


if __name__ == "__main__":
    main()
