#!/usr/bin/env python
"""Top stalled SASS instructions (with source line) of the k-th kernel (default: first) in an .ncu-rep; also key raw metrics.
    python tools/ncu_hot.py file.ncu-rep [topN] [k]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30; kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for r in rows[2 + kidx:3 + kidx]:
    print(r[h.index('Kernel Name')][:100])
    for w in want:
        if w in h: print(f"   {w} = {r[h.index(w)]} {rows[1][h.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(kidx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc = h.index('Source'); isamp = h.index('# Samples'); iex = h.index('Instructions Executed')
stalls = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for r in rows[2:]:
    if len(r) < len(h) or r[0] == 'Kernel Name': break
    try: s = int(r[isamp])
    except ValueError: continue
    data.append((s, r))
tot = sum(s for s, _ in data)
agg = collections.Counter()
for s, r in data:
    for i in stalls:
        if r[i]: agg[h[i]] += int(r[i])
print('total samples', tot, 'by reason:', agg.most_common(8))
for s, r in sorted(data, key=lambda x: -x[0])[:top]:
    st = {h[i]: int(r[i]) for i in stalls if r[i] and int(r[i]) > 0}
    t3 = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{s:6d} {100 * s / tot:5.1f}% ex={r[iex]:>8} {r[isrc][:64]:64s} {t3}")
