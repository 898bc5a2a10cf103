#!/usr/bin/env python3
"""Per-kernel summary of an ncu launch list (tools/gpu_check.sh):
   python tools/launch_summary.py gpurun_out/launches.csv [last_n_launches] > profiles/rX_launches_summary.txt"""
import collections, csv, statistics, sys
path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 1626
rows, hdr = [], None
for r in csv.reader(open(path)):
    if len(r) < 6:
        continue
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    try:
        name, val, unit = r[hdr.index("Kernel Name")], float(r[hdr.index("Metric Value")]), r[hdr.index("Metric Unit")]
    except Exception:
        continue
    val = val / 1e3 if unit == "ns" else (val * 1e3 if unit == "ms" else val)
    rows.append((name.split("(")[0], val))
sel = rows[-last:]
data = collections.defaultdict(list)
for k, v in sel:
    data[k].append(v)
tot = sum(v for _, v in sel)
print(sys.argv[3] if len(sys.argv) > 3 else "ncu --metrics gpu__time_duration.sum --clock-control none python bench.py ...")
print(f"last {len(sel)} of {len(rows)} launches; cold-cache, serialised launches: compare SHARES; total {tot/1e3:.2f} ms")
for k, v in sorted(data.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:56]:56s} n={len(v):5d} sum={sum(v)/1e3:8.2f} ms  share={100*sum(v)/tot:5.1f}%  median={statistics.median(v):7.1f} us  min={min(v):6.1f} max={max(v):7.1f}")
