#!/usr/bin/env python3
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py gpurun_out/prof_k_segment.ncu-rep [top_n] [launch_index]
Aggregates warp-stall samples and executed warp instructions of the SASS rows onto CUDA lines."""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
tables, cur, fname = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1]; cur = None; continue
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        cur = {"file": fname, "hdr": r, "rows": []}; tables.append(cur); continue
    if cur is not None and r:
        cur["rows"].append(r)
# tables repeat per launch; group by first-file recurrence
first = tables[0]["file"] if tables else None
groups, g = [], []
for t in tables:
    if t["file"] == first and g:
        groups.append(g); g = []
    g.append(t)
if g: groups.append(g)
sel = groups[launch] if launch < len(groups) else []
agg = defaultdict(lambda: [0.0, 0.0, ""])
stall_cols = None
stalls = defaultdict(float)
tot_s = tot_i = 0.0
for t in sel:
    h = t["hdr"]
    ci = h.index("Instructions Executed"); cs = h.index("# Samples")
    sc = [(i, n) for i, n in enumerate(h) if n.startswith("stall_")]
    line = None
    for r in t["rows"]:
        if r[0]:
            line = (t["file"].split("/")[-1], r[0]); agg[line][2] = r[1].strip()[:100]
        if line is None or len(r) <= max(ci, cs): continue
        try:
            ins = float(r[ci] or 0); smp = float(r[cs] or 0)
        except ValueError:
            continue
        agg[line][0] += smp; agg[line][1] += ins; tot_s += smp; tot_i += ins
        for i, n in sc:
            try: stalls[n] += float(r[i] or 0)
            except (ValueError, IndexError): pass
print(f"launch {launch}/{len(groups)}: samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
print("stalls:", ", ".join(f"{n[6:]} {100*v/max(tot_s,1):.0f}%" for n, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))
for (f, ln), (smp, ins, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*smp/max(tot_s,1):5.1f}% smp {100*ins/max(tot_i,1):5.1f}% ins  {f}:{ln}  {src}")
