#!/usr/bin/env python
"""Top stall sites of one launch of an ncu report (source page, SASS): python tools/ncu_hot.py rep launch_index [n]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:120])
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[col["# Samples"]].isdigit()]
tot = sum(int(r[col["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print("samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][col["# Samples"]]))[:n]:
    s = int(r[col["# Samples"]])
    top = sorted(((int(r[col[x]] or 0), x) for x in stalls), reverse=True)[:2]
    print(f"{i:5d} {s:6d} {100*s/tot:5.1f}%  {r[col['Source']].strip()[:70]:70s} {top}")
