#!/usr/bin/env python
"""In-order launch list of an ncu `--metrics gpu__time_duration.sum --csv` log, run-length compressed:
    python tools/seq_launches.py FILE.csv"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = None
seq = []
for r in rows:
    if "Kernel Name" in r:
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d["Metric Value"].replace(",", ""))
        except ValueError: continue
        u = d["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|wdg::|void ", "", d["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        seq.append((name, v, d.get("Grid Size", "")))
tot = sum(v for _, v, _ in seq)
print(f"total {tot/1e3:.3f} ms over {len(seq)} launches")
i = 0
while i < len(seq):
    # detect a repeating block of period p starting at i
    best = (1, 1)
    for p in range(1, 9):
        reps = 1
        while i + (reps + 1) * p <= len(seq) and [s[0] for s in seq[i + reps * p: i + (reps + 1) * p]] == [s[0] for s in seq[i: i + p]]:
            reps += 1
        if reps > 1 and reps * p > best[0] * best[1]:
            best = (p, reps)
    p, reps = best
    if reps > 1:
        blk = seq[i: i + p * reps]
        t = sum(v for _, v, _ in blk)
        print(f"  [{reps} x]  {t/1e3:8.3f} ms  " + " | ".join(f"{seq[i+k][0]} {sum(blk[j*p+k][1] for j in range(reps))/reps:.1f}us g{seq[i+k][2]}" for k in range(p)))
        i += p * reps
    else:
        print(f"            {seq[i][1]/1e3:8.3f} ms  {seq[i][0]} g{seq[i][2]}")
        i += 1
