"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
Usage: python tools/agg_launches.py FILE.csv [top_n]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr = None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            u = d["Metric Unit"]
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)   # -> us
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|wdg::", "", d["Kernel Name"])
            name = re.sub(r"\(.*", "", name)
            agg[name][0] += 1
            agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f"total {tot / 1e3:.2f} ms over {n} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1] / 1e3:9.2f} ms {v[0]:6d} x {v[1] / v[0]:9.1f} us {v[1] / tot * 100:5.1f}%  {k}")


if __name__ == "__main__":
    main()
