"""Turns an `ncu --set full` report into the small per-kernel table committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("lts__t_bytes.sum", "l2_MB"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [n for _, n in METRICS]
    print("| # | kernel | " + " | ".join(names) + " |")
    print("|---|---|" + "---|" * len(names))
    for k, r in enumerate(data):
        kn = r[col["Kernel Name"]]
        kn = kn.replace("void wdg::", "").split("(")[0]
        vals = []
        for m, n in METRICS:
            if m not in col:
                vals.append("n/a")
                continue
            v, u = r[col[m]], units[col[m]]
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                vals.append(v)
                continue
            if n == "time_us":
                f = f / 1e3 if u in ("ns", "nsecond") else (f * 1e3 if u in ("ms", "msecond") else f)
            if n.endswith("_MB"):
                f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * f
            if n == "sm_ghz":
                f = {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0}.get(u, 1.0) * f
            vals.append(f"{f:.3g}" if abs(f) < 1000 else f"{f:.0f}")
        print(f"| {k} | {kn} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
