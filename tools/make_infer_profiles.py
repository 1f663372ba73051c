"""Regenerates the inference-path files under profiles/ from raw captures in gpurun_out/:
  launches_<tag>.csv   ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv  python bench.py --steps 2 --warmup 3 --no-cpu-baseline
  prof_<tag>.ncu-rep  ncu --set full --clock-control none --import-source on -c 17            python bench.py --steps 1 --warmup 1 --no-cpu-baseline
  bench_<tag>.json / bench_ref_<tag>.json   python bench.py [--impl reference]
Usage: python tools/make_infer_profiles.py [tag]   (tag = r1, r2, ...: the round the captures belong to; default r2)"""
import collections
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
STAGES = ["pack_input", "conv8x8s2", "conv4x4s2"] + ["convlstm"] * 8 + ["conv3x3", "convT2x2s2", "border_fix", "border_fix",
                                                                       "upconvT5x5", "conv3x3_out"]


def clean(n):
    n = re.sub(r"\(anonymous namespace\)::|<unnamed>::|wdg::|void ", "", n)
    return re.sub(r"\(.*", "", n)


def launch_list():
    rows = list(csv.reader(open(os.path.join(G, f"launches_{TAG}.csv"), errors="replace")))
    hdr, agg = None, collections.OrderedDict()
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
            v = v / 1e3 if d["Metric Unit"] == "ns" else v
            k = clean(d["Kernel Name"])
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# {TAG} launch list (ncu --metrics gpu__time_duration.sum --clock-control none -c 400, `bench.py --steps 2 --warmup 3 --no-cpu-baseline`)\n",
           "Per-launch times are cold-cache and serialised: compare SHARES with the CUDA-event stage shares of profiles/" + TAG + "_bench_n1.json, not absolutes.",
           "conv_umma_kernel<BN,EPI>: EPI 0 = bias/LeakyReLU/BN affine, 1 = ConvLSTM gates. halo_conv_kernel<BN,NCHUNK,NTAP,TPS,EPI>: <128,3,8,2,1> = 8x8 s2 conv,",
           "<64,3,16,4,0> = fused upsample + 5x5 transposed conv, <16,1,9,3,2> = final 3x3 conv in super-pixel form; a trailing `true` = CTA-pair form.",
           "conv_pair_kernel<BN,EPI,PREC>: CTA-pair (tcgen05.mma.cta_group::2) implicit GEMM: <256,1> = all ConvLSTM steps in one persistent launch,",
           "<128,0> = 4x4 s2 conv.\n",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |")
    open(os.path.join(P, f"{TAG}_launch_list.md"), "w").write("\n".join(out) + "\n")
    shutil.copy(os.path.join(G, f"launches_{TAG}.csv"), os.path.join(P, f"{TAG}_launches.csv"))


def ncu_tables():
    rep = os.path.join(G, f"prof_{TAG}.ncu-rep")
    md = subprocess.run([sys.executable, os.path.join(P, "summarize_ncu.py"), rep], capture_output=True, text=True).stdout
    head = (f"# {TAG} — ncu `--set full --clock-control none --import-source on` of the generator forward (bench.py workload: 64 sequences x 8\n"
            "timesteps = 512 fields), B200.  Kernel order = launch order of the forward (pack, 8x8 s2, 4x4 s2, ConvLSTM -- all 8 steps in one\n"
            "persistent launch --, 3x3, convT 2x2, edge lines, border GEMM, fused upsample conv, final conv).  Cold first forward: use the ratios,\n"
            "not the absolute times.\n\n")
    open(os.path.join(P, f"{TAG}_ncu_kernels.md"), "w").write(head + md.replace("(anonymous namespace)::", ""))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def dram(r):
        t = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            t += float(r[col[m]].replace(",", "")) * scale.get(units[col[m]], 1.0)
        return t
    # first forward of the capture: from the first input-packing kernel to the one before the next
    names = [clean(r[col["Kernel Name"]]) for r in data]
    starts = [i for i, n in enumerate(names) if n.startswith("pack_input")]
    assert starts, names[:5]
    fwd = list(range(starts[0], starts[1] if len(starts) > 1 else len(data)))

    def stage_of(name, seen):
        if name.startswith("pack_input"): return "pack_input"
        if name.startswith("halo_conv_kernel<128"): return "conv8x8s2"
        if name.startswith(("conv_pair_kernel<256", "conv_umma_kernel<256")): return "convlstm"
        if name.startswith("conv_pair_kernel<128"): return "conv4x4s2"
        if name.startswith("conv_umma_kernel<128"): return "convT2x2s2" if "conv4x4s2" in seen else "conv4x4s2"
        if re.match(r"halo_conv_kernel<64, [24], 9, 3", name): return "conv3x3"
        if name.startswith(("edge_lines", "conv_umma_kernel<48")): return "border_fix"
        if re.match(r"halo_conv_kernel<64, [35], 16, 4", name): return "upconvT5x5"
        if name.startswith(("halo_conv_kernel<16", "final_conv3x3")): return "conv3x3_out"
        return None
    per = collections.OrderedDict()
    for i in fwd:
        st = stage_of(names[i], per)
        if st is not None:
            per.setdefault(st, []).append((names[i], dram(data[i])))
    out = collections.OrderedDict()
    for st, lst in per.items():
        if st == "convlstm" and len(lst) > 1:
            later = [b for _, b in lst[1:]]
            out[st] = {"dram_bytes_per_launch": sum(later) / len(later), "kernel": lst[0][0] + " (steps t>0)"}
        elif st == "convlstm":
            out[st] = {"dram_bytes_per_launch": lst[0][1], "kernel": lst[0][0] + " (all T steps in one persistent launch)"}
        elif st == "border_fix":
            out[st] = {"dram_bytes_per_launch": sum(b for _, b in lst), "kernel": " + ".join(k for k, _ in lst)}
        else:
            out[st] = {"dram_bytes_per_launch": lst[0][1], "kernel": lst[0][0]}
    out["_source"] = "ncu --set full --clock-control none (profiles/" + TAG + "_ncu_kernels.md), bench.py workload, per launch"
    json.dump(out, open(os.path.join(P, f"{TAG}_traffic.json"), "w"), indent=1)


def main():
    launch_list()
    ncu_tables()
    shutil.copy(os.path.join(G, f"bench_{TAG}.json"), os.path.join(P, f"{TAG}_bench_n1.json"))
    if os.path.exists(os.path.join(G, f"bench_ref_{TAG}.json")):
        shutil.copy(os.path.join(G, f"bench_ref_{TAG}.json"), os.path.join(P, f"{TAG}_bench_reference_arm.json"))
    print("profiles/ (inference) refreshed")


if __name__ == "__main__":
    main()
