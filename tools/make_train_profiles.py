"""Regenerates the training-path files under profiles/ from the raw captures in gpurun_out/ (see the gpurun command
in each file's header).  Usage: python tools/make_train_profiles.py"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def clean(s):
    return s.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")


def summarize(rep):
    return clean(subprocess.run([sys.executable, os.path.join(P, "summarize_ncu.py"), os.path.join(G, rep)],
                                capture_output=True, text=True).stdout)


def conv_table():
    rows = [json.loads(l) for l in open(os.path.join(G, "tc_conv_bench_final.jsonl"))]
    by = {}
    for r in rows:
        by.setdefault(r["layer"], {})[r["prec"]] = r
    out = ["# r1 — training convolution GEMMs per layer shape (B200, `python tools/bench_tc_conv.py`, CUDA events, 10 reps)\n",
           "Layer shapes of one WGAN step at B=8, T=24 (192 images; '1 step' rows: the 8 images of one ConvLSTM timestep).",
           "fp32 = CUDA-core implicit GEMM (`train_ops.cu`); tf32 / bf16 = tcgen05 implicit GEMM (`train_gemm_tc.cu`), fp32 accumulate.",
           "Numbers are TFLOP/s of the dense convolution FLOPs (2·N·Ho·Wo·k²·Ci·Co); tf32 ms in the last column.\n",
           "| layer | GFLOP | forward fp32 / tf32 / bf16 | backward-data | backward-weight | tf32 ms (fwd / bwd-data / bwd-weight) |",
           "|---|---|---|---|---|---|"]
    for k, v in by.items():
        f = lambda key: " / ".join(f"{v[p][key]:.0f}" if v[p][key] >= 10 else f"{v[p][key]:.1f}" for p in ("fp32", "tf32", "bf16"))
        t = v["tf32"]
        out.append(f"| {k} | {t['gflop']:.1f} | {f('fwd_tflops')} | {f('bwd_data_tflops')} | {f('bwd_weight_tflops')} | "
                   f"{t['fwd_ms']:.3f} / {t['bwd_data_ms']:.3f} / {t['bwd_weight_ms']:.3f} |")
    open(os.path.join(P, "r1_train_conv_table.md"), "w").write("\n".join(out) + "\n")


def ncu_kernels():
    out = ["# r1 — ncu `--set full --clock-control none` of the tcgen05 training GEMM (`tc_gemm_kernel`), B200\n",
           "Command: `ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 3 python tools/ncu_tc_conv.py <prec> N H W Ci k Co s p`",
           "(second forward / backward-data / backward-weight launch of the given layer shape; per-launch times are cold-cache).\n",
           "## ConvLSTM input convolution of the generator, 192 x 24x24, 128 -> 512, 3x3 — tf32", summarize("prof_tc_lstm.ncu-rep"),
           "## same layer — bf16", summarize("prof_tc_lstm_bf16.ncu-rep"),
           "## critic ConvLSTM(16) recurrent convolution, one timestep: 8 x 96x96, 16 -> 64, 3x3 — tf32", summarize("prof_tc_dl.ncu-rep")]
    tail = open(os.path.join(P, "r1_train_ncu_kernels.md")).read()
    tail = tail[tail.index("## Stall picture"):] if "## Stall picture" in tail else ""
    open(os.path.join(P, "r1_train_ncu_kernels.md"), "w").write("\n".join(out) + "\n" + tail)


def launches():
    agg = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "agg_launches.py"),
                          os.path.join(G, "train_step_tf32_launches.csv"), "40"], capture_output=True, text=True).stdout
    first = agg.splitlines()[0]
    tot_ms, n = float(first.split()[1]), int(first.split()[4])
    step = json.load(open(os.path.join(G, "bench_train_tf32.json")))["ms_per_step"]
    out = ["# r1 — launch list of ONE WGAN train step (tf32), B200\n",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv python bench_train.py --precision tf32 --steps 1 --warmup 1 --no-cpu-baseline`",
           f"(the warm-up step and the timed step are both in the list: {n // 2} launches and {tot_ms / 2:.0f} ms of summed kernel time per",
           f"step — serialised, cold-cache — against {step:.0f} ms per step measured with CUDA events in the same build: the step is",
           "GPU-bound, launch gaps are hidden; SHARES are what this list is for).\n", "```", agg.rstrip(), "```"]
    open(os.path.join(P, "r1_train_launches.md"), "w").write("\n".join(out) + "\n")


def main():
    conv_table()
    ncu_kernels()
    launches()
    for prec in ("fp32", "tf32", "bf16"):
        shutil.copy(os.path.join(G, f"bench_train_{prec}.json"), os.path.join(P, f"r1_bench_train_{prec}.json"))
    print("profiles/ refreshed")


if __name__ == "__main__":
    main()
