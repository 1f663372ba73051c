#!/usr/bin/env python
"""Benchmark of the generator forward hot path (BASELINE.json metric: fields/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one generator forward over one batch of synthetic input.  Workload at every N:
BASELINE.json configs[1] -- 64 sequences x 8 timesteps x 96x96, 3 image + 20 noise channels,
random-init weights with non-trivial BatchNorm statistics -- PER GPU (weak scaling: inference
shards by independent sequences with no data-path collective).

Prints ONE JSON line (rank 0).  `value` = fields/s with inputs resident in HBM; `e2e` = the same
metric through the host-buffer C-ABI call (pinned host -> device copies and the device -> host
read of the result inside the timed region); `roofline` = the dominant kernel against the
measured bf16 peak (MEASURED_PEAKS.json); `cpu_baseline` = the CPU restatement of the reference
graph (oracle/torch_port.py, "port": TensorFlow 2.4.3 cannot be installed here) on a bounded
sample, timed on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- workload (BASELINE.json configs[1])
B, T, S, CIN, CNOISE, COUT = 64, 8, 96, 3, 20, 2
FIELDS_PER_STEP = B * T
FLOP_PER_FIELD = 3_827_367_936            # BASELINE.md §2 (dense formulation of the reference graph)
# per-(b,t) MMAC of each stage on the reference's dense formulation
STAGE_MMAC = {"conv8x8s2": 434.110464, "conv4x4s2": 150.994944, "convlstm": 679.477248, "conv3x3": 42.467328,
              "convT2x2s2": 14.155776, "upconvT5x5": 589.824, "conv3x3_out": 2.654208}
# algorithmic HBM bytes per field of the bandwidth-bound stages
STAGE_BYTES = {"pack_input": 96 * 96 * 23 * 4 + 102 * 102 * 24 * 2,
               "border_fix": 2 * (4 * 104 * 160 * 2) + 2 * (96 * 192 * 4),
               "conv3x3_out": 96 * 96 * 16 * 2 + 96 * 96 * 2 * 4}
CPU_SAMPLE_B = 8                            # bounded CPU sample: 8 sequences x 8 timesteps


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def synth_inputs(seed):
    rng = np.random.default_rng(seed)
    image = rng.standard_normal((B, T, S, S, CIN), dtype=np.float32)
    noise = rng.standard_normal((B, T, S, S, CNOISE), dtype=np.float32) * np.float32(0.1)
    return image, noise


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_fields_per_s(steps, warmup):
    """Times the CPU restatement of the reference generator graph on all host threads."""
    import torch
    from oracle.generator import synthetic_generator_weights
    from oracle.torch_port import TorchGenerator
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = TorchGenerator(synthetic_generator_weights(0), torch.float32)
    rng = np.random.default_rng(1)
    image = torch.from_numpy(rng.standard_normal((CPU_SAMPLE_B, T, S, S, CIN), dtype=np.float32))
    noise = torch.from_numpy(rng.standard_normal((CPU_SAMPLE_B, T, S, S, CNOISE), dtype=np.float32) * np.float32(0.1))
    for _ in range(warmup):
        gen.forward(image, noise)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        gen.forward(image, noise)
        times.append(time.perf_counter() - t0)
    fields = CPU_SAMPLE_B * T
    return fields / (sum(times) / len(times)), cores, times


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, cores, times = cpu_port_fields_per_s(args.steps, min(args.warmup, 2))
    sample = (f"{CPU_SAMPLE_B} sequences x {T} timesteps x {S}x{S} per step (1/{B // CPU_SAMPLE_B} of the GPU batch), "
              f"fp32 torch-CPU restatement of models.py:9-73 on {cores} threads ({cpu_model_name()}); "
              "TensorFlow 2.4.3 not installable")
    line = {"impl": "reference", "metric": "generator_fields_per_sec", "value": v, "unit": "fields/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 2),
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"generator inference, {B} sequences x {T} timesteps x {S}x{S}, "
                                   f"{CIN}+{CNOISE} input channels, fixed noise (BASELINE configs[1])",
                       "step_sample": f"{CPU_SAMPLE_B}x{T} fields per timed step"},
            "cpu_baseline": {"value": v, "unit": "fields/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "fields/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from wind_downscaling_gan_b200.gan.models import make_generator
    from oracle.generator import synthetic_generator_weights  # weights only; the oracle is not on the timed path

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    peaks = load_peaks()
    gen = make_generator(S, CIN, CNOISE, COUT, T)
    gen.set_weights(synthetic_generator_weights(0))
    image_h, noise_h = synth_inputs(100 + rank)
    image_p = torch.from_numpy(image_h).pin_memory()
    noise_p = torch.from_numpy(noise_h).pin_memory()
    out_p = torch.empty((B, T, S, S, COUT), dtype=torch.float32).pin_memory()
    image_d = image_p.cuda(non_blocking=True)
    noise_d = noise_p.cuda(non_blocking=True)
    out_d = torch.empty((B, T, S, S, COUT), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput (+ per-stage events for the roofline)
    gen.set_profiling(True)
    for _ in range(args.warmup):
        gen.forward_device(image_d, noise_d, out_d)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stage_acc = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        gen.forward_device(image_d, noise_d, out_d)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    # stage times: re-run K profiled steps, reading the events after each (reads sync, so not in the timed loop)
    for _ in range(args.steps):
        gen.forward_device(image_d, noise_d, out_d)
        for k, v in gen.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    gen.set_profiling(False)
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
    ms_per_step = ms_total / args.steps
    value = world * FIELDS_PER_STEP / (ms_per_step * 1e-3)

    # ---------------- end to end through the host-buffer C-ABI call
    for _ in range(2):
        gen.predict_host(image_p, noise_p, out_p)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        gen.predict_host(image_p, noise_p, out_p)
    ev1.record()
    barrier()
    e2e_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    e2e_value = world * FIELDS_PER_STEP / (e2e_ms * 1e-3)
    checksum = float(out_p.double().abs().mean())
    # same call with the noise drawn on the device by the library's FlexibleNoiseGenerator (api.py:136 does this in TF)
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    ng = FlexibleNoiseGenerator((B, T, S, S, CNOISE), std=0.1, random_seed=7)
    for _ in range(2):
        gen.predict_host_gen_noise(image_p, ng, out_p)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        gen.predict_host_gen_noise(image_p, ng, out_p)
    ev1.record()
    barrier()
    e2e_gn_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps

    if rank == 0:
        # dominant kernel = stage with the largest device time
        stages = {}
        for name, ms in stage_ms.items():
            launches = T if name == "convlstm" else 1
            ent = {"ms_per_step": ms, "launches_per_step": launches, "share": ms / sum(stage_ms.values())}
            if name in STAGE_MMAC and name != "conv3x3_out":
                tf = 2 * STAGE_MMAC[name] * 1e6 * FIELDS_PER_STEP / (ms * 1e-3) / 1e12
                ent.update(bound="tensor", achieved=tf, unit="TFLOP/s", frac=tf / peaks["bf16_tflops_sustained"])
            else:
                gbs = STAGE_BYTES[name] * FIELDS_PER_STEP / (ms * 1e-3) / 1e9
                ent.update(bound="hbm", achieved=gbs, unit="GB/s", frac=gbs / peaks["hbm_gbs"])
            stages[name] = ent
        dom = max(stage_ms, key=stage_ms.get)
        d = stages[dom]
        traffic = None
        try:   # DRAM bytes per launch of that kernel from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[dom]["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"],
                    "peak": peaks["bf16_tflops_sustained"] if d["bound"] == "tensor" else peaks["hbm_gbs"],
                    "unit": d["unit"], "frac": d["frac"], "traffic": traffic,
                    "avg_launch_ms": d["ms_per_step"] / d["launches_per_step"], "peak_source": peaks["source"] +
                    (" (sustained bf16: kernel timed inside a long step)" if d["bound"] == "tensor" else ""),
                    "whole_forward": {"achieved": FLOP_PER_FIELD * value / world / 1e12, "unit": "TFLOP/s",
                                      "frac": FLOP_PER_FIELD * value / world / 1e12 / peaks["bf16_tflops_sustained"]},
                    "stages": stages}
        if args.no_cpu_baseline:
            cpu = None
        else:
            v, cores, _ = cpu_port_fields_per_s(3, 1)
            cpu = {"value": v, "unit": "fields/s", "cores": cores, "kind": "port",
                   "sample": f"{CPU_SAMPLE_B} sequences x {T} timesteps x {S}x{S} (1/{B // CPU_SAMPLE_B} of the GPU batch), "
                             f"mean of 3 after 1 warm-up, fp32 torch-CPU restatement of the reference graph on {cores} "
                             f"threads ({cpu_model_name()}); TensorFlow 2.4.3 not installable"}
        line = {"metric": "generator_fields_per_sec", "value": value, "unit": "fields/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"generator inference, {B} sequences x {T} timesteps x {S}x{S} per GPU, "
                                       f"{CIN}+{CNOISE} input channels, fixed noise (BASELINE configs[1])",
                           "fields_per_step_per_gpu": FIELDS_PER_STEP, "weights": "synthetic, seed 0, non-trivial BN stats",
                           "parallelism": f"independent sequences x{world}, no collective",
                           "l2": "inputs 434 MB + 1.6 GB activations per step exceed the 126 MB L2",
                           "tolerance": "rel-L2 <= 1e-2 vs float64 oracle (bf16 operands, fp32 accumulate)"},
                "e2e": {"value": e2e_value, "unit": "fields/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(image_p.numel() * 4 + noise_p.numel() * 4),
                        "d2h_bytes_per_step": int(out_p.numel() * 4), "api": "wdg_generator_predict_host (pinned host buffers)"},
                "e2e_device_noise": {"value": world * FIELDS_PER_STEP / (e2e_gn_ms * 1e-3), "unit": "fields/s", "ms_per_step": e2e_gn_ms,
                                     "h2d_bytes_per_step": int(image_p.numel() * 4), "d2h_bytes_per_step": int(out_p.numel() * 4),
                                     "api": "wdg_generator_predict_host_gen_noise (noise generated on the device, as api.py:136 does)"},
                "gpu_launches": gen.launches_per_forward() * args.steps,
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "output_abs_mean": checksum}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
