#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric: generator fields/sec; configs[3]: the WGAN training step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload infer|train] [--precision bf16|tf32 | fp32|tf32|bf16]

--workload infer (default): a "step" is one generator forward over one batch of synthetic input.  Workload at every
N: BASELINE.json configs[1] -- 64 sequences x 8 timesteps x 96x96, 3 image + 20 noise channels, random-init weights
with non-trivial BatchNorm statistics -- PER GPU (weak scaling: inference shards by independent sequences with no
data-path collective).  ONE JSON line (rank 0):
  value     fields/s with image AND noise resident in HBM (the fixed-noise call of configs[1]);
  e2e       fields/s through the host-buffer C-ABI call the reference's `gen.predict([tensor, noise_generator(...)])`
            maps to (api.py:136-137): pinned host image -> device, noise drawn on the device inside the input-packing
            kernel, forward, result -> pinned host; `e2e_host_noise` is the same with the noise tensor ALSO coming from
            the host (283 MB more H2D per step);
  roofline  the dominant kernel (largest share of the step) against the measured bf16 peak of MEASURED_PEAKS.json --
            the BURST figure when the SM clock sampled during the timed region is >= 0.9 x max (a short timed region
            runs at boost clocks), the SUSTAINED one otherwise; FLOPs are the EXECUTED ones of the reference's dense
            formulation (the t = 0 ConvLSTM step has no recurrent half);
  tf32      value / e2e of the same workload with kind::tf32 operands (rel-L2 <= 1e-3 of the fp32 reference);
  cpu_baseline  the CPU restatement of the reference graph (oracle/torch_port.py, "port": TensorFlow 2.4.3 cannot be
            installed here) on a bounded sample, on this box's host cores.
--precision tf32 makes the tf32 arithmetic the primary line (and reports bf16 beside it).

--workload train: a step is one `GAN.train_step` (ganbase.py:21-94: 3 critic updates, 1 generator update, metric
recompute) on batch 8 x 24 timesteps x 96x96 PER GPU, data parallel with an NCCL gradient all-reduce and synchronised
BatchNorm inside the timed step.  metric = wgan_train_samples_per_sec; e2e includes the host -> device upload of each
step's batch from pinned memory and the device -> host read of the step's metrics.

--impl reference: the CPU port of the same workload on the box's host cores (rank 0 only), same metric / config keys.
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

# ---- inference workload (BASELINE.json configs[1])
B, T, S, CIN, CNOISE, COUT = 64, 8, 96, 3, 20, 2
FIELDS_PER_STEP = B * T
# per-(b,t) MMAC of each stage on the reference's dense formulation.  ConvLSTM: the input
# convolution runs at every step, the recurrent one at t >= 1 only (h_0 = 0): executed = 339.74 * (2 - 1/T).
LSTM_HALF_MMAC = 339.738624
STAGE_MMAC = {"conv8x8s2": 434.110464, "conv4x4s2": 150.994944, "convlstm": LSTM_HALF_MMAC * (2 - 1.0 / T),
              "conv3x3": 42.467328, "convT2x2s2": 14.155776, "upconvT5x5": 589.824, "conv3x3_out": 2.654208}
FLOP_PER_FIELD_DENSE = 3_827_367_936                       # BASELINE.md §2 (every step with both ConvLSTM halves)
FLOP_PER_FIELD = 2e6 * sum(STAGE_MMAC.values())            # executed at this T
CPU_SAMPLE_B = 8                                           # bounded CPU sample: 8 sequences x 8 timesteps
WORKLOAD = (f"generator inference, {B} sequences x {T} timesteps x {S}x{S} per GPU, {CIN}+{CNOISE} input channels "
            "(BASELINE configs[1])")
# ---- training workload (BASELINE.json configs[3])
TB, TT = 8, 24
TRAIN_FLOP_PER_SAMPLE = 1.0e12
TRAIN_WORKLOAD = f"WGAN train_step, batch {TB} per GPU x {TT} timesteps x {S}x{S} (BASELINE configs[3])"


TOLERANCE = {"bf16": "rel-L2 <= 1e-2 vs float64 oracle (bf16 operands, fp32 accumulate)",
             "tf32": "rel-L2 <= 1e-3 vs float64 oracle (tf32 operands, fp32 accumulate, fp32 output conv)"}


def infer_config(world, precision):
    """The `config` object of the inference line -- identical in the b200 arm and the reference arm."""
    return {"workload": WORKLOAD, "fields_per_step_per_gpu": FIELDS_PER_STEP,
            "weights": "synthetic, seed 0, non-trivial BN stats",
            "parallelism": f"independent sequences x{world}, no collective",
            "l2": "inputs 434 MB + >1.6 GB activations per step exceed the 126 MB L2",
            "tolerance": TOLERANCE[precision]}


def train_config(world):
    return {"workload": TRAIN_WORKLOAD,
            "parallelism": f"data parallel x{world}: NCCL gradient all-reduce (4 per step) + synchronised BatchNorm",
            "l2": "activations of one step (> 10 GB) exceed the 126 MB L2"}


def stage_bytes(esz):
    """Algorithmic HBM bytes per field of the bandwidth-bound stages (esz = bytes per activation element)."""
    return {"pack_input": 96 * 96 * 23 * 4 + 102 * 102 * 24 * esz,
            "border_fix": 2 * (4 * 104 * 160 * esz) + 2 * (96 * 192 * 4),
            "conv3x3_out": 96 * 96 * 16 * esz + 96 * 96 * 2 * 4}


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
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)
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


def tensor_peak(peaks, clocks, precision):
    """Measured bf16 peak that applies to this run (burst at boost clocks, sustained otherwise); tf32 = half of it."""
    boost = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.9 * clocks["sm_max_mhz"])
    peak = peaks["bf16_tflops"] if boost else peaks["bf16_tflops_sustained"]
    kind = f"{peaks['source']} bf16 {'burst' if boost else 'sustained'} peak (SM clock sampled {clocks.get('sm_mhz') if clocks else None} MHz)"
    if precision == "tf32":
        return peak / 2, kind + " / 2 for tf32 (assumed: no measured tf32 figure)"
    return peak, kind


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ============================================================================================ CPU legs (the oracle port)
_AFFINITY0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None


def _all_host_cores():
    """The CPU legs use every core the process started with (the GPU legs may have bound it to one NUMA node)."""
    if _AFFINITY0 is not None:
        try:
            os.sched_setaffinity(0, _AFFINITY0)
        except OSError:
            pass
    return len(_AFFINITY0) if _AFFINITY0 else (os.cpu_count() or 1)


def cpu_port_fields_per_s(steps, warmup):
    """Times the CPU restatement of the reference generator graph on all host threads."""
    import torch
    from oracle.generator import synthetic_generator_weights
    from oracle.torch_port import TorchGenerator
    cores = _all_host_cores()
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


def cpu_port_train_samples_per_s(steps, warmup, Bc=1, Tc=4):
    """One WGAN step of the torch-CPU fp32 autograd restatement (oracle/torch_train.py) on a bounded sample: Bc
    sequences x Tc timesteps; converted to samples/s of the full T = 24 step by the timestep ratio (the step's cost is
    linear in T)."""
    import torch
    from oracle import torch_train as tt
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    tt.DT = torch.float32
    cores = _all_host_cores()
    torch.set_num_threads(cores)
    rng = np.random.default_rng(0)
    lr = rng.standard_normal((Bc, Tc, S, S, 3)).astype(np.float32)
    hr = rng.standard_normal((Bc, Tc, S, S, 2)).astype(np.float32)
    draws = []
    for _ in range(3):
        draws += [0.1 * rng.standard_normal((Bc, Tc, S, S, 20)), rng.uniform(0, 1, (Bc,)),
                  0.1 * rng.standard_normal((Bc, Tc, S, S, 2)), 0.1 * rng.standard_normal((Bc, Tc, S, S, 2))]
    draws += [0.1 * rng.standard_normal((Bc, Tc, S, S, 20))] * 2
    st = tt.State(synthetic_generator_weights(0), synthetic_critic_weights(1, size=S))
    for _ in range(warmup):
        tt.train_step(st, lr, hr, draws)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tt.train_step(st, lr, hr, draws)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return Bc * (Tc / TT) / dt, cores, times, f"{Bc} sequence x {Tc} of {TT} timesteps x {S}x{S} per step, scaled by {Tc}/{TT}"


def run_reference(args):
    """The reference arm: the CPU port of the same workload, same metric / config keys as the b200 arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "train":
        steps = min(args.steps, 3)
        v, cores, times, sample = cpu_port_train_samples_per_s(steps, 1)
        line = {"impl": "reference", "metric": "wgan_train_samples_per_sec", "value": v, "unit": "samples/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": 1e3 * sum(times) / len(times),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": train_config(int(os.environ.get("WORLD_SIZE", "1"))),
                "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                                 "sample": sample + f"; torch-CPU fp32 autograd restatement of ganbase.py:21-94 on {cores} threads "
                                                    f"({cpu_model_name()}); TensorFlow 2.4.3 not installable"},
                "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    v, cores, times = cpu_port_fields_per_s(args.steps, args.warmup)
    sample = (f"{CPU_SAMPLE_B} sequences x {T} timesteps x {S}x{S} per timed step (1/{B // CPU_SAMPLE_B} of the GPU batch), "
              f"fp32 torch-CPU restatement of models.py:9-73 on {cores} threads ({cpu_model_name()}); "
              "TensorFlow 2.4.3 not installable")
    line = {"impl": "reference", "metric": "generator_fields_per_sec", "value": v, "unit": "fields/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": infer_config(int(os.environ.get("WORLD_SIZE", "1")), args.precision),
            "cpu_baseline": {"value": v, "unit": "fields/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "fields/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ============================================================================================ distributed plumbing
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
        torch.cuda.set_device(self.local_rank)
        # page-locked host buffers of the e2e legs are first touched on the NUMA node of this rank's GPU
        from wind_downscaling_gan_b200.hostmem import bind_to_gpu_numa_node
        self.numa = bind_to_gpu_numa_node(self.local_rank) if os.environ.get("WDG_BIND_NUMA", "1") != "0" else None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """K calls of fn between barriers, CUDA events on the launching stream, max over ranks -> ms per step."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        self.barrier()
        return self.max_over_ranks(ev0.elapsed_time(ev1)) / steps

    def close(self):
        """Tear the process group down; never hang on it.  (Seen once at N = 2: destroy_process_group() did not return
        after a run that mixed CUDA-graph-captured and eager NCCL calls on one communicator -- the JSON line was already
        out.)  All ranks meet at a barrier, flush, and a watchdog ends the process if the teardown stalls."""
        if self.world > 1:
            import gc
            import threading
            sys.stdout.flush()
            sys.stderr.flush()
            self.barrier()
            gc.collect()
            threading.Timer(15.0, lambda: os._exit(0)).start()
            self.dist.destroy_process_group()
            os._exit(0)


# ============================================================================================ inference
def measure_inference(D, gen, precision, args, image_p, noise_p, out_p, image_d, noise_d, out_d, with_stages):
    """Device-resident and end-to-end rates of one precision.  Returns a dict (rank-independent numbers)."""
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    torch = D.torch
    gen.set_precision(precision)
    gen.set_profiling(with_stages)
    for _ in range(args.warmup):
        gen.forward_device(image_d, noise_d, out_d)
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    ms_step = D.timed(lambda: gen.forward_device(image_d, noise_d, out_d), args.steps)
    clocks = sampler.stop() if D.rank == 0 else None
    stage_ms = {}
    if with_stages:   # re-run K profiled steps, reading the events after each (reads sync, so not in the timed loop)
        for _ in range(args.steps):
            gen.forward_device(image_d, noise_d, out_d)
            for k, v in gen.stage_ms().items():
                stage_ms[k] = stage_ms.get(k, 0.0) + v / args.steps
        gen.set_profiling(False)
    res = {"precision": precision, "ms_per_step": ms_step, "value": D.world * FIELDS_PER_STEP / (ms_step * 1e-3),
           "clocks": clocks, "stage_ms": stage_ms, "lstm_launches": gen.launches_per_forward() - 9}
    # end to end, the reference's call pattern: host image in, noise drawn on the device, host result out
    ng = FlexibleNoiseGenerator((B, T, S, S, CNOISE), std=0.1, random_seed=7 + D.rank)
    for _ in range(2):
        gen.predict_host_gen_noise(image_p, ng, out_p)
    ms = D.timed(lambda: gen.predict_host_gen_noise(image_p, ng, out_p), args.steps)
    h2d, d2h = int(image_p.numel() * 4), int(out_p.numel() * 4)
    res["e2e"] = {"value": D.world * FIELDS_PER_STEP / (ms * 1e-3), "unit": "fields/s", "ms_per_step": ms,
                  "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "h2d_gbs_per_gpu": h2d / (ms * 1e-3) / 1e9, "d2h_gbs_per_gpu": d2h / (ms * 1e-3) / 1e9,
                  "api": "wdg_generator_predict_host_gen_noise: pinned host image in, noise drawn in the packing kernel "
                         "(api.py:136), pinned host result out"}
    res["checksum_gen_noise"] = float(out_p.double().abs().mean())
    # the same with the noise tensor coming from the host as well
    for _ in range(2):
        gen.predict_host(image_p, noise_p, out_p)
    ms = D.timed(lambda: gen.predict_host(image_p, noise_p, out_p), args.steps)
    h2d = int(image_p.numel() * 4 + noise_p.numel() * 4)
    res["e2e_host_noise"] = {"value": D.world * FIELDS_PER_STEP / (ms * 1e-3), "unit": "fields/s", "ms_per_step": ms,
                             "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                             "h2d_gbs_per_gpu": h2d / (ms * 1e-3) / 1e9,
                             "api": "wdg_generator_predict_host: pinned host image AND noise in, pinned host result out"}
    res["checksum"] = float(out_p.double().abs().mean())
    return res


def inference_roofline(res, peaks, world):
    esz = 4 if res["precision"] == "tf32" else 2
    peak, peak_kind = tensor_peak(peaks, res["clocks"], res["precision"])
    sb = stage_bytes(esz)
    total = sum(res["stage_ms"].values())
    stages = {}
    for name, ms in res["stage_ms"].items():
        # ConvLSTM: all T steps in one persistent launch (or one per step with WDG_NO_LSTM_PERSIST=1)
        launches = res.get("lstm_launches", T) if name == "convlstm" else (2 if name == "border_fix" else 1)
        ent = {"ms_per_step": ms, "launches_per_step": launches, "share": ms / total}
        if name in STAGE_MMAC and name != "conv3x3_out":
            tf = 2 * STAGE_MMAC[name] * 1e6 * FIELDS_PER_STEP / (ms * 1e-3) / 1e12
            ent.update(bound="tensor", achieved=tf, unit="TFLOP/s", frac=tf / peak)
        else:
            gbs = sb[name] * FIELDS_PER_STEP / (ms * 1e-3) / 1e9
            ent.update(bound="hbm", achieved=gbs, unit="GB/s", frac=gbs / peaks["hbm_gbs"])
        stages[name] = ent
    dom = max(stages, key=lambda k: stages[k]["share"])     # dominant kernel = largest share of the step
    d = stages[dom]
    traffic = None
    for fn in ("r2_traffic.json", "r1_traffic.json"):       # DRAM bytes per launch from the committed ncu --set full capture
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", fn)))[dom]["dram_bytes_per_launch"]
            break
        except Exception:
            pass
    whole = FLOP_PER_FIELD * res["value"] / world / 1e12
    return {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": peak if d["bound"] == "tensor" else peaks["hbm_gbs"],
            "unit": d["unit"], "frac": d["frac"], "traffic": traffic if res["precision"] == "bf16" else None,
            "avg_launch_ms": d["ms_per_step"] / d["launches_per_step"], "share_of_step": d["share"],
            "peak_source": peak_kind if d["bound"] == "tensor" else f"{peaks['source']} HBM copy bandwidth",
            "flops": "executed FLOPs of the reference's dense formulation (t = 0 ConvLSTM step counted without its recurrent half)",
            "whole_forward": {"achieved": whole, "unit": "TFLOP/s", "frac": whole / peak,
                              "dense_equivalent": FLOP_PER_FIELD_DENSE * res["value"] / world / 1e12},
            "stages": stages}


def run_inference(args):
    import torch
    from wind_downscaling_gan_b200.gan.models import make_generator
    from oracle.generator import synthetic_generator_weights  # weights only; the oracle is not on the timed path
    D = Dist()
    peaks = load_peaks()
    gen = make_generator(S, CIN, CNOISE, COUT, T)
    gen.set_weights(synthetic_generator_weights(0))
    image_h, noise_h = synth_inputs(100 + D.rank)
    image_p = torch.from_numpy(image_h).pin_memory()
    noise_p = torch.from_numpy(noise_h).pin_memory()
    out_p = torch.empty((B, T, S, S, COUT), dtype=torch.float32).pin_memory()
    image_d = image_p.cuda(non_blocking=True)
    noise_d = noise_p.cuda(non_blocking=True)
    out_d = torch.empty((B, T, S, S, COUT), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    other = "tf32" if args.precision == "bf16" else "bf16"
    main = measure_inference(D, gen, args.precision, args, image_p, noise_p, out_p, image_d, noise_d, out_d, True)
    side = measure_inference(D, gen, other, args, image_p, noise_p, out_p, image_d, noise_d, out_d, True)
    gen.set_precision(args.precision)
    if D.rank == 0:
        tol = {"bf16": "rel-L2 <= 1e-2 vs float64 oracle (bf16 operands, fp32 accumulate)",
               "tf32": "rel-L2 <= 1e-3 vs float64 oracle (tf32 operands, fp32 accumulate, fp32 output conv)"}
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, _ = cpu_port_fields_per_s(3, 1)
            cpu = {"value": v, "unit": "fields/s", "cores": cores, "kind": "port",
                   "sample": f"{CPU_SAMPLE_B} sequences x {T} timesteps x {S}x{S} (1/{B // CPU_SAMPLE_B} of the GPU batch), "
                             f"mean of 3 after 1 warm-up, fp32 torch-CPU restatement of the reference graph on {cores} "
                             f"threads ({cpu_model_name()}); TensorFlow 2.4.3 not installable"}
        side_roof = inference_roofline(side, peaks, D.world)
        line = {"metric": "generator_fields_per_sec", "value": main["value"], "unit": "fields/s", "n_gpus": D.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
                "config": infer_config(D.world, args.precision),
                "e2e": main["e2e"], "e2e_host_noise": main["e2e_host_noise"],
                "gpu_launches": gen.launches_per_forward() * args.steps,
                "clocks": main["clocks"], "host_numa": D.numa, "roofline": inference_roofline(main, peaks, D.world),
                other: {"dtype": other, "value": side["value"], "ms_per_step": side["ms_per_step"], "e2e": side["e2e"],
                        "e2e_host_noise": side["e2e_host_noise"], "tolerance": tol[other], "clocks": side["clocks"],
                        "roofline": {k: side_roof[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "peak_source", "whole_forward")},
                        "stage_ms": side["stage_ms"]},
                "cpu_baseline": cpu, "output_abs_mean": main["checksum"]}
        print(json.dumps(line), flush=True)
    D.close()


# ============================================================================================ training
def run_training(args):
    import torch
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    from wind_downscaling_gan_b200.train.dist import Comm
    from oracle.critic import synthetic_critic_weights        # weights only
    from oracle.generator import synthetic_generator_weights
    D = Dist()
    peaks = load_peaks()
    comm = Comm() if D.world > 1 else None
    gen, disc = make_generator(S, 3, 20, 2, TT), make_discriminator(S, S, 3, 2, TT)
    gen.set_weights(synthetic_generator_weights(0))      # identical replicas on every rank
    disc.set_weights(synthetic_critic_weights(1, size=S))
    gan = GAN(gen, disc, FlexibleNoiseGenerator((TB, TT, S, S, 20), std=0.1, random_seed=100 + D.rank))
    gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                discriminator_loss=train.discriminator_loss, train_precision=args.precision)
    rng = np.random.default_rng(200 + D.rank)
    lr_p = torch.from_numpy(rng.standard_normal((TB, TT, S, S, 3), dtype=np.float32)).pin_memory()
    hr_p = torch.from_numpy(rng.standard_normal((TB, TT, S, S, 2), dtype=np.float32)).pin_memory()
    lr_d, hr_d = lr_p.cuda(), hr_p.cuda()
    last = {}

    def step_resident():
        last["m"] = gan.train_step((lr_d, hr_d), comm=comm)

    def step_e2e():      # the call a user makes: host batch in (pinned), metrics dict out (device -> host read inside)
        last["m"] = gan.train_step((lr_p, hr_p), comm=comm)

    for _ in range(args.warmup):
        step_resident()
    if comm is not None:
        comm.reset_timers()
    sampler = ClockSampler(D.local_rank)
    if D.rank == 0:
        sampler.start()
    ms_step = D.timed(step_resident, args.steps)
    clocks = sampler.stop() if D.rank == 0 else None
    # The timed steps replay ONE CUDA graph: the collectives are nodes of it and no Python-side event brackets them.  Their
    # device time is measured on two extra EAGER steps of the same model (CUDA events around every NCCL call).
    ar_ms, ar_calls = 0.0, 0
    if comm is not None:
        graph_was = gan.use_cuda_graph
        gan.use_cuda_graph = False
        gan.train_step((lr_d, hr_d), comm=comm)
        comm.reset_timers()
        for _ in range(2):
            gan.train_step((lr_d, hr_d), comm=comm)
        ar_calls = len(comm._events) // 2
        ar_ms = comm.collective_ms() / 2
        comm.timing = False
        gan.use_cuda_graph = graph_was
    launches = gan.launches_per_step() if hasattr(gan, "launches_per_step") else None
    step_e2e()
    ms_e2e = D.timed(step_e2e, args.steps)
    # extension, reported beside the headline: the gradient-penalty passes whose result nothing uses are skipped
    # (bit-identical weights and metrics, tests/test_train_gpu.py); the headline `value` runs every pass of ganbase.py:21-94
    gen2, disc2 = make_generator(S, 3, 20, 2, TT), make_discriminator(S, S, 3, 2, TT)
    gen2.set_weights(synthetic_generator_weights(0))
    disc2.set_weights(synthetic_critic_weights(1, size=S))
    gan2 = GAN(gen2, disc2, FlexibleNoiseGenerator((TB, TT, S, S, 20), std=0.1, random_seed=100 + D.rank), skip_dead_gradient_penalty=True)
    gan2.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                 discriminator_loss=train.discriminator_loss, train_precision=args.precision)
    for _ in range(args.warmup):
        gan2.train_step((lr_d, hr_d), comm=comm)
    ms_skip = D.timed(lambda: gan2.train_step((lr_d, hr_d), comm=comm), args.steps)
    if D.rank == 0:
        samples = D.world * TB / (ms_step * 1e-3)
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12     # FMA lanes x 2 flop x max SM clock
        if args.precision == "fp32":
            peak, peak_kind, bound = fp32_peak, "fp32 FMA peak of the SMs at max clock", "fp32-cuda-core"
        else:
            peak, peak_kind = tensor_peak(peaks, clocks, args.precision)
            bound = "tensor"
        achieved = TRAIN_FLOP_PER_SAMPLE * samples / D.world / 1e12
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, _, sample = cpu_port_train_samples_per_s(1, 1)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": sample + f"; torch-CPU fp32 autograd restatement of ganbase.py:21-94 on {cores} threads ({cpu_model_name()})"}
        line = {"metric": "wgan_train_samples_per_sec", "value": samples, "unit": "samples/s", "n_gpus": D.world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
                "config": train_config(D.world),
                "e2e": {"value": D.world * TB / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(lr_p.numel() * 4 + hr_p.numel() * 4), "d2h_bytes_per_step": 64,
                        "api": "GAN.train_step((low_res, high_res)) with pinned host batches; metrics dict read back"},
                "gpu_launches": launches * args.steps if launches else None,
                "collective": {"allreduce_ms_per_step": ar_ms, "share_of_step": ar_ms / ms_step if ms_step else None,
                               "nccl_calls_per_step": ar_calls,
                               "note": "device time of the NCCL collectives (flat gradient buffers + BatchNorm statistics), CUDA events "
                                       "around every call on the launch stream, measured on 2 eager steps after the timed "
                                       "CUDA-graph region (inside the graph they are nodes no event brackets)"},
                "dead_gradient_penalty_skipped": {"value": D.world * TB / (ms_skip * 1e-3), "unit": "samples/s", "ms_per_step": ms_skip,
                                                  "note": "GAN(..., skip_dead_gradient_penalty=True): same weights and metrics bit for bit; NOT the headline"},
                "cuda_graph": bool(getattr(gan, "_graphed", None) is not None and gan._graphed.graph is not None),
                "clocks": clocks,
                "roofline": {"kernel": "whole train_step (many kernels)", "bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                             "frac": achieved / peak, "traffic": None, "peak_source": peak_kind,
                             "flops": "~1.0 TFLOP per sample (BASELINE.md §2)"},
                "cpu_baseline": cpu, "last_metrics": {k: v for k, v in last["m"].items() if v is not None}}
        print(json.dumps(line), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train"])
    ap.add_argument("--precision", default=None, help="infer: bf16 (default) | tf32; train: tf32 (default) | bf16 | fp32")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 200 if args.workload == "infer" else 10
    if args.warmup is None:
        args.warmup = 5 if args.workload == "infer" else 3
    if args.warmup < 3:
        print(f"bench.py: --warmup {args.warmup} raised to 3 (timing rules)", file=sys.stderr)
        args.warmup = 3
    if args.precision is None:
        args.precision = "bf16" if args.workload == "infer" else "tf32"
    ok = ("bf16", "tf32") if args.workload == "infer" else ("fp32", "tf32", "bf16")
    if args.precision not in ok:
        raise SystemExit(f"--precision for --workload {args.workload} must be one of {ok}")
    if args.impl == "reference":
        return run_reference(args)
    return run_inference(args) if args.workload == "infer" else run_training(args)


if __name__ == "__main__":
    main()
