#!/usr/bin/env python
"""Benchmark of the WGAN training step (BASELINE.json configs[3]): batch 8 per GPU x 24 timesteps x 96x96, one
`GAN.train_step` (3 critic updates + 1 generator update + metric recompute, ganbase.py:21-94), data-parallel over
N GPUs with an NCCL gradient all-reduce and synchronised BatchNorm (weak scaling).

    python bench_train.py [--steps K] [--warmup W]          (N > 1: launch with torch.distributed.run)

Prints one JSON line (rank 0).  --precision selects the arithmetic of the convolution GEMMs: fp32 (CUDA cores),
tf32 (default) or bf16 (tcgen05 tensor cores, fp32 accumulation); the roofline entry is reported against the
matching peak.
The CPU baseline is the torch-CPU float32 restatement of the same step (oracle/torch_train.py) on a bounded sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
B, T, S = 8, 24, 96
FLOP_PER_SAMPLE = 1.0e12          # ~1.0 TFLOP per sample at T = 24


def cpu_baseline(Bc=1, Tc=4):
    import torch
    from oracle import torch_train as tt
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    tt.DT = torch.float32
    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(0)
    lr = rng.standard_normal((Bc, Tc, S, S, 3)).astype(np.float32)
    hr = rng.standard_normal((Bc, Tc, S, S, 2)).astype(np.float32)
    draws = []
    for _ in range(3):
        draws += [0.1 * rng.standard_normal((Bc, Tc, S, S, 20)), rng.uniform(0, 1, (Bc,)),
                  0.1 * rng.standard_normal((Bc, Tc, S, S, 2)), 0.1 * rng.standard_normal((Bc, Tc, S, S, 2))]
    draws += [0.1 * rng.standard_normal((Bc, Tc, S, S, 20))] * 2
    st = tt.State(synthetic_generator_weights(0), synthetic_critic_weights(1, size=S))
    tt.train_step(st, lr, hr, draws)
    t0 = time.perf_counter()
    tt.train_step(st, lr, hr, draws)
    dt = time.perf_counter() - t0
    return Bc * Tc / dt, os.cpu_count() or 1, f"{Bc} sequence x {Tc} timesteps x {S}x{S}, one step after one warm-up"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32", "bf16"],
                    help="arithmetic of the convolution GEMMs: fp32 = CUDA cores, tf32 / bf16 = tcgen05 (fp32 accumulate)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    from wind_downscaling_gan_b200.train.dist import Comm
    from wind_downscaling_gan_b200.train import ops as train_ops
    train_ops.set_precision(args.precision)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    comm = Comm() if world > 1 else None
    gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    gen.set_weights(synthetic_generator_weights(0))      # identical replicas on every rank
    disc.set_weights(synthetic_critic_weights(1, size=S))
    gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=100 + rank))
    gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                discriminator_loss=train.discriminator_loss)
    g = torch.Generator(device="cuda").manual_seed(rank)
    lr = torch.randn((B, T, S, S, 3), device="cuda", generator=g)
    hr = torch.randn((B, T, S, S, 2), device="cuda", generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        m = gan.train_step((lr, hr), comm=comm)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m = gan.train_step((lr, hr), comm=comm)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        ms_step = ms / args.steps
        samples = world * B / (ms_step * 1e-3)
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12     # FMA lanes x 2 flop x max SM clock
        try:
            bf16_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
        except Exception:
            bf16_peak = 1377.3
        peak = {"fp32": fp32_peak, "tf32": bf16_peak / 2, "bf16": bf16_peak}[args.precision]
        bound = {"fp32": "fp32-cuda-core", "tf32": "tensor", "bf16": "tensor"}[args.precision]
        note = {"fp32": "implicit-GEMM convs on CUDA cores (fp32 FMA peak of the SMs)",
                "tf32": "tcgen05 kind::tf32 gather-fed implicit GEMMs; peak = half the measured sustained bf16 rate",
                "bf16": "tcgen05 kind::f16 (bf16) gather-fed implicit GEMMs; peak = measured sustained bf16 rate"}[args.precision]
        cpu = None
        if not args.no_cpu_baseline:
            v, cores, sample = cpu_baseline()
            cpu = {"value": v, "unit": "sequence-timesteps/s", "cores": cores, "kind": "port", "sample": sample + "; torch-CPU fp32 autograd restatement of ganbase.py:21-94"}
        line = {"metric": "wgan_train_samples_per_sec", "value": samples, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "tf32": "tf32", "bf16": "bf16"}[args.precision], "data": "synthetic",
                "config": {"workload": f"WGAN train_step, batch {B} per GPU x {T} timesteps x {S}x{S} (BASELINE configs[3])",
                           "parallelism": f"data parallel x{world}: NCCL flat-bucket gradient all-reduce + synchronised BatchNorm"},
                "sequence_timesteps_per_sec": samples * T,
                "roofline": {"bound": bound, "achieved": FLOP_PER_SAMPLE * samples / world / 1e12, "peak": peak,
                             "unit": "TFLOP/s", "frac": FLOP_PER_SAMPLE * samples / world / 1e12 / peak, "traffic": None,
                             "note": note},
                "cpu_baseline": cpu, "last_metrics": {k: v for k, v in m.items() if v is not None}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
