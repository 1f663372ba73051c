#!/usr/bin/env python
"""End-to-end tiled downscaling of a whole domain through the public multi-window API
(`wind_downscaling_gan_b200.engine.downscale_series`): BASELINE.json configs[2] (Switzerland, one synthetic year, sharded
by 24-h window) and configs[4] (100-member noise ensemble over the Western-Europe box, sharded by member).

    python bench_domain.py [--config switzerland|weurope] [--windows W] [--members M] [--precision bf16|tf32]
    (N > 1: python -m torch.distributed.run --nproc-per-node N ... ; one process per GPU, no collective on the data path)

What is timed (wall clock, max over ranks, after one warm-up group): coarse inputs already uploaded (once, reported as
`setup_s`) -> per window: device regrid + patch gather + (column, channel) normalisation -> generator with the noise
drawn in the packing kernel -> crop + overlap-mean stitch -> the stitched hourly maps in page-locked HOST memory.
`kernel_rate` is the generator alone on the same resident batch; `api_over_kernel` = value / (world * kernel_rate).
Prints one JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def swiss_inputs(days, seed):
    """Coarse ERA5 over the Swiss box (0.25 deg: 25 x 13 points), a synthetic DEM raster and a 294 x 429 template
    (the size of the reference's switzerland_cosmo_map.nc grid)."""
    from tests.synth import synthetic_dem, synthetic_era5
    from wind_downscaling_gan_b200.grid import GridDataset
    era = synthetic_era5(lon0=5.0, lon1=11.0, lat0=45.0, lat1=48.0, hours=24 * days, seed=seed)
    dem = synthetic_dem(lon0=4.5, lon1=11.5, lat0=44.5, lat1=48.5, seed=seed + 1, n=1200)
    tpl = GridDataset({}, {"lat_1": np.linspace(45.0, 48.0, 294), "lon_1": np.linspace(5.0, 11.0, 429)})
    return era, dem, tpl


def weurope_inputs(seed):
    from tests.synth import synthetic_dem, synthetic_era5
    era = synthetic_era5(lon0=-4.96, lon1=8.3, lat0=42.2, lat1=51.3, hours=24, seed=seed)
    dem = synthetic_dem(lon0=-5.5, lon1=9.0, lat0=41.5, lat1=52.0, seed=seed + 1, n=2000)
    assert (len(era.coords["longitude"]), len(era.coords["latitude"])) == (54, 37)
    return era, dem, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="switzerland", choices=["switzerland", "weurope"])
    ap.add_argument("--windows", type=int, default=365, help="24-h windows in total (switzerland)")
    ap.add_argument("--members", type=int, default=100, help="noise ensemble members (weurope)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from wind_downscaling_gan_b200 import api, engine
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    from wind_downscaling_gan_b200.hostmem import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(torch.cuda.current_device())      # page-locked result buffers on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    so, devnull = sys.stdout, open(os.devnull, "w")
    sys.stdout = devnull                       # get_network() prints like the reference
    net = api.get_network()
    sys.stdout = so
    net.generator.set_precision(args.precision)
    t0 = time.perf_counter()
    if args.config == "switzerland":
        era, dem, tpl = swiss_inputs(args.windows, 100)
        eng = engine.downscaler_from_era5(era, dem, net, overlap_factor=0.05, high_res_template=tpl)
        a, b = engine.shard_units(eng.n_windows, rank, world)
        wins, members, units = list(range(a, b)), 1, args.windows
        label = f"Switzerland 294x429 px, overlap 0.05, {args.windows} windows x 24 h sharded by window (BASELINE configs[2])"
    else:
        era, dem, tpl = weurope_inputs(300)
        eng = engine.downscaler_from_era5(era, dem, net, overlap_factor=0.05)
        a, b = engine.shard_units(args.members, rank, world)
        wins, members, units = [0], b - a, args.members
        label = f"W-Europe 962x972 px, overlap 0.05, one 24-h window x {args.members} noise members sharded by member (BASELINE configs[4])"
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t0
    n_units_here = len(wins) * members
    out = eng.allocate_output(len(wins), members)
    eng.run(wins[:1], 1 if args.config == "switzerland" else min(2, members), out=out)      # warm-up: plan, workspace
    # the generator alone on the resident batch (kernel rate at this batch shape)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        eng._forward()
    ev1.record()
    torch.cuda.synchronize()
    fields_per_forward = eng.tensors.shape[0] * 24
    kernel_rate = fields_per_forward * reps / (ev0.elapsed_time(ev1) * 1e-3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run(wins, members, out=out)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, kernel_rate], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dt, kernel_rate = float(tmax[0].item()), float(t[1].item()) / world
    if rank == 0:
        fields = units * eng.n_patches * 24
        value = fields / dt
        print(json.dumps({
            "metric": "domain_downscaling_fields_per_sec", "value": value, "unit": "fields/s", "n_gpus": world,
            "domain_timesteps_per_sec": units * 24 / dt, "seconds": dt, "setup_s": setup_s, "scaling": "strong",
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": label, "patches_per_window": eng.n_patches, "windows_per_forward": eng.wpf,
                       "sequences_per_forward": int(eng.tensors.shape[0]), "units_total": units, "units_rank0": n_units_here,
                       "output_per_unit": list(eng.out_shape)},
            "kernel_rate": {"value": kernel_rate, "unit": "fields/s per GPU",
                            "note": "generator forward alone (noise drawn in-kernel) on the resident batch of this run"},
            "api_over_kernel": value / (world * kernel_rate),
            "d2h_bytes_total": int(np.prod(eng.out_shape)) * 4 * units,
            "host_numa": numa,
            "path": "engine.downscale_series: resident coarse inputs -> device regrid+gather/normalise per window -> generator "
                    "(in-kernel noise) -> device stitch -> page-locked host maps (copy stream, overlapped)",
            "timing": "wall clock around engine.run() on every rank, max over ranks"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
