#!/usr/bin/env python
"""End-to-end tiled downscaling of a whole domain (BASELINE.json configs[2] and [4]) through the public API
(`wind_downscaling_gan_b200.predict`): coarse u10/v10 + DEM on the host -> patch gather/normalise -> generator (noise
drawn on the device) -> crop + overlap-mean stitch -> stitched hi-res fields back on the host.

    python bench_domain.py [--config switzerland|weurope] [--windows W] [--members M]
    (N > 1: torch.distributed.run; windows / (window, member) pairs are sharded over ranks, no collective)

configs[2]: Switzerland 294 x 429 px, overlap 0.05 -> 20 patches per 24-h window; one synthetic year = 365 windows
            (default here: 16 windows per rank so the run stays short; --windows 365 for the whole year).
configs[4]: W-Europe box 962 x 972 px -> 169 patches per window, M noise members for one window.
Prints one JSON line: fields/s (96x96x2 generator outputs) and domain-timesteps/s (stitched hourly maps).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def synthetic_domain(H, W, hours, seed):
    """Fields already regridded to the hi-res template (process_era5 / process_topo output): blocky nearest-neighbour
    copies of a coarse N(0, 5^2) wind and a U(0, 3000) m DEM."""
    from wind_downscaling_gan_b200.grid import GridDataset
    rng = np.random.default_rng(seed)
    ch, cw = -(-H // 26), -(-W // 18)
    def wind():
        c = 5.0 * rng.standard_normal((hours, ch, cw)).astype(np.float32)
        return np.repeat(np.repeat(c, 26, 1), 18, 2)[:, :H, :W]
    lat, lon = np.linspace(45.0, 48.0, H), np.linspace(5.0, 11.0, W)
    time_ = np.datetime64("2016-01-01T00") + np.arange(hours) * np.timedelta64(1, "h")
    era = GridDataset({"u10": (("time", "lat_1", "lon_1"), wind()), "v10": (("time", "lat_1", "lon_1"), wind())},
                      {"time": time_, "lat_1": lat, "lon_1": lon})
    topo = GridDataset({"elevation": (("lat_1", "lon_1"), rng.uniform(0, 3000, (H, W)).astype(np.float32))},
                       {"lat_1": lat, "lon_1": lon})
    tpl = GridDataset({}, {"time": time_, "lat_1": lat, "lon_1": lon})
    return era, topo, tpl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="switzerland", choices=["switzerland", "weurope"])
    ap.add_argument("--windows", type=int, default=None, help="24-h windows in total (default 16 per rank)")
    ap.add_argument("--members", type=int, default=4, help="noise ensemble members per window (weurope)")
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from wind_downscaling_gan_b200 import api
    from wind_downscaling_gan_b200.train.dist import shard_batch
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    if args.config == "switzerland":
        H, W, ov, per_window = 294, 429, 0.05, 20
        units = args.windows if args.windows else 16 * world
        label = f"Switzerland 294x429 px, overlap 0.05, {units} windows x 24 h (BASELINE configs[2])"
    else:
        H, W, ov, per_window = 962, 972, 0.05, 169
        units = args.members * (args.windows or 1)
        label = f"W-Europe 962x972 px, overlap 0.05, {units} (window, member) pairs x 24 h (BASELINE configs[4])"
    a, b = shard_batch(units, rank, world)
    net = api.get_network()
    devnull = open(os.devnull, "w")
    if args.config == "switzerland":     # custom (COSMO) template: predict() on fields already on the hi-res grid
        era, topo, tpl = synthetic_domain(H, W, 24, 100 + rank)
        call = lambda: api.predict(era, topo, tpl, overlap_factor=ov, network=net, group_size=per_window)
        path = "api.predict: host hi-res fields -> device gather/normalise -> generator (device noise) -> device stitch -> host"
    else:                                # the CLI path: coarse ERA5 (54 x 37 points) + DEM raster through downscale()
        from tests.synth import synthetic_dem, synthetic_era5
        era5 = synthetic_era5(lon0=-4.96, lon1=8.3, lat0=42.2, lat1=51.3, hours=24, seed=100 + rank)
        dem = synthetic_dem(lon0=-5.5, lon1=9.0, lat0=41.5, lat1=52.0, seed=200 + rank, n=2000)
        assert (len(era5.coords["longitude"]), len(era5.coords["latitude"])) == (54, 37)
        call = lambda: api.downscale(era5, dem, overlap_factor=ov, network=net, group_size=per_window)
        path = ("api.downscale: host coarse ERA5 + DEM raster -> device regrid+gather/normalise -> generator (device noise) -> "
                "device stitch -> host")

    def one():
        so = sys.stdout
        sys.stdout = devnull        # predict() prints progress like the reference
        try:
            return call()
        finally:
            sys.stdout = so

    out = one()                     # warm-up: plans, workspaces
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a, b):
        out = one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        fields = units * per_window * 24
        print(json.dumps({"metric": "domain_downscaling_fields_per_sec", "value": fields / dt, "unit": "fields/s", "n_gpus": world,
                          "domain_timesteps_per_sec": units * 24 / dt, "seconds": dt, "scaling": "strong", "dtype": "bf16",
                          "data": "synthetic", "config": {"workload": label, "patches_per_window": per_window,
                                                          "output_shape": list(out["u10"].shape)},
                          "path": path, "timing": "wall clock around the public API calls, max over ranks"}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
