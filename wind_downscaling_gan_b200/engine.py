"""Back-to-back tiled downscaling of MANY 24-hour windows and/or noise members over one domain -- the loop a user of the
reference writes around `downscale()` / `predict()` (one call per date, as the CLI does: `cli.py:19-24`), and the
100-member ensemble of the notebook's Western-Europe box, without paying the per-call host costs again and again.

What `predict()` (api.py:89-152) does per call is kept exactly -- patch grid, reversed-latitude gather, (column, channel)
normalisation over the call's patches, groups through the generator with fresh N(0, 0.1^2) noise, 2-px crop,
overlap mean, labels sorted ascending -- but

  * the coarse ERA5 fields, the DEM and the nearest-neighbour index maps (api.py:31-43) are uploaded ONCE and stay in
    HBM; a window's patches are gathered straight from them (wdg_gather_normalise_regrid), so no input crosses PCIe
    per window;
  * an ensemble gathers / normalises a window once and sweeps the noise members over the resident tensor (the noise is
    drawn inside the generator's packing kernel);
  * several windows share one generator launch set (the statistics stay per window), the stitched maps leave through a
    ring of device buffers on a copy stream into page-locked host memory while the next windows compute, and the
    ascending-label ordering is folded into the stitch kernel's row index instead of a host-side re-sort;
  * units (windows, or members) are split contiguously over data-parallel ranks -- one process per GPU, no collective:
    windows and members are independent.
"""
import ctypes as C

import numpy as np

from . import _lib, tiling
from .grid import GridDataset, from_xarray, nearest_index

SEQ, IMG, CROP, NOISE_CHANNELS, NB_IN, NB_OUT = 24, 96, 2, 20, 3, 2


def shard_units(n, rank, world):
    """Contiguous split of n units: (start, stop) of this rank."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class DomainDownscaler:
    """Device-resident inputs + plan for one domain.  See the module docstring.

    u10c / v10c: (T_total, uh, uw) fields on their own (coarse or hi-res) grid; uv_rows / uv_cols: for every template
    row / column the source row / column (identity for fields already on the template); dem (eh, ew) with dem_rows /
    dem_cols likewise; lat / lon: template coordinate labels."""

    def __init__(self, network, lat, lon, overlap_factor, u10c, v10c, uv_rows, uv_cols, dem, dem_rows, dem_cols,
                 dem_divisor=1e3, windows_per_forward=None, times=None):
        import torch
        self.torch = torch
        self.net = network
        self.gen = network.generator
        self.lat, self.lon = np.asarray(lat), np.asarray(lon)
        self.H, self.W = len(self.lat), len(self.lon)
        self.times = times
        self.sx, self.sy = tiling.patch_grid(self.H, self.W, overlap_factor, IMG)                 # api.py:101-116
        self.nx, self.ny = len(self.sx), len(self.sy)
        self.n_patches = self.nx * self.ny
        u10c, v10c = np.ascontiguousarray(u10c, np.float32), np.ascontiguousarray(v10c, np.float32)
        dem = np.ascontiguousarray(dem, np.float32)
        self.T_total, self.uh, self.uw = u10c.shape
        self.eh, self.ew = dem.shape
        self.n_windows = self.T_total // SEQ
        self.dem_divisor = float(dem_divisor)
        dev = "cuda"
        self.d_u, self.d_v, self.d_dem = (torch.from_numpy(a).to(dev) for a in (u10c, v10c, dem))
        i32 = lambda a: torch.as_tensor(np.asarray(a, np.int32), device=dev)
        self.d_maps = [i32(m) for m in (uv_rows, uv_cols, dem_rows, dem_cols)]
        self.d_sx, self.d_sy = i32(self.sx), i32(self.sy)
        rows = tiling.covered_rows(self.sy, IMG, CROP)
        cols = tiling.covered_cols(self.sx, IMG, CROP)
        # groupby(...).mean() returns the labels sorted ascending (api.py:150): fold that order into the stitch's index arrays
        self.rows = rows[np.argsort(self.lat[rows], kind="stable")]
        self.cols = cols[np.argsort(self.lon[cols], kind="stable")]
        self.d_rows, self.d_cols = i32(self.rows), i32(self.cols)
        self.out_lat, self.out_lon = self.lat[self.rows], self.lon[self.cols]
        if windows_per_forward is None:      # fill the 148 SMs: ~60+ sequences per generator launch set
            windows_per_forward = max(1, min(8, -(-60 // self.n_patches)))
        self.wpf = max(1, min(windows_per_forward, max(1, self.T_total // SEQ)))     # no padded slots for a single window
        L = _lib.lib()
        nb = C.c_size_t()
        _lib.check(L.wdg_patch_scratch_bytes(self.nx, self.ny, 1, IMG, C.byref(nb)))
        self.scratch = torch.empty(nb.value, dtype=torch.uint8, device=dev)
        self.mean = torch.empty((IMG, 3), dtype=torch.float64, device=dev)
        self.std = torch.empty((IMG, 3), dtype=torch.float64, device=dev)
        B = self.wpf * self.n_patches
        self.tensors = torch.zeros((B, SEQ, IMG, IMG, NB_IN), dtype=torch.float32, device=dev)
        self.preds = torch.empty((B, SEQ, IMG, IMG, NB_OUT), dtype=torch.float32, device=dev)
        self.out_shape = (NB_OUT, SEQ, len(self.rows), len(self.cols))
        self.ring = [torch.empty(self.out_shape, dtype=torch.float32, device=dev) for _ in range(4)]
        self.ring_free = [torch.cuda.Event() for _ in self.ring]
        self.copy_stream = torch.cuda.Stream()
        self._ring_i = 0

    # ------------------------------------------------------------------------------------------------------------
    def _gather(self, window, slot):
        """Patches of one 24-h window, normalised with that window's own statistics, into slot `slot` of the batch."""
        t = self.torch
        L = _lib.lib()
        off = window * SEQ * self.uh * self.uw * 4
        out = self.tensors[slot * self.n_patches:(slot + 1) * self.n_patches]
        stream = C.c_void_p(t.cuda.current_stream().cuda_stream)
        _lib.check(L.wdg_gather_normalise_regrid(
            C.c_void_p(self.d_u.data_ptr() + off), C.c_void_p(self.d_v.data_ptr() + off), SEQ, self.uh, self.uw,
            self.d_maps[0].data_ptr(), self.d_maps[1].data_ptr(), self.d_dem.data_ptr(), self.eh, self.ew,
            self.d_maps[2].data_ptr(), self.d_maps[3].data_ptr(), C.c_float(self.dem_divisor), self.H, self.W,
            self.d_sx.data_ptr(), self.nx, self.d_sy.data_ptr(), self.ny, SEQ, IMG, self.mean.data_ptr(), self.std.data_ptr(),
            out.data_ptr(), self.scratch.data_ptr(), stream))

    def _stitch_to_host(self, slot, host_view):
        """Crop + overlap mean of slot `slot`'s predictions (api.py:148-150) -> ring buffer -> page-locked host view."""
        t = self.torch
        L = _lib.lib()
        i = self._ring_i
        self._ring_i = (i + 1) % len(self.ring)
        main = t.cuda.current_stream()
        main.wait_event(self.ring_free[i])          # the copy that last used this buffer has finished
        buf = self.ring[i]
        pred = self.preds[slot * self.n_patches:(slot + 1) * self.n_patches]
        _lib.check(L.wdg_stitch(pred.data_ptr(), self.d_sx.data_ptr(), self.nx, self.d_sy.data_ptr(), self.ny, 1, SEQ, IMG, CROP,
                                NB_OUT, self.d_rows.data_ptr(), len(self.rows), self.d_cols.data_ptr(), len(self.cols),
                                buf.data_ptr(), C.c_void_p(main.cuda_stream)))
        done = t.cuda.Event()
        done.record(main)
        with t.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(done)
            host_view.copy_(buf, non_blocking=True)
            self.ring_free[i].record(self.copy_stream)

    def _forward(self):
        ng = self.net.noise_generator
        if hasattr(ng, "reserve"):
            self.gen.forward_device_gen_noise(self.tensors, ng, out=self.preds)      # api.py:136-137, noise drawn in-kernel
        else:
            self.gen.forward_device(self.tensors, ng(bs=self.tensors.shape[0], channels=NOISE_CHANNELS), out=self.preds)

    def allocate_output(self, n_windows, members=1):
        """Page-locked result buffer (windows, members, 2, 24, rows, cols)."""
        return self.torch.empty((n_windows, members) + self.out_shape, dtype=self.torch.float32).pin_memory()

    def run(self, windows=None, members=1, out=None):
        """Downscales the given windows (default: all), `members` noise draws each.  Returns the page-locked tensor
        (len(windows), members, 2 [u10, v10], 24, rows, cols); rows / cols carry the labels `out_lat` / `out_lon`."""
        t = self.torch
        windows = list(range(self.n_windows)) if windows is None else list(windows)
        if out is None:
            out = self.allocate_output(len(windows), members)
        if members == 1:
            for g0 in range(0, len(windows), self.wpf):
                grp = windows[g0:g0 + self.wpf]
                for slot, w in enumerate(grp):
                    self._gather(w, slot)
                self._forward()
                for slot in range(len(grp)):
                    self._stitch_to_host(slot, out[g0 + slot, 0])
        else:                                   # ensemble: gather once per window group, sweep the members
            for g0 in range(0, len(windows), self.wpf):
                grp = windows[g0:g0 + self.wpf]
                for slot, w in enumerate(grp):
                    self._gather(w, slot)
                for m in range(members):
                    self._forward()
                    for slot in range(len(grp)):
                        self._stitch_to_host(slot, out[g0 + slot, m])
        self.copy_stream.synchronize()
        t.cuda.current_stream().synchronize()
        return out

    def as_dataset(self, out, window_pos=0, member=0, window=None):
        """One (window, member) of `run`'s result as the GridDataset `predict()` returns (views, no copy)."""
        a = out[window_pos, member].numpy()
        w = window if window is not None else window_pos
        times = None if self.times is None else np.asarray(self.times)[w * SEQ:(w + 1) * SEQ]
        dims = ("time", "lat_1", "lon_1")
        coords = {"lat_1": self.out_lat, "lon_1": self.out_lon}
        if times is not None:
            coords["time"] = times
        return GridDataset({"u10": (dims, a[0]), "v10": (dims, a[1])}, coords)


def downscaler_from_era5(era5, raster_topo, network, range_lon=None, range_lat=None, overlap_factor=0.05,
                         high_res_template=None, windows_per_forward=None):
    """The engine for `downscale(era5, raster_topo, ...)` inputs (api.py:155-160): coarse ERA5 + DEM raster; the template
    is the reference's 18x / 26x linspace unless a custom one (e.g. the COSMO-1 Swiss grid) is given."""
    from . import api
    era5, raster_topo = from_xarray(era5), from_xarray(raster_topo)
    tpl = from_xarray(high_res_template) if high_res_template is not None else \
        api.build_high_res_template_from_era5(era5, range_lon=range_lon, range_lat=range_lat)
    lon_coord, lat_coord = api._coord_names(tpl)
    lat_t, lon_t = tpl.coords[lat_coord], tpl.coords[lon_coord]
    u10c = era5.transposed("u10", ("time", "latitude", "longitude"))
    v10c = era5.transposed("v10", ("time", "latitude", "longitude"))
    uv_r, uv_c = nearest_index(era5.coords["latitude"], lat_t), nearest_index(era5.coords["longitude"], lon_t)
    name = next(iter(raster_topo.data_vars))
    dims, dem = raster_topo.var_dims(name), raster_topo[name]
    if "band" in dims:
        dem = np.take(dem, 0, axis=dims.index("band"))
        dims = tuple(d for d in dims if d != "band")
    if dims != ("y", "x"):
        dem = np.transpose(dem, [dims.index("y"), dims.index("x")])
    e_r, e_c = nearest_index(raster_topo.coords["y"], lat_t), nearest_index(raster_topo.coords["x"], lon_t)
    return DomainDownscaler(network, lat_t, lon_t, overlap_factor, u10c, v10c, uv_r, uv_c, dem, e_r, e_c,
                            windows_per_forward=windows_per_forward, times=era5.coords.get("time"))


def downscale_series(era5, raster_topo, range_lon=None, range_lat=None, overlap_factor=0.05, network=None, members=1,
                     rank=0, world=1, high_res_template=None, windows_per_forward=None, bind_numa=None):
    """`downscale()` once per 24-hour window of `era5` (and `members` noise draws per window), back to back on this GPU.
    With world > 1 (one process per GPU) the windows -- or, for a single window, the members -- are split contiguously
    over the ranks and this call computes rank's share.  Returns (engine, units, result) where `units` lists the
    (window, member range) this rank computed and `result` is engine.run()'s page-locked tensor.  bind_numa (default: on
    when world > 1): restrict the process to the CPUs of its GPU's NUMA node first, so the page-locked result buffers of
    the ranks are spread over the sockets instead of piling onto one (hostmem.py)."""
    from . import api
    if bind_numa if bind_numa is not None else world > 1:
        import torch
        from .hostmem import bind_to_gpu_numa_node
        bind_to_gpu_numa_node(torch.cuda.current_device())
    network = network if network is not None else api.get_network()
    eng = downscaler_from_era5(era5, raster_topo, network, range_lon, range_lat, overlap_factor, high_res_template,
                               windows_per_forward)
    if eng.n_windows > 1 or members == 1:
        a, b = shard_units(eng.n_windows, rank, world)
        wins, mem = list(range(a, b)), members
    else:
        a, b = shard_units(members, rank, world)
        wins, mem = [0], b - a
    out = eng.run(wins, mem)
    return eng, {"windows": wins, "members": mem}, out
