"""The reference's public inference API (`/root/reference/src/downscaling/api.py`) on the B200 path.

Same names, arguments and error behaviour: `downscale`, `predict`, `process_era5`, `process_topo`,
`build_high_res_template_from_era5`, `get_network` and the module constants (api.py:21-28).
Gridded inputs are `GridDataset`s (numpy stand-in for xarray, see grid.py); xarray objects are
converted when xarray is importable.  Patch gather / normalise / stitch and the generator run on the GPU.
"""
import ctypes as C
import math  # noqa: F401
from pathlib import Path

import numpy as np

from . import _lib, tiling
from .data.data_generator import FlexibleNoiseGenerator
from .gan import metrics, train
from .gan.ganbase import GAN
from .gan.models import make_discriminator, make_generator
from .grid import GridDataset, from_xarray, nearest_index

__all__ = ["WEIGHTS_PATH", "SEQUENCE_LENGTH", "IMG_SIZE", "BATCH_SIZE", "NOISE_CHANNELS", "NOISE_STD", "NB_INPUTS",
           "NB_OUTPUTS", "process_topo", "process_era5", "build_high_res_template_from_era5", "get_network",
           "predict", "downscale", "GridDataset", "make_generator", "make_discriminator", "GAN",
           "FlexibleNoiseGenerator"]

WEIGHTS_PATH = (Path(__file__) / '../weights-55.ckpt').resolve()   # api.py:21
SEQUENCE_LENGTH = 24
IMG_SIZE = 96
BATCH_SIZE = 8
NOISE_CHANNELS = 20
NOISE_STD = 0.1
NB_INPUTS = 3
NB_OUTPUTS = 2
CROP = 2                 # api.py:148


def _coord_names(template: GridDataset):
    lon = [c for c in template.coords if c.startswith('lon')][0]
    lat = [c for c in template.coords if c.startswith('lat')][0]
    return lon, lat


def process_topo(raster_topo, high_res_template):
    """api.py:31-37: band 0 of the DEM raster, nearest-neighbour sampled on the template grid."""
    raster_topo, high_res_template = from_xarray(raster_topo), from_xarray(high_res_template)
    lon_coord, lat_coord = _coord_names(high_res_template)
    name = next(iter(raster_topo.data_vars))
    dims = raster_topo.var_dims(name)
    dem = raster_topo[name]
    if 'band' in dims:
        dem = np.take(dem, 0, axis=dims.index('band'))
        dims = tuple(d for d in dims if d != 'band')
    if dims != ('y', 'x'):
        dem = np.transpose(dem, [dims.index('y'), dims.index('x')])
    iy = nearest_index(raster_topo.coords['y'], high_res_template.coords[lat_coord])
    ix = nearest_index(raster_topo.coords['x'], high_res_template.coords[lon_coord])
    return GridDataset({'elevation': ((lat_coord, lon_coord), dem[np.ix_(iy, ix)])},
                       {lat_coord: high_res_template.coords[lat_coord], lon_coord: high_res_template.coords[lon_coord]})


def process_era5(ds_era5, high_res_template):
    """api.py:40-43: u10 / v10 nearest-neighbour sampled on the template grid."""
    ds_era5, high_res_template = from_xarray(ds_era5), from_xarray(high_res_template)
    lon_coord, lat_coord = _coord_names(high_res_template)
    iy = nearest_index(ds_era5.coords['latitude'], high_res_template.coords[lat_coord])
    ix = nearest_index(ds_era5.coords['longitude'], high_res_template.coords[lon_coord])
    out = {}
    for v in ('u10', 'v10'):
        a = ds_era5.transposed(v, ('time', 'latitude', 'longitude'))
        out[v] = (('time', lat_coord, lon_coord), a[:, iy][:, :, ix])
    return GridDataset(out, {'time': ds_era5.coords['time'], lat_coord: high_res_template.coords[lat_coord],
                             lon_coord: high_res_template.coords[lon_coord]})


def build_high_res_template_from_era5(ds_era5, range_lon=None, range_lat=None):
    """api.py:46-62: 18x (lon) / 26x (lat) linspace over the selected ERA5 box."""
    ds_era5 = from_xarray(ds_era5)
    upsampling_lat = 26
    upsampling_lon = 18
    lon, lat = ds_era5.coords['longitude'], ds_era5.coords['latitude']
    if not range_lon:
        range_lon = (float(lon.min()), float(lon.max()))
        nb_lon = len(lon)
    else:
        nb_lon = int(np.count_nonzero((lon >= range_lon[0]) & (lon <= range_lon[1])))   # .sel(slice(lo, hi))
    if not range_lat:
        range_lat = (float(lat.min()), float(lat.max()))
        nb_lat = len(lat)
    else:
        nb_lat = int(np.count_nonzero((lat >= range_lat[0]) & (lat <= range_lat[1])))   # .sel(slice(hi, lo)), descending
    new_longitudes = np.linspace(range_lon[0], range_lon[1], upsampling_lon * nb_lon)
    new_latitudes = np.linspace(range_lat[0], range_lat[1], upsampling_lat * nb_lat)
    return GridDataset({}, {'time': ds_era5.coords['time'], 'lon_1': new_longitudes, 'lat_1': new_latitudes})


def get_network(weights_path=None):
    """api.py:65-86.  Builds the GAN with the api.py constants and loads `weights-55.ckpt` when its tensor
    data is available (the reference repository ships only the .index files); otherwise the
    network keeps its Keras-default random initialisation."""
    print('Loading network...')
    generator = make_generator(image_size=IMG_SIZE, in_channels=NB_INPUTS, noise_channels=NOISE_CHANNELS,
                               out_channels=NB_OUTPUTS, n_timesteps=SEQUENCE_LENGTH)
    discriminator = make_discriminator(low_res_size=IMG_SIZE, high_res_size=IMG_SIZE, low_res_channels=NB_INPUTS,
                                       high_res_channels=NB_OUTPUTS, n_timesteps=SEQUENCE_LENGTH)
    noise_shape = (BATCH_SIZE, SEQUENCE_LENGTH, IMG_SIZE, IMG_SIZE, NOISE_CHANNELS)
    gan = GAN(generator, discriminator, noise_generator=FlexibleNoiseGenerator(noise_shape, std=NOISE_STD))
    gan.compile(generator_optimizer=train.generator_optimizer(),
                generator_metrics=[metrics.AngularCosineDistance(), metrics.LogSpectralDistance(),
                                   metrics.WeightedRMSEForExtremes(), metrics.WindSpeedWeightedRMSE(), metrics.SpatialKS()],
                discriminator_optimizer=train.discriminator_optimizer(), discriminator_loss=train.discriminator_loss,
                metrics=[metrics.discriminator_score_fake(), metrics.discriminator_score_real()])
    path = Path(weights_path) if weights_path is not None else WEIGHTS_PATH
    try:
        gan.load_weights(path)
    except FileNotFoundError:
        print(f'  no tensor data under {path}: using randomly initialised weights')
    return gan


def _dev_i32(a):
    import torch
    return torch.as_tensor(np.asarray(a, np.int32), device="cuda")


_pinned = {}


def _pinned_out(shape):
    """Page-locked host buffer for the stitched result (re-used across calls of the same shape)."""
    import torch
    buf = _pinned.get(shape)
    if buf is None:
        _pinned.clear()
        buf = _pinned[shape] = torch.empty(shape, dtype=torch.float32).pin_memory()
    return buf


def _run_tiles(gather, time_window, pixels_lat, pixels_lon, lat_vals, lon_vals, times, lat_coord_hr, lon_coord_hr,
               overlap_factor, network, noise, group_size):
    """api.py:98-151 on the device.  `gather(d_sx, nx, d_sy, ny, mean, std, tensors, scratch, stream)` launches the
    patch gather + normalisation (from hi-res fields, or straight from the coarse grids with the regrid folded in)."""
    import torch
    ntimeseq = time_window // SEQUENCE_LENGTH
    starts_x, starts_y = tiling.patch_grid(pixels_lat, pixels_lon, overlap_factor, IMG_SIZE)          # api.py:101-116
    nx, ny = len(starts_x), len(starts_y)
    n_patches = nx * ny * ntimeseq
    print(f'Applying model to {n_patches} patches')
    L = _lib.lib()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d_sx, d_sy = _dev_i32(starts_x), _dev_i32(starts_y)
    nbytes = C.c_size_t()
    _lib.check(L.wdg_patch_scratch_bytes(nx, ny, max(ntimeseq, 1), IMG_SIZE, C.byref(nbytes)))
    scratch = torch.empty(nbytes.value, dtype=torch.uint8, device="cuda")
    mean = torch.empty((IMG_SIZE, 3), dtype=torch.float64, device="cuda")
    std = torch.empty((IMG_SIZE, 3), dtype=torch.float64, device="cuda")
    tensors = torch.empty((n_patches, SEQUENCE_LENGTH, IMG_SIZE, IMG_SIZE, NB_INPUTS), dtype=torch.float32, device="cuda")
    gather(d_sx, nx, d_sy, ny, mean, std, tensors, scratch, stream)
    gen = network.generator
    preds = torch.empty((n_patches, SEQUENCE_LENGTH, IMG_SIZE, IMG_SIZE, NB_OUTPUTS), dtype=torch.float32, device="cuda")
    group_size = group_size or BATCH_SIZE * 2                                                        # api.py:132
    num_groups = math.ceil(n_patches / group_size)
    for t in range(num_groups):
        sl = slice(t * group_size, min((t + 1) * group_size, n_patches))
        tensor = tensors[sl]
        ng = network.noise_generator
        if noise is None and hasattr(ng, "reserve") and ng.noise_shape[1:4] == tuple(tensor.shape[1:4]):
            # api.py:136 draws this group's noise from the generator: drawn inside the packing kernel instead
            gen.forward_device_gen_noise(tensor, ng, out=preds[sl])
        else:
            if noise is None:
                nz = ng(bs=tensor.shape[0], channels=NOISE_CHANNELS)
            else:
                nz = torch.as_tensor(noise[sl], dtype=torch.float32).cuda()
            gen.forward_device(tensor, nz, out=preds[sl])
        print(f'Predicted {(t + 1) / num_groups:.0%}')
    rows = tiling.covered_rows(starts_y, IMG_SIZE, CROP)
    cols = tiling.covered_cols(starts_x, IMG_SIZE, CROP)
    out = torch.empty((NB_OUTPUTS, ntimeseq * SEQUENCE_LENGTH, len(rows), len(cols)), dtype=torch.float32, device="cuda")
    d_rows, d_cols = _dev_i32(rows), _dev_i32(cols)   # keep alive until the kernel has run
    _lib.check(L.wdg_stitch(preds.data_ptr(), d_sx.data_ptr(), nx, d_sy.data_ptr(), ny, ntimeseq, SEQUENCE_LENGTH,
                            IMG_SIZE, CROP, NB_OUTPUTS, d_rows.data_ptr(), len(rows), d_cols.data_ptr(),
                            len(cols), out.data_ptr(), stream))
    host = _pinned_out(tuple(out.shape))
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    out = host.numpy()
    # groupby(...).mean() returns the labels sorted ascending (api.py:150)
    lat_vals, lon_vals = np.asarray(lat_vals)[rows], np.asarray(lon_vals)[cols]
    oy, ox = np.argsort(lat_vals, kind="stable"), np.argsort(lon_vals, kind="stable")
    if not (np.array_equal(oy, np.arange(len(oy))) and np.array_equal(ox, np.arange(len(ox)))):
        out = out[:, :, oy][:, :, :, ox]
        lat_vals, lon_vals = lat_vals[oy], lon_vals[ox]
    else:
        out = out.copy()   # detach from the re-used pinned staging buffer
    times = np.asarray(times)[:ntimeseq * SEQUENCE_LENGTH]
    dims = ('time', lat_coord_hr, lon_coord_hr)
    return GridDataset({'u10': (dims, out[0]), 'v10': (dims, out[1])},
                       {'time': times, lat_coord_hr: lat_vals, lon_coord_hr: lon_vals})


def predict(inputs_era5, inputs_topo, high_res_template, overlap_factor=0.05, network=None, noise=None,
            group_size=None):
    """api.py:89-152.  Extra keyword arguments (not in the reference): `network` re-uses a GAN instead of
    calling get_network(); `noise` (N,24,96,96,20) replaces the generator's own draws (parity runs);
    `group_size` overrides the reference's 16 patches per generator call."""
    import torch
    inputs_era5, inputs_topo, high_res_template = (from_xarray(x) for x in (inputs_era5, inputs_topo, high_res_template))
    lat_coord_hr = [c for c in high_res_template.coords if c.startswith('lat') or c.startswith('y')][0]
    lon_coord_hr = [c for c in high_res_template.coords if c.startswith('lon') or c.startswith('x')][0]
    network = network if network is not None else get_network()
    order = ('time', lat_coord_hr, lon_coord_hr)
    u10 = np.ascontiguousarray(inputs_era5.transposed('u10', order), np.float32)
    v10 = np.ascontiguousarray(inputs_era5.transposed('v10', order), np.float32)
    elev_km = np.ascontiguousarray(inputs_topo.transposed('elevation', order[1:]) / 1e3, np.float32)   # api.py:96
    time_window, pixels_lat, pixels_lon = u10.shape
    d_u, d_v, d_e = (torch.from_numpy(a).cuda() for a in (u10, v10, elev_km))

    def gather(d_sx, nx, d_sy, ny, mean, std, tensors, scratch, stream):
        _lib.check(_lib.lib().wdg_gather_normalise(d_u.data_ptr(), d_v.data_ptr(), d_e.data_ptr(), time_window, pixels_lat,
                                                   pixels_lon, d_sx.data_ptr(), nx, d_sy.data_ptr(), ny, SEQUENCE_LENGTH, IMG_SIZE,
                                                   mean.data_ptr(), std.data_ptr(), tensors.data_ptr(), scratch.data_ptr(), stream))

    return _run_tiles(gather, time_window, pixels_lat, pixels_lon, high_res_template.coords[lat_coord_hr],
                      high_res_template.coords[lon_coord_hr], inputs_era5.coords['time'], lat_coord_hr, lon_coord_hr,
                      overlap_factor, network, noise, group_size)


def downscale(era5, raster_topo, range_lon=None, range_lat=None, overlap_factor=0.05, network=None, noise=None,
              group_size=None):
    """api.py:155-160.  Same result as process_era5 / process_topo / predict in sequence, but the nearest-neighbour
    regridding (api.py:31-43) is folded into the device-side patch gather: only the coarse ERA5 fields and the DEM
    raster cross PCIe instead of their 18x26-fold replicated hi-res copies."""
    import torch
    era5, raster_topo = from_xarray(era5), from_xarray(raster_topo)
    tpl = build_high_res_template_from_era5(era5, range_lon=range_lon, range_lat=range_lat)
    lon_coord, lat_coord = _coord_names(tpl)
    lat_t, lon_t = tpl.coords[lat_coord], tpl.coords[lon_coord]
    network = network if network is not None else get_network()
    u10c = np.ascontiguousarray(era5.transposed('u10', ('time', 'latitude', 'longitude')), np.float32)
    v10c = np.ascontiguousarray(era5.transposed('v10', ('time', 'latitude', 'longitude')), np.float32)
    uv_r, uv_c = nearest_index(era5.coords['latitude'], lat_t), nearest_index(era5.coords['longitude'], lon_t)
    name = next(iter(raster_topo.data_vars))
    dims, dem = raster_topo.var_dims(name), raster_topo[name]
    if 'band' in dims:
        dem = np.take(dem, 0, axis=dims.index('band'))
        dims = tuple(d for d in dims if d != 'band')
    if dims != ('y', 'x'):
        dem = np.transpose(dem, [dims.index('y'), dims.index('x')])
    dem = np.ascontiguousarray(dem, np.float32)
    e_r, e_c = nearest_index(raster_topo.coords['y'], lat_t), nearest_index(raster_topo.coords['x'], lon_t)
    d_u, d_v, d_dem = (torch.from_numpy(a).cuda() for a in (u10c, v10c, dem))
    d_maps = [_dev_i32(m) for m in (uv_r, uv_c, e_r, e_c)]
    time_window, pixels_lat, pixels_lon = u10c.shape[0], len(lat_t), len(lon_t)

    def gather(d_sx, nx, d_sy, ny, mean, std, tensors, scratch, stream):
        _lib.check(_lib.lib().wdg_gather_normalise_regrid(
            d_u.data_ptr(), d_v.data_ptr(), time_window, u10c.shape[1], u10c.shape[2], d_maps[0].data_ptr(), d_maps[1].data_ptr(),
            d_dem.data_ptr(), dem.shape[0], dem.shape[1], d_maps[2].data_ptr(), d_maps[3].data_ptr(), C.c_float(1e3),
            pixels_lat, pixels_lon, d_sx.data_ptr(), nx, d_sy.data_ptr(), ny, SEQUENCE_LENGTH, IMG_SIZE, mean.data_ptr(),
            std.data_ptr(), tensors.data_ptr(), scratch.data_ptr(), stream))

    return _run_tiles(gather, time_window, pixels_lat, pixels_lon, lat_t, lon_t, era5.coords['time'], lat_coord, lon_coord,
                      overlap_factor, network, noise, group_size)
