"""`downscale` console script with the reference's flags (`cli.py:9-26`, `setup.py:12-16`).

netCDF4 / rasterio are not available in this image, so `--era` and `--dem` take `.npz` files written by
`GridDataset.to_npz` (ERA5: u10, v10 on (time, latitude, longitude); DEM: elevation on (band, y, x));
when xarray is importable, netCDF / GeoTIFF inputs are opened exactly as the reference does."""
import argparse
from pathlib import Path

from . import downscale
from .grid import GridDataset


def _open_era(folder, date):
    files = sorted(Path(folder).glob(f'{date}*surface*.npz'))
    if files:
        return GridDataset.from_npz(files[0])
    import xarray as xr  # reference behaviour (cli.py:22)
    return xr.open_mfdataset(Path(folder).glob(f'{date}*surface*.nc'))


def _open_dem(path):
    if str(path).endswith('.npz'):
        return GridDataset.from_npz(path)
    import xarray as xr  # reference behaviour (cli.py:23)
    return xr.open_rasterio(path)


def main(argv=None):
    parser = argparse.ArgumentParser(description='Downscale ER5 wind fields')
    parser.add_argument('--era', help='path to folder with ERA5 data', required=True)
    parser.add_argument('--dem', help='path to DEM data file', required=True)
    parser.add_argument('--date', help='date to downscale in YYYYMMDD format', required=True)
    parser.add_argument('--lon', default=None, help='longitude range (ex: 45.6:46.2)')
    parser.add_argument('--lat', default=None, help='latitude range (ex: 45.6:46.2)')
    parser.add_argument('-o', '--output', help='output path for the downscaled map (*.nc)', default='downscaled.nc')
    args = parser.parse_args(argv)

    longitude_r = tuple(map(float, args.lon.split(':'))) if args.lon else None
    latitude_r = tuple(map(float, args.lat.split(':'))) if args.lat else None

    era5 = _open_era(args.era, args.date)
    raster_topo = _open_dem(args.dem)
    downscaled_maps = downscale(era5, raster_topo, range_lon=longitude_r, range_lat=latitude_r, overlap_factor=0.01)
    out = str(args.output)
    if out.endswith('.nc'):
        try:
            import xarray as xr
            xr.Dataset({k: (d, a) for k, (d, a) in downscaled_maps.data_vars.items()},
                       coords=downscaled_maps.coords).to_netcdf(out)
            return
        except ImportError:
            out = out[:-3] + '.npz'
            print(f'xarray/netCDF4 unavailable: writing {out}')
    downscaled_maps.to_npz(out)


if __name__ == '__main__':
    main()
