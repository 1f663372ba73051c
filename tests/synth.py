"""Synthetic ERA5 / DEM inputs shaped like BASELINE configs."""
import numpy as np

from wind_downscaling_gan_b200.grid import GridDataset


def synthetic_era5(lon0=-1.0, lon1=3.0, lat0=48.0, lat1=50.0, hours=24, seed=0, step=0.25):
    """u10, v10 ~ smooth N(0, 5^2) m/s on a 0.25 degree grid, latitude descending as in ERA5 files."""
    rng = np.random.default_rng(seed)
    lon = np.arange(lon0, lon1 + 1e-9, step)
    lat = np.arange(lat1, lat0 - 1e-9, -step)
    def field():
        a = rng.standard_normal((hours, len(lat), len(lon)))
        a = (a + np.roll(a, 1, 1) + np.roll(a, 1, 2) + np.roll(a, 1, 0)) / 2.0
        return (5.0 * a).astype(np.float32)
    time = np.datetime64("2016-04-01T00") + np.arange(hours) * np.timedelta64(1, "h")
    return GridDataset({"u10": (("time", "latitude", "longitude"), field()),
                        "v10": (("time", "latitude", "longitude"), field())},
                       {"time": time, "latitude": lat, "longitude": lon})


def synthetic_dem(lon0=-1.5, lon1=3.5, lat0=47.5, lat1=50.5, seed=1, n=600):
    """elevation ~ U(0, 3000) m on an n x n raster (band, y, x), y descending like a GeoTIFF."""
    rng = np.random.default_rng(seed)
    x = np.linspace(lon0, lon1, n)
    y = np.linspace(lat1, lat0, n)
    return GridDataset({"elevation": (("band", "y", "x"), rng.uniform(0, 3000, (1, n, n)).astype(np.float32))},
                       {"x": x, "y": y})
