"""Generates tests/golden/patch_grid.json by EXECUTING the reference's own tiling lines
(`/root/reference/src/downscaling/api.py:101-124`: grid arithmetic, slice construction and patch
ordering) against a recording stand-in for the xarray dataset.  The lines are read from the
reference at generation time; nothing of the reference is copied into this repository.

Run in the build container:  python tests/golden/make_patch_grid.py
"""
import json
import math
import os
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
API = "/root/reference/src/downscaling/api.py"
FIRST, LAST = 101, 124   # min_cols ... positions = {...}

CASES = [  # (name, pixels_lat, pixels_lon, time_window, overlap_factor)
    ("cfg1_cli_306x234_ov0.01", 234, 306, 24, 0.01),
    ("cfg3_switzerland_429x294_ov0.05", 294, 429, 48, 0.05),
    ("cfg5_weurope_972x962_ov0.05", 962, 972, 24, 0.05),
    ("small_200x130_ov0.3", 130, 200, 72, 0.3),
    ("exact_multiple_192x192_ov0.05", 192, 192, 24, 0.05),
    ("default_overlap_500x333", 333, 500, 25, 0.05),
]


class Rec:
    """Stands in for `ds_in`: records the isel() slices."""

    def __init__(self, sl=None):
        self.sl = sl

    def isel(self, d):
        return Rec(d)

    def __getitem__(self, names):
        assert names == ["u10", "v10", "elevation"]
        return self


def run_case(pixels_lat, pixels_lon, time_window, overlap_factor):
    src = open(API).read().splitlines()[FIRST - 1:LAST]
    code = textwrap.dedent("\n".join(src))
    npx = types.SimpleNamespace(**{k: getattr(np, k) for k in ("concatenate", "ones", "zeros")})
    npx.math = math                      # np.math was removed from numpy 2.x; the reference pins numpy 1.19
    ns = dict(np=npx, IMG_SIZE=96, SEQUENCE_LENGTH=24, pixels_lat=pixels_lat, pixels_lon=pixels_lon,
              overlap_factor=overlap_factor, ntimeseq=time_window // 24, ds_in=Rec(), lon_coord_hr="lon",
              lat_coord_hr="lat", print=lambda *a, **k: None)
    exec(code, ns)
    patches = []
    for (sx, sy, k), rec in ns["squares"].items():
        lat = list(range(pixels_lat))[rec.sl["lat"]]
        lon = list(range(pixels_lon))[rec.sl["lon"]]
        t = list(range(time_window))[rec.sl["time"]]
        patches.append({"index": ns["positions"][(sx, sy, k)], "sx": sx, "sy": sy, "k": k,
                        "lat_first": lat[0], "lat_last": lat[-1], "n_lat": len(lat),
                        "lon_first": lon[0], "lon_last": lon[-1], "n_lon": len(lon),
                        "t_first": t[0], "t_last": t[-1]})
    return {"pixels_lat": pixels_lat, "pixels_lon": pixels_lon, "time_window": time_window,
            "overlap_factor": overlap_factor, "slices_start_x": ns["slices_start_x"],
            "slices_start_y": ns["slices_start_y"], "n_patches": len(patches), "patches": patches}


def main():
    out = {name: run_case(*args) for name, *args in CASES}
    # error behaviour of api.py:102-106
    errs = {}
    for name, (la, lo) in {"lon_too_small": (300, 96), "lat_never_checked": (97, 300)}.items():
        try:
            run_case(la, lo, 24, 0.05)
            errs[name] = None
        except Exception as e:  # noqa: BLE001
            errs[name] = type(e).__name__
    out["_errors"] = errs
    with open(os.path.join(HERE, "patch_grid.json"), "w") as f:
        json.dump(out, f)
    print({k: (v["n_patches"] if isinstance(v, dict) and "n_patches" in v else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
