"""Pins oracle/patches.py:stitch against REAL pandas (run in this container: pandas 3.0.2; xarray is absent).

The reference stitches with (api.py:140-151)

    predicted_squares = {key: xr.Dataset(... coords=squares[key].coords).isel(lon: slice(2,-2), lat: slice(2,-2))}
    bigdata = pd.concat([s.to_dataframe() for s in predicted_squares.values()])
    unique  = bigdata.groupby(level=['time', lat, lon]).mean()

`Dataset.to_dataframe()` of a (time, lat, lon) dataset is a frame with one float32 column per variable and the
product MultiIndex of its coordinate arrays in dimension order -- built here with `MultiIndex.from_product`
(the one xarray call that cannot be executed).  `pd.concat` and `groupby(level=...).mean()` are executed for real,
on the float32 frames and on the same frames cast to float64 (the float64 instantiation of the same Cython
`group_mean`, which is what pandas 1.3.3 -- the reference's pin -- runs for float32 input after `ensure_float64`).

Writes tests/golden/stitch_pandas.npz: for each case the integer geometry, the input seed, and either the full
stitched arrays (small cases) or their SHA-256 (the 96-px case).
"""
import hashlib
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import patches as P  # noqa: E402  (integer grid only: pinned separately by patch_grid.json)

# name: (H, W, overlap, img, seq, ntimeseq, seed, store_full)
CASES = {
    "img32_ov90": (150, 160, 0.9, 32, 2, 2, 0, False),     # up to ~280 overlapping patches per pixel
    "img32_ov30": (101, 77, 0.3, 32, 2, 1, 1, True),
    "cfg1_img96": (234, 306, 0.01, 96, 24, 1, 2, False),   # BASELINE configs[0] domain: 12 patches
    "img96_ov20": (300, 250, 0.2, 96, 2, 2, 3, False),
}


def pandas_stitch(pred, sx, sy, ntimeseq, seq, img, crop, H, W, dtype):
    lat = np.linspace(50.0, 40.0, H)     # descending, as in ERA5 / the hi-res template
    lon = np.linspace(-5.0, 9.0, W)
    time = np.datetime64("2016-04-01T00") + np.arange(seq * ntimeseq) * np.timedelta64(1, "h")
    frames, n = [], 0
    for x in sx:
        for y in sy:
            rows = P.patch_rows(y, img)[crop:img - crop]
            cols = np.arange(x, x + img)[crop:img - crop]
            for k in range(ntimeseq):
                idx = pd.MultiIndex.from_product([time[k * seq:(k + 1) * seq], lat[rows], lon[cols]],
                                                 names=["time", "latitude", "longitude"])
                p = pred[n, :, crop:img - crop, crop:img - crop]
                frames.append(pd.DataFrame({"u10": p[..., 0].reshape(-1).astype(dtype),
                                            "v10": p[..., 1].reshape(-1).astype(dtype)}, index=idx))
                n += 1
    bigdata = pd.concat(frames)                                                 # api.py:149
    unique = bigdata.groupby(level=["time", "latitude", "longitude"]).mean()    # api.py:150
    rows = P.covered(sy, True, img, crop)
    cols = P.covered(sx, False, img, crop)
    assert len(unique) == seq * ntimeseq * len(rows) * len(cols)
    out = unique.to_numpy().reshape(seq * ntimeseq, len(rows), len(cols), 2)
    # ascending latitude label = descending domain row
    return np.ascontiguousarray(out[:, ::-1]).astype(np.float32)


def main():
    blob = {"pandas_version": np.array(pd.__version__)}
    for name, (H, W, ov, img, seq, nts, seed, full) in CASES.items():
        sx, sy = P.patch_grid(H, W, ov, img=img)
        N = len(sx) * len(sy) * nts
        pred = (np.random.default_rng(seed).standard_normal((N, seq, img, img, 2)) * 5).astype(np.float32)
        o32 = pandas_stitch(pred, sx, sy, nts, seq, img, 2, H, W, np.float32)
        o64 = pandas_stitch(pred, sx, sy, nts, seq, img, 2, H, W, np.float64)
        blob[name + "_meta"] = np.array([H, W, int(round(ov * 100)), img, seq, nts, seed, int(full)])
        if full:
            blob[name + "_f32"], blob[name + "_f64"] = o32, o64
        blob[name + "_sha_f32"] = np.array(hashlib.sha256(o32.tobytes()).hexdigest())
        blob[name + "_sha_f64"] = np.array(hashlib.sha256(o64.tobytes()).hexdigest())
        print(name, len(sx), "x", len(sy), "patches", o32.shape, "f32 != f64 in", int((o32 != o64).sum()), "of", o32.size)
    np.savez_compressed(os.path.join(HERE, "stitch_pandas.npz"), **blob)


if __name__ == "__main__":
    main()
