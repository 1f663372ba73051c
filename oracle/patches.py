"""Numpy restatement of the reference's tiling driver (`api.py:98-151`):
patch grid, reversed-latitude slicing, (column, channel) normalisation,
2-px crop and label-keyed overlap mean.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Integer parts are pinned by
the known answers evaluated from api.py:98-116 (tests/golden/patch_grid.json).
"""
import math

import numpy as np

IMG = 96  # api.py:23
SEQ = 24  # api.py:22
CROP = 2  # api.py:148


def _starts(pixels, n, img):
    """api.py:111-116 for one axis."""
    dist = (pixels - img) // (n - 1)
    left = pixels - ((n - 1) * dist + img)
    vec = np.concatenate([[0], np.ones(left), np.zeros(n - left - 1)]).cumsum()
    return [int(i * dist + x) for i, x in zip(range(n), vec)]


def patch_grid(pixels_lat, pixels_lon, overlap_factor=0.05, img=IMG):
    """api.py:101-116.  Returns (slices_start_x, slices_start_y).

    Reproduces F11: the latitude size is never validated (the lon test is repeated)."""
    min_cols, max_cols = math.ceil(pixels_lon / img), pixels_lon - img
    if max_cols < min_cols:
        raise RuntimeError(f"Lon dimension too small: got {pixels_lon} pixels, need at least {img}")
    min_rows, max_rows = math.ceil(pixels_lat / img), pixels_lat - img
    if max_cols < min_cols:  # sic (api.py:105)
        raise RuntimeError(f"Lat dimension too small: got {pixels_lat} pixels, need at least {img}")
    assert 0 <= overlap_factor <= 1, "overlap_factor must be in [0,1] range"
    ncols = math.floor(min_cols + overlap_factor ** 2 * (max_cols - min_cols))
    nrows = math.floor(min_rows + overlap_factor ** 2 * (max_rows - min_rows))
    return _starts(pixels_lon, ncols, img), _starts(pixels_lat, nrows, img)


def patch_rows(sy, img=IMG):
    """Domain row of patch row p (api.py:119): rows sy+95..sy, or 96..1 when sy == 0 (F10)."""
    if sy != 0:
        return np.arange(sy + img - 1, sy - 1, -1)
    return np.arange(img, 0, -1)


def gather_patches(u10, v10, elevation_km, starts_x, starts_y, seq=SEQ, img=IMG):
    """api.py:117-127.  u10, v10: (T, H, W); elevation_km: (H, W) (already / 1e3, api.py:96).

    Returns float64 (N, seq, img, img, 3), N ordered sx-major, then sy, then time block k."""
    T = u10.shape[0]
    ntimeseq = T // seq
    out = []
    for sx in starts_x:
        for sy in starts_y:
            rows = patch_rows(sy, img)
            for k in range(ntimeseq):
                ts = slice(k * seq, (k + 1) * seq)
                u = u10[ts][:, rows][:, :, sx:sx + img]
                v = v10[ts][:, rows][:, :, sx:sx + img]
                e = np.broadcast_to(elevation_km[rows][:, sx:sx + img], u.shape)
                out.append(np.stack([u, v, e], -1).astype(np.float64))
    return np.stack(out, 0)


def normalise(tensors):
    """api.py:128-129: statistics over axes (0,1,2) of the 5-D array -> per (column, channel) (F9)."""
    mean = np.nanmean(tensors, axis=(0, 1, 2), keepdims=True)
    std = np.nanstd(tensors, axis=(0, 1, 2), keepdims=True)
    return (tensors - mean) / std, mean, std


def covered(starts, first_is_zero_shift, img=IMG, crop=CROP):
    """Sorted domain indices covered by the cropped patches along one axis.

    Columns: [sx+2, sx+94).  Rows: [sy+2, sy+94), except sy == 0 covers [3, 95) (F10)."""
    idx = set()
    for s in starts:
        lo = s + crop
        if first_is_zero_shift and s == 0:
            lo += 1
        idx.update(range(lo, lo + img - 2 * crop))
    return np.array(sorted(idx), dtype=np.int64)


def stitch(pred, starts_x, starts_y, ntimeseq, seq=SEQ, img=IMG, crop=CROP, accum="float64"):
    """api.py:140-151.  pred: float32 (N, seq, img, img, C) in gather order.

    Crop [2:-2] in both patch axes, key every value by (time, lat, lon), mean over duplicates, sorted ascending by
    (time, row, col).  The mean is pandas' `groupby(level=...).mean()` = Cython `group_mean`
    (pandas/_libs/groupby.pyx): per key, Kahan-compensated running sum in order of appearance (= patch index),
    divided by the count, in the template's floating type:
      accum="float64": pandas==1.3.3, the reference's pin (requirements.txt:11) -- `_call_cython_op` upcasts float32
                       columns with `ensure_float64`, runs group_mean[float64] and casts the result back to float32;
      accum="float32": pandas >= 1.5 (3.0.2 is what this image has) -- group_mean[float32] directly.
    PINNED: tests/golden/stitch_pandas.npz holds the output of REAL pandas 3.0.2 `concat(...).groupby(level=...).mean()`
    on float32 frames (-> the float32 mode, bit for bit) and on the same frames cast to float64 (-> the float64
    template of the same loop, bit for bit); tests/golden/make_stitch_golden.py is the generating script.
    Returns (rows, cols, out) with out float32 (ntimeseq*seq, len(rows), len(cols), C)."""
    dt = {"float64": np.float64, "float32": np.float32}[accum]
    pred = np.asarray(pred, np.float32)
    C = pred.shape[-1]
    rows = covered(starts_y, True, img, crop)
    cols = covered(starts_x, False, img, crop)
    rpos = {int(r): i for i, r in enumerate(rows)}
    cpos = {int(c): i for i, c in enumerate(cols)}
    shape = (ntimeseq * seq, len(rows), len(cols), C)
    sumx, comp = np.zeros(shape, dt), np.zeros(shape, dt)
    cnt = np.zeros((len(rows), len(cols)), np.int64)
    n = 0
    for sx in starts_x:
        ci = np.array([cpos[sx + p] for p in range(crop, img - crop)])
        for sy in starts_y:
            prow = patch_rows(sy, img)[crop:img - crop]
            ri = np.array([rpos[int(r)] for r in prow])
            cnt[np.ix_(ri, ci)] += 1
            for k in range(ntimeseq):
                sl = (slice(k * seq, (k + 1) * seq), ri[:, None], ci[None, :])
                val = pred[n, :, crop:img - crop, crop:img - crop].astype(dt)
                y = val - comp[sl]                 # group_mean: y = val - compensation
                t = sumx[sl] + y                   #             t = sumx + y
                comp[sl] = (t - sumx[sl]) - y      #             compensation = t - sumx - y
                sumx[sl] = t
                n += 1
    assert n == pred.shape[0]
    out = (sumx / cnt[None, :, :, None].astype(dt)).astype(np.float32)
    return rows, cols, out
