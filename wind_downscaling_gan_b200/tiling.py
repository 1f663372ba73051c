"""Host-side integer logic of the reference's tiling driver (`api.py:98-124,140-151`): where the
96x96 patches start, in which order they are enumerated, and which domain rows/columns the cropped
patches cover.  The floating-point work (gather, normalise, stitch) runs on the GPU (csrc/wdg_patches.cu).
"""
import math

import numpy as np


def axis_starts(pixels: int, n: int, img: int):
    """api.py:111-116 for one axis: n patch starts, spaced (pixels-img)//(n-1) apart, the remainder
    handed out one pixel at a time to the first gaps."""
    step = (pixels - img) // (n - 1)          # ZeroDivisionError for n == 1, like the reference
    spare = pixels - ((n - 1) * step + img)
    return [i * step + min(i, spare) for i in range(n)]


def patch_grid(pixels_lat: int, pixels_lon: int, overlap_factor: float, img: int):
    """api.py:101-116 -> (slices_start_x, slices_start_y)."""
    min_cols, max_cols = math.ceil(pixels_lon / img), pixels_lon - img
    if max_cols < min_cols:
        raise RuntimeError(f'Lon dimension too small: got {pixels_lon} pixels, need at least {img}')
    min_rows, max_rows = math.ceil(pixels_lat / img), pixels_lat - img
    if max_cols < min_cols:  # the reference repeats the longitude test here (api.py:105); kept bug-compatible
        raise RuntimeError(f'Lat dimension too small: got {pixels_lat} pixels, need at least {img}')
    assert 0 <= overlap_factor <= 1, 'overlap_factor must be in [0,1] range'
    ncols = math.floor(min_cols + overlap_factor ** 2 * (max_cols - min_cols))
    nrows = math.floor(min_rows + overlap_factor ** 2 * (max_rows - min_rows))
    return axis_starts(pixels_lon, ncols, img), axis_starts(pixels_lat, nrows, img)


def covered_cols(starts_x, img: int, crop: int):
    """Domain columns present in the stitched output: union of [sx+crop, sx+img-crop)."""
    keep = np.zeros(max(starts_x) + img, bool)
    for s in starts_x:
        keep[s + crop:s + img - crop] = True
    return np.flatnonzero(keep).astype(np.int32)


def covered_rows(starts_y, img: int, crop: int):
    """Domain rows present in the stitched output.  Patch row p is domain row sy+img-1-p, except for
    sy == 0 where the reference's slice(img, 0, -1) shifts the block up by one row (api.py:119)."""
    keep = np.zeros(max(starts_y) + img + 1, bool)
    for s in starts_y:
        top = (s + img - 1) if s != 0 else img
        keep[top - (img - crop - 1):top - crop + 1] = True
    return np.flatnonzero(keep).astype(np.int32)
