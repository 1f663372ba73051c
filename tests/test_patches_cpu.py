"""CPU tests of the tiling restatement (oracle/patches.py) against golden vectors produced by
executing the reference's own lines api.py:101-124 (tests/golden/make_patch_grid.py)."""
import json
import os

import numpy as np
import pytest

from oracle import patches as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "patch_grid.json")))
CASES = [k for k in GOLD if not k.startswith("_")]


@pytest.mark.parametrize("name", CASES)
def test_patch_grid_matches_reference_lines(name):
    g = GOLD[name]
    sx, sy = P.patch_grid(g["pixels_lat"], g["pixels_lon"], g["overlap_factor"])
    assert sx == g["slices_start_x"] and sy == g["slices_start_y"]
    ntimeseq = g["time_window"] // 24
    assert len(sx) * len(sy) * ntimeseq == g["n_patches"]
    # ordering (sx-major, sy, k) and the reversed-latitude slice incl. the sy == 0 quirk (F10)
    idx = 0
    for x in sx:
        for y in sy:
            rows = P.patch_rows(y)
            for k in range(ntimeseq):
                p = g["patches"][idx]
                assert (p["index"], p["sx"], p["sy"], p["k"]) == (idx, x, y, k)
                assert (p["lat_first"], p["lat_last"], p["n_lat"]) == (int(rows[0]), int(rows[-1]), len(rows))
                assert (p["lon_first"], p["lon_last"], p["n_lon"]) == (x, x + 95, 96)
                assert (p["t_first"], p["t_last"]) == (24 * k, 24 * k + 23)
                idx += 1


def test_known_answers_from_survey():
    assert P.patch_grid(234, 306, 0.01) == ([0, 70, 140, 210], [0, 69, 138])
    assert P.patch_grid(294, 429, 0.05) == ([0, 84, 167, 250, 333], [0, 66, 132, 198])
    sx, sy = P.patch_grid(962, 972, 0.05)
    assert len(sx) == len(sy) == 13 and sx[1] - sx[0] == 73 and sy[:4] == [0, 73, 146, 218] and sy[-1] == 866


def test_error_behaviour_matches_reference():
    with pytest.raises(RuntimeError):
        P.patch_grid(300, 96, 0.05)
    assert GOLD["_errors"]["lon_too_small"] == "RuntimeError"
    with pytest.raises(ZeroDivisionError):      # F11: latitude is never validated
        P.patch_grid(97, 300, 0.05)
    assert GOLD["_errors"]["lat_never_checked"] == "ZeroDivisionError"


def test_gather_normalise_stitch_roundtrip():
    rng = np.random.default_rng(0)
    H, W, T = 234, 306, 48
    u = rng.standard_normal((T, H, W)).astype(np.float32)
    v = rng.standard_normal((T, H, W)).astype(np.float32)
    e = rng.uniform(0, 3, (H, W))
    sx, sy = P.patch_grid(H, W, 0.01)
    t = P.gather_patches(u, v, e, sx, sy)
    assert t.shape == (len(sx) * len(sy) * 2, 24, 96, 96, 3)
    # first patch (sx=0, sy=0): rows 96..1 reversed (F10)
    assert np.array_equal(t[0, 3, :, :, 0], u[3, 96:0:-1, 0:96].astype(np.float64))
    # second y start: rows sy+95..sy
    n1 = 1 * 2  # (ix=0, iy=1, k=0)
    assert np.array_equal(t[n1, 0, :, :, 1], v[0, sy[1] + 95:sy[1] - 1:-1, 0:96].astype(np.float64))
    tn, mean, std = P.normalise(t)
    assert mean.shape == (1, 1, 1, 96, 3)       # F9: per (column, channel)
    assert np.allclose(tn.mean((0, 1, 2)), 0, atol=1e-9) and np.allclose(tn.std((0, 1, 2)), 1, atol=1e-9)
    # stitching the gathered u/v back is the identity on the covered region (mean of equal values)
    rows, cols, out = P.stitch(t[..., :2].astype(np.float32), sx, sy, 2)
    assert rows[0] == 3 and cols[0] == 2          # uncovered border (F10 shifts the first row block by one)
    assert np.array_equal(out[..., 0], u[:, rows][:, :, cols])
    assert np.array_equal(out[..., 1], v[:, rows][:, :, cols])


def test_stitch_overlap_mean_in_patch_order():
    sx, sy = P.patch_grid(130, 200, 0.3)
    n = len(sx) * len(sy)
    rng = np.random.default_rng(1)
    pred = rng.standard_normal((n, 24, 96, 96, 2)).astype(np.float32)
    rows, cols, out = P.stitch(pred, sx, sy, 1)
    # brute-force one overlapped pixel
    r, c = int(rows[len(rows) // 2]), int(cols[len(cols) // 2])
    vals, i = [], 0
    for x in sx:
        for y in sy:
            prow = P.patch_rows(y)
            if x + 2 <= c < x + 94 and r in prow[2:94]:
                vals.append(pred[i, 5, list(prow).index(r), c - x, 0])
            i += 1
    assert len(vals) > 1
    acc = np.float64(0)
    for v in vals:
        acc += np.float64(v)
    assert out[5, list(rows).index(r), list(cols).index(c), 0] == np.float32(acc / len(vals))


# ---------------------------------------------------------------------------------------------------------------
# The stitch mean pinned against REAL pandas (tests/golden/stitch_pandas.npz, made by make_stitch_golden.py with
# `pd.concat(...).groupby(level=[time, lat, lon]).mean()` as api.py:149-150 calls them).
def _stitch_cases():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "stitch_pandas.npz"))
    for key in z.files:
        if key.endswith("_meta"):
            yield key[:-5], z


def test_stitch_oracle_is_bit_exact_against_real_pandas():
    import hashlib
    n = 0
    for name, z in _stitch_cases():
        H, W, ov100, img, seq, nts, seed, full = (int(v) for v in z[name + "_meta"])
        sx, sy = P.patch_grid(H, W, ov100 / 100.0, img=img)
        N = len(sx) * len(sy) * nts
        pred = (np.random.default_rng(seed).standard_normal((N, seq, img, img, 2)) * 5).astype(np.float32)
        for accum, tag in (("float32", "f32"), ("float64", "f64")):
            rows, cols, out = P.stitch(pred, sx, sy, nts, seq=seq, img=img, accum=accum)
            assert hashlib.sha256(out.tobytes()).hexdigest() == str(z[f"{name}_sha_{tag}"]), (name, accum)
            if full:
                assert np.array_equal(out, z[f"{name}_{tag}"])
        n += 1
    assert n >= 4


def test_stitch_accumulation_modes_differ_only_in_the_last_bit():
    name, z = next(c for c in _stitch_cases() if c[0] == "img32_ov30")
    a, b = z[name + "_f32"], z[name + "_f64"]
    assert not np.array_equal(a, b)
    assert np.max(np.abs(a.astype(np.float64) - b)) < 2e-6       # inputs ~ N(0, 5^2): one float32 ulp of the partial sums
