"""GPU tests of the tiling driver around the generator (api.py:98-151 on the device): gather + normalise
within float tolerance, stitch BIT-EXACT against the oracle restatement, and `predict` / `downscale` / CLI
end to end on the BASELINE configs[0] domain."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _dev_i32(a):
    import torch
    return torch.as_tensor(np.asarray(a, np.int32), device="cuda")


@pytest.mark.parametrize("H,W,T,ov", [(234, 306, 48, 0.01), (294, 429, 24, 0.05), (130, 200, 30, 0.3)])
def test_gather_normalise_matches_oracle(H, W, T, ov):
    import torch
    from oracle import patches as P
    from wind_downscaling_gan_b200 import _lib, tiling
    rng = np.random.default_rng(0)
    u = (5 * rng.standard_normal((T, H, W))).astype(np.float32)
    v = (5 * rng.standard_normal((T, H, W))).astype(np.float32)
    e = rng.uniform(0, 3, (H, W)).astype(np.float32)
    sx, sy = tiling.patch_grid(H, W, ov, 96)
    ref, mean_ref, std_ref = P.normalise(P.gather_patches(u, v, e.astype(np.float64), sx, sy))
    nts = T // 24
    L = _lib.lib()
    nb = C.c_size_t()
    _lib.check(L.wdg_patch_scratch_bytes(len(sx), len(sy), nts, 96, C.byref(nb)))
    scratch = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    mean = torch.empty((96, 3), dtype=torch.float64, device="cuda")
    std = torch.empty((96, 3), dtype=torch.float64, device="cuda")
    out = torch.empty((len(sx) * len(sy) * nts, 24, 96, 96, 3), dtype=torch.float32, device="cuda")
    d = [torch.from_numpy(a).cuda() for a in (u, v, e)]
    dsx, dsy = _dev_i32(sx), _dev_i32(sy)
    _lib.check(L.wdg_gather_normalise(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), T, H, W, dsx.data_ptr(), len(sx),
                                      dsy.data_ptr(), len(sy), 24, 96, mean.data_ptr(), std.data_ptr(), out.data_ptr(),
                                      scratch.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.allclose(mean.cpu().numpy(), mean_ref.reshape(96, 3), rtol=1e-12, atol=1e-12)
    assert np.allclose(std.cpu().numpy(), std_ref.reshape(96, 3), rtol=1e-12, atol=1e-12)
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    assert np.allclose(got, ref.astype(np.float32), rtol=2e-7, atol=2e-7)


@pytest.mark.parametrize("H,W,nts,ov", [(234, 306, 2, 0.01), (294, 429, 1, 0.05), (130, 200, 1, 0.3), (192, 192, 1, 0.05)])
def test_stitch_bit_exact(H, W, nts, ov):
    import torch
    from oracle import patches as P
    from wind_downscaling_gan_b200 import _lib, tiling
    sx, sy = tiling.patch_grid(H, W, ov, 96)
    n = len(sx) * len(sy) * nts
    rng = np.random.default_rng(1)
    pred = (3 * rng.standard_normal((n, 24, 96, 96, 2))).astype(np.float32)
    rows_ref, cols_ref, ref = P.stitch(pred, sx, sy, nts)
    rows, cols = tiling.covered_rows(sy, 96, 2), tiling.covered_cols(sx, 96, 2)
    assert np.array_equal(rows, rows_ref) and np.array_equal(cols, cols_ref)
    out = torch.empty((2, nts * 24, len(rows), len(cols)), dtype=torch.float32, device="cuda")
    dp = torch.from_numpy(pred).cuda()
    dsx, dsy, dr, dc = _dev_i32(sx), _dev_i32(sy), _dev_i32(rows), _dev_i32(cols)
    _lib.check(_lib.lib().wdg_stitch(dp.data_ptr(), dsx.data_ptr(), len(sx), dsy.data_ptr(), len(sy), nts, 24, 96, 2, 2,
                                     dr.data_ptr(), len(rows), dc.data_ptr(), len(cols), out.data_ptr(), None))
    got = out.cpu().numpy()
    assert np.array_equal(got[0], ref[..., 0]) and np.array_equal(got[1], ref[..., 1])   # bit-exact


@pytest.mark.parametrize("accum", ["float64", "float32"])
def test_stitch_bit_exact_against_real_pandas(accum):
    """tests/golden/stitch_pandas.npz = output of real pandas `concat(...).groupby(level=...).mean()` (api.py:149-150)
    on float32 frames (pandas >= 1.5 semantics) and float64 frames (pandas 1.3.3, the reference's pin)."""
    import hashlib
    import os
    import torch
    from wind_downscaling_gan_b200 import _lib, tiling
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "stitch_pandas.npz"))
    tag, mode = {"float64": ("f64", 0), "float32": ("f32", 1)}[accum]
    for name in [k[:-5] for k in z.files if k.endswith("_meta")]:
        H, W, ov100, img, seq, nts, seed, full = (int(v) for v in z[name + "_meta"])
        sx, sy = tiling.patch_grid(H, W, ov100 / 100.0, img)
        N = len(sx) * len(sy) * nts
        pred = (np.random.default_rng(seed).standard_normal((N, seq, img, img, 2)) * 5).astype(np.float32)
        rows, cols = tiling.covered_rows(sy, img, 2), tiling.covered_cols(sx, img, 2)
        out = torch.empty((2, nts * seq, len(rows), len(cols)), dtype=torch.float32, device="cuda")
        dp = torch.from_numpy(pred).cuda()
        dsx, dsy, dr, dc = _dev_i32(sx), _dev_i32(sy), _dev_i32(rows), _dev_i32(cols)
        _lib.check(_lib.lib().wdg_stitch_accum(dp.data_ptr(), dsx.data_ptr(), len(sx), dsy.data_ptr(), len(sy), nts, seq, img, 2, 2,
                                               dr.data_ptr(), len(rows), dc.data_ptr(), len(cols), out.data_ptr(), mode, None))
        got = np.ascontiguousarray(out.cpu().numpy().transpose(1, 2, 3, 0))     # (T, rows, cols, C) like the golden
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(z[f"{name}_sha_{tag}"]), (name, accum)
        if full:
            assert np.array_equal(got, z[f"{name}_{tag}"])


def _cfg1():
    from tests.synth import synthetic_dem, synthetic_era5
    return synthetic_era5(), synthetic_dem()


def test_predict_cfg1_matches_oracle_pipeline():
    """BASELINE configs[0]: one date, lon -1:3, lat 48:50 -> 306 x 234 px, 12 patches x 24 h, fixed noise."""
    import torch
    from oracle import patches as P
    from oracle.generator import synthetic_generator_weights
    from oracle.torch_port import TorchGenerator
    from wind_downscaling_gan_b200 import api
    era, dem = _cfg1()
    tpl = api.build_high_res_template_from_era5(era, range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0))
    e5, topo = api.process_era5(era, tpl), api.process_topo(dem, tpl)
    net = api.get_network()
    w = synthetic_generator_weights(7)
    net.generator.set_weights(w)
    noise = (0.1 * np.random.default_rng(5).standard_normal((12, 24, 96, 96, 20))).astype(np.float32)
    out = api.predict(e5, topo, tpl, overlap_factor=0.01, network=net, noise=noise)
    # oracle pipeline
    sx, sy = P.patch_grid(234, 306, 0.01)
    assert (sx, sy) == ([0, 70, 140, 210], [0, 69, 138])
    t = P.gather_patches(e5["u10"], e5["v10"], topo["elevation"] / 1e3, sx, sy)
    tn, _, _ = P.normalise(t)
    pred = TorchGenerator(w, torch.float64).forward(tn.astype(np.float32), noise).numpy().astype(np.float32)
    rows, cols, ref = P.stitch(pred, sx, sy, 1)
    assert out["u10"].shape == (24, len(rows), len(cols)) == (24, 229, 302)
    assert np.array_equal(out.coords["lat_1"], tpl.coords["lat_1"][rows])
    assert np.array_equal(out.coords["lon_1"], tpl.coords["lon_1"][cols])
    for i, v in enumerate(("u10", "v10")):
        rel = np.linalg.norm(out[v].astype(np.float64) - ref[..., i]) / np.linalg.norm(ref[..., i])
        assert rel < 1e-2, (v, rel)


def test_fused_regrid_downscale_equals_predict_on_regridded_inputs():
    """downscale() folds the nearest-neighbour regrid (api.py:31-43) into the device gather; the result must equal
    process_era5 / process_topo on the host followed by predict(), bit for bit."""
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200 import api
    era, dem = _cfg1()
    net = api.get_network()
    net.generator.set_weights(synthetic_generator_weights(7))
    noise = (0.1 * np.random.default_rng(5).standard_normal((12, 24, 96, 96, 20))).astype(np.float32)
    tpl = api.build_high_res_template_from_era5(era, range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0))
    a = api.predict(api.process_era5(era, tpl), api.process_topo(dem, tpl), tpl, overlap_factor=0.01, network=net, noise=noise)
    b = api.downscale(era, dem, range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0), overlap_factor=0.01, network=net, noise=noise)
    assert np.array_equal(a["u10"], b["u10"]) and np.array_equal(a["v10"], b["v10"])
    assert np.array_equal(a.coords["lat_1"], b.coords["lat_1"]) and np.array_equal(a.coords["lon_1"], b.coords["lon_1"])


def test_downscale_and_cli(tmp_path):
    from wind_downscaling_gan_b200 import cli, downscale
    from wind_downscaling_gan_b200.grid import GridDataset
    era, dem = _cfg1()
    res = downscale(era, dem, range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0), overlap_factor=0.01)
    assert res["u10"].shape == (24, 229, 302) and np.isfinite(res["v10"]).all()
    (tmp_path / "era").mkdir()
    era.to_npz(tmp_path / "era" / "20160401_era5_surface_hourly.npz")
    dem.to_npz(tmp_path / "dem.npz")
    out = tmp_path / "downscaled.npz"
    cli.main(["--era", str(tmp_path / "era"), "--dem", str(tmp_path / "dem.npz"), "--date", "20160401",
              "--lon=-1:3", "--lat", "48:50", "-o", str(out)])   # argparse needs '=' for a leading minus, as in the reference
    back = GridDataset.from_npz(out)
    assert back["u10"].shape == (24, 229, 302) and back.var_dims("u10") == ("time", "lat_1", "lon_1")


def test_noise_generator_shapes_and_std():
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    g = FlexibleNoiseGenerator((8, 24, 96, 96, 20), std=0.1, random_seed=3)
    a = g(bs=2, channels=20)
    assert tuple(a.shape) == (2, 24, 96, 96, 20) and a.is_cuda
    assert abs(float(a.std()) - 0.1) < 2e-3 and abs(float(a.mean())) < 1e-3
    assert tuple(g(channels=2).shape) == (8, 24, 96, 96, 2)
    b = FlexibleNoiseGenerator((8, 24, 96, 96, 20), std=0.1, random_seed=3)(bs=2, channels=20)
    assert bool((a == b).all())
    c = g(bs=2, channels=20)                       # the stream advances between calls
    assert not bool((a == c).all())
    z = (a / 0.1).double()
    assert abs(float((z ** 4).mean()) - 3.0) < 0.05 and abs(float((z ** 3).mean())) < 0.02   # Gaussian moments


def test_predict_host_with_device_noise_matches_explicit_noise():
    """The generated-noise host entry point equals predict_host fed with the same Philox stream (both chunked paths)."""
    import torch
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan.models import make_generator
    for B in (3, 35):
        T, S = 2, 32
        gen = make_generator(S, 3, 20, 2, T)
        gen.set_weights(synthetic_generator_weights(9))
        image = torch.randn((B, T, S, S, 3)).pin_memory()
        noise = FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=5)()
        ref = gen.predict_host(image, noise.cpu())
        got = gen.predict_host_gen_noise(image, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=5))
        assert torch.equal(ref, got)
