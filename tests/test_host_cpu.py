"""CPU tests of the host-side mirror of api.py: regridding, template, tiling logic, error behaviour."""
import json
import os

import numpy as np
import pytest

from wind_downscaling_gan_b200 import api, tiling
from wind_downscaling_gan_b200.grid import GridDataset, nearest_index
from tests.synth import synthetic_dem, synthetic_era5

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "patch_grid.json")))


def test_constants_match_reference():
    # api.py:22-28
    assert (api.SEQUENCE_LENGTH, api.IMG_SIZE, api.BATCH_SIZE, api.NOISE_CHANNELS, api.NOISE_STD, api.NB_INPUTS,
            api.NB_OUTPUTS) == (24, 96, 8, 20, 0.1, 3, 2)
    assert api.WEIGHTS_PATH.name == "weights-55.ckpt"


@pytest.mark.parametrize("name", [k for k in GOLD if not k.startswith("_")])
def test_product_tiling_matches_reference_lines(name):
    g = GOLD[name]
    sx, sy = tiling.patch_grid(g["pixels_lat"], g["pixels_lon"], g["overlap_factor"], 96)
    assert sx == g["slices_start_x"] and sy == g["slices_start_y"]
    lat_cov = set()
    for p in g["patches"]:
        lo, hi = sorted((p["lat_first"], p["lat_last"]))
        lat_cov.update(range(lo + 2, hi - 1))
    assert sorted(lat_cov) == list(tiling.covered_rows(sy, 96, 2))
    lon_cov = set()
    for p in g["patches"]:
        lon_cov.update(range(p["lon_first"] + 2, p["lon_last"] - 1))
    assert sorted(lon_cov) == list(tiling.covered_cols(sx, 96, 2))


def test_tiling_errors():
    with pytest.raises(RuntimeError, match="Lon dimension too small"):
        tiling.patch_grid(300, 96, 0.05, 96)
    with pytest.raises(ZeroDivisionError):
        tiling.patch_grid(97, 300, 0.05, 96)
    with pytest.raises(AssertionError):
        tiling.patch_grid(300, 300, 1.5, 96)


def test_nearest_index_matches_pandas():
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(0)
    for labels in (np.arange(48.0, 50.01, 0.25), np.arange(50.0, 47.99, -0.25), np.array([1.0, 2.0, 4.0, 8.0])):
        targets = np.concatenate([rng.uniform(labels.min() - 1, labels.max() + 1, 200),
                                  (labels[:-1] + labels[1:]) / 2, labels])
        ref = pd.Index(labels).get_indexer(targets, method="nearest")
        assert np.array_equal(nearest_index(labels, targets), ref)


def test_template_and_regrid_cfg1():
    era = synthetic_era5()
    dem = synthetic_dem()
    tpl = api.build_high_res_template_from_era5(era, range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0))
    assert len(tpl.coords["lon_1"]) == 18 * 17 == 306 and len(tpl.coords["lat_1"]) == 26 * 9 == 234
    e = api.process_era5(era, tpl)
    assert e["u10"].shape == (24, 234, 306)
    # nearest neighbour: template point -> closest ERA5 node
    j = int(np.argmin(np.abs(era.coords["latitude"] - tpl.coords["lat_1"][100])))
    i = int(np.argmin(np.abs(era.coords["longitude"] - tpl.coords["lon_1"][200])))
    assert e["u10"][5, 100, 200] == era["u10"][5, j, i]
    t = api.process_topo(dem, tpl)
    assert t["elevation"].shape == (234, 306)
    jj = int(np.argmin(np.abs(dem.coords["y"] - tpl.coords["lat_1"][7])))
    ii = int(np.argmin(np.abs(dem.coords["x"] - tpl.coords["lon_1"][9])))
    assert t["elevation"][7, 9] == dem["elevation"][0, jj, ii]


def test_grid_dataset_npz_roundtrip(tmp_path):
    era = synthetic_era5(hours=2)
    era.to_npz(tmp_path / "a.npz")
    back = GridDataset.from_npz(tmp_path / "a.npz")
    assert np.array_equal(back["u10"], era["u10"]) and back.var_dims("v10") == ("time", "latitude", "longitude")
    assert np.array_equal(back.coords["time"], era.coords["time"])


def test_gan_compile_selects_training_precision():
    """GAN.compile(..., train_precision=...) is an extension over ganbase.py:115-124: it selects the arithmetic of the
    training convolution GEMMs (process-wide setting of the C ABI, wdg_train_set_precision)."""
    import pytest
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.train import ops

    class Stub:
        def compile(self, *a, **k):
            self.args = (a, k)

    gan = GAN(Stub(), Stub(), noise_generator=None)
    try:
        gan.compile(generator_optimizer="g", discriminator_optimizer="d", train_precision="tf32")
        assert ops.get_precision() == "tf32"
        gan.compile(generator_optimizer="g", discriminator_optimizer="d", train_precision="bf16")
        assert ops.get_precision() == "bf16"
        with pytest.raises(KeyError):
            gan.compile(generator_optimizer="g", discriminator_optimizer="d", train_precision="fp8")
    finally:
        ops.set_precision("fp32")
    assert ops.get_precision() == "fp32"


def test_shard_units_covers_everything_once():
    from wind_downscaling_gan_b200.engine import shard_units
    for n in (0, 1, 5, 8, 100, 365):
        for world in (1, 2, 3, 8):
            spans = [shard_units(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_numa_binding_is_a_noop_without_topology_information(monkeypatch):
    """hostmem.bind_to_gpu_numa_node: parses sysfs cpulists, never widens the affinity mask, and does nothing when the box
    does not say which node a GPU sits on (this container: no GPU at all)."""
    import os
    from wind_downscaling_gan_b200 import hostmem
    before = os.sched_getaffinity(0)
    rep = hostmem.bind_to_gpu_numa_node(0)
    assert rep["bound"] is False and rep["numa_node"] is None and os.sched_getaffinity(0) == before
    # a node whose CPU list covers this process: binding keeps the intersection only
    monkeypatch.setattr(hostmem, "gpu_numa_node", lambda i: 0)
    some = set(sorted(before)[:max(1, len(before) // 2)])
    monkeypatch.setattr(hostmem, "node_cpus", lambda n: some | {10_000})
    try:
        rep = hostmem.bind_to_gpu_numa_node(0)
        assert os.sched_getaffinity(0) == some and rep["bound"] == (some != before) and rep["cpus_on_node"] == len(some)
    finally:
        os.sched_setaffinity(0, before)


def test_numa_cpulist_parsing(monkeypatch):
    import builtins
    import io
    from wind_downscaling_gan_b200 import hostmem
    real_open = builtins.open
    monkeypatch.setattr(builtins, "open", lambda p, *a, **k: io.StringIO("0-3,8,10-11\n") if str(p).endswith("node7/cpulist")
                        else real_open(p, *a, **k))
    assert hostmem.node_cpus(7) == {0, 1, 2, 3, 8, 10, 11}
    monkeypatch.undo()
    assert hostmem.node_cpus(123456) == set()
