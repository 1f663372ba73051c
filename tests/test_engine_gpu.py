"""GPU tests of the multi-window / ensemble engine (wind_downscaling_gan_b200/engine.py) against the single-call public API
(`downscale`, which tests/test_predict_gpu.py checks against the oracle pipeline)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(days):
    from tests.synth import synthetic_dem, synthetic_era5
    return synthetic_era5(hours=24 * days, seed=3), synthetic_dem(seed=4)


def _network(seed, std=0.1):
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200 import api
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    net = api.get_network()
    net.generator.set_weights(synthetic_generator_weights(7))
    net.noise_generator = FlexibleNoiseGenerator((8, 24, 96, 96, 20), std=std, random_seed=seed)
    return net


def _day(era, k):
    from wind_downscaling_gan_b200.grid import GridDataset
    sl = slice(24 * k, 24 * (k + 1))
    return GridDataset({v: (era.var_dims(v), era[v][sl]) for v in ("u10", "v10")},
                       {"time": era.coords["time"][sl], "latitude": era.coords["latitude"], "longitude": era.coords["longitude"]})


def test_series_equals_one_downscale_call_per_day_bit_for_bit():
    """3 days through the engine (one window per generator launch set) == downscale() on each day, same noise stream."""
    from wind_downscaling_gan_b200 import api, engine
    era, dem = _inputs(3)
    box = dict(range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0), overlap_factor=0.01)
    eng, units, out = engine.downscale_series(era, dem, network=_network(11), windows_per_forward=1, **box)
    assert units == {"windows": [0, 1, 2], "members": 1} and tuple(out.shape) == (3, 1, 2, 24, 229, 302)
    ref_net = _network(11)
    for k in range(3):
        ref = api.downscale(_day(era, k), dem, network=ref_net, group_size=12, **box)
        got = eng.as_dataset(out, k)
        assert np.array_equal(got["u10"], ref["u10"]) and np.array_equal(got["v10"], ref["v10"]), k
        assert np.array_equal(got.coords["lat_1"], ref.coords["lat_1"]) and np.array_equal(got.coords["lon_1"], ref.coords["lon_1"])
        assert np.array_equal(got.coords["time"], ref.coords["time"])


def test_windows_per_forward_and_sharding_do_not_change_results():
    """Several windows per generator launch set, and a 2-rank split, give the same maps (zero noise: deterministic)."""
    from wind_downscaling_gan_b200 import engine
    era, dem = _inputs(5)
    box = dict(range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0), overlap_factor=0.01)
    _, _, one = engine.downscale_series(era, dem, network=_network(1, std=0.0), windows_per_forward=1, **box)
    _, _, three = engine.downscale_series(era, dem, network=_network(1, std=0.0), windows_per_forward=3, **box)
    assert np.array_equal(one.numpy(), three.numpy())
    parts = [engine.downscale_series(era, dem, network=_network(1, std=0.0), rank=r, world=2, **box) for r in range(2)]
    assert [p[1]["windows"] for p in parts] == [[0, 1, 2], [3, 4]]
    assert np.array_equal(np.concatenate([p[2].numpy() for p in parts]), one.numpy())


def test_ensemble_members_share_one_gather():
    """members > 1: every member is a fresh noise draw over the same normalised patches; member m of the ensemble equals
    the m-th successive downscale() call on that day (same stream), and the members differ from each other."""
    from wind_downscaling_gan_b200 import api, engine
    era, dem = _inputs(1)
    box = dict(range_lon=(-1.0, 3.0), range_lat=(48.0, 50.0), overlap_factor=0.01)
    eng, units, out = engine.downscale_series(era, dem, network=_network(21), members=3, **box)
    assert units == {"windows": [0], "members": 3} and tuple(out.shape)[:2] == (1, 3)
    ref_net = _network(21)
    for m in range(3):
        ref = api.downscale(era, dem, network=ref_net, group_size=12, **box)
        assert np.array_equal(eng.as_dataset(out, 0, m)["u10"], ref["u10"]), m
    assert not np.array_equal(out[0, 0].numpy(), out[0, 1].numpy())
    # members split over two ranks: 2 + 1
    a = engine.downscale_series(era, dem, network=_network(21), members=3, rank=0, world=2, **box)
    b = engine.downscale_series(era, dem, network=_network(21), members=3, rank=1, world=2, **box)
    assert (a[1]["members"], b[1]["members"]) == (2, 1)
