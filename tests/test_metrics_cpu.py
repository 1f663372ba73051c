"""CPU tests of the metrics oracle (known answers) and of the host-side metric helpers."""
import numpy as np


def test_oracle_known_answers():
    from oracle import metrics as om
    rng = np.random.default_rng(0)
    real = rng.standard_normal((2, 2, 8, 10, 2)).astype(np.float32)
    # identical fields: all distances zero
    for fn in (om.wind_speed_weighted_rmse, om.wind_speed_rmse, om.extreme_weighted_rmse, om.log_spectral_distance):
        assert np.allclose(fn(real, real), 0)
    assert np.allclose(om.angular_cosine_distance(real, real), 0, atol=2e-4)
    # opposite vectors: angular distance 1, "opposite cosine similarity" .5 * (1 - (-1)) = 1
    assert np.allclose(om.angular_cosine_distance(real, -real), 1, atol=2e-4)
    assert np.allclose(om.opposite_cosine_similarity(real, -real), 1, atol=1e-6)
    # orthogonal vectors: acd = 0.5
    rot = np.stack([-real[..., 1], real[..., 0]], -1)
    assert np.allclose(om.angular_cosine_distance(real, rot), 0.5, atol=1e-6)
    # under-estimating by a constant factor: tau = 1 - t, beta from the definition
    u = np.full((1, 1, 4, 4, 2), 3.0, np.float32)
    est, rea = np.sqrt(2) * 1.5, np.sqrt(2) * 3.0
    beta = (4 + rea) / (4 + est)
    want = np.sqrt(0.575 * 2 * (1.5 - beta * 3.0) ** 2)
    assert np.allclose(om.wind_speed_weighted_rmse(u, 0.5 * u), want, rtol=1e-6)
    assert np.allclose(om.wind_speed_rmse(u, 0.5 * u), rea - est, rtol=1e-6)


def test_oracle_lsd_is_over_width_and_channels():
    """tf.signal.rfft2d transforms the two innermost axes of the (B,T,H,W,C) tensor: rows of H are independent."""
    from oracle import metrics as om
    rng = np.random.default_rng(1)
    real = rng.standard_normal((1, 1, 6, 16, 2)).astype(np.float32)
    fake = rng.standard_normal((1, 1, 6, 16, 2)).astype(np.float32)
    full = om.log_spectral_distance(real, fake)[0] ** 2
    per_row = [om.log_spectral_distance(real[:, :, h:h + 1], fake[:, :, h:h + 1])[0] ** 2 for h in range(6)]
    assert np.allclose(full, np.mean(per_row))
    # scaling the fake field by 10 multiplies every power by 100: 20 dB everywhere
    assert np.allclose(om.log_spectral_distance(10 * real, real), 20.0, rtol=1e-4)


def test_oracle_ks_known_answers():
    from oracle import metrics as om
    real = np.full((1, 1, 12, 12, 1), -5.0, np.float32)
    fake = np.full((1, 1, 12, 12, 1), 5.0, np.float32)
    ks = om.spatially_convolved_ks_stat(real, fake, 3)
    assert ks.shape == (10, 10) and np.all(ks == 1.0)            # disjoint supports
    assert np.all(om.spatially_convolved_ks_stat(real, real, 3) == 0)
    half = fake.copy()
    half[0, 0, :6] = -5.0                                        # windows fully in the upper half agree with `real`
    ks = om.spatially_convolved_ks_stat(real, half, 3)
    assert np.all(ks[:4] == 0) and np.all(ks[6:] == 1.0) and np.allclose(ks[4], 1 / 3) and np.allclose(ks[5], 2 / 3)


def test_host_helpers():
    from wind_downscaling_gan_b200.gan import metrics as gm
    rng = np.random.default_rng(2)
    real = {"U_10M": rng.standard_normal((3, 8, 8)), "V_10M": rng.standard_normal((3, 8, 8))}
    fake = {"u10": real["U_10M"].copy(), "v10": real["V_10M"].copy()}
    assert np.allclose(gm.cosine_similarity_from_xarray(real, fake), 1.0)
    assert np.allclose(gm.log_spectral_distance_from_xarray(real, fake), 0.0)
    fake["u10"] = fake["u10"] * 0.5
    t = gm.tanh_wind_speed_weighted_rmse_from_xarray(real, fake)
    assert t.shape == (3, 8, 8) and np.all((t >= 0) & (t <= 1))
    a = rng.standard_normal((2, 3, 4, 4, 2))
    assert np.allclose(gm.rmse_from_xarray(a, a + 1.0), np.sqrt(2.0))
