"""GPU parity tests of the evaluation metrics (csrc/wdg_metrics.cu) against oracle/metrics.py.
Tolerances: fused fp32 reductions 1e-5 relative (angular metrics 1e-4: acos is ill-conditioned near +-1 in fp32),
log spectral distance 1e-4, KS image exact up to fp32 rounding of count / P^2 (1e-6 absolute)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def fields(B, T, H, W, seed, nan=False):
    rng = np.random.default_rng(seed)
    real = (6 * rng.standard_normal((B, T, H, W, 2))).astype(np.float32)
    fake = (real + 2 * rng.standard_normal((B, T, H, W, 2))).astype(np.float32)
    fake[0, 0, :3] = real[0, 0, :3]                  # identical vectors: cos = 1 exactly
    real[-1, -1, 2, 5] = 0                           # zero vector: the 1e-12 clamp of l2_normalize
    if nan:
        real[0, 0, 1, 1, 0] = np.nan
    return real, fake


@pytest.mark.parametrize("B,T,H,W", [(2, 3, 24, 20), (3, 2, 96, 96)])
def test_pointwise_metrics(B, T, H, W):
    from oracle import metrics as om
    from wind_downscaling_gan_b200.gan import metrics as gm
    real, fake = fields(B, T, H, W, 0)
    got = gm.pointwise_metrics(real, fake)
    ref = {"ws_weighted_rmse": om.wind_speed_weighted_rmse(real, fake), "ws_rmse": om.wind_speed_rmse(real, fake),
           "extreme_rmse": om.extreme_weighted_rmse(real, fake), "acd": om.angular_cosine_distance(real, fake),
           "opposite_cosine_similarity": om.opposite_cosine_similarity(real, fake)}
    for k, r in ref.items():
        tol = 1e-4 if k in ("acd", "opposite_cosine_similarity") else 1e-5
        assert got[k].shape == (B,)
        np.testing.assert_allclose(got[k], r, rtol=tol, atol=1e-7, err_msg=k)
    # the single-metric entry points of the reference return the same numbers
    np.testing.assert_array_equal(gm.wind_speed_weighted_rmse(real, fake), got["ws_weighted_rmse"])
    np.testing.assert_array_equal(gm.angular_cosine_distance(real, fake), got["acd"])
    # identical inputs: every distance is zero (acd up to the clamp)
    same = gm.pointwise_metrics(real, real)
    assert np.all(same["ws_weighted_rmse"] == 0) and np.all(same["ws_rmse"] == 0) and np.all(same["extreme_rmse"] == 0)
    np.testing.assert_allclose(same["acd"], om.angular_cosine_distance(real, real), atol=2e-4)   # fp32 acos near 1


def test_pointwise_nan_is_dropped():
    from oracle import metrics as om
    from wind_downscaling_gan_b200.gan import metrics as gm
    real, fake = fields(2, 2, 16, 16, 1, nan=True)
    got = gm.pointwise_metrics(real, fake)
    np.testing.assert_allclose(got["ws_weighted_rmse"], om.wind_speed_weighted_rmse(real, fake), rtol=1e-5)
    np.testing.assert_allclose(got["ws_rmse"], om.wind_speed_rmse(real, fake), rtol=1e-5)
    assert np.isfinite(got["ws_weighted_rmse"]).all() and np.isfinite(got["ws_rmse"]).all()


@pytest.mark.parametrize("B,T,H,W", [(2, 2, 12, 20), (2, 3, 96, 96)])
def test_log_spectral_distance(B, T, H, W):
    from oracle import metrics as om
    from wind_downscaling_gan_b200.gan import metrics as gm
    real, fake = fields(B, T, H, W, 2)
    got = gm.log_spectral_distance(real, fake)
    np.testing.assert_allclose(got, om.log_spectral_distance(real, fake), rtol=1e-4)
    assert np.all(gm.log_spectral_distance(real, real) == 0)


@pytest.mark.parametrize("B,T,H,W,P", [(2, 2, 20, 24, None), (1, 2, 96, 96, None), (2, 1, 30, 30, 5)])
def test_spatial_ks(B, T, H, W, P):
    from oracle import metrics as om
    from wind_downscaling_gan_b200.gan import metrics as gm
    real, fake = fields(B, T, H, W, 3)
    real[0, 0, 0, :4, 0] = om.KS_POINTS[[0, 17, 50, 99]]        # values exactly ON comparison points
    fake[0, 0, 1, :3, 1] = [-31.0, 30.0, 45.0]                   # outside the point range
    got = gm.spatially_convolved_ks_stat(real, fake, P)
    ref = om.spatially_convolved_ks_stat(real, fake, P)
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, atol=1e-6, rtol=0)
    assert np.all(gm.spatially_convolved_ks_stat(real, real, P) == 0)


def test_metric_wrappers():
    from wind_downscaling_gan_b200.gan import metrics as gm
    real, fake = fields(2, 2, 16, 16, 4)
    m = gm.WindSpeedWeightedRMSE()
    assert m.name == "ws_weighted_rmse"
    m.update_state(real, fake)
    m.update_state(real, fake)
    np.testing.assert_allclose(m.result(), gm.wind_speed_weighted_rmse(real, fake).mean(), rtol=1e-6)
    m.reset_states()
    assert m.result() == 0.0
    d = gm.discriminator_score_real()
    d.update_state(np.array([[1.0], [3.0]]), np.array([[5.0], [7.0]]))
    assert d.result() == 2.0 and gm.discriminator_score_fake().name == "d_fake"
