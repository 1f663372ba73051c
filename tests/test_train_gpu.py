"""GPU parity tests of the training path (critic forward/backward, training-mode generator, WGAN train_step) against
the float64 oracles (oracle/critic.py numpy, oracle/torch_train.py torch autograd).  The CUDA path is fp32:
tolerances are relative L2 <= 1e-4 for activations / gradients and 1e-3 for quantities after several updates."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if hasattr(b, "detach") else b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def data(B, T, S, seed):
    rng = np.random.default_rng(seed)
    lr = rng.standard_normal((B, T, S, S, 3)).astype(np.float32)
    hr = rng.standard_normal((B, T, S, S, 2)).astype(np.float32)
    return rng, lr, hr


@pytest.mark.parametrize("B,T,S", [(1, 2, 96), (2, 3, 32)])
def test_critic_forward_matches_oracle(B, T, S):
    from oracle.critic import critic_forward, synthetic_critic_weights
    from wind_downscaling_gan_b200.gan.models import make_discriminator
    _, lr, hr = data(B, T, S, 0)
    w = synthetic_critic_weights(1, size=S)
    d = make_discriminator(S, S, 3, 2, T)
    assert set(d.weight_names()) == set(w)
    d.set_weights(w)
    out = d([lr, hr], training=False)
    assert tuple(out.shape) == (B, 1)
    assert rel(out, critic_forward(w, lr, hr)) < 1e-4
    with pytest.raises(NotImplementedError):
        make_discriminator(96, 48, 3, 2, T)          # models.py:89-91


def test_generator_training_forward_backward():
    import torch
    from oracle import torch_train as tt
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.train.nets import GenNet, to_device
    B, T, S = 2, 2, 32
    rng, lr, _ = data(B, T, S, 2)
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    gw = synthetic_generator_weights(3)
    ref_w = {k: tt.T(v).clone() for k, v in gw.items()}
    out_ref, reads = tt.generator(ref_w, tt.T(lr), tt.T(noise), training=True)
    dout = rng.standard_normal(out_ref.shape)
    names = tt.trainable(ref_w)
    grads_ref = torch.autograd.grad((out_ref * tt.T(dout)).sum(), [reads[n] for n in names])
    w = to_device(gw)
    net = GenNet(w)
    out = net.forward(torch.from_numpy(lr).cuda(), torch.from_numpy(noise).cuda(), training=True)
    assert rel(out, out_ref) < 1e-4
    # spectral norm (in place) and BatchNorm moving statistics were updated like the reference's
    for k in ("layer_with_weights-0/layer/w", "layer_with_weights-7/layer/sn_u", "layer_with_weights-1/moving_mean",
              "layer_with_weights-10/moving_variance"):
        assert rel(w[k], ref_w[k]) < 1e-4, k
    grads = net.backward(torch.from_numpy(dout.astype(np.float32)).cuda())
    assert set(grads) == set(names)
    for n, gr in zip(names, grads_ref):
        assert rel(grads[n], gr) < 2e-4, n


def test_critic_backward_weights_and_input():
    import torch
    from oracle import torch_train as tt
    from oracle.critic import synthetic_critic_weights
    from wind_downscaling_gan_b200.train.nets import CriticNet, to_device
    B, T, S = 2, 2, 32
    _, lr, hr = data(B, T, S, 4)
    dw = synthetic_critic_weights(5, size=S)
    ref_w = {k: tt.T(v).clone() for k, v in dw.items()}
    hr_t = tt.T(hr).requires_grad_(True)
    s_ref, reads = tt.critic(ref_w, tt.T(lr), hr_t, training=True)
    ds = np.array([[0.7], [-1.3]])
    names = tt.trainable(ref_w)
    gr = torch.autograd.grad((s_ref * tt.T(ds)).sum(), [reads[n] for n in names] + [hr_t])
    w = to_device(dw)
    net = CriticNet(w, S)
    s = net.forward(torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda(), training=True)
    assert rel(s, s_ref) < 1e-4
    g, dhr = net.backward(torch.from_numpy(ds.astype(np.float32)).cuda(), need_weight_grads=True, need_input_grad=True)
    assert set(g) == set(names)
    for n, r in zip(names, gr[:-1]):
        assert rel(g[n], r) < 2e-4, n
    assert rel(dhr, gr[-1]) < 2e-4


def test_train_step_matches_oracle():
    """One full WGAN step (3 critic updates + 1 generator update + metric recompute) with shared random draws."""
    from oracle import torch_train as tt
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    B, T, S = 2, 2, 32
    rng, lr, hr = data(B, T, S, 6)
    gw, dw = synthetic_generator_weights(7), synthetic_critic_weights(8, size=S)
    draws = []
    for _ in range(3):
        draws += [0.1 * rng.standard_normal((B, T, S, S, 20)), rng.uniform(0, 1, (B,)),
                  0.1 * rng.standard_normal((B, T, S, S, 2)), 0.1 * rng.standard_normal((B, T, S, S, 2))]
    draws += [0.1 * rng.standard_normal((B, T, S, S, 20)), 0.1 * rng.standard_normal((B, T, S, S, 20))]
    draws = [np.asarray(d, np.float32) for d in draws]
    st = tt.State(gw, dw)
    m_ref = tt.train_step(st, lr, hr, draws)
    gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
    gen.set_weights(gw)
    disc.set_weights(dw)
    gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1))
    gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                discriminator_loss=train.discriminator_loss)
    m = gan.train_step((lr, hr), draws=draws)
    for k in ("g_loss", "g_disc_loss", "d_loss", "d_gradient_pen", "g_gradient_param", "d_gradient_param"):
        assert abs(m[k] - m_ref[k]) <= 2e-3 * max(1.0, abs(m_ref[k])), (k, m[k], m_ref[k])
    gan.sync_weights()
    new_g, new_d = gen.get_weights(), disc.get_weights()
    for k, v in st.g.items():
        assert rel(new_g[k], v) < 1e-3, k
    for k, v in st.d.items():
        assert rel(new_d[k], v) < 1e-3, k
    # the weights really moved (Adam + spectral normalisation)
    assert rel(new_g["layer_with_weights-4/cell/kernel"], gw["layer_with_weights-4/cell/kernel"]) > 1e-5
    t = gan.test_step((lr, hr), draws=[draws[-1]])
    assert set(t) >= {"loss"} and np.isfinite(t["loss"])


def test_train_step_s96_t24_matches_golden():
    """The reference's real training shape (96 px, 24 timesteps, api.py:22-23): one full WGAN step on the fp32 CUDA
    path against tests/golden/train_step_s96_t24.npz (float64 torch-autograd oracle, make_train_golden.py): the 98 ->
    31 -> 9 -> 2 critic pyramid backward and 24-step BPTT of both ConvLSTMs included.  Per tensor: norm of the updated
    weights, norm of their change over the step and 8 seeded projections of that change.  Tolerance on the change:
    3x the distance at which the oracle's own float32 run lands from its float64 run (`floor` in the fixture: ~2e-2
    for the generator, ~2e-3 for the critic -- fp32 rounding amplified by the piecewise-linear units), at least 2e-3."""
    import os
    from tests.golden.make_train_golden import case, projections
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_step_s96_t24.npz"))
    B, T, S = (int(v) for v in z["meta"][:3])
    assert (T, S) == (24, 96)
    lr, hr, draws, gw, dw = case(B, T, S)
    gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
    gen.set_weights(gw)
    disc.set_weights(dw)
    gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1))
    gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                discriminator_loss=train.discriminator_loss)
    m = gan.train_step((lr, hr), draws=draws)
    for k in ("g_loss", "g_disc_loss", "d_loss", "d_gradient_pen", "g_gradient_param", "d_gradient_param", "d_real", "d_fake"):
        ref = float(z["metric/" + k])
        assert abs(m[k] - ref) <= 2e-3 * max(1.0, abs(ref)), (k, m[k], ref)
    gan.sync_weights()
    worst = 0.0
    for prefix, old, new in (("g", gw, gen.get_weights()), ("d", dw, disc.get_weights())):
        for k in sorted(new):
            a, b = np.asarray(old[k], np.float64), np.asarray(new[k], np.float64)
            norm, dnorm = float(z[f"{prefix}/{k}/norm"]), float(z[f"{prefix}/{k}/dnorm"])
            assert abs(np.linalg.norm(b) - norm) <= 1e-4 * norm + 1e-7, (prefix, k)
            # a difference e between the two changes moves each projection by ~|e| (max of 8 Gaussians: ~2|e|)
            slack = max(3 * float(z[f"{prefix}/{k}/floor"]), 2e-3) * dnorm + 2e-7 * norm
            err = np.abs(projections(f"{prefix}/{k}", b - a) - z[f"{prefix}/{k}/proj"]).max()
            worst = max(worst, err / max(slack, 1e-30))
            assert err <= 2 * slack, (prefix, k, err, slack)
    print("worst projection error / slack:", worst)


def test_graphed_train_step_equals_eager_and_syncs_weights():
    """GAN.train_step captures the step into one CUDA graph after two eager calls.  Two GANs with identical weights and
    noise seeds -- one replaying the graph, one forced eager -- stay bit-identical over 5 steps (same kernels, same
    device-resident random stream, same Adam clocks); the handles see the trained weights without an explicit sync, and
    writing weights into a handle restarts the optimizer state."""
    import torch
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    B, T, S = 2, 2, 32
    rng, lr, hr = data(B, T, S, 9)
    gw, dw = synthetic_generator_weights(7), synthetic_critic_weights(8, size=S)

    def build(graph):
        gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
        gen.set_weights(gw)
        disc.set_weights(dw)
        gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=5))
        gan.use_cuda_graph = graph
        gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                    discriminator_loss=train.discriminator_loss)
        return gan

    a, b = build(True), build(False)
    for step in range(5):
        ma, mb = a.train_step((lr, hr)), b.train_step((lr, hr))
        assert ma == mb, (step, ma, mb)
    assert a._graphed.graph is not None and b._graphed is None
    assert a.noise_generator._offset == b.noise_generator._offset
    assert a.generator.optimizer.iterations == b.generator.optimizer.iterations == 5
    assert a.discriminator.optimizer.iterations == 15
    wa, wb = a.generator.get_weights(), b.generator.get_weights()          # no sync_weights() call: pulled in lazily
    assert all(np.array_equal(wa[k], wb[k]) for k in wa)
    assert not np.array_equal(wa["layer_with_weights-4/cell/kernel"], gw["layer_with_weights-4/cell/kernel"])
    # inference through the handle runs the TRAINED weights
    noise = (0.1 * rng.standard_normal((B, T, S, S, 20))).astype(np.float32)
    out_a = a.generator.predict([lr, noise])
    fresh = make_generator(S, 3, 20, 2, T)
    fresh.set_weights(wa)
    assert np.array_equal(out_a, fresh.predict([lr, noise]))
    # writing weights into a handle invalidates the training state (fresh Adam slots, clocks back to 0)
    a.generator.set_weights(gw)
    assert a._train is None and a.generator.optimizer.iterations == 0
    m = a.train_step((lr, hr))
    assert np.isfinite(m["d_loss"]) and a.generator.optimizer.iterations == 1


def test_skipping_the_dead_gradient_penalty_passes_changes_nothing():
    """GAN(..., skip_dead_gradient_penalty=True) drops the interpolate -> critic -> input-gradient passes of all but the last
    critic iteration (their result reaches neither the weights nor the log): metrics and weights bit-identical."""
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    B, T, S = 2, 2, 32
    _, lr, hr = data(B, T, S, 10)
    gw, dw = synthetic_generator_weights(7), synthetic_critic_weights(8, size=S)
    out = []
    for skip in (False, True):
        gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
        gen.set_weights(gw)
        disc.set_weights(dw)
        gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=3), skip_dead_gradient_penalty=skip)
        gan.use_cuda_graph = False
        gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                    discriminator_loss=train.discriminator_loss)
        ms = [gan.train_step((lr, hr)) for _ in range(2)]
        out.append((ms, gen.get_weights(), disc.get_weights(), gan.launches_per_step()))
    (ma, ga, da, la), (mb, gb, db, lb) = out
    assert ma == mb
    assert all(np.array_equal(ga[k], gb[k]) for k in ga) and all(np.array_equal(da[k], db[k]) for k in da)
    assert lb < la


def test_train_step_updates_the_compiled_metrics():
    """ganbase.py:71-72, :82-93 with the metric set of api.py:76-84: every step updates the five generator metrics with
    (high_res, recomputed fake) and the two score metrics with (D(real), D(fake)); the dict carries their running means as
    g_<name> / <name>.  Checked against oracle/metrics.py on the recomputed tensors of each step."""
    from oracle import metrics as om
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import metrics, train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    B, T, S = 2, 2, 32
    gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T)
    gen.set_weights(synthetic_generator_weights(7))
    disc.set_weights(synthetic_critic_weights(8, size=S))
    gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=1))
    gan.use_cuda_graph = False
    gan.compile(generator_optimizer=train.generator_optimizer(),
                generator_metrics=[metrics.AngularCosineDistance(), metrics.LogSpectralDistance(), metrics.WeightedRMSEForExtremes(),
                                   metrics.WindSpeedWeightedRMSE(), metrics.SpatialKS()],
                discriminator_optimizer=train.discriminator_optimizer(), discriminator_loss=train.discriminator_loss,
                metrics=[metrics.discriminator_score_fake(), metrics.discriminator_score_real()])
    fns = {"g_acd": om.angular_cosine_distance, "g_lsd": om.log_spectral_distance, "g_extreme_rmse": om.extreme_weighted_rmse,
           "g_ws_weighted_rmse": om.wind_speed_weighted_rmse, "g_spatial_ks": om.spatially_convolved_ks_stat}
    acc = {k: [] for k in list(fns) + ["d_real", "d_fake"]}
    for step in range(2):
        _, lr, hr = data(B, T, S, 20 + step)
        m = gan.train_step((lr, hr))
        fake, s_real, s_fake = (t.cpu().numpy() for t in gan._last_recompute)
        for k, fn in fns.items():
            acc[k] += list(np.asarray(fn(hr, fake), np.float64).ravel())
        acc["d_real"] += list(s_real.ravel())
        acc["d_fake"] += list(s_fake.ravel())
        assert set(m) >= {"g_loss", "g_disc_loss", "g_reco_loss", "d_loss", "d_gradient_pen", "g_gradient_param", "d_gradient_param"} | set(acc)
        for k, v in acc.items():
            assert abs(m[k] - np.mean(v)) <= 1e-4 * max(1.0, abs(np.mean(v))), (step, k, m[k], np.mean(v))
    t = gan.test_step((lr, hr))
    assert {"loss", "d_real", "d_fake"} <= set(t)
    with pytest.raises(NotImplementedError):
        gan.train_step((lr, hr, np.ones(B, np.float32)))
