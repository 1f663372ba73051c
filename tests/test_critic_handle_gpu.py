"""GPU tests of the `wdg_critic` handle (csrc/wdg_critic.cu; C ABI in include/wdg.h): the C++ graph walker against
(1) the same graph walked op by op from Python over the op-level C ABI -- bit for bit, every arithmetic mode -- and
(2) the float64 torch-autograd oracle for the checkpoint (shortcut) topology of SURVEY F6, which only the handle builds.
Tolerances: fp32 path relative L2 <= 1e-4 (score) / 2e-4 (gradients)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, np.float64)
    b = np.asarray(b.detach().cpu().numpy() if hasattr(b, "detach") else b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def data(B, T, S, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((B, T, S, S, 3)).astype(np.float32), rng.standard_normal((B, T, S, S, 2)).astype(np.float32))


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
@pytest.mark.parametrize("training", [True, False])
def test_handle_is_the_op_walk_bit_for_bit(prec, training):
    import torch
    from oracle.critic import synthetic_critic_weights
    from tests.critic_op_walk import OpWalkCritic
    from wind_downscaling_gan_b200.train import ops
    from wind_downscaling_gan_b200.train.nets import CriticNet, to_device
    B, T, S = 2, 3, 32
    lr, hr = data(B, T, S, 11)
    lr, hr = torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda()
    dw = synthetic_critic_weights(3, size=S)
    ds = torch.tensor([[0.7], [-1.3]], device="cuda")
    ops.set_precision(prec)
    try:
        wa, wb = to_device(dw), to_device(dw)
        a, b = OpWalkCritic(wa, S), CriticNet(wb, S)
        sa, sb = a.forward(lr, hr, training), b.forward(lr, hr, training)
        ga, dha = a.backward(ds, need_weight_grads=True, need_input_grad=True)
        gb, dhb = b.backward(ds, need_weight_grads=True, need_input_grad=True)
        # second call on the SAME variables: input gradient only (the gradient-penalty pass, ganbase.py:32-35)
        a2, b2 = OpWalkCritic(wa, S), CriticNet(wb, S)
        a2.forward(lr, hr, training)
        b2.forward(lr, hr, training)
        _, dha2 = a2.backward(ds, need_weight_grads=False, need_input_grad=True)
        g2, dhb2 = b2.backward(ds, need_weight_grads=False, need_input_grad=True)
        torch.cuda.synchronize()
    finally:
        ops.set_precision("fp32")
    assert torch.equal(sa, sb)
    assert set(ga) == set(gb) and g2 == {}
    for n in ga:
        assert torch.equal(ga[n], gb[n]), n
    assert torch.equal(dha, dhb) and torch.equal(dha2, dhb2)
    for n in wa:      # spectral-norm power iterations wrote the same variables back (twice in training mode)
        assert torch.equal(wa[n], wb[n]), n
    moved = not torch.equal(wa["layer_with_weights-2/layer/sn_u"].cpu(), torch.from_numpy(dw["layer_with_weights-2/layer/sn_u"]))
    assert moved == training


@pytest.mark.parametrize("B,T,S", [(2, 2, 32), (1, 2, 96)])
def test_checkpoint_topology_matches_oracle(B, T, S):
    """The graph revision the shipped discriminator checkpoint was written from (shortcut conv 6x6 / stride 11 / pad 4 +
    LayerNorm + add around the last 7x7 stage): score, every weight gradient and the input gradient vs torch autograd."""
    import torch
    from oracle import torch_train as tt
    from oracle.critic import critic_forward, critic_weight_shapes, synthetic_critic_weights
    from wind_downscaling_gan_b200.gan.models import make_discriminator
    from wind_downscaling_gan_b200.train.nets import CriticNet, to_device
    lr, hr = data(B, T, S, 5)
    dw = synthetic_critic_weights(9, size=S, ckpt_topology=True)
    assert "layer_with_weights-14/layer/kernel" in dw if S == 96 else True
    _, P = critic_weight_shapes(S, 3, 2, 16, True)
    # tf_utils.py:22-25 at 96 px: 9x9 -> 2x2 through a 6x6 kernel, stride 11, padding 4 (the checkpoint's w[6,6,128,256]);
    # at 32 px the source is 10x10: stride 12
    assert P["shortcut"] is not None and (P["shortcut"]["k"], P["shortcut"]["stride"], P["shortcut"]["pad"]) == ((6, 11, 4) if S == 96 else (6, 12, 4))
    # inference-mode forward through the Keras-like object against the numpy oracle
    d = make_discriminator(S, S, 3, 2, T, ckpt_topology=True)
    assert set(d.weight_names()) == set(dw)
    d.set_weights(dw)
    assert rel(d([lr, hr], training=False), critic_forward(dw, lr, hr, ckpt_topology=True)) < 1e-4
    # training-mode forward + both backward passes against autograd
    ref_w = {k: tt.T(v).clone() for k, v in dw.items()}
    hr_t = tt.T(hr).requires_grad_(True)
    s_ref, reads = tt.critic(ref_w, tt.T(lr), hr_t, training=True, P=P)
    ds = np.linspace(-1.3, 0.7, B).reshape(B, 1)
    names = tt.trainable(ref_w)
    gr = torch.autograd.grad((s_ref * tt.T(ds)).sum(), [reads[n] for n in names] + [hr_t])
    w = to_device(dw)
    net = CriticNet(w, S)
    assert net.h.ckpt_topology
    s = net.forward(torch.from_numpy(lr).cuda(), torch.from_numpy(hr).cuda(), training=True)
    assert rel(s, s_ref) < 1e-4
    g, dhr = net.backward(torch.from_numpy(ds.astype(np.float32)).cuda(), need_weight_grads=True, need_input_grad=True)
    assert set(g) == set(names)
    for n, r in zip(names, gr[:-1]):
        assert rel(g[n], r) < 2e-4, n
    assert rel(dhr, gr[-1]) < 2e-4
    for n in ref_w:   # in-place spectral normalisation of every wrapped layer, the shortcut conv included
        assert rel(w[n], ref_w[n]) < 1e-4, n


def test_train_step_with_checkpoint_topology_runs_and_moves_the_shortcut():
    from oracle.critic import synthetic_critic_weights
    from oracle.generator import synthetic_generator_weights
    from wind_downscaling_gan_b200.data.data_generator import FlexibleNoiseGenerator
    from wind_downscaling_gan_b200.gan import train
    from wind_downscaling_gan_b200.gan.ganbase import GAN
    from wind_downscaling_gan_b200.gan.models import make_discriminator, make_generator
    B, T, S = 2, 2, 32
    lr, hr = data(B, T, S, 6)
    gw, dw = synthetic_generator_weights(7), synthetic_critic_weights(8, size=S, ckpt_topology=True)
    gen, disc = make_generator(S, 3, 20, 2, T), make_discriminator(S, S, 3, 2, T, ckpt_topology=True)
    gen.set_weights(gw)
    disc.set_weights(dw)
    gan = GAN(gen, disc, FlexibleNoiseGenerator((B, T, S, S, 20), std=0.1, random_seed=3))
    gan.compile(generator_optimizer=train.generator_optimizer(), discriminator_optimizer=train.discriminator_optimizer(),
                discriminator_loss=train.discriminator_loss)
    m = gan.train_step((lr, hr))
    assert all(np.isfinite(v) for k, v in m.items() if v is not None)
    new = disc.get_weights()
    sc = [n for n in new if new[n].shape[:2] == (6, 6)]
    assert len(sc) == 1 and rel(new[sc[0]], dw[sc[0]]) > 1e-6


def test_handle_rejects_undersized_buffers():
    import ctypes as C
    import torch
    from wind_downscaling_gan_b200 import _lib
    from wind_downscaling_gan_b200.train.nets import CriticHandle
    h = CriticHandle.get(32, 3, 2, 16, False)
    B, T = 1, 2
    flat = torch.zeros(h.n_total, device="cuda")
    lr, hr = torch.zeros(B, T, 32, 32, 3, device="cuda"), torch.zeros(B, T, 32, 32, 2, device="cuda")
    score = torch.zeros(B, 1, device="cuda")
    nb, sb = C.c_size_t(), C.c_size_t()
    _lib.check(_lib.lib().wdg_critic_context_bytes(h.h, B, T, 1, C.byref(nb)))
    _lib.check(_lib.lib().wdg_critic_scratch_bytes(h.h, B, T, C.byref(sb)))
    ctx, sc = torch.empty(nb.value, dtype=torch.uint8, device="cuda"), torch.empty(sb.value, dtype=torch.uint8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = _lib.lib().wdg_critic_forward(h.h, p(flat), p(lr), p(hr), p(score), B, T, 1, p(ctx), nb.value - 256, p(sc), sb.value, None)
    assert rc != 0 and b"context too small" in _lib.lib().wdg_last_error()
    rc = _lib.lib().wdg_critic_backward(h.h, p(flat), p(ctx), B, T, 1, p(score), None, None, p(sc), sb.value, None)
    assert rc != 0 and b"nothing to compute" in _lib.lib().wdg_last_error()
