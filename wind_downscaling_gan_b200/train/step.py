"""WGAN training / evaluation step of the reference (`gan/ganbase.py:21-113`) on the fp32 CUDA training kernels."""
import numpy as np
import torch

from . import ops
from .nets import CriticNet, GenNet, to_device, trainable_names


def _dev(x):
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=torch.float32).contiguous()
    return torch.as_tensor(np.asarray(x, np.float32)).cuda().contiguous()


class TrainState:
    """Device-side fp32 weights, Adam slots and step counters of both models."""

    def __init__(self, generator, discriminator, g_opt, d_opt):
        self.generator, self.discriminator = generator, discriminator
        self.g = to_device(generator.get_weights())
        self.d = to_device(discriminator.get_weights())
        self.g_opt, self.d_opt = g_opt, d_opt
        self.g_slots = {n: (torch.zeros_like(self.g[n]), torch.zeros_like(self.g[n])) for n in trainable_names(self.g)}
        self.d_slots = {n: (torch.zeros_like(self.d[n]), torch.zeros_like(self.d[n])) for n in trainable_names(self.d)}
        self.size = generator.image_size

    def push_weights(self):
        self.generator.set_weights({k: v.cpu().numpy() for k, v in self.g.items()})
        self.discriminator.set_weights({k: v.cpu().numpy() for k, v in self.d.items()})


def adam_apply(weights, slots, grads, opt):
    """Keras Adam (train.py:35,58): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t * m / (sqrt(v) + eps)."""
    opt.iterations += 1
    t = opt.iterations
    lr_t = opt.lr * np.sqrt(1.0 - opt.beta_2 ** t) / (1.0 - opt.beta_1 ** t)
    for n, g in grads.items():
        m, v = slots[n]
        ops.adam(weights[n], m, v, g, float(lr_t), opt.beta_1, opt.beta_2, opt.epsilon)


def _mean_sq(grads):
    return float(np.mean([float(ops.reduce(g, 2, scale=1.0 / g.numel()).item()) for g in grads.values()]))


def train_step(st, low_res, high_res, noise_generator, n_critic=3, draws=None, gamma=100.0, comm=None):
    """One WGAN step on this rank's share of the batch.  With a communicator (train/dist.py) the gradients are
    all-reduced after every backward pass and BatchNorm statistics are synchronised, so `world` ranks x local batch
    reproduce the reference's single process at the global batch."""
    ops.use_current_stream()
    low_res, high_res = _dev(low_res), _dev(high_res)
    B = low_res.shape[0]
    world = comm.world if comm is not None else 1
    Bg = B * world                                    # global batch: the loss means run over it
    it = iter(draws) if draws is not None else None

    def noise(channels=None):
        if it is not None:
            return _dev(next(it))
        return noise_generator(B) if channels is None else noise_generator(B, channels=channels)

    def uniform():
        if it is not None:
            return _dev(next(it)).reshape(B)
        return torch.rand(B, device="cuda", dtype=torch.float32)

    gen = GenNet(st.g, comm)
    out_ch = high_res.shape[-1]
    def const(v):
        return torch.full((B, 1), float(v), device="cuda", dtype=torch.float32)

    def mean(t):
        return float(ops.reduce(t, 0, scale=1.0 / t.numel()).item())

    ones = const(1.0)
    for _ in range(n_critic):                                                        # ganbase.py:26
        fake = gen.forward(low_res, noise(), training=True)                           # :28-29
        eps = uniform()                                                               # :30
        combined = torch.empty_like(high_res)
        ops.lerp_batch(combined, high_res, fake, eps)                                 # :31
        d_gp = CriticNet(st.d, st.size)
        d_gp.forward(low_res, combined, training=True)                                # :32-34
        _, g_img = d_gp.backward(ones, need_weight_grads=False, need_input_grad=True) # :35
        norms = ops.empty(B, out_ch)
        ops.gp_norm(g_img, norms)                                                     # :36 (reduced over T, H, W only)
        nrm = norms.cpu().numpy().astype(np.float64)
        gradient_reg = gamma * np.mean((nrm - 1.0) ** 2)                              # :37 (a constant for the weights)
        hr_n = torch.empty_like(high_res)
        ops.axpby(ops.full(hr_n), ops.full(high_res), 1.0, ops.full(noise(out_ch)), 1.0)   # :40
        d_real = CriticNet(st.d, st.size)
        s_real = d_real.forward(low_res, hr_n, training=True)                         # :41
        fhr = torch.empty_like(fake)
        ops.axpby(ops.full(fhr), ops.full(fake), 1.0, ops.full(noise(out_ch)), 1.0)   # :42
        d_fake = CriticNet(st.d, st.size)
        s_fake = d_fake.forward(low_res, fhr, training=True)                          # :43
        # d_loss = -(mean(real) - mean(fake)) + gradient_reg                          # :44-45, train.py:11-12
        g1, _ = d_real.backward(const(-1.0 / Bg))
        g2, _ = d_fake.backward(const(1.0 / Bg))
        d_grads = {}
        for n in g1:
            ops.axpby(ops.full(g1[n]), ops.full(g1[n]), 1.0, ops.full(g2[n]), 1.0)
            d_grads[n] = g1[n]
        if comm is not None:
            comm.allreduce_grads(d_grads)
        adam_apply(st.d, st.d_slots, d_grads, st.d_opt)                               # :46-47
    fake = gen.forward(low_res, noise(), training=True)                               # :51-52
    d_g = CriticNet(st.d, st.size)
    score = d_g.forward(low_res, fake, training=True)                                 # :53
    gen_disc_loss = -mean(score)                                       # :54
    _, dfake = d_g.backward(const(-1.0 / Bg), need_weight_grads=False, need_input_grad=True)
    g_grads = gen.backward(dfake)                                                     # :60
    if comm is not None:   # BatchNorm gamma/beta gradients are already global sums (synchronised BN)
        comm.allreduce_grads(g_grads, skip=[n for n in g_grads if n.endswith(("gamma", "beta"))])
    adam_apply(st.g, st.g_slots, g_grads, st.g_opt)                                   # :61
    # metric recompute, inference mode                                               # :64-68
    d_eval = CriticNet(st.d, st.size)
    s_real = d_eval.forward(low_res, high_res, training=False)
    fake_m = gen.forward(low_res, noise(), training=False)
    s_fake = d_eval.forward(low_res, fake_m, training=False)
    real_m, fake_mm = mean(s_real), mean(s_fake)
    return {"g_loss": -fake_mm, "g_disc_loss": gen_disc_loss, "g_reco_loss": None, "d_loss": -(real_m - fake_mm),
            "d_gradient_pen": float(nrm.mean()), "g_gradient_param": _mean_sq(g_grads), "d_gradient_param": _mean_sq(d_grads),
            "d_real": real_m, "d_fake": fake_mm, "d_gradient_reg": float(gradient_reg)}


def test_step(st, x, y, noise_generator, draws=None):
    ops.use_current_stream()
    x, y = _dev(x), _dev(y)
    B = x.shape[0]
    nz = _dev(draws[0]) if draws is not None else noise_generator(B)
    d = CriticNet(st.d, st.size)
    s_real = d.forward(x, y, training=False)
    fake = GenNet(st.g).forward(x, nz, training=False)
    s_fake = d.forward(x, fake, training=False)
    real_m = float(ops.reduce(s_real, 0, scale=1.0 / B).item())
    fake_m = float(ops.reduce(s_fake, 0, scale=1.0 / B).item())
    return {"loss": -(real_m - fake_m), "d_real": real_m, "d_fake": fake_m}
