"""WGAN training / evaluation step of the reference (`gan/ganbase.py:21-113`) on the CUDA training kernels.

Nothing in `train_step` synchronises with the host: the random draws come from device-resident Philox state, the Adam
step counters / bias-corrected learning rates live in device memory, and every logged scalar is reduced on the device
into one small buffer that is read back once, after the step.  That makes the whole step capturable as ONE CUDA graph
(`GraphedStep`): ~5 000 kernel launches replayed without their Python / ctypes / tensor-map-encoding host cost."""
import numpy as np
import torch

from . import ops
from .nets import CriticNet, GenNet, to_device, trainable_names

METRIC_KEYS = ("g_loss", "g_disc_loss", "d_loss", "d_gradient_pen", "g_gradient_param", "d_gradient_param", "d_real",
               "d_fake", "d_gradient_reg")


def _dev(x):
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=torch.float32).contiguous()
    return torch.as_tensor(np.asarray(x, np.float32)).cuda().contiguous()


class TrainState:
    """Device-side fp32 weights, Adam slots and step counters of both models."""

    def __init__(self, generator, discriminator, g_opt, d_opt):
        self.generator, self.discriminator = generator, discriminator
        self.g = to_device(generator.get_weights())
        # critic: ONE flat variable buffer laid out by the wdg_critic handle (trainable variables first), so its Adam
        # step is one launch and its gradient all-reduce one collective over a contiguous range
        self.d = discriminator._handle().pack(discriminator.get_weights())
        self.g_opt, self.d_opt = g_opt, d_opt
        self.g_slots = {n: (torch.zeros_like(self.g[n]), torch.zeros_like(self.g[n])) for n in trainable_names(self.g)}
        nt = self.d.handle.n_train
        self.d_m, self.d_v = torch.zeros(nt, dtype=torch.float32, device="cuda"), torch.zeros(nt, dtype=torch.float32, device="cuda")
        self.size = generator.image_size
        # device-resident optimizer clocks: [0] generator, [1] critic
        self.steps = torch.tensor([g_opt.iterations, d_opt.iterations], dtype=torch.int32, device="cuda")
        self.lr_t = torch.zeros(2, dtype=torch.float32, device="cuda")
        self.dirty = False      # the model handles hold older weights than this state

    def push_weights(self):
        self.generator.set_weights({k: v.cpu().numpy() for k, v in self.g.items()}, _from_state=True)
        self.discriminator.set_weights({k: v.cpu().numpy() for k, v in self.d.items()}, _from_state=True)
        self.dirty = False


def adam_apply(st, which, weights, slots, grads, opt):
    """Keras Adam (train.py:35,58): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); w -= lr_t * m / (sqrt(v) + eps).  t and lr_t are
    device scalars (graph capturable); `opt.iterations` mirrors t on the host."""
    opt.iterations += 1
    ops.adam_lr(st.lr_t[which:which + 1], st.steps[which:which + 1], opt.lr, opt.beta_1, opt.beta_2)
    for n, g in grads.items():
        m, v = slots[n]
        ops.adam_dev(weights[n], m, v, g, st.lr_t[which:which + 1], opt.beta_1, opt.beta_2, opt.epsilon)


def adam_apply_flat(st, which, w_flat, m_flat, v_flat, g_flat, opt):
    """Same update over a whole model held in one flat buffer: one launch (padding between variables stays 0)."""
    opt.iterations += 1
    ops.adam_lr(st.lr_t[which:which + 1], st.steps[which:which + 1], opt.lr, opt.beta_1, opt.beta_2)
    ops.adam_dev(w_flat[:g_flat.numel()], m_flat, v_flat, g_flat, st.lr_t[which:which + 1], opt.beta_1, opt.beta_2, opt.epsilon)


def _mean_sq_into(grads, out):
    """out[0] = mean over tensors of mean(g^2) (ganbase.py:79-81), reduced on the device."""
    per = ops.empty(len(grads))
    for i, g in enumerate(grads.values()):
        ops.reduce_into(per[i:i + 1], g, 2, scale=1.0 / g.numel())
    ops.reduce_into(out, per, 0, scale=1.0 / len(grads))


def train_step_device(st, low_res, high_res, noise_generator, n_critic=3, draws=None, gamma=100.0, comm=None,
                      skip_dead_gp=False):
    """One WGAN step on this rank's share of the batch; returns the metrics as a device tensor (order METRIC_KEYS) and the
    tensors of the inference-mode recompute (ganbase.py:64-66: generated images, critic scores of real and generated) that
    the compiled metric objects are updated with (:71-72).
    With a communicator (train/dist.py) the gradients are all-reduced after every backward pass and BatchNorm
    statistics are synchronised, so `world` ranks x local batch reproduce the reference's single process at the global
    batch.

    skip_dead_gp (extension, OFF by default): the gradient penalty never reaches the weights (ganbase.py:32-37 computes it
    outside `disc_tape`) and only the LAST critic iteration's norms are logged (:85-88), so the interpolate ->
    critic forward -> backward-to-input passes of the earlier iterations influence nothing but the critic's spectral-norm
    state.  With the flag those iterations run just that power iteration (and still draw eps, so the random stream is
    unchanged): weights and metrics come out bit-identical, two critic forward + input-backward passes cheaper."""
    ops.use_current_stream()
    B = low_res.shape[0]
    world = comm.world if comm is not None else 1
    Bg = B * world                                    # global batch: the loss means run over it
    it = iter(draws) if draws is not None else None
    M = ops.zeros(len(METRIC_KEYS))
    slot = {k: M[i:i + 1] for i, k in enumerate(METRIC_KEYS)}

    def noise(channels=None):
        if it is not None:
            return _dev(next(it))
        return noise_generator(B) if channels is None else noise_generator(B, channels=channels)

    def uniform():
        if it is not None:
            return _dev(next(it)).reshape(B)
        if hasattr(noise_generator, "uniform"):
            return noise_generator.uniform(B)      # the library's device-resident Philox stream (graph capturable)
        return torch.rand(B, device="cuda", dtype=torch.float32)

    gen = GenNet(st.g, comm)
    out_ch = high_res.shape[-1]

    def const(v):
        return torch.full((B, 1), float(v), device="cuda", dtype=torch.float32)

    ones = const(1.0)
    norms = None
    for ci in range(n_critic):                                                       # ganbase.py:26
        fake = gen.forward(low_res, noise(), training=True, keep_context=False)       # :28-29 (no backward through it)
        eps = uniform()                                                               # :30
        if skip_dead_gp and ci < n_critic - 1:
            CriticNet(st.d, st.size).spectral_norm_step()                             # the only live effect of :32-35 here
        else:
            combined = torch.empty_like(high_res)
            ops.lerp_batch(combined, high_res, fake, eps)                             # :31
            d_gp = CriticNet(st.d, st.size)
            d_gp.forward(low_res, combined, training=True)                            # :32-34
            _, g_img = d_gp.backward(ones, need_weight_grads=False, need_input_grad=True)   # :35
            norms = ops.empty(B, out_ch)
            ops.gp_norm(g_img, norms)                                                 # :36 (reduced over T, H, W only)
            del d_gp, g_img, combined
        hr_n = torch.empty_like(high_res)
        ops.axpby(ops.full(hr_n), ops.full(high_res), 1.0, ops.full(noise(out_ch)), 1.0)   # :40
        d_real = CriticNet(st.d, st.size)
        d_real.forward(low_res, hr_n, training=True)                                  # :41
        fhr = torch.empty_like(fake)
        ops.axpby(ops.full(fhr), ops.full(fake), 1.0, ops.full(noise(out_ch)), 1.0)   # :42
        d_fake = CriticNet(st.d, st.size)
        d_fake.forward(low_res, fhr, training=True)                                   # :43
        # d_loss = -(mean(real) - mean(fake)) + gradient_reg                          # :44-45, train.py:11-12
        g1, _ = d_real.backward(const(-1.0 / Bg))
        g2, _ = d_fake.backward(const(1.0 / Bg))
        f1, f2 = g1.flat.view(-1, 64), g2.flat.view(-1, 64)                           # both backward passes, every variable
        ops.axpby(ops.full(f1), ops.full(f1), 1.0, ops.full(f2), 1.0)
        d_grads = g1
        if comm is not None:
            comm.allreduce_grads(d_grads)
        adam_apply_flat(st, 1, st.d.flat, st.d_m, st.d_v, d_grads.flat, st.d_opt)     # :46-47
    # gradient penalty of the LAST critic iteration (:36-37; a constant for the weights): gamma * mean((norm - 1)^2)
    nm1 = ops.empty(B, out_ch)
    ops.axpby(ops.full(nm1), ops.full(norms), 1.0, ops.full(torch.ones_like(norms)), -1.0)
    ops.reduce_into(slot["d_gradient_reg"], nm1, 2, scale=gamma / norms.numel())
    ops.reduce_into(slot["d_gradient_pen"], norms, 0, scale=1.0 / norms.numel())
    fake = gen.forward(low_res, noise(), training=True)                               # :51-52
    d_g = CriticNet(st.d, st.size)
    score = d_g.forward(low_res, fake, training=True)                                 # :53
    ops.reduce_into(slot["g_disc_loss"], score, 0, scale=-1.0 / score.numel())        # :54
    _, dfake = d_g.backward(const(-1.0 / Bg), need_weight_grads=False, need_input_grad=True)
    g_grads = gen.backward(dfake)                                                     # :60
    if comm is not None:   # BatchNorm gamma/beta gradients are already global sums (synchronised BN)
        comm.allreduce_grads(g_grads, skip=[n for n in g_grads if n.endswith(("gamma", "beta"))])
    adam_apply(st, 0, st.g, st.g_slots, g_grads, st.g_opt)                            # :61
    # metric recompute, inference mode                                               # :64-68
    d_eval = CriticNet(st.d, st.size)
    s_real = d_eval.forward(low_res, high_res, training=False)
    fake_m = gen.forward(low_res, noise(), training=False, keep_context=False)
    s_fake = d_eval.forward(low_res, fake_m, training=False)
    ops.reduce_into(slot["d_real"], s_real, 0, scale=1.0 / s_real.numel())
    ops.reduce_into(slot["d_fake"], s_fake, 0, scale=1.0 / s_fake.numel())
    ops.axpby(ops.full(slot["g_loss"]), ops.full(slot["d_fake"]), -1.0)               # g_loss = -mean(D(G))
    ops.axpby(ops.full(slot["d_loss"]), ops.full(slot["d_fake"]), 1.0, ops.full(slot["d_real"]), -1.0)   # -(real - fake)
    _mean_sq_into(g_grads, slot["g_gradient_param"])
    _mean_sq_into(d_grads, slot["d_gradient_param"])
    if comm is not None and comm.world > 1:      # the logged scalars are means over the GLOBAL batch / all replicas
        comm.allreduce_sum(M)
        ops.axpby(ops.full(M), ops.full(M), 1.0 / comm.world)
    st.dirty = True
    return M, (fake_m, s_real, s_fake)


def metrics_dict(M):
    vals = M.detach().cpu().numpy().astype(np.float64)
    out = {k: float(v) for k, v in zip(METRIC_KEYS, vals)}
    out["g_reco_loss"] = None
    return out


def train_step(st, low_res, high_res, noise_generator, n_critic=3, draws=None, gamma=100.0, comm=None, skip_dead_gp=False):
    """Eager form: runs the step and reads the metrics back (one device -> host copy).  Returns (dict, recompute tensors)."""
    M, extras = train_step_device(st, _dev(low_res), _dev(high_res), noise_generator, n_critic, draws, gamma, comm, skip_dead_gp)
    return metrics_dict(M), extras


class GraphedStep:
    """`train_step` captured into one CUDA graph for a fixed (batch, T, S) and replayed: the inputs are copied into
    static device buffers, the graph is launched, and the metric buffer is read back.  The first `WARMUP` calls run
    eagerly (they size every scratch buffer and the library's packed-weight arena; nothing may allocate with cudaMalloc
    while capturing)."""
    WARMUP = 2

    def __init__(self, st, noise_generator, n_critic, comm=None, skip_dead_gp=False):
        self.st, self.ng, self.n_critic, self.comm, self.skip_dead_gp = st, noise_generator, n_critic, comm, skip_dead_gp
        self.calls = 0
        self.graph = None
        self.shape = None
        self.failed = False

    def _capture(self, low_res, high_res):
        self.lr_buf, self.hr_buf = low_res.clone(), high_res.clone()
        off0 = self.ng._offset
        it0 = (self.st.g_opt.iterations, self.st.d_opt.iterations)
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            self.M, self.extras = train_step_device(self.st, self.lr_buf, self.hr_buf, self.ng, self.n_critic, None, comm=self.comm,
                                                    skip_dead_gp=self.skip_dead_gp)
        ops.use_current_stream()
        # capturing does not execute: undo the host mirrors it advanced, replays re-apply them
        self.blocks_per_step = self.ng._offset - off0
        self.ng._offset = off0
        self.iters_per_step = (self.st.g_opt.iterations - it0[0], self.st.d_opt.iterations - it0[1])
        self.st.g_opt.iterations, self.st.d_opt.iterations = it0
        self.graph = g
        self.shape = (tuple(low_res.shape), tuple(high_res.shape))

    def __call__(self, low_res, high_res):
        low_res, high_res = _dev(low_res), _dev(high_res)
        self.calls += 1
        shape = (tuple(low_res.shape), tuple(high_res.shape))
        if self.failed or self.calls <= self.WARMUP or (self.graph is not None and shape != self.shape):
            M, extras = train_step_device(self.st, low_res, high_res, self.ng, self.n_critic, None, comm=self.comm, skip_dead_gp=self.skip_dead_gp)
            return metrics_dict(M), extras
        if self.graph is None:
            try:
                self._capture(low_res, high_res)
            except Exception as exc:      # capture not possible in this environment: stay eager, say so once
                import warnings
                warnings.warn(f"CUDA-graph capture of train_step failed ({exc!r}); running eagerly")
                self.failed = True
                torch.cuda.synchronize()
                ops.use_current_stream()
                M, extras = train_step_device(self.st, low_res, high_res, self.ng, self.n_critic, None, comm=self.comm, skip_dead_gp=self.skip_dead_gp)
                return metrics_dict(M), extras
        self.lr_buf.copy_(low_res, non_blocking=True)
        self.hr_buf.copy_(high_res, non_blocking=True)
        self.graph.replay()
        self.ng._offset += self.blocks_per_step
        self.st.g_opt.iterations += self.iters_per_step[0]
        self.st.d_opt.iterations += self.iters_per_step[1]
        self.st.dirty = True
        return metrics_dict(self.M), self.extras


def test_step(st, x, y, noise_generator, draws=None):
    ops.use_current_stream()
    x, y = _dev(x), _dev(y)
    B = x.shape[0]
    nz = _dev(draws[0]) if draws is not None else noise_generator(B)
    d = CriticNet(st.d, st.size)
    s_real = d.forward(x, y, training=False)
    fake = GenNet(st.g).forward(x, nz, training=False, keep_context=False)
    s_fake = d.forward(x, fake, training=False)
    real_m = float(ops.reduce(s_real, 0, scale=1.0 / B).item())
    fake_m = float(ops.reduce(s_fake, 0, scale=1.0 / B).item())
    return {"loss": -(real_m - fake_m), "d_real": real_m, "d_fake": fake_m}
