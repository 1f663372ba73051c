"""Torch-CPU (float64, autograd) restatement of the WGAN training step of the reference
(`gan/ganbase.py:21-94`, `gan/train.py`, models.py in training mode, TFA SpectralNormalization,
Keras BatchNormalization / Adam semantics of SURVEY.md §8(c)).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED.

Random draws are not generated here: the caller passes the list `draws` of arrays consumed in the
order the reference consumes its generator (per critic iteration: G noise, eps, noise on real, noise on
fake; then G noise for the generator update; then G noise for the metric recompute).
"""
import numpy as np
import torch
import torch.nn.functional as Fn

from .critic import critic_weight_shapes

LW = "layer_with_weights-%d/"
DT = torch.float64
# Operand rounding of the convolution GEMMs, to check the tensor-core training kernels decision-for-decision:
# None (exact), "tf32" (fp32 -> 10-bit mantissa, round to nearest, ties away: PTX cvt.rna.tf32.f32) or "bf16"
# (fp32 -> bf16, round to nearest even).  Every GEMM of the product path rounds BOTH operands: the forward
# (x, w), the backward-data (dy, w) and the backward-weight (x, dy) one; accumulation stays exact here.
OPERAND = None
# Which GEMMs round: None = all of them, or a predicate (kind, ci, co, k, stride) -> bool with kind in
# {"fwd", "bwd_data", "bwd_weight"} (of the underlying convolution), ci/co its input/output channels, k its kernel size.
# The product keeps a few narrow 3x3 layers on exact-fp32 CUDA-core kernels (csrc/train_ops.cu: direct_ok); a test that
# wants decision-for-decision agreement passes that map here.
OPERAND_POLICY = None


def _rounds(kind, ci, co, k, stride):
    return OPERAND is not None and (OPERAND_POLICY is None or bool(OPERAND_POLICY(kind, ci, co, k, stride)))


def rnd(x, on=True):
    if OPERAND is None or not on:
        return x
    x32 = x.detach().to(torch.float32)
    if OPERAND == "bf16":
        return x32.to(torch.bfloat16).to(DT)
    bits = x32.contiguous().view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF          # sign-magnitude: rounds the magnitude, ties away from zero
    return bits.view(torch.float32).to(DT)


class _RoundedConv(torch.autograd.Function):
    """conv2d / conv_transpose2d whose three GEMMs see rounded operands (bias and accumulation exact)."""

    @staticmethod
    def forward(ctx, x, w, transposed, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (transposed, stride, padding)
        if transposed:   # torch weight (in, out, kh, kw): the transposed layer is the backward-data of a conv out -> in
            on = _rounds("bwd_data", w.shape[1], w.shape[0], w.shape[2], stride)
            return Fn.conv_transpose2d(rnd(x, on), rnd(w, on), None, stride=stride, padding=padding)
        on = _rounds("fwd", w.shape[1], w.shape[0], w.shape[2], stride)
        return Fn.conv2d(rnd(x, on), rnd(w, on), None, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        transposed, stride, padding = ctx.cfg
        if transposed:   # y = conv_bwd_data(x, w): dx = conv_fwd(dy, w), dw = conv_bwd_weight(input = dy, grad = x)
            ci, co, k = w.shape[1], w.shape[0], w.shape[2]
            a, b = _rounds("fwd", ci, co, k, stride), _rounds("bwd_weight", ci, co, k, stride)
            dx = Fn.conv2d(rnd(dy, a), rnd(w, a), None, stride=stride, padding=padding)
            dw = torch.nn.grad.conv2d_weight(rnd(dy, b), w.shape, rnd(x, b), stride=stride, padding=padding)
        else:
            ci, co, k = w.shape[1], w.shape[0], w.shape[2]
            a, b = _rounds("bwd_data", ci, co, k, stride), _rounds("bwd_weight", ci, co, k, stride)
            dx = torch.nn.grad.conv2d_input(x.shape, rnd(w, a), rnd(dy, a), stride=stride, padding=padding)
            dw = torch.nn.grad.conv2d_weight(rnd(x, b), w.shape, rnd(dy, b), stride=stride, padding=padding)
        return dx, dw, None, None, None


def conv2d(x, w, b=None, stride=1, padding=0):
    if OPERAND is None:
        return Fn.conv2d(x, w, b, stride=stride, padding=padding)
    y = _RoundedConv.apply(x, w, False, stride, padding)
    return y if b is None else y + b.view(1, -1, 1, 1)


def conv_transpose2d(x, w, b=None, stride=1, padding=0):
    if OPERAND is None:
        return Fn.conv_transpose2d(x, w, b, stride=stride, padding=padding)
    y = _RoundedConv.apply(x, w, True, stride, padding)
    return y if b is None else y + b.view(1, -1, 1, 1)


def T(a):
    return torch.as_tensor(np.asarray(a), dtype=DT)


def l2n(x, eps=1e-12):
    return x / torch.sqrt(torch.clamp((x * x).sum(), min=eps))


def hwio(w):      # Conv2D kernel HWIO -> torch OIHW
    return w.permute(3, 2, 0, 1)


def convt_w(w):   # Conv2DTranspose kernel (kh, kw, out, in) -> torch (in, out, kh, kw)
    return w.permute(3, 2, 0, 1)


class State:
    """Mutable training state: weights (incl. sn_u, BN moving stats), Adam slots, step counters."""

    def __init__(self, gen_w, disc_w):
        self.g = {k: T(v).clone() for k, v in gen_w.items()}
        self.d = {k: T(v).clone() for k, v in disc_w.items()}
        self.g_adam = {"t": 0, "m": {}, "v": {}}
        self.d_adam = {"t": 0, "m": {}, "v": {}}


def trainable(names):
    return [n for n in names if not n.endswith(("sn_u", "moving_mean", "moving_variance"))]


def sn_update(w, name_w, name_u):
    """TFA 0.14 normalize_weights(): in-place w <- w / sigma, u <- u' (no gradient through sigma)."""
    with torch.no_grad():
        W = w[name_w].reshape(-1, w[name_w].shape[-1])
        u = w[name_u]
        v = l2n(u @ W.T)
        u2 = l2n(v @ W)
        sigma = (v @ W @ u2.T).reshape(())
        w[name_w] = w[name_w] / sigma
        w[name_u] = u2


def bn(x, w, i, training, momentum=0.99, eps=1e-3):
    """x: (N, C, H, W).  Training: biased batch variance for the normalisation; moving variance updated with the
    Bessel-corrected one (fused kernel behaviour)."""
    p = LW % i
    if training:
        mean = x.mean((0, 2, 3))
        var = x.var((0, 2, 3), unbiased=False)
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            w[p + "moving_mean"] = w[p + "moving_mean"] * momentum + mean.detach() * (1 - momentum)
            w[p + "moving_variance"] = w[p + "moving_variance"] * momentum + var.detach() * (n / (n - 1)) * (1 - momentum)
    else:
        mean, var = w[p + "moving_mean"], w[p + "moving_variance"]
    g, b = w[p + "gamma"], w[p + "beta"]
    return (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + eps) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def ln(x, w, i, eps=1e-3):
    """LayerNormalization over the channel axis of (N, C, H, W)."""
    p = LW % i
    m = x.mean(1, keepdim=True)
    v = x.var(1, unbiased=False, keepdim=True)
    return (x - m) / torch.sqrt(v + eps) * w[p + "gamma"].view(1, -1, 1, 1) + w[p + "beta"].view(1, -1, 1, 1)


def lrelu(x):
    return Fn.leaky_relu(x, float(np.float32(0.2)))


def conv_lstm(x, K, R, b):
    """x (B, T, C, H, W) -> all h_t (B, T, F, H, W); gates i, f, c, o; hard_sigmoid / tanh."""
    B, Tn = x.shape[:2]
    Fc = R.shape[2]
    h = torch.zeros(B, Fc, x.shape[3], x.shape[4], dtype=DT)
    c = torch.zeros_like(h)
    outs = []
    for t in range(Tn):
        z = conv2d(x[:, t], hwio(K), b, padding=1) + conv2d(h, hwio(R), None, padding=1)
        zi, zf, zc, zo = z.split(Fc, 1)
        i = torch.clamp(0.2 * zi + 0.5, 0, 1)
        f = torch.clamp(0.2 * zf + 0.5, 0, 1)
        c = f * c + i * torch.tanh(zc)
        o = torch.clamp(0.2 * zo + 0.5, 0, 1)
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1)


def leafify(w, names):
    """Fresh differentiable reads of the variables (one per forward call, like TF variable reads)."""
    out = dict(w)
    for n in names:
        out[n] = w[n].detach().clone().requires_grad_(True)
    return out


def generator(w, image, noise, training):
    """w: dict of tensors (mutated for SN / BN moving stats when training).  image/noise (B,T,S,S,C).
    Returns (output (B,T,S,S,Cout), reads) where `reads` holds the differentiable weight reads."""
    if training:
        for i in (0, 2, 5, 7):
            sn_update(w, (LW % i) + "layer/w", (LW % i) + "layer/sn_u")
    r = leafify(w, trainable(w)) if training else w
    B, Tn, S = image.shape[:3]
    x = torch.cat([image, noise], -1).reshape(B * Tn, S, S, -1).permute(0, 3, 1, 2)
    x = bn(lrelu(conv2d(x, hwio(r[(LW % 0) + "layer/w"]), r[(LW % 0) + "layer/layer/bias"], stride=2, padding=3)), r, 1, training)
    res2 = x
    x = bn(lrelu(conv2d(x, hwio(r[(LW % 2) + "layer/w"]), r[(LW % 2) + "layer/layer/bias"], stride=2, padding=1)), r, 3, training)
    res4 = x
    s4 = x.shape[-1]
    x = conv_lstm(x.reshape(B, Tn, -1, s4, s4), r[(LW % 4) + "cell/kernel"], r[(LW % 4) + "cell/recurrent_kernel"], r[(LW % 4) + "cell/bias"])
    x = x.reshape(B * Tn, -1, s4, s4)
    x = bn(lrelu(conv2d(x, hwio(r[(LW % 5) + "layer/w"]), r[(LW % 5) + "layer/layer/bias"], padding=1)), r, 6, training)
    x = torch.cat([x, res4], 1)
    x = bn(lrelu(conv_transpose2d(x, convt_w(r[(LW % 7) + "layer/w"]), r[(LW % 7) + "layer/layer/bias"], stride=2)), r, 8, training)
    x = torch.cat([x, res2], 1)
    x = Fn.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    x = bn(lrelu(conv_transpose2d(x, convt_w(r[(LW % 9) + "layer/kernel"]), r[(LW % 9) + "layer/bias"], padding=2)), r, 10, training)
    x = conv2d(x, hwio(r[(LW % 11) + "layer/kernel"]), r[(LW % 11) + "layer/bias"], padding=1)
    if training:   # BN moving statistics were written into r (a copy): propagate
        for k in w:
            if k.endswith(("moving_mean", "moving_variance")):
                w[k] = r[k]
    return x.permute(0, 2, 3, 1).reshape(B, Tn, S, S, -1), r


def critic(w, low_res, high_res, training, P=None):
    """(B,T,S,S,3), (B,T,S,S,2) -> (B,1).  Current-code topology (no shortcut) unless P has one."""
    B, Tn, S = low_res.shape[:3]
    if P is None:
        F = w[(LW % 2) + "layer/w"].shape[-1]
        _, P = critic_weight_shapes(S, low_res.shape[-1], high_res.shape[-1], F, False)
    sn_layers = [2, 3] + [e["idx"] for e in P["pyramid"] + P["loop2"] + P["tail"]] + ([P["shortcut"]["idx"]] if P["shortcut"] else [])
    if training:
        for i in sn_layers:
            sn_update(w, (LW % i) + "layer/w", (LW % i) + "layer/sn_u")
    r = leafify(w, trainable(w)) if training else w

    def snconv(x, i, stride=1, pad=0):
        return lrelu(conv2d(x, hwio(r[(LW % i) + "layer/w"]), r[(LW % i) + "layer/layer/bias"], stride=stride, padding=pad))

    def cf(x):   # (B,T,S,S,C) -> (B,T,C,S,S)
        return x.permute(0, 1, 4, 2, 3)

    hr = conv_lstm(cf(high_res), r[(LW % 0) + "cell/kernel"], r[(LW % 0) + "cell/recurrent_kernel"], r[(LW % 0) + "cell/bias"])
    hr = ln(snconv(hr.reshape(B * Tn, -1, S, S), 2, pad=1), r, 4)
    mix = conv_lstm(cf(torch.cat([low_res, high_res], -1)), r[(LW % 1) + "cell/kernel"], r[(LW % 1) + "cell/recurrent_kernel"], r[(LW % 1) + "cell/bias"])
    mix = ln(snconv(mix.reshape(B * Tn, -1, S, S), 3, pad=1), r, 5)
    x = torch.cat([hr, mix], 1)
    for e in P["pyramid"]:
        x = ln(snconv(x, e["idx"], stride=3, pad=1), r, e["ln"])
    sc_in = x
    for e in P["loop2"]:
        x = ln(snconv(x, e["idx"], stride=3, pad=1), r, e["ln"])
    if P["shortcut"] is not None:
        sc = P["shortcut"]
        x = x + ln(snconv(sc_in, sc["idx"], stride=sc["stride"], pad=sc["pad"]), r, sc["ln"])
    for e in P["tail"]:
        x = ln(snconv(x, e["idx"], stride=2), r, e["ln"])
    x = x.permute(0, 2, 3, 1).reshape(B, Tn, -1)
    x = x @ r[(LW % P["dense"]) + "layer/kernel"] + r[(LW % P["dense"]) + "layer/bias"]
    return x.mean(1), r


def adam_apply(w, slots, grads, lr, b1=0.5, b2=0.9, eps=0.1):
    """Keras Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); var -= lr_t * m / (sqrt(v) + eps)."""
    slots["t"] += 1
    t = slots["t"]
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    with torch.no_grad():
        for n, g in grads.items():
            m = slots["m"].get(n, torch.zeros_like(g))
            v = slots["v"].get(n, torch.zeros_like(g))
            m = m + (g - m) * (1 - b1)
            v = v + (g * g - v) * (1 - b2)
            slots["m"][n], slots["v"][n] = m, v
            w[n] = w[n] - lr_t * m / (torch.sqrt(v) + eps)


def train_step(st, low_res, high_res, draws, n_critic=3, gamma=100.0, lr_g=1e-4, lr_d=4e-4):
    """ganbase.py:21-94.  Mutates `st`; returns the metrics dict (python floats)."""
    low_res, high_res = T(low_res), T(high_res)
    draws = [T(d) for d in draws]
    di = iter(draws)
    B = low_res.shape[0]
    for _ in range(n_critic):
        noise = next(di)
        with torch.no_grad():
            fake, _ = generator(st.g, low_res, noise, training=True)
        eps = next(di).reshape(B, 1, 1, 1, 1)
        combined = (eps * high_res + (1 - eps) * fake).detach().requires_grad_(True)
        out, _ = critic(st.d, low_res, combined, training=True)
        g_img, = torch.autograd.grad(out.sum(), combined)
        g_norm = torch.sqrt((g_img ** 2).sum((1, 2, 3)))                 # (B, C): reduced over T, H, W only (F4)
        gradient_reg = gamma * ((g_norm - 1) ** 2).mean()
        hr_n = high_res + next(di)
        s_real, r1 = critic(st.d, low_res, hr_n, training=True)
        fhr = fake + next(di)
        s_fake, r2 = critic(st.d, low_res, fhr, training=True)
        d_loss = -(s_real.mean() - s_fake.mean()) + gradient_reg.detach()   # the penalty is a constant here (F3)
        names = trainable(st.d)
        g1 = torch.autograd.grad(d_loss, [r1[n] for n in names], retain_graph=True, allow_unused=True)
        g2 = torch.autograd.grad(d_loss, [r2[n] for n in names], allow_unused=True)
        grads = {n: (a if a is not None else 0) + (b if b is not None else 0) for n, a, b in zip(names, g1, g2)}
        d_grads = grads
        adam_apply(st.d, st.d_adam, grads, lr_d)
    noise = next(di)
    fake, rg = generator(st.g, low_res, noise, training=True)
    score, _ = critic(st.d, low_res, fake, training=True)
    gen_loss = -score.mean()
    names = trainable(st.g)
    gg = torch.autograd.grad(gen_loss, [rg[n] for n in names])
    g_grads = dict(zip(names, gg))
    adam_apply(st.g, st.g_adam, g_grads, lr_g)
    with torch.no_grad():
        s_real, _ = critic(st.d, low_res, high_res, training=False)
        fake_m, _ = generator(st.g, low_res, next(di), training=False)
        s_fake, _ = critic(st.d, low_res, fake_m, training=False)
        d_loss_m = -(s_real.mean() - s_fake.mean())
        g_loss_m = -s_fake.mean()
    return {"g_loss": float(g_loss_m), "g_disc_loss": float(gen_loss), "d_loss": float(d_loss_m),
            "d_gradient_pen": float(g_norm.mean()),
            "g_gradient_param": float(np.mean([float((g ** 2).mean()) for g in g_grads.values()])),
            "d_gradient_param": float(np.mean([float((g ** 2).mean()) for g in d_grads.values()])),
            "d_real": float(s_real.mean()), "d_fake": float(s_fake.mean())}
