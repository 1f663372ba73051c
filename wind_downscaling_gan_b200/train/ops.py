"""Python wrappers of the training kernels (include/wdg.h, "building blocks of the WGAN training step").  Tensors are contiguous fp32
CUDA torch tensors used as raw device buffers; `View` addresses a channel slice of a wider channels-last buffer."""
import ctypes as C

import torch

from .. import _lib

F32 = torch.float32


_stream = [None]


def use_current_stream():
    """Pins the torch current stream for the following op calls (querying it per call costs ~13 us of host time and a
    training step issues ~6000 ops); call again whenever the caller switches streams."""
    _stream[0] = C.c_void_p(torch.cuda.current_stream().cuda_stream)


PRECISIONS = {"fp32": 0, "tf32": 1, "bf16": 2}


def set_precision(name):
    """Arithmetic of the convolution GEMMs: "fp32" (CUDA cores), "tf32" or "bf16" (tcgen05, fp32 accumulate)."""
    _lib.check(_lib.lib().wdg_train_set_precision(PRECISIONS[name]))


def get_precision():
    mode = _lib.lib().wdg_train_get_precision()
    return {v: k for k, v in PRECISIONS.items()}[mode]


def _s():
    if _stream[0] is None:
        use_current_stream()
    return _stream[0]


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def empty(*shape):
    return torch.empty(shape, dtype=F32, device="cuda")


def zeros(*shape):
    return torch.zeros(shape, dtype=F32, device="cuda")


class View:
    """Channels [co, co+C) of a channels-last buffer whose pixel pitch is cs."""

    def __init__(self, t, C_, cs=None, co=0):
        self.t, self.C, self.cs, self.co = t, C_, (cs if cs is not None else t.shape[-1]), co

    @property
    def rows(self):
        return self.t.numel() // self.t.shape[-1]


def full(t):
    return View(t, t.shape[-1])


_scratch = {}


def scratch(nbytes, tag="default"):
    buf = _scratch.get(tag)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device="cuda")
        _scratch[tag] = buf
    return buf


def _geo(N, H, W, Ci, kh, kw, Co, stride, pad_t, pad_l, Ho, Wo, xv, yv):
    return (C.c_int * 16)(N, H, W, Ci, kh, kw, Co, stride, pad_t, pad_l, Ho, Wo, xv.cs, xv.co, yv.cs, yv.co)


def conv_out(n, k, s, pad_lo, pad_hi):
    return (n + pad_lo + pad_hi - k) // s + 1


def conv2d_fwd(xv, w, bias, yv, N, H, W, stride, pad, Ho, Wo, accumulate=False, alpha=1.0):
    """y = leaky_alpha((y +) conv(x, w) + bias); alpha = 1: linear."""
    kh, kw, Ci, Co = w.shape
    g = _geo(N, H, W, Ci, kh, kw, Co, stride, pad, pad, Ho, Wo, xv, yv)
    _lib.check(_lib.lib().wdg_conv2d_fwd_act(_p(xv.t), _p(w), _p(bias), _p(yv.t), g, int(accumulate), alpha, _s()))


def conv2d_bwd_data(dyv, w, dxv, N, H, W, stride, pad, Ho, Wo, accumulate=False):
    kh, kw, Ci, Co = w.shape
    g = _geo(N, H, W, Ci, kh, kw, Co, stride, pad, pad, Ho, Wo, dxv, dyv)
    _lib.check(_lib.lib().wdg_conv2d_bwd_data(_p(dyv.t), _p(w), _p(dxv.t), g, int(accumulate), _s()))


def conv2d_bwd_weight(xv, dyv, dw, N, H, W, stride, pad, Ho, Wo, accumulate=False):
    kh, kw, Ci, Co = dw.shape
    g = _geo(N, H, W, Ci, kh, kw, Co, stride, pad, pad, Ho, Wo, xv, dyv)
    nb = C.c_size_t()
    _lib.check(_lib.lib().wdg_conv2d_bwd_weight_scratch(g, C.byref(nb), None))
    sc = scratch(nb.value, "wgrad")
    _lib.check(_lib.lib().wdg_conv2d_bwd_weight(_p(xv.t), _p(dyv.t), _p(dw), g, _p(sc), int(accumulate), _s()))


def colsum(av, out, mode=0, bv=None, accumulate=False):
    sc = scratch(512 * max(av.C, 32) * 4, "colsum")
    b = bv if bv is not None else av
    _lib.check(_lib.lib().wdg_colsum(mode, _p(av.t), av.cs, av.co, _p(b.t), b.cs, b.co, av.rows, av.C, _p(out), _p(sc),
                                     int(accumulate), _s()))


def leaky_fwd(x, alpha=0.2):
    _lib.check(_lib.lib().wdg_leaky_relu_fwd(_p(x), x.numel(), alpha, _s()))


def leaky_bwd(dy, y, alpha=0.2):
    _lib.check(_lib.lib().wdg_leaky_relu_bwd(_p(dy), _p(y), dy.numel(), alpha, _s()))


def axpby(outv, xv, a=1.0, yv=None, b=0.0, accumulate=False):
    y = yv if yv is not None else xv
    _lib.check(_lib.lib().wdg_axpby(_p(outv.t), outv.cs, outv.co, _p(xv.t), xv.cs, xv.co, a, _p(yv.t) if yv is not None else None,
                                    y.cs, y.co, b, xv.rows, xv.C, int(accumulate), _s()))


def lerp_batch(out, real, fake, eps):
    n = real.numel()
    _lib.check(_lib.lib().wdg_lerp_batch(_p(out), _p(real), _p(fake), _p(eps), n // real.shape[0], n, _s()))


def bn_train_fwd(x, y, gamma, beta, mm, mv, save_mean, save_invstd, eps=1e-3, momentum=0.99):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    sc = scratch((512 * max(Cc, 32) + 2 * Cc) * 4, "bn")
    _lib.check(_lib.lib().wdg_bn_train_fwd(_p(x), _p(y), _p(gamma), _p(beta), _p(mm), _p(mv), _p(save_mean), _p(save_invstd), rows,
                                           Cc, eps, momentum, _p(sc), _s()))


def bn_infer(x, y, gamma, beta, mean, var, eps=1e-3):
    Cc = x.shape[-1]
    sc = scratch(Cc * 4, "bn")
    _lib.check(_lib.lib().wdg_bn_infer(_p(x), _p(y), _p(gamma), _p(beta), _p(mean), _p(var), x.numel() // Cc, Cc, eps, _p(sc), _s()))


def bn_train_bwd(dy, x, gamma, save_mean, save_invstd, dx, dgamma, dbeta, act_alpha=1.0):
    """act_alpha != 1: x is the output of LeakyReLU(act_alpha); its backward is folded into dx."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    sc = scratch((rows * Cc + 512 * max(Cc, 32)) * 4, "norm_bwd")
    _lib.check(_lib.lib().wdg_bn_train_bwd(_p(dy), _p(x), _p(gamma), _p(save_mean), _p(save_invstd), _p(dx), _p(dgamma), _p(dbeta),
                                           rows, Cc, act_alpha, _p(sc), _s()))


def ln_fwd(x, yv, gamma, beta, save_mean, save_invstd, eps=1e-3):
    Cc = x.shape[-1]
    _lib.check(_lib.lib().wdg_ln_fwd(_p(x), _p(yv.t), yv.cs, yv.co, _p(gamma), _p(beta), _p(save_mean), _p(save_invstd),
                                     x.numel() // Cc, Cc, eps, _s()))


def ln_bwd(dyv, x, gamma, save_mean, save_invstd, dx, dgamma, dbeta, act_alpha=1.0):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    sc = scratch((rows * Cc + 512 * max(Cc, 32)) * 4, "norm_bwd")
    _lib.check(_lib.lib().wdg_ln_bwd(_p(dyv.t), dyv.cs, dyv.co, _p(x), _p(gamma), _p(save_mean), _p(save_invstd), _p(dx), _p(dgamma),
                                     _p(dbeta), rows, Cc, act_alpha, _p(sc), _s()))


def lstm_gates_fwd(z, c_prev, c_out, h_out):
    Fc = c_out.shape[-1]
    _lib.check(_lib.lib().wdg_lstm_gates_fwd(_p(z), _p(c_prev), _p(c_out), _p(h_out), c_out.numel() // Fc, Fc, _s()))


def lstm_gates_bwd(gates, c_prev, c_cur, dh, dc, dh_rec=None):
    Fc = c_cur.shape[-1]
    _lib.check(_lib.lib().wdg_lstm_gates_bwd(_p(gates), _p(c_prev), _p(c_cur), _p(dh), _p(dh_rec), _p(dc),
                                              c_cur.numel() // Fc, Fc, _s()))


SMALL_LSTM_FILTERS = (1, 2, 4)


def lstm_small_fwd(z, h_prev, R, c_prev, c_out, h_out):
    """Fused recurrent 3x3 conv + gates for cells with 1/2/4 filters; z [N,H,W,4F] holds x-conv + bias, becomes the gates."""
    N, H, W, Fc = c_out.shape
    _lib.check(_lib.lib().wdg_lstm_small_fwd(_p(z), _p(h_prev), _p(R), _p(c_prev), _p(c_out), _p(h_out), N, H, W, Fc, _s()))


def lstm_small_bwd_data(dz, R, dh_rec):
    N, H, W, Fc = dh_rec.shape
    _lib.check(_lib.lib().wdg_lstm_small_bwd_data(_p(dz), _p(R), _p(dh_rec), N, H, W, Fc, _s()))


def lstm16_fused():
    """The 16-filter cell runs its recurrent steps as fused tcgen05 launches in the tensor-core training modes."""
    import os
    return get_precision() != "fp32" and not os.environ.get("WDG_NO_LSTM16")


def lstm16_pack(R):
    """R [3,3,16,64] -> packed forward / backward-data operands of the fused recurrent steps (tf32)."""
    packed = torch.empty(64 * 9 * 32 + 16 * 18 * 32, dtype=F32, device="cuda")
    _lib.check(_lib.lib().wdg_lstm16_pack(_p(R), _p(packed), _s()))
    return packed


def round_tf32(x):
    _lib.check(_lib.lib().wdg_round_tf32(_p(x), x.numel(), _s()))


def lstm16_fwd_step(gates, h_prev, packed, c_prev, c_out, h_out):
    N, H, W, _ = c_out.shape
    _lib.check(_lib.lib().wdg_lstm16_fwd_step(_p(gates), _p(h_prev), _p(packed), _p(c_prev), _p(c_out), _p(h_out), N, H, W, _s()))


def lstm16_bwd_step(dz_next, packed, gates_s, c_prev, c_cur, dh, dc):
    N, H, W, _ = c_cur.shape
    _lib.check(_lib.lib().wdg_lstm16_bwd_step(_p(dz_next), _p(packed), _p(gates_s), _p(c_prev), _p(c_cur), _p(dh), _p(dc), N, H, W, _s()))


def lstm128_pack(R):
    """R [3,3,128,512] -> packed forward operand of the fused recurrent steps of the 128-filter cell (tf32)."""
    packed = torch.empty(512 * 36 * 32, dtype=F32, device="cuda")
    _lib.check(_lib.lib().wdg_lstm128_pack(_p(R), _p(packed), _s()))
    return packed


def lstm128_fwd_step(gates, h_prev, packed, c_prev, c_out, h_out):
    N, H, W, _ = c_out.shape
    _lib.check(_lib.lib().wdg_lstm128_fwd_step(_p(gates), _p(h_prev), _p(packed), _p(c_prev), _p(c_out), _p(h_out), N, H, W, _s()))


def lstm128_fused():
    """The generator's 128-filter cell runs its forward recurrent steps as fused tcgen05 launches in the tensor-core modes."""
    import os
    return get_precision() != "fp32" and not os.environ.get("WDG_NO_LSTM128")


def upconv_fused():
    """The generator's concat -> bilinear x2 -> 5x5 transposed conv block runs as the fused tcgen05 kernel in the tensor-core
    training modes."""
    import os
    return get_precision() != "fp32" and not os.environ.get("WDG_NO_FUSED_UPCONV")


def upconv5x5_fwd(a, b, w, bias, out):
    """a [N,h,h,32], b [N,h,h,128] -> out [N,2h,2h,16] = leaky(convT5x5(upsample2x(concat(a, b))) + bias)."""
    N, h = a.shape[0], a.shape[1]
    nb = C.c_size_t()
    _lib.check(_lib.lib().wdg_upconv5x5_workspace_bytes(N, h, C.byref(nb)))
    ws = scratch(nb.value + 1024, "upconv")
    base = (ws.data_ptr() + 1023) // 1024 * 1024
    _lib.check(_lib.lib().wdg_upconv5x5_fwd(_p(a), _p(b), _p(w), _p(bias), _p(out), N, h, C.c_void_p(base), nb.value, _s()))


def upsample2x_fwd(x, y):
    n, h, w, Cc = x.shape
    _lib.check(_lib.lib().wdg_upsample2x_fwd(_p(x), _p(y), n, h, w, Cc, _s()))


def upsample2x_bwd(dy, dx):
    n, h, w, Cc = dx.shape
    _lib.check(_lib.lib().wdg_upsample2x_bwd(_p(dy), _p(dx), n, h, w, Cc, _s()))


def dense_mean_fwd(flat, w, bias, score, B, T, D):
    _lib.check(_lib.lib().wdg_dense_mean_fwd(_p(flat), _p(w), _p(bias), _p(score), B, T, D, _s()))


def dense_mean_bwd(dscore, flat, w, dflat, dw, dbias, B, T, D):
    _lib.check(_lib.lib().wdg_dense_mean_bwd(_p(dscore), _p(flat), _p(w), _p(dflat), _p(dw), _p(dbias), B, T, D, _s()))


def reduce(a, mode=0, b=None, scale=1.0):
    """Returns a 1-element CUDA tensor: scale * sum(a | a*b | a*a)."""
    out = empty(1)
    sc = scratch(1024 * 8, "reduce")
    _lib.check(_lib.lib().wdg_reduce(mode, _p(a), _p(b), a.numel(), scale, _p(out), _p(sc), _s()))
    return out


def reduce_into(out, a, mode=0, b=None, scale=1.0):
    """out[0] = scale * sum(a | a*b | a*a), `out` a 1-element view of a device buffer (no host round trip)."""
    sc = scratch(1024 * 8, "reduce")
    _lib.check(_lib.lib().wdg_reduce(mode, _p(a), _p(b), a.numel(), scale, _p(out), _p(sc), _s()))


def gp_norm(g, out):
    B, Cc = g.shape[0], g.shape[-1]
    _lib.check(_lib.lib().wdg_gp_norm(_p(g), _p(out), B, g.numel() // (B * Cc), Cc, _s()))


def adam(w, m, v, g, lr_t, b1, b2, eps):
    _lib.check(_lib.lib().wdg_adam(_p(w), _p(m), _p(v), _p(g), w.numel(), lr_t, b1, b2, eps, _s()))


def adam_lr(lr_t, step, lr, b1, b2):
    """step += 1; lr_t = lr * sqrt(1 - b2^step) / (1 - b1^step), both device scalars."""
    _lib.check(_lib.lib().wdg_adam_lr(_p(lr_t), _p(step), lr, b1, b2, _s()))


def adam_dev(w, m, v, g, lr_t, b1, b2, eps):
    _lib.check(_lib.lib().wdg_adam_dev(_p(w), _p(m), _p(v), _p(g), w.numel(), _p(lr_t), b1, b2, eps, _s()))


def sn_update(w, u):
    Cl = w.shape[-1]
    R = w.numel() // Cl
    sc = scratch((R + 64 * Cl + 4) * 4, "sn")
    _lib.check(_lib.lib().wdg_sn_update(_p(w), _p(u), R, Cl, _p(sc), _s()))


def bias_act(xv, bias, alpha=1.0):
    """x = leaky(x + bias, alpha) in place on a channel view (alpha = 1: linear)."""
    _lib.check(_lib.lib().wdg_bias_act(_p(xv.t), xv.cs, xv.co, _p(bias), xv.rows, xv.C, alpha, _s()))


def conv_transpose_s1_fwd(xv, w, yv, N, H, W, pad):
    """Stride-1 Conv2DTranspose forward as one 1x1 GEMM into per-pixel tap columns + overlap-add (wdg_col2im): the wide
    input is read once instead of once per tap.  w: (kh, kw, out, in); xv: view with `in` channels; yv: `out` channels."""
    kh, kw, cout, cin = w.shape
    wmat = empty(cin, kh * kw * cout)                       # [in][(ky, kx, out)]
    _lib.check(_lib.lib().wdg_transpose01(_p(w), _p(wmat), kh * kw * cout, cin, 1, _s()))
    nb = N * H * W * kh * kw * cout * 4
    cols = scratch(nb, "col2im")[:nb].view(F32).view(N, H, W, kh * kw * cout)
    conv2d_fwd(xv, wmat.view(1, 1, cin, kh * kw * cout), None, full(cols), N, H, W, 1, 0, H, W)
    _lib.check(_lib.lib().wdg_col2im(_p(cols), _p(yv.t), N, H, W, kh, kw, cout, pad, yv.cs, yv.co, _s()))


def transpose01(x):
    """[A][B][...] -> [B][A][...] (new tensor)."""
    A, B = x.shape[0], x.shape[1]
    out = torch.empty((B, A) + tuple(x.shape[2:]), dtype=F32, device="cuda")
    _lib.check(_lib.lib().wdg_transpose01(_p(x), _p(out), A, B, x.numel() // (A * B), _s()))
    return out


def bn_finalize_apply(x, y, gamma, beta, mm, mv, s1, s2, save_mean, save_invstd, rows_global, eps=1e-3, momentum=0.99):
    Cc = x.shape[-1]
    _lib.check(_lib.lib().wdg_bn_finalize_apply(_p(x), _p(y), _p(gamma), _p(beta), _p(mm), _p(mv), _p(s1), _p(s2), _p(save_mean),
                                                _p(save_invstd), x.numel() // Cc, rows_global, Cc, eps, momentum, _s()))


def bn_bwd_sums(dy, x, save_mean, save_invstd, dgamma, dbeta):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    sc = scratch((rows * Cc + 512 * max(Cc, 32)) * 4, "norm_bwd")
    _lib.check(_lib.lib().wdg_bn_bwd_sums(_p(dy), _p(x), _p(save_mean), _p(save_invstd), _p(dgamma), _p(dbeta), rows, Cc, _p(sc), _s()))


def bn_bwd_dx(dy, x, gamma, save_mean, save_invstd, dgamma, dbeta, dx, rows_global, act_alpha=1.0):
    Cc = x.shape[-1]
    _lib.check(_lib.lib().wdg_bn_bwd_dx(_p(dy), _p(x), _p(gamma), _p(save_mean), _p(save_invstd), _p(dgamma), _p(dbeta), _p(dx),
                                        x.numel() // Cc, rows_global, Cc, act_alpha, _s()))
