"""Training-path generator and critic: explicit forward / backward over the kernels of train/ops.py (fp32 tensors;
convolution GEMMs in the precision selected by `ops.set_precision`).

Graphs follow the reference's `make_generator` / `make_discriminator` (`gan/models.py:9-142`) in TRAINING mode:
SpectralNormalization performs one in-place power iteration per call (TFA 0.14), BatchNormalization uses batch
statistics and updates its moving averages, LayerNormalization / ConvLSTM2D as in inference.  Weights are fp32 CUDA
tensors keyed by the checkpoint variable names.  Activations are channels-last [N = B*T, H, W, C].
"""

import numpy as np
import torch

from . import ops
from .ops import View, full

LW = "layer_with_weights-%d/"
ALPHA = float(np.float32(0.2))


def to_device(weights):
    return {k: torch.as_tensor(np.asarray(v, np.float32)).cuda().contiguous() for k, v in weights.items()}


def trainable_names(weights):
    return [n for n in weights if not n.endswith(("sn_u", "moving_mean", "moving_variance"))]


class Conv:
    """Conv2D (HWIO kernel) or Conv2DTranspose (kernel (kh,kw,out,in) = HWIO of the conv it transposes)."""

    def __init__(self, w, b, stride, pad, transposed=False, leaky=True):
        self.w, self.b, self.stride, self.pad, self.T, self.leaky = w, b, stride, pad, transposed, leaky

    def out_hw(self, H, W):
        k, s, p = self.w.shape[0], self.stride, self.pad
        if self.T:
            return (H - 1) * s + k - 2 * p, (W - 1) * s + k - 2 * p
        return ops.conv_out(H, k, s, p, p), ops.conv_out(W, k, s, p, p)

    def forward(self, xv, N, H, W, out=None):
        """xv: View of the input ([N,H,W,*]); returns the activation tensor [N,Ho,Wo,Cout] (or writes into `out` view)."""
        Ho, Wo = self.out_hw(H, W)
        cout = self.w.shape[2] if self.T else self.w.shape[3]
        y = out if out is not None else full(ops.empty(N, Ho, Wo, cout))
        if self.T and self.stride == 1 and self.w.shape[0] > 1 and xv.C >= 64:
            # wide input, many taps (generator: 5x5 over 160 channels): 1x1 GEMM + overlap-add instead of a tap gather
            ops.conv_transpose_s1_fwd(xv, self.w, y, N, H, W, self.pad)
        elif self.T:   # forward of the transposed conv = backward-data of the conv (conv input dims = our output dims)
            ops.conv2d_bwd_data(xv, self.w, y, N, Ho, Wo, self.stride, self.pad, H, W)
        else:             # bias + LeakyReLU fused into the GEMM epilogue
            ops.conv2d_fwd(xv, self.w, self.b, y, N, H, W, self.stride, self.pad, Ho, Wo, alpha=ALPHA if self.leaky else 1.0)
        if self.T:
            ops.bias_act(y, self.b, ALPHA if self.leaky else 1.0)
        self.ctx = (xv, N, H, W, Ho, Wo, y)
        return y

    def backward(self, dy, need_dx=True, dx_out=None, accumulate_dx=False, need_dw=True, act_done=False):
        """dy: dense tensor [N,Ho,Wo,Cout] (modified in place by the activation backward).  Returns (dx view, dw, db).
        act_done: the LeakyReLU backward was already folded into dy by the normalisation layer's backward."""
        xv, N, H, W, Ho, Wo, y = self.ctx
        dyv = full(dy)
        if self.leaky and not act_done:
            if y.cs != y.C:
                raise NotImplementedError("leaky backward needs a dense activation")
            ops.leaky_bwd(dy, y.t, ALPHA)
        db = dw = None
        if need_dw:
            db = ops.empty(self.b.shape[0])
            ops.colsum(dyv, db)
            dw = ops.empty(*self.w.shape)
        dx = None
        if self.T:
            if need_dw:
                ops.conv2d_bwd_weight(dyv, xv, dw, N, Ho, Wo, self.stride, self.pad, H, W)
            if need_dx:
                dx = dx_out if dx_out is not None else full(ops.empty(N, H, W, xv.C))
                ops.conv2d_fwd(dyv, self.w, None, dx, N, Ho, Wo, self.stride, self.pad, H, W, accumulate=accumulate_dx)
        else:
            if need_dw:
                ops.conv2d_bwd_weight(xv, dyv, dw, N, H, W, self.stride, self.pad, Ho, Wo)
            if need_dx:
                dx = dx_out if dx_out is not None else full(ops.empty(N, H, W, xv.C))
                ops.conv2d_bwd_data(dyv, self.w, dx, N, H, W, self.stride, self.pad, Ho, Wo, accumulate=accumulate_dx)
        return dx, dw, db


class BatchNorm:
    """BatchNormalization (axis -1).  With a communicator the batch statistics (and the two backward sums) are
    all-reduced over the data-parallel ranks, which makes N ranks x local batch equivalent to the reference's single
    process at the global batch; dgamma / dbeta returned by backward() are then already global sums."""

    def __init__(self, w, i, comm=None):
        p = LW % i
        self.names = (p + "gamma", p + "beta", p + "moving_mean", p + "moving_variance")
        self.w, self.comm = w, comm

    def forward(self, x, training):
        g, b, mm, mv = (self.w[n] for n in self.names)
        y = torch.empty_like(x)
        if training:
            C = x.shape[-1]
            rows = x.numel() // C
            self.sm, self.si = ops.empty(C), ops.empty(C)
            if self.comm is None or self.comm.world == 1:
                ops.bn_train_fwd(x, y, g, b, mm, mv, self.sm, self.si)
                self.rows_global = rows
            else:
                sums = ops.empty(2, C)
                ops.colsum(full(x), sums[0], mode=0)
                ops.colsum(full(x), sums[1], mode=2)
                self.comm.allreduce_sum(sums)
                self.rows_global = rows * self.comm.world
                ops.bn_finalize_apply(x, y, g, b, mm, mv, sums[0], sums[1], self.sm, self.si, self.rows_global)
            self.x = x
        else:
            ops.bn_infer(x, y, g, b, mm, mv)
        return y

    def backward(self, dy, act_alpha=1.0):
        """act_alpha != 1: the input of this layer is the output of LeakyReLU(act_alpha); its backward is folded into dx."""
        C = self.x.shape[-1]
        dx = torch.empty_like(self.x)
        if self.comm is None or self.comm.world == 1:
            dg, db = ops.empty(C), ops.empty(C)
            ops.bn_train_bwd(dy, self.x, self.w[self.names[0]], self.sm, self.si, dx, dg, db, act_alpha)
        else:
            sums = ops.empty(2, C)
            ops.bn_bwd_sums(dy, self.x, self.sm, self.si, sums[0], sums[1])
            self.comm.allreduce_sum(sums)
            ops.bn_bwd_dx(dy, self.x, self.w[self.names[0]], self.sm, self.si, sums[0], sums[1], dx, self.rows_global, act_alpha)
            dg, db = sums[0], sums[1]
        return dx, {self.names[0]: dg, self.names[1]: db}


class LayerNorm:
    def __init__(self, w, i):
        p = LW % i
        self.names = (p + "gamma", p + "beta")
        self.w = w

    def forward(self, x, out=None):
        rows = x.numel() // x.shape[-1]
        self.sm, self.si, self.x = ops.empty(rows), ops.empty(rows), x
        y = out if out is not None else full(torch.empty_like(x))
        ops.ln_fwd(x, y, self.w[self.names[0]], self.w[self.names[1]], self.sm, self.si)
        return y

    def backward(self, dyv, act_alpha=1.0):
        C = self.x.shape[-1]
        dx, dg, db = torch.empty_like(self.x), ops.empty(C), ops.empty(C)
        ops.ln_bwd(dyv, self.x, self.w[self.names[0]], self.sm, self.si, dx, dg, db, act_alpha)
        return dx, {self.names[0]: dg, self.names[1]: db}


class ConvLSTM:
    """ConvLSTM2D(F, 3x3, same, return_sequences): x-conv with bias + recurrent conv, gates i,f,c,o."""

    def __init__(self, K, R, b):
        self.K, self.R, self.b = K, R, b

    def forward(self, x, B, T):
        """x: [B*T, H, W, Cin] batch-major -> h sequence [B*T, H, W, F] batch-major."""
        _, H, W, Cin = x.shape
        F = self.R.shape[2]
        xt = ops.transpose01(x.view(B, T, H, W, Cin))            # [T, B, H, W, Cin]
        hs = ops.empty(T, B, H, W, F)
        cs = ops.empty(T, B, H, W, F)
        gates = ops.empty(T, B, H, W, 4 * F)
        # the input convolution (+ bias) of every timestep is one GEMM; only the recurrent convolution is serial
        ops.conv2d_fwd(full(xt.view(T * B, H, W, Cin)), self.K, self.b, full(gates.view(T * B, H, W, 4 * F)), T * B, H, W, 1, 1, H, W)
        small = F in ops.SMALL_LSTM_FILTERS and tuple(self.R.shape[:2]) == (3, 3)
        fused16 = F == 16 and tuple(self.R.shape[:2]) == (3, 3) and ops.lstm16_fused()
        packed = ops.lstm16_pack(self.R) if fused16 else None
        fused128 = F == 128 and tuple(self.R.shape[:2]) == (3, 3) and ops.lstm128_fused()
        packed128 = ops.lstm128_pack(self.R) if fused128 and T > 1 else None
        for t in range(T):
            if small:      # recurrent conv + gates in one bandwidth-bound pass (critic high-resolution branch)
                ops.lstm_small_fwd(gates[t], hs[t - 1] if t > 0 else None, self.R, cs[t - 1] if t > 0 else None, cs[t], hs[t])
                continue
            if fused128:   # generator cell: recurrent conv (tcgen05, K = 1152, N = 512) + gates in ONE launch per step
                if t == 0:
                    ops.lstm_gates_fwd(gates[0], None, cs[0], hs[0])
                    ops.round_tf32(hs[0])
                else:
                    ops.lstm128_fwd_step(gates[t], hs[t - 1], packed128, cs[t - 1], cs[t], hs[t])
                continue
            if fused16:    # critic mixed branch: recurrent conv (tcgen05) + gates in ONE launch per step
                if t == 0:
                    ops.lstm_gates_fwd(gates[0], None, cs[0], hs[0])
                    ops.round_tf32(hs[0])
                else:
                    ops.lstm16_fwd_step(gates[t], hs[t - 1], packed, cs[t - 1], cs[t], hs[t])
                continue
            if t > 0:
                ops.conv2d_fwd(full(hs[t - 1]), self.R, None, full(gates[t]), B, H, W, 1, 1, H, W, accumulate=True)
            ops.lstm_gates_fwd(gates[t], cs[t - 1] if t > 0 else None, cs[t], hs[t])
        self.ctx = (xt, hs, cs, gates, B, T, H, W, Cin, F)
        self.packed = packed
        return ops.transpose01(hs).view(B * T, H, W, F)

    def backward(self, dh_seq, need_dx=True, need_dw=True):
        """dh_seq: [B*T, H, W, F] batch-major.  Returns (dx batch-major or None, dK, dR, db)."""
        xt, hs, cs, gates, B, T, H, W, Cin, F = self.ctx
        dhs = ops.transpose01(dh_seq.view(B, T, H, W, F))        # [T, B, ...]
        dK = dR = db = None
        dc = ops.zeros(B, H, W, F)
        dh_rec = ops.empty(B, H, W, F)
        small = F in ops.SMALL_LSTM_FILTERS and tuple(self.R.shape[:2]) == (3, 3)
        fused16 = getattr(self, "packed", None) is not None
        if fused16:      # step T-1 has no recurrent gradient; every earlier step is one fused launch:
            # recurrent backward-data of dz_{t+1} (tcgen05) + gate backward of step t in the epilogue
            ops.lstm_gates_bwd(gates[T - 1], cs[T - 2] if T > 1 else None, cs[T - 1], dhs[T - 1], dc, None)
            ops.round_tf32(gates[T - 1])
            for t in range(T - 2, -1, -1):
                ops.lstm16_bwd_step(gates[t + 1], self.packed, gates[t], cs[t - 1] if t > 0 else None, cs[t], dhs[t], dc)
        for t in range(T - 1, -1, -1) if not fused16 else ():            # serial part: gate backward and the recurrent backward-data
            ops.lstm_gates_bwd(gates[t], cs[t - 1] if t > 0 else None, cs[t], dhs[t], dc,
                               dh_rec if t < T - 1 else None)                            # gates[t] now holds dz_t
            if t > 0:
                if small:
                    ops.lstm_small_bwd_data(gates[t], self.R, dh_rec)
                else:
                    ops.conv2d_bwd_data(full(gates[t]), self.R, full(dh_rec), B, H, W, 1, 1, H, W)
        # everything that only needs all dz_t: one GEMM each over the T*B images
        dz_all = full(gates.view(T * B, H, W, 4 * F))
        if need_dw:
            dK, dR, db = torch.empty_like(self.K), torch.empty_like(self.R), torch.empty_like(self.b)
            ops.conv2d_bwd_weight(full(xt.view(T * B, H, W, Cin)), dz_all, dK, T * B, H, W, 1, 1, H, W)
            if T > 1:
                ops.conv2d_bwd_weight(full(hs[:T - 1].view((T - 1) * B, H, W, F)), full(gates[1:].view((T - 1) * B, H, W, 4 * F)),
                                      dR, (T - 1) * B, H, W, 1, 1, H, W)
            else:
                dR.zero_()
            ops.colsum(dz_all, db)
        dxt = None
        if need_dx:
            dxt = ops.empty(T, B, H, W, Cin)
            ops.conv2d_bwd_data(dz_all, self.K, full(dxt.view(T * B, H, W, Cin)), T * B, H, W, 1, 1, H, W)
        dx = ops.transpose01(dxt).view(B * T, H, W, Cin) if need_dx else None
        return dx, dK, dR, db


def sn_step(w, idx):
    """One training-mode call of SpectralNormalization on layer `idx`: w <- w / sigma, u <- u' (in place)."""
    ops.sn_update(w[(LW % idx) + "layer/w"], w[(LW % idx) + "layer/sn_u"])


# =============================================================================================== generator
class GenNet:
    SN_LAYERS = (0, 2, 5, 7)

    def __init__(self, weights, comm=None):
        self.w = weights   # dict name -> CUDA tensor (shared with the caller; updated in place)
        self.comm = comm   # data-parallel communicator: synchronised BatchNorm statistics

    def forward(self, image, noise, training, keep_context=True):
        """image [B,T,S,S,Cin], noise [B,T,S,S,Cn] -> [B,T,S,S,Cout].  Keeps the context for backward() unless
        keep_context is False (the critic-loop and metric forwards of train_step are never differentiated: dropping their
        saved activations returns ~3 GB per call to the allocator)."""
        w = self.w
        B, T, S = image.shape[:3]
        N = B * T
        if training:
            for i in self.SN_LAYERS:
                sn_step(w, i)
        cin, cn = image.shape[-1], noise.shape[-1]
        x0 = ops.empty(N, S, S, cin + cn)                                                      # models.py:28
        ops.axpby(View(x0, cin, cin + cn, 0), full(image.reshape(N, S, S, cin)))
        ops.axpby(View(x0, cn, cin + cn, cin), full(noise.reshape(N, S, S, cn)))
        L = {}
        L["c0"] = Conv(w[(LW % 0) + "layer/w"], w[(LW % 0) + "layer/layer/bias"], 2, 3)        # :32-33
        a0 = L["c0"].forward(full(x0), N, S, S)
        L["bn1"] = BatchNorm(w, 1, self.comm)
        r2 = L["bn1"].forward(a0.t, training)                                                  # :34
        S2 = S // 2
        L["c2"] = Conv(w[(LW % 2) + "layer/w"], w[(LW % 2) + "layer/layer/bias"], 2, 1)        # :38-39
        a2 = L["c2"].forward(full(r2), N, S2, S2)
        L["bn3"] = BatchNorm(w, 3, self.comm)
        r4 = L["bn3"].forward(a2.t, training)                                                  # :40
        S4 = S // 4
        L["lstm"] = ConvLSTM(w[(LW % 4) + "cell/kernel"], w[(LW % 4) + "cell/recurrent_kernel"], w[(LW % 4) + "cell/bias"])
        hseq = L["lstm"].forward(r4, B, T)                                                     # :45
        L["c5"] = Conv(w[(LW % 5) + "layer/w"], w[(LW % 5) + "layer/layer/bias"], 1, 1)        # :49
        a5 = L["c5"].forward(full(hseq), N, S4, S4)
        L["bn6"] = BatchNorm(w, 6, self.comm)
        r5 = L["bn6"].forward(a5.t, training)                                                  # :50
        F = r4.shape[-1]
        cat7 = ops.empty(N, S4, S4, F // 2 + F)                                                # :54
        ops.axpby(View(cat7, F // 2, F // 2 + F, 0), full(r5))
        ops.axpby(View(cat7, F, F // 2 + F, F // 2), full(r4))
        L["c7"] = Conv(w[(LW % 7) + "layer/w"], w[(LW % 7) + "layer/layer/bias"], 2, 0, transposed=True)   # :55
        a7 = L["c7"].forward(full(cat7), N, S4, S4)
        L["bn8"] = BatchNorm(w, 8, self.comm)
        r7 = L["bn8"].forward(a7.t, training)                                                  # :56
        C9 = F // 4 + r2.shape[-1]
        L["c9"] = Conv(w[(LW % 9) + "layer/kernel"], w[(LW % 9) + "layer/bias"], 1, 2, transposed=True)     # :63-64
        if ops.upconv_fused() and F // 4 == 32 and r2.shape[-1] == 128 and S2 <= 48:
            # :60-64 in one fused tcgen05 pass: neither the concat, nor the 96x96x160 upsampled tensor, nor tap columns exist
            a9 = full(ops.empty(N, S, S, F // 8))
            ops.upconv5x5_fwd(r7, r2, L["c9"].w, L["c9"].b, a9.t)
            L["c9"].ctx = None
            self._up_inputs = (r7, r2, a9) if keep_context else None       # backward rebuilds the upsampled tensor from these
        else:
            cat9 = ops.empty(N, S2, S2, C9)                                                    # :60
            ops.axpby(View(cat9, F // 4, C9, 0), full(r7))
            ops.axpby(View(cat9, r2.shape[-1], C9, F // 4), full(r2))
            up = ops.empty(N, S, S, C9)
            ops.upsample2x_fwd(cat9, up)                                                       # :62
            a9 = L["c9"].forward(full(up), N, S, S)
            self._up_inputs = None
        L["bn10"] = BatchNorm(w, 10, self.comm)
        r9 = L["bn10"].forward(a9.t, training)                                                 # :69
        L["c11"] = Conv(w[(LW % 11) + "layer/kernel"], w[(LW % 11) + "layer/bias"], 1, 1, leaky=False)      # :70
        out = L["c11"].forward(full(r9), N, S, S)
        self.L, self.dims = (L if keep_context else None), (B, T, S, N, F, r2.shape[-1])
        return out.t.view(B, T, S, S, -1)

    def backward(self, dout):
        """dout [B,T,S,S,Cout] -> dict of weight gradients (checkpoint names)."""
        L = self.L
        B, T, S, N, F, C2 = self.dims
        S2, S4 = S // 2, S // 4
        g = {}

        def put(i, leafw, leafb, dw, db):
            g[(LW % i) + leafw], g[(LW % i) + leafb] = dw, db

        d = dout.reshape(N, S, S, -1).contiguous().clone()
        dr9, dw, db = L["c11"].backward(d)
        put(11, "layer/kernel", "layer/bias", dw, db)
        da9, gb = L["bn10"].backward(dr9.t, ALPHA)
        g.update(gb)
        if getattr(self, "_up_inputs", None) is not None:      # fused forward: materialise the upsampled tensor now, once
            r7, r2, a9 = self._up_inputs
            C9 = F // 4 + C2
            cat9 = ops.empty(N, S2, S2, C9)
            ops.axpby(View(cat9, F // 4, C9, 0), full(r7))
            ops.axpby(View(cat9, C2, C9, F // 4), full(r2))
            up = ops.empty(N, S, S, C9)
            ops.upsample2x_fwd(cat9, up)
            L["c9"].ctx = (full(up), N, S, S, S, S, a9)
            del cat9
        dup, dw, db = L["c9"].backward(da9, act_done=True)
        put(9, "layer/kernel", "layer/bias", dw, db)
        dcat9 = ops.empty(N, S2, S2, F // 4 + C2)
        ops.upsample2x_bwd(dup.t, dcat9)
        dr7 = ops.empty(N, S2, S2, F // 4)
        ops.axpby(full(dr7), View(dcat9, F // 4, F // 4 + C2, 0))
        dr2 = ops.empty(N, S2, S2, C2)
        ops.axpby(full(dr2), View(dcat9, C2, F // 4 + C2, F // 4))
        da7, gb = L["bn8"].backward(dr7, ALPHA)
        g.update(gb)
        dcat7, dw, db = L["c7"].backward(da7, act_done=True)
        put(7, "layer/w", "layer/layer/bias", dw, db)
        dr5 = ops.empty(N, S4, S4, F // 2)
        ops.axpby(full(dr5), View(dcat7.t, F // 2, F // 2 + F, 0))
        dr4 = ops.empty(N, S4, S4, F)
        ops.axpby(full(dr4), View(dcat7.t, F, F // 2 + F, F // 2))
        da5, gb = L["bn6"].backward(dr5, ALPHA)
        g.update(gb)
        dh, dw, db = L["c5"].backward(da5, act_done=True)
        put(5, "layer/w", "layer/layer/bias", dw, db)
        dx, dK, dR, dbl = L["lstm"].backward(dh.t)
        g[(LW % 4) + "cell/kernel"], g[(LW % 4) + "cell/recurrent_kernel"], g[(LW % 4) + "cell/bias"] = dK, dR, dbl
        ops.axpby(full(dr4), full(dr4), 1.0, full(dx), 1.0)
        da2, gb = L["bn3"].backward(dr4, ALPHA)
        g.update(gb)
        dr2b, dw, db = L["c2"].backward(da2, act_done=True)
        put(2, "layer/w", "layer/layer/bias", dw, db)
        ops.axpby(full(dr2), full(dr2), 1.0, dr2b, 1.0)
        da0, gb = L["bn1"].backward(dr2, ALPHA)
        g.update(gb)
        _, dw, db = L["c0"].backward(da0, need_dx=False, act_done=True)
        put(0, "layer/w", "layer/layer/bias", dw, db)
        return g


# =============================================================================================== critic
class FlatVars(dict):
    """Checkpoint variable name -> view into ONE flat fp32 device buffer laid out by the `wdg_critic` handle (trainable
    variables first, `sn_u` vectors after them; include/wdg.h).  `.flat` is the whole buffer, `.handle` the CriticHandle."""

    def __init__(self, handle, flat, names):
        super().__init__()
        self.handle, self.flat = handle, flat
        for name in names:
            shape, off, _ = handle.table[name]
            self[name] = flat[off:off + int(np.prod(shape))].view(*shape)


class CriticHandle:
    """`wdg_critic*` (csrc/wdg_critic.cu): layer plan + variable table of `make_discriminator` (models.py:76-142)."""

    _cache = {}

    def __init__(self, size, lr_ch, hr_ch, n_timesteps, F, ckpt_topology):
        import ctypes as C
        from .. import _lib
        self.key = (size, lr_ch, hr_ch, F, bool(ckpt_topology))
        h = C.c_void_p()
        rc = _lib.lib().wdg_critic_create(C.byref(h), size, size, lr_ch, hr_ch, n_timesteps, F, int(bool(ckpt_topology)))
        if rc != 0:
            raise ValueError(_lib.lib().wdg_last_error().decode())
        self.h = h
        self.table = {}                       # name -> (shape, offset in floats, trainable), creation order
        for i in range(_lib.lib().wdg_critic_num_weights(h)):
            name, dims, nd, off, tr = C.c_char_p(), (C.c_int64 * 4)(), C.c_int(), C.c_int64(), C.c_int()
            _lib.check(_lib.lib().wdg_critic_weight_info(h, i, C.byref(name), dims, C.byref(nd), C.byref(off), C.byref(tr)))
            self.table[name.value.decode()] = (tuple(int(d) for d in dims[:nd.value]), int(off.value), bool(tr.value))
        self.n_total = int(_lib.lib().wdg_critic_num_floats(h))
        self.n_train = int(_lib.lib().wdg_critic_num_trainable_floats(h))
        self.size, self.lr_ch, self.hr_ch, self.F, self.ckpt_topology = size, lr_ch, hr_ch, F, bool(ckpt_topology)

    @classmethod
    def get(cls, size, lr_ch, hr_ch, F, ckpt_topology=False, n_timesteps=24):
        key = (size, lr_ch, hr_ch, F, bool(ckpt_topology))
        if key not in cls._cache:
            cls._cache[key] = cls(size, lr_ch, hr_ch, n_timesteps, F, ckpt_topology)
        return cls._cache[key]

    @classmethod
    def for_weights(cls, weights, size):
        """The handle whose variable table matches a name -> array dict (either topology)."""
        F = weights[(LW % 2) + "layer/w"].shape[-1]
        hr_ch = weights[(LW % 0) + "cell/kernel"].shape[2]
        lr_ch = weights[(LW % 1) + "cell/kernel"].shape[2] - hr_ch
        for topo in (False, True):
            h = cls.get(size, lr_ch, hr_ch, F, topo)
            if set(h.table) == set(weights) and all(tuple(weights[n].shape) == h.table[n][0] for n in h.table):
                return h
        raise ValueError("critic weights match neither the current-code nor the checkpoint (shortcut) topology at size %d" % size)

    def shapes(self):
        return {n: s for n, (s, _, _) in self.table.items()}

    def trainable(self):
        return [n for n, (_, _, tr) in self.table.items() if tr]

    def pack(self, weights):
        """name -> numpy / tensor dict -> FlatVars on the current device."""
        flat = torch.zeros(self.n_total, dtype=torch.float32, device="cuda")
        fv = FlatVars(self, flat, list(self.table))
        for n, v in fv.items():
            src = weights[n]
            v.copy_(src if isinstance(src, torch.Tensor) else torch.as_tensor(np.asarray(src, np.float32)))
        return fv

    def grads(self, flat):
        return FlatVars(self, flat, self.trainable())


class CriticNet:
    """One differentiable call of the critic on the `wdg_critic` handle: forward keeps a context buffer, backward
    consumes it.  `weights`: FlatVars (used in place) or a name -> CUDA tensor dict (packed; the spectral-norm update of a
    training-mode call is written back into its tensors, as the in-place Keras wrapper would)."""

    def __init__(self, weights, size):
        if isinstance(weights, FlatVars):
            self.vars, self._writeback = weights, None
        else:
            self.vars, self._writeback = CriticHandle.for_weights(weights, size).pack(weights), weights
        self.h = self.vars.handle
        self.w = self.vars

    def _write_back(self):
        if self._writeback is not None:
            for n, v in self.vars.items():
                self._writeback[n].copy_(v)

    def _scratch(self, B, T):
        import ctypes as C
        from .. import _lib
        nb = C.c_size_t()
        _lib.check(_lib.lib().wdg_critic_scratch_bytes(self.h.h, B, T, C.byref(nb)))
        return ops.scratch(nb.value, "critic"), nb.value

    def spectral_norm_step(self):
        """What a training-mode call does to the critic's VARIABLES (one power iteration per wrapped layer), without the
        forward pass."""
        from .. import _lib
        sc, nb = self._scratch(1, 1)
        _lib.check(_lib.lib().wdg_critic_sn_update(self.h.h, ops._p(self.vars.flat), ops._p(sc), nb, ops._s()))
        self._write_back()

    def forward(self, low_res, high_res, training):
        """[B,T,S,S,3], [B,T,S,S,2] -> score [B,1]."""
        import ctypes as C
        from .. import _lib
        B, T = low_res.shape[:2]
        low_res, high_res = low_res.contiguous(), high_res.contiguous()
        nb = C.c_size_t()
        _lib.check(_lib.lib().wdg_critic_context_bytes(self.h.h, B, T, int(bool(training)), C.byref(nb)))
        self.ctx = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
        sc, sb = self._scratch(B, T)
        score = ops.empty(B, 1)
        _lib.check(_lib.lib().wdg_critic_forward(self.h.h, ops._p(self.vars.flat), ops._p(low_res), ops._p(high_res), ops._p(score),
                                                 B, T, int(bool(training)), ops._p(self.ctx), nb.value, ops._p(sc), sb, ops._s()))
        self.dims = (B, T, int(bool(training)), tuple(high_res.shape))
        if training:
            self._write_back()
        return score

    def backward(self, dscore, need_weight_grads=True, need_input_grad=False):
        """dscore [B,1].  Returns (weight grads: FlatVars name -> view of one flat buffer, or {}; d high_res or None)."""
        from .. import _lib
        B, T, training, hr_shape = self.dims
        gflat = ops.zeros(self.h.n_train) if need_weight_grads else None
        dhr = torch.empty(hr_shape, dtype=torch.float32, device="cuda") if need_input_grad else None
        sc, sb = self._scratch(B, T)
        _lib.check(_lib.lib().wdg_critic_backward(self.h.h, ops._p(self.vars.flat), ops._p(self.ctx), B, T, training,
                                                  ops._p(dscore.contiguous()), ops._p(gflat), ops._p(dhr), ops._p(sc), sb, ops._s()))
        self.ctx = None
        return (self.h.grads(gflat) if need_weight_grads else {}), dhr
