"""TEST HELPER: the critic graph walked op by op from Python over the op-level C ABI (wdg_conv2d_*, wdg_ln_*, wdg_lstm_*,
...), as the product did before the graph moved behind the `wdg_critic` handle (csrc/wdg_critic.cu).  The handle must
reproduce this walk bit for bit (tests/test_critic_handle_gpu.py): same kernels, same order, same arithmetic.  Current-code
topology only (no shortcut branch)."""
import torch

from wind_downscaling_gan_b200.train import ops
from wind_downscaling_gan_b200.train.nets import ALPHA, LW, Conv, ConvLSTM, LayerNorm, sn_step
from wind_downscaling_gan_b200.train.ops import View, full


def critic_plan(size, F):
    """Shapes of the pyramid built by models.py:111-136 (current code: the `i > 1` shortcut never triggers)."""
    convs, idx, s, c = [], 6, size, 2 * F
    while s >= 16:
        so = (s + 2 - 7) // 3 + 1
        convs.append(dict(idx=idx, ln=idx + 1, k=7, stride=3, pad=1, cin=c, cout=2 * c, size_in=s, size_out=so))
        s, c, idx = so, 2 * c, idx + 2
    i = 0
    while s >= 4:
        so = (s + 2 - 7) // 3 + 1
        if so < 1:
            raise ValueError("invalid image size for the critic")
        convs.append(dict(idx=idx, ln=idx + 1, k=7, stride=3, pad=1, cin=c, cout=2 * c, size_in=s, size_out=so))
        s, c, idx = so, 2 * c, idx + 2
        i += 1
    if i > 1:
        raise NotImplementedError("shortcut branch (models.py:127-130) is unreachable for valid sizes")
    while s > 2:
        so = (s - 3) // 2 + 1
        convs.append(dict(idx=idx, ln=idx + 1, k=3, stride=2, pad=0, cin=c, cout=2 * c, size_in=s, size_out=so))
        s, c, idx = so, 2 * c, idx + 2
    return convs, idx, s * s * c


class OpWalkCritic:
    def __init__(self, weights, size):
        self.w = weights
        self.F = weights[(LW % 2) + "layer/w"].shape[-1]
        self.convs, self.dense_idx, self.flat = critic_plan(size, self.F)
        self.sn_layers = [2, 3] + [e["idx"] for e in self.convs]

    def spectral_norm_step(self):
        """What a training-mode call does to the critic's VARIABLES (one power iteration per wrapped layer), without the
        forward pass."""
        for i in self.sn_layers:
            sn_step(self.w, i)

    def forward(self, low_res, high_res, training):
        """[B,T,S,S,3], [B,T,S,S,2] -> score [B,1]."""
        w, F = self.w, self.F
        B, T, S = low_res.shape[:3]
        N = B * T
        if training:
            for i in self.sn_layers:
                sn_step(w, i)
            # every training-mode call reads the variables at their current value (later in-place SN updates of the
            # shared tensors must not leak into this call's backward): snapshot them
            w = {k: (v.clone() if k.endswith(("/w", "kernel", "bias", "gamma", "beta")) else v) for k, v in w.items()}
        cl, ch = low_res.shape[-1], high_res.shape[-1]
        hr_in = high_res.reshape(N, S, S, ch).contiguous()
        mix_in = ops.empty(N, S, S, cl + ch)                                                   # models.py:100
        ops.axpby(View(mix_in, cl, cl + ch, 0), full(low_res.reshape(N, S, S, cl)))
        ops.axpby(View(mix_in, ch, cl + ch, cl), full(hr_in))
        L = {}
        L["lstm_hr"] = ConvLSTM(w[(LW % 0) + "cell/kernel"], w[(LW % 0) + "cell/recurrent_kernel"], w[(LW % 0) + "cell/bias"])
        h1 = L["lstm_hr"].forward(hr_in, B, T)                                                 # :93
        L["c_hr"] = Conv(w[(LW % 2) + "layer/w"], w[(LW % 2) + "layer/layer/bias"], 1, 1)      # :94-96
        a_hr = L["c_hr"].forward(full(h1), N, S, S)
        x = ops.empty(N, S, S, 2 * F)                                                          # :108
        L["ln_hr"] = LayerNorm(w, 4)
        L["ln_hr"].forward(a_hr.t, View(x, F, 2 * F, 0))                                       # :97
        L["lstm_mix"] = ConvLSTM(w[(LW % 1) + "cell/kernel"], w[(LW % 1) + "cell/recurrent_kernel"], w[(LW % 1) + "cell/bias"])
        h2 = L["lstm_mix"].forward(mix_in, B, T)                                               # :101
        L["c_mix"] = Conv(w[(LW % 3) + "layer/w"], w[(LW % 3) + "layer/layer/bias"], 1, 1)     # :102-104
        a_mix = L["c_mix"].forward(full(h2), N, S, S)
        L["ln_mix"] = LayerNorm(w, 5)
        L["ln_mix"].forward(a_mix.t, View(x, F, 2 * F, F))                                     # :105
        cur, size = x, S
        for n, e in enumerate(self.convs):                                                     # :111-136
            c = Conv(w[(LW % e["idx"]) + "layer/w"], w[(LW % e["idx"]) + "layer/layer/bias"], e["stride"], e["pad"])
            a = c.forward(full(cur), N, size, size)
            ln = LayerNorm(w, e["ln"])
            cur = ln.forward(a.t).t
            L["pc%d" % n], L["pln%d" % n] = c, ln
            size = e["size_out"]
        D = self.flat
        score = ops.empty(B, 1)
        dk, db_ = w[(LW % self.dense_idx) + "layer/kernel"], w[(LW % self.dense_idx) + "layer/bias"]
        ops.dense_mean_fwd(cur, dk, db_, score, B, T, D)                                       # :137-140
        self.L, self.dims, self.flat_act, self.w_used = L, (B, T, S, N, cl, ch), cur, w
        return score

    def backward(self, dscore, need_weight_grads=True, need_input_grad=False):
        """dscore [B,1].  Returns (weight grads dict or {}, d high_res [B,T,S,S,ch] or None)."""
        w, F, L = self.w_used, self.F, self.L
        B, T, S, N, cl, ch = self.dims
        g = {}
        nw = need_weight_grads
        D = self.flat
        dk = w[(LW % self.dense_idx) + "layer/kernel"]
        dflat = torch.empty_like(self.flat_act)
        ddk, ddb = ops.empty(*dk.shape), ops.empty(1)
        ops.dense_mean_bwd(dscore, self.flat_act, dk, dflat, ddk, ddb, B, T, D)
        g[(LW % self.dense_idx) + "layer/kernel"], g[(LW % self.dense_idx) + "layer/bias"] = ddk, ddb
        d = dflat
        for n in range(len(self.convs) - 1, -1, -1):
            e = self.convs[n]
            da, gb = L["pln%d" % n].backward(full(d), ALPHA)
            g.update(gb)
            dxv, dw, db = L["pc%d" % n].backward(da, need_dw=nw, act_done=True)
            g[(LW % e["idx"]) + "layer/w"], g[(LW % e["idx"]) + "layer/layer/bias"] = dw, db
            d = dxv.t
        # d: [N,S,S,2F] gradient of the concat(hr, mix)
        d_hr_in = ops.zeros(N, S, S, ch) if need_input_grad else None
        for tag, ln_i, conv_i, lstm_i, off in (("hr", 4, 2, 0, 0), ("mix", 5, 3, 1, F)):
            da, gb = L["ln_" + tag].backward(View(d, F, 2 * F, off), ALPHA)
            g.update(gb)
            dh, dw, db = L["c_" + tag].backward(da, need_dw=nw, act_done=True)
            g[(LW % conv_i) + "layer/w"], g[(LW % conv_i) + "layer/layer/bias"] = dw, db
            dx, dK, dR, dbl = L["lstm_" + tag].backward(dh.t, need_dx=need_input_grad, need_dw=nw)
            g[(LW % lstm_i) + "cell/kernel"], g[(LW % lstm_i) + "cell/recurrent_kernel"], g[(LW % lstm_i) + "cell/bias"] = dK, dR, dbl
            if need_input_grad:
                if tag == "hr":
                    ops.axpby(full(d_hr_in), full(d_hr_in), 1.0, full(dx), 1.0)
                else:
                    ops.axpby(full(d_hr_in), full(d_hr_in), 1.0, View(dx, ch, cl + ch, cl), 1.0)
        return (g if need_weight_grads else {}), (d_hr_in.view(B, T, S, S, ch) if need_input_grad else None)
