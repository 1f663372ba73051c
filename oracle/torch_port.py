"""Second, independently written CPU restatement of the reference generator
(`models.py:9-73`) on torch-CPU library ops (conv2d / conv_transpose2d /
interpolate), multi-threaded.  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Two uses:
  * cross-check of oracle/generator.py (two restatements written against
    different primitives must agree to float64 round-off);
  * the timed CPU baseline of bench.py ("port": TF 2.4.3 is not installable here).

`emulate_bf16=True` rounds every convolution operand (activations and weights)
to bfloat16 and keeps fp32 accumulation -- the arithmetic the B200 kernels use --
so the tolerance written in the GPU parity tests can be budgeted on CPU.
"""
import numpy as np
import torch
import torch.nn.functional as F

LW = "layer_with_weights-%d/"


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


class TorchGenerator:
    def __init__(self, weights, dtype=torch.float32, emulate_bf16=False):
        self.dtype = dtype
        self.bf16 = emulate_bf16
        w = {k: _t(v, dtype) for k, v in weights.items()}
        q = self._q
        # Conv2D kernels HWIO -> OIHW
        self.w0 = q(w[(LW % 0) + "layer/w"].permute(3, 2, 0, 1).contiguous())
        self.b0 = w[(LW % 0) + "layer/layer/bias"]
        self.w2 = q(w[(LW % 2) + "layer/w"].permute(3, 2, 0, 1).contiguous())
        self.b2 = w[(LW % 2) + "layer/layer/bias"]
        self.wk = q(w[(LW % 4) + "cell/kernel"].permute(3, 2, 0, 1).contiguous())
        self.wr = q(w[(LW % 4) + "cell/recurrent_kernel"].permute(3, 2, 0, 1).contiguous())
        self.bl = w[(LW % 4) + "cell/bias"]
        self.w5 = q(w[(LW % 5) + "layer/w"].permute(3, 2, 0, 1).contiguous())
        self.b5 = w[(LW % 5) + "layer/layer/bias"]
        # Conv2DTranspose kernels (kh, kw, out, in) -> torch conv_transpose2d weight (in, out, kh, kw)
        self.w7 = q(w[(LW % 7) + "layer/w"].permute(3, 2, 0, 1).contiguous())
        self.b7 = w[(LW % 7) + "layer/layer/bias"]
        self.w9 = q(w[(LW % 9) + "layer/kernel"].permute(3, 2, 0, 1).contiguous())
        self.b9 = w[(LW % 9) + "layer/bias"]
        self.w11 = q(w[(LW % 11) + "layer/kernel"].permute(3, 2, 0, 1).contiguous())
        self.b11 = w[(LW % 11) + "layer/bias"]
        self.bn = {}
        for i in (1, 3, 6, 8, 10):
            p = LW % i
            scale = w[p + "gamma"] / torch.sqrt(w[p + "moving_variance"] + 1e-3)
            shift = w[p + "beta"] - w[p + "moving_mean"] * scale
            self.bn[i] = (scale.view(1, -1, 1, 1), shift.view(1, -1, 1, 1))

    def _q(self, x):
        if self.bf16:
            return x.to(torch.bfloat16).to(self.dtype)
        return x

    def _bn(self, x, i):
        s, t = self.bn[i]
        return x * s + t

    @torch.no_grad()
    def forward(self, image, noise):
        """image (B,T,S,S,Cin), noise (B,T,S,S,Cn) numpy/torch -> torch (B,T,S,S,Cout)."""
        q = self._q
        image = _t(image, self.dtype)
        noise = _t(noise, self.dtype)
        B, T, S = image.shape[:3]
        x = torch.cat([image, noise], -1).reshape(B * T, S, S, -1).permute(0, 3, 1, 2)
        x = x.contiguous(memory_format=torch.channels_last)
        x = self._bn(F.leaky_relu(F.conv2d(q(x), self.w0, self.b0, stride=2, padding=3), 0.2), 1)
        res_2 = x
        x = self._bn(F.leaky_relu(F.conv2d(q(x), self.w2, self.b2, stride=2, padding=1), 0.2), 3)
        res_4 = x
        s4 = x.shape[-1]
        Fc = self.wr.shape[1]
        xs = x.reshape(B, T, Fc, s4, s4)
        # input convolution for every timestep at once, recurrent one step by step
        zx = F.conv2d(q(x), self.wk, self.bl, padding=1).reshape(B, T, 4 * Fc, s4, s4)
        h = torch.zeros(B, Fc, s4, s4, dtype=self.dtype)
        c = torch.zeros_like(h)
        hs = []
        for t in range(T):
            z = zx[:, t] + F.conv2d(q(h), self.wr, None, padding=1)
            zi, zf, zc, zo = z.split(Fc, 1)
            i = torch.clamp(0.2 * zi + 0.5, 0, 1)
            f = torch.clamp(0.2 * zf + 0.5, 0, 1)
            c = f * c + i * torch.tanh(zc)
            o = torch.clamp(0.2 * zo + 0.5, 0, 1)
            h = o * torch.tanh(c)
            hs.append(h)
        del xs
        x = torch.stack(hs, 1).reshape(B * T, Fc, s4, s4)
        x = self._bn(F.leaky_relu(F.conv2d(q(x), self.w5, self.b5, padding=1), 0.2), 6)
        x = torch.cat([x, res_4], 1)
        x = self._bn(F.leaky_relu(F.conv_transpose2d(q(x), self.w7, self.b7, stride=2), 0.2), 8)
        x = torch.cat([x, res_2], 1)
        x = F.interpolate(q(x), scale_factor=2, mode="bilinear", align_corners=False)
        x = self._bn(F.leaky_relu(F.conv_transpose2d(q(x), self.w9, self.b9, stride=1, padding=2), 0.2), 10)
        x = F.conv2d(q(x), self.w11, self.b11, padding=1)
        return x.permute(0, 2, 3, 1).reshape(B, T, S, S, -1)
