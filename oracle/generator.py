"""Float64 restatement of the reference generator graph (`models.py:9-73`).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED.

Weights are a dict keyed by the reference checkpoint's own variable names
(`weights-55.ckpt/generator.index`, e.g. ``layer_with_weights-0/layer/w``) with
the reference's array layouts (Conv2D HWIO, Conv2DTranspose (kh, kw, out, in)).
"""
import numpy as np

from . import layers as L

LW = "layer_with_weights-%d/"


def generator_weight_shapes(in_channels=3, noise_channels=20, out_channels=2, feature_channels=128):
    """Name -> shape table of `make_generator` (models.py:28-71); equals the
    reference ckpt's `generator.index` at api.py:22-28 constants."""
    F = feature_channels
    cin = in_channels + noise_channels
    f0 = cin * 8 if cin * 8 <= F else F  # models.py:31
    s = {}

    def bn(i, c):
        for k in ("gamma", "beta", "moving_mean", "moving_variance"):
            s[(LW % i) + k] = (c,)

    s[(LW % 0) + "layer/w"] = (8, 8, cin, f0)
    s[(LW % 0) + "layer/layer/bias"] = (f0,)
    s[(LW % 0) + "layer/sn_u"] = (1, f0)
    bn(1, f0)
    s[(LW % 2) + "layer/w"] = (4, 4, f0, F)
    s[(LW % 2) + "layer/layer/bias"] = (F,)
    s[(LW % 2) + "layer/sn_u"] = (1, F)
    bn(3, F)
    s[(LW % 4) + "cell/kernel"] = (3, 3, F, 4 * F)
    s[(LW % 4) + "cell/recurrent_kernel"] = (3, 3, F, 4 * F)
    s[(LW % 4) + "cell/bias"] = (4 * F,)
    s[(LW % 5) + "layer/w"] = (3, 3, F, F // 2)
    s[(LW % 5) + "layer/layer/bias"] = (F // 2,)
    s[(LW % 5) + "layer/sn_u"] = (1, F // 2)
    bn(6, F // 2)
    s[(LW % 7) + "layer/w"] = (2, 2, F // 4, F // 2 + F)
    s[(LW % 7) + "layer/layer/bias"] = (F // 4,)
    s[(LW % 7) + "layer/sn_u"] = (1, F // 2 + F)
    bn(8, F // 4)
    s[(LW % 9) + "layer/kernel"] = (5, 5, F // 8, F // 4 + f0)
    s[(LW % 9) + "layer/bias"] = (F // 8,)
    bn(10, F // 8)
    s[(LW % 11) + "layer/kernel"] = (3, 3, F // 8, out_channels)
    s[(LW % 11) + "layer/bias"] = (out_channels,)
    return s


def synthetic_generator_weights(seed=0, scale=1.0, **kw):
    """Deterministic synthetic weights with non-trivial BN statistics:
    conv kernels ~ N(0, 1/fan_in)*scale-ish, BN mu~N(0,.5), var~U(.5,2), gamma~U(.5,1.5), beta~N(0,.2)."""
    rng = np.random.default_rng(seed)
    w = {}
    for name, shp in generator_weight_shapes(**kw).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("w", "kernel", "recurrent_kernel"):
            if name.startswith(LW % 7) or name.startswith(LW % 9):
                fan_in = shp[0] * shp[1] * shp[3]  # ConvT layout (kh, kw, out, in)
                if name.startswith(LW % 7):
                    fan_in = shp[3]  # 2x2 s2: one tap per output pixel
            else:
                fan_in = shp[0] * shp[1] * shp[2]
            a = rng.standard_normal(shp) * (scale * 1.4 / np.sqrt(fan_in))
        elif leaf == "bias":
            a = rng.standard_normal(shp) * 0.1
            if "cell" in name:  # unit_forget_bias
                F = shp[0] // 4
                a[F:2 * F] += 1.0
        elif leaf == "sn_u":
            a = rng.standard_normal(shp) * 0.02
        elif leaf == "gamma":
            a = rng.uniform(0.5, 1.5, shp)
        elif leaf == "beta":
            a = rng.standard_normal(shp) * 0.2
        elif leaf == "moving_mean":
            a = rng.standard_normal(shp) * 0.5
        elif leaf == "moving_variance":
            a = rng.uniform(0.5, 2.0, shp)
        else:
            raise KeyError(name)
        w[name] = a.astype(np.float32)
    return w


def _bn(x, w, i):
    p = LW % i
    return L.batchnorm_infer(x, w[p + "gamma"], w[p + "beta"], w[p + "moving_mean"], w[p + "moving_variance"])


def generator_forward(w, image, noise, return_intermediates=False):
    """Inference-mode forward (`gen.predict([image, noise])`, api.py:137).

    image (B, T, S, S, Cin), noise (B, T, S, S, Cn) -> (B, T, S, S, Cout), float64.
    Inference mode: BN uses moving statistics, SpectralNormalization is the identity
    on the stored `w`.
    """
    image = np.asarray(image, L.F64)
    noise = np.asarray(noise, L.F64)
    B, T, S, _, _ = image.shape
    inter = {}
    x = np.concatenate([image, noise], -1).reshape(B * T, S, S, -1)          # models.py:28
    x = L.zero_pad(x, 3)                                                      # :32
    x = L.leaky_relu(L.conv2d(x, w[(LW % 0) + "layer/w"], w[(LW % 0) + "layer/layer/bias"], stride=2))  # :33
    x = _bn(x, w, 1)                                                          # :34
    res_2 = x
    inter["res_2"] = x
    x = L.zero_pad(x, 1)                                                      # :38
    x = L.leaky_relu(L.conv2d(x, w[(LW % 2) + "layer/w"], w[(LW % 2) + "layer/layer/bias"], stride=2))  # :39
    x = _bn(x, w, 3)                                                          # :40
    res_4 = x
    inter["res_4"] = x
    s4 = x.shape[1]
    x = L.conv_lstm2d(x.reshape(B, T, s4, s4, -1), w[(LW % 4) + "cell/kernel"],
                      w[(LW % 4) + "cell/recurrent_kernel"], w[(LW % 4) + "cell/bias"])  # :45
    x = x.reshape(B * T, s4, s4, -1)
    inter["lstm"] = x
    x = L.leaky_relu(L.conv2d(x, w[(LW % 5) + "layer/w"], w[(LW % 5) + "layer/layer/bias"], padding="same"))  # :49
    x = _bn(x, w, 6)                                                          # :50
    inter["g5"] = x
    x = np.concatenate([x, res_4], -1)                                        # :54
    x = L.leaky_relu(L.conv2d_transpose_s2k2(x, w[(LW % 7) + "layer/w"], w[(LW % 7) + "layer/layer/bias"]))  # :55
    x = _bn(x, w, 8)                                                          # :56
    inter["g7"] = x
    x = np.concatenate([x, res_2], -1)                                        # :60
    x = L.upsample_bilinear_x2(x)                                             # :62
    x = L.leaky_relu(L.conv2d_transpose_same_s1(x, w[(LW % 9) + "layer/kernel"], w[(LW % 9) + "layer/bias"]))  # :63
    x = _bn(x, w, 10)                                                         # :69
    inter["g9"] = x
    x = L.conv2d(x, w[(LW % 11) + "layer/kernel"], w[(LW % 11) + "layer/bias"], padding="same")  # :70
    out = x.reshape(B, T, S, S, -1)
    if return_intermediates:
        return out, inter
    return out
