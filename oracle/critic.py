"""Float64 numpy restatement of the reference critic graph (`models.py:76-142`, `tf_utils.py:7-32`).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (shapes pinned by
tests/golden/ckpt_manifest.json -- note SURVEY F6: the shipped checkpoint has an extra shortcut conv that the
current code never builds; `ckpt_topology=True` reproduces it).

Weights: dict keyed by the checkpoint's variable names (`layer_with_weights-N/...`), reference layouts.
"""
import math

import numpy as np

from . import layers as L

LW = "layer_with_weights-%d/"


def critic_plan(size=96, lr_ch=3, hr_ch=2, F=16, ckpt_topology=False):
    """Walks the graph-building loops of models.py:111-139 and returns the list of layers with shapes.

    Each pyramid entry: dict(kind, idx (layer_with_weights index of the conv), ln (index of its LayerNorm),
    k, stride, pad, cin, cout, size_in, size_out)."""
    plan = []
    idx = 4  # 0: hr ConvLSTM, 1: mix ConvLSTM, 2: hr conv, 3: mix conv  (creation order of weighted layers)
    ln_hr, ln_mix = idx, idx + 1
    idx += 2
    s, c = size, 2 * F
    pyramid = []

    def conv7(s, c, idx):
        so = (s + 2 - 7) // 3 + 1
        return dict(kind="conv7", idx=idx, ln=idx + 1, k=7, stride=3, pad=1, cin=c, cout=2 * c, size_in=s, size_out=so)

    while s >= 16:                                   # models.py:111-116
        e = conv7(s, c, idx)
        pyramid.append(e)
        s, c, idx = e["size_out"], e["cout"], idx + 2
    shortcut_from = (s, c)
    i = 0
    loop2 = []
    while s >= 4:                                    # models.py:120-126
        e = conv7(s, c, idx)
        if e["size_out"] < 1:
            raise ValueError("invalid image size for the critic (negative conv output)")
        loop2.append(e)
        s, c, idx = e["size_out"], e["cout"], idx + 2
        i += 1
    # the checkpoint orders: conv (idx), [shortcut conv], LN of conv, [LN of shortcut]
    shortcut = None
    if (i > 1) or (ckpt_topology and i >= 1):        # models.py:127 (`i > 1` never triggers; ckpt built with >= 1)
        hs, hc = shortcut_from
        t = s
        if t == 1:
            k, st, pd = hs, 1, 0
        else:
            st = int(math.ceil((2 + hs) / (t - 1)))
            pd = int(math.ceil((st * (t - 1) - hs) / 2) + 1 + 2)
            k = int(st * (1 - t) + hs + 2 * pd)
        shortcut = dict(kind="shortcut", k=k, stride=st, pad=pd, cin=hc, cout=c, size_in=hs, size_out=t)
    tail = []
    while s > 2:                                     # models.py:132-136
        so = (s - 3) // 2 + 1
        tail.append(dict(kind="conv3", k=3, stride=2, pad=0, cin=c, cout=2 * c, size_in=s, size_out=so))
        s, c = so, 2 * c
    return dict(ln_hr=ln_hr, ln_mix=ln_mix, pyramid=pyramid, loop2=loop2, shortcut=shortcut, tail=tail,
                flat=s * s * c, F=F, lr_ch=lr_ch, hr_ch=hr_ch, size=size)


def critic_weight_shapes(size=96, lr_ch=3, hr_ch=2, F=16, ckpt_topology=False):
    """Name -> shape in checkpoint naming.  With ckpt_topology=True and size 96 this equals discriminator.index."""
    P = critic_plan(size, lr_ch, hr_ch, F, ckpt_topology)
    s = {}
    s[(LW % 0) + "cell/kernel"] = (3, 3, hr_ch, 4 * hr_ch)
    s[(LW % 0) + "cell/recurrent_kernel"] = (3, 3, hr_ch, 4 * hr_ch)
    s[(LW % 0) + "cell/bias"] = (4 * hr_ch,)
    s[(LW % 1) + "cell/kernel"] = (3, 3, lr_ch + hr_ch, 4 * F)
    s[(LW % 1) + "cell/recurrent_kernel"] = (3, 3, F, 4 * F)
    s[(LW % 1) + "cell/bias"] = (4 * F,)
    for i, cin in ((2, hr_ch), (3, F)):
        s[(LW % i) + "layer/w"] = (3, 3, cin, F)
        s[(LW % i) + "layer/layer/bias"] = (F,)
        s[(LW % i) + "layer/sn_u"] = (1, F)
    for i in (4, 5):
        s[(LW % i) + "gamma"] = (F,)
        s[(LW % i) + "beta"] = (F,)
    idx = 6
    convs = P["pyramid"] + P["loop2"]
    for n, e in enumerate(convs):
        last = n == len(convs) - 1
        s[(LW % idx) + "layer/w"] = (7, 7, e["cin"], e["cout"])
        s[(LW % idx) + "layer/layer/bias"] = (e["cout"],)
        s[(LW % idx) + "layer/sn_u"] = (1, e["cout"])
        e["idx"] = idx
        idx += 1
        if last and P["shortcut"] is not None:
            sc = P["shortcut"]
            s[(LW % idx) + "layer/w"] = (sc["k"], sc["k"], sc["cin"], sc["cout"])
            s[(LW % idx) + "layer/layer/bias"] = (sc["cout"],)
            s[(LW % idx) + "layer/sn_u"] = (1, sc["cout"])
            sc["idx"] = idx
            idx += 1
        s[(LW % idx) + "gamma"] = (e["cout"],)
        s[(LW % idx) + "beta"] = (e["cout"],)
        e["ln"] = idx
        idx += 1
        if last and P["shortcut"] is not None:
            s[(LW % idx) + "gamma"] = (P["shortcut"]["cout"],)
            s[(LW % idx) + "beta"] = (P["shortcut"]["cout"],)
            P["shortcut"]["ln"] = idx
            idx += 1
    for e in P["tail"]:
        s[(LW % idx) + "layer/w"] = (3, 3, e["cin"], e["cout"])
        s[(LW % idx) + "layer/layer/bias"] = (e["cout"],)
        s[(LW % idx) + "layer/sn_u"] = (1, e["cout"])
        e["idx"] = idx
        s[(LW % (idx + 1)) + "gamma"] = (e["cout"],)
        s[(LW % (idx + 1)) + "beta"] = (e["cout"],)
        e["ln"] = idx + 1
        idx += 2
    s[(LW % idx) + "layer/kernel"] = (P["flat"], 1)
    s[(LW % idx) + "layer/bias"] = (1,)
    P["dense"] = idx
    return s, P


def synthetic_critic_weights(seed=0, **kw):
    rng = np.random.default_rng(seed)
    shapes, _ = critic_weight_shapes(**kw)
    w = {}
    for name, shp in shapes.items():
        leaf = name.rsplit("/", 1)[1]
        if leaf in ("w", "kernel", "recurrent_kernel"):
            fan_in = int(np.prod(shp[:-1]))
            a = rng.standard_normal(shp) * (1.3 / np.sqrt(fan_in))
        elif leaf == "bias":
            a = rng.standard_normal(shp) * 0.1
            if "cell" in name:
                F = shp[0] // 4
                a[F:2 * F] += 1.0
        elif leaf == "sn_u":
            a = rng.standard_normal(shp) * 0.02
        elif leaf == "gamma":
            a = rng.uniform(0.5, 1.5, shp)
        elif leaf == "beta":
            a = rng.standard_normal(shp) * 0.2
        else:
            raise KeyError(name)
        w[name] = a.astype(np.float32)
    return w


def critic_forward(w, low_res, high_res, ckpt_topology=False):
    """`discriminator([low_res, high_res], training=False)`: (B,T,S,S,3), (B,T,S,S,2) -> (B,1).
    Inference mode: SpectralNormalization is the identity on the stored `w`."""
    low_res = np.asarray(low_res, L.F64)
    high_res = np.asarray(high_res, L.F64)
    B, T, S = low_res.shape[:3]
    F = w[(LW % 2) + "layer/w"].shape[-1]
    _, P = critic_weight_shapes(S, low_res.shape[-1], high_res.shape[-1], F, ckpt_topology)

    def ln(x, i):
        return L.layernorm(x, w[(LW % i) + "gamma"], w[(LW % i) + "beta"])

    def snconv(x, i, stride=1, padding="valid"):
        return L.leaky_relu(L.conv2d(x, w[(LW % i) + "layer/w"], w[(LW % i) + "layer/layer/bias"], stride, padding))

    hr = L.conv_lstm2d(high_res, w[(LW % 0) + "cell/kernel"], w[(LW % 0) + "cell/recurrent_kernel"], w[(LW % 0) + "cell/bias"])
    hr = ln(snconv(hr.reshape(B * T, S, S, -1), 2, padding="same"), 4)                        # models.py:93-97
    mix = np.concatenate([low_res, high_res], -1)                                              # :100
    mix = L.conv_lstm2d(mix, w[(LW % 1) + "cell/kernel"], w[(LW % 1) + "cell/recurrent_kernel"], w[(LW % 1) + "cell/bias"])
    mix = ln(snconv(mix.reshape(B * T, S, S, -1), 3, padding="same"), 5)                      # :101-105
    x = np.concatenate([hr, mix], -1)                                                          # :108
    for e in P["pyramid"]:
        x = ln(snconv(L.zero_pad(x, 1), e["idx"], stride=3), e["ln"])                          # :111-116
    shortcut = x
    for e in P["loop2"]:
        x = ln(snconv(L.zero_pad(x, 1), e["idx"], stride=3), e["ln"])                          # :120-126
    if P["shortcut"] is not None:                                                              # :127-130, tf_utils.py:15-32
        sc = P["shortcut"]
        x = x + ln(snconv(L.zero_pad(shortcut, sc["pad"]), sc["idx"], stride=sc["stride"]), sc["ln"])
    for e in P["tail"]:
        x = ln(snconv(x, e["idx"], stride=2), e["ln"])                                         # :132-136
    x = x.reshape(B, T, -1)                                                                    # Flatten (h, w, c)
    x = x @ np.asarray(w[(LW % P["dense"]) + "layer/kernel"], L.F64) + np.asarray(w[(LW % P["dense"]) + "layer/bias"], L.F64)
    return x.mean(1)                                                                           # GlobalAveragePooling1D
