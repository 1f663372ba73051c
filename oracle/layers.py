"""Float64 numpy layer primitives (channels-last), restating the Keras / TFA
semantics the reference's graphs rely on (SURVEY.md §8(c) checklist).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED.

Every function takes and returns numpy float64 arrays laid out (N, H, W, C),
the layout of `models.py:24-25` after TimeDistributed folds (B, T) into N.
"""
import numpy as np

F64 = np.float64


def zero_pad(x, p):
    """`kl.ZeroPadding2D(padding=p)` (models.py:32,38,112,121): symmetric pad."""
    if p == 0:
        return x
    return np.pad(x, ((0, 0), (p, p), (p, p), (0, 0)))


def conv2d(x, w, b=None, stride=1, padding="valid"):
    """`kl.Conv2D`: cross-correlation, kernel HWIO (models.py:33,39,49,70).

    VALID: out = floor((n - k) / s) + 1.  SAME (stride 1 only here): pad k//2.
    Accumulated tap by tap so no im2col buffer is materialised.
    """
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    kh, kw, cin, cout = w.shape
    assert x.shape[-1] == cin, (x.shape, w.shape)
    if padding == "same":
        assert stride == 1 and kh % 2 == 1 and kw % 2 == 1
        x = np.pad(x, ((0, 0), (kh // 2, kh // 2), (kw // 2, kw // 2), (0, 0)))
    n, h, wd, _ = x.shape
    oh = (h - kh) // stride + 1
    ow = (wd - kw) // stride + 1
    out = np.zeros((n, oh, ow, cout), F64)
    for ky in range(kh):
        for kx in range(kw):
            xs = x[:, ky:ky + (oh - 1) * stride + 1:stride, kx:kx + (ow - 1) * stride + 1:stride, :]
            out += xs @ w[ky, kx]
    if b is not None:
        out += np.asarray(b, F64)
    return out


def conv2d_transpose_s2k2(x, w, b=None):
    """`kl.Conv2DTranspose(F, (2,2), strides=2)` VALID (models.py:55).

    Kernel layout (kh, kw, out, in).  Non-overlapping:
    out[2y+ky, 2x+kx, o] = sum_i in[y, x, i] * W[ky, kx, o, i] + b[o].
    """
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    kh, kw, cout, cin = w.shape
    assert (kh, kw) == (2, 2) and x.shape[-1] == cin
    n, h, wd, _ = x.shape
    out = np.zeros((n, 2 * h, 2 * wd, cout), F64)
    for ky in range(2):
        for kx in range(2):
            out[:, ky::2, kx::2, :] = x @ w[ky, kx].T
    if b is not None:
        out += np.asarray(b, F64)
    return out


def conv2d_transpose_same_s1(x, w, b=None):
    """`kl.Conv2DTranspose(F, (5,5), padding='same')`, stride 1 (models.py:63-64).

    Gradient-of-convolution definition: with kernel layout (kh, kw, out, in) and
    p = k // 2,  out[q, o] = sum_{k, i} in[q - k + p, i] * W[k, o, i] + b[o],
    i.e. a SAME cross-correlation with the spatially flipped kernel and the two
    channel axes swapped.
    """
    w = np.asarray(w, F64)
    w_corr = np.transpose(w[::-1, ::-1], (0, 1, 3, 2))  # -> (kh, kw, in, out)
    return conv2d(x, w_corr, b, stride=1, padding="same")


def upsample_bilinear_x2(x):
    """`kl.UpSampling2D((2,2), interpolation='bilinear')` (models.py:62).

    TF2 `tf.image.resize` bilinear: half-pixel centres, no align_corners, edge
    clamp.  Source coordinate of output j is j/2 - 1/4, so
      out[2k]   = .25 in[k-1] + .75 in[k]   (in[-1] := in[0])
      out[2k+1] = .75 in[k]   + .25 in[k+1] (in[n]  := in[n-1])
    applied separably along H then W.
    """
    x = np.asarray(x, F64)

    def up_axis(a, axis):
        a = np.moveaxis(a, axis, 0)
        n = a.shape[0]
        prev = np.concatenate([a[:1], a[:-1]], 0)
        nxt = np.concatenate([a[1:], a[-1:]], 0)
        out = np.empty((2 * n,) + a.shape[1:], F64)
        out[0::2] = 0.25 * prev + 0.75 * a
        out[1::2] = 0.75 * a + 0.25 * nxt
        return np.moveaxis(out, 0, axis)

    return up_axis(up_axis(x, 1), 2)


def leaky_relu(x, alpha=0.2):
    """`LeakyReLU(0.2)` used as `activation=` (models.py:33 etc.); alpha is stored fp32."""
    a = F64(np.float32(alpha))
    return np.where(x >= 0, x, a * x)


def hard_sigmoid(x):
    """TF<=2.x Keras `hard_sigmoid`: clip(0.2 x + 0.5, 0, 1) (ConvLSTM2D recurrent_activation default)."""
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)


def batchnorm_infer(x, gamma, beta, mean, var, eps=1e-3):
    """`kl.BatchNormalization()` in inference mode (models.py:34,40,50,56,69): axis -1, eps 1e-3."""
    g, b, m, v = (np.asarray(a, F64) for a in (gamma, beta, mean, var))
    return g * (x - m) / np.sqrt(v + eps) + b


def batchnorm_train(x, gamma, beta, eps=1e-3):
    """Training-mode BN: biased batch variance over every axis but the last.

    Returns (y, batch_mean, batch_var_biased)."""
    axes = tuple(range(x.ndim - 1))
    m = x.mean(axes)
    v = x.var(axes)
    y = np.asarray(gamma, F64) * (x - m) / np.sqrt(v + eps) + np.asarray(beta, F64)
    return y, m, v


def layernorm(x, gamma, beta, eps=1e-3):
    """`kl.LayerNormalization()` (models.py:97,105,116,125,136): axis -1 only, biased var, eps 1e-3."""
    m = x.mean(-1, keepdims=True)
    v = x.var(-1, keepdims=True)
    return np.asarray(gamma, F64) * (x - m) / np.sqrt(v + eps) + np.asarray(beta, F64)


def conv_lstm2d(x, kernel, recurrent_kernel, bias):
    """`kl.ConvLSTM2D(F, (3,3), padding='same', return_sequences=True)` (models.py:45,93,101).

    x: (B, T, H, W, Cin).  kernel (3,3,Cin,4F), recurrent_kernel (3,3,F,4F), bias (4F,).
    Gate order on the last axis is i, f, c, o.  h0 = c0 = 0.  Per step
      z   = conv_same(x_t, K) + b + conv_same(h_{t-1}, R)
      i,f,o = hard_sigmoid(z_i), hard_sigmoid(z_f), hard_sigmoid(z_o)
      c_t = f * c_{t-1} + i * tanh(z_c);  h_t = o * tanh(c_t)
    Returns all h_t: (B, T, H, W, F).
    """
    x = np.asarray(x, F64)
    B, T, H, W, _ = x.shape
    F = recurrent_kernel.shape[2]
    assert kernel.shape[-1] == 4 * F and recurrent_kernel.shape[-1] == 4 * F
    h = np.zeros((B, H, W, F), F64)
    c = np.zeros((B, H, W, F), F64)
    out = np.empty((B, T, H, W, F), F64)
    for t in range(T):
        z = conv2d(x[:, t], kernel, bias, padding="same") + conv2d(h, recurrent_kernel, None, padding="same")
        zi, zf, zc, zo = z[..., :F], z[..., F:2 * F], z[..., 2 * F:3 * F], z[..., 3 * F:]
        i = hard_sigmoid(zi)
        f = hard_sigmoid(zf)
        c = f * c + i * np.tanh(zc)
        o = hard_sigmoid(zo)
        h = o * np.tanh(c)
        out[:, t] = h
    return out


def l2_normalize(x, eps=1e-12):
    """`tf.math.l2_normalize`: x / sqrt(max(sum x^2, eps))."""
    return x / np.sqrt(max(float(np.sum(x * x)), eps))


def spectral_norm_step(w, u):
    """One training-mode call of TFA 0.14 `SpectralNormalization.normalize_weights`
    (power_iterations=1).  W = reshape(w, (-1, C_last)); v = l2n(u W^T); u' = l2n(v W);
    sigma = v W u'^T; returns (w / sigma, u')."""
    w = np.asarray(w, F64)
    u = np.asarray(u, F64)
    W = w.reshape(-1, w.shape[-1])
    v = l2_normalize(u @ W.T)
    u2 = l2_normalize(v @ W)
    sigma = float((v @ W @ u2.T).squeeze())
    return w / sigma, u2
