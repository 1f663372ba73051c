"""numpy restatement of the reference's evaluation metrics (`gan/metrics.py`).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED: TensorFlow / tensorflow-probability are not
importable here; each function follows the reference lines it cites plus the published semantics of the TF ops it
calls (keras `cosine_similarity` = -sum(l2_normalize(a) * l2_normalize(b)) with l2_normalize(x) = x * rsqrt(max(sum
x^2, 1e-12)); `tf.signal.rfft2d` = FFT over the two innermost axes; tfp `Empirical.cdf(p)` = fraction of samples <= p).
Elementwise math is float32 like TF's, sums are accumulated in float64.
"""
import numpy as np

F32 = np.float32


def _nan0(x):
    return np.where(np.isnan(x), F32(0), x)


def wind_speed_weighted_rmse(real, fake):
    """metrics.py:32-45 -> (B,)"""
    real, fake = np.asarray(real, F32), np.asarray(fake, F32)
    u, v, uh, vh = real[..., 0], real[..., 1], fake[..., 0], fake[..., 1]
    est, rea = np.sqrt(uh * uh + vh * vh), np.sqrt(u * u + v * v)
    beta = (F32(4) + rea) / (F32(4) + est)
    tau = np.where(est >= rea, F32(0.425), F32(1) - F32(0.425))
    res = _nan0(tau * ((uh - beta * u) ** 2 + (vh - beta * v) ** 2))
    return np.sqrt(res.astype(np.float64).mean(axis=(1, 2, 3)))


def wind_speed_rmse(real, fake):
    """metrics.py:81-91 -> (B,)"""
    real, fake = np.asarray(real, F32), np.asarray(fake, F32)
    est = np.sqrt(fake[..., 0] ** 2 + fake[..., 1] ** 2)
    rea = np.sqrt(real[..., 0] ** 2 + real[..., 1] ** 2)
    return np.sqrt(_nan0((rea - est) ** 2).astype(np.float64).mean(axis=(1, 2, 3)))


def extreme_weighted_rmse(real, fake):
    """metrics.py:66-73 -> (B,): weights = real^2 / sum(real^2) over the WHOLE tensor."""
    real, fake = np.asarray(real, np.float64), np.asarray(fake, np.float64)
    sq = real ** 2
    tot = sq.sum()
    w = sq / tot if tot != 0 else np.zeros_like(sq)
    res = w * (real - fake) ** 2
    res = np.where(np.isnan(res), 0.0, res)
    return np.sqrt(res.sum(axis=(1, 2, 3, 4)))


def _cos(real, fake):
    real, fake = np.asarray(real, F32), np.asarray(fake, F32)
    rr, ff = (real * real).sum(-1), (fake * fake).sum(-1)
    inv = F32(1) / np.sqrt(np.maximum(rr, F32(1e-12))) * (F32(1) / np.sqrt(np.maximum(ff, F32(1e-12))))
    return (real * fake).sum(-1) * inv


def angular_cosine_distance(real, fake):
    """metrics.py:97-105 -> (B,)"""
    c = np.clip(_cos(real, fake), -1, 1).astype(np.float64)
    return (np.arccos(c) / np.pi).mean(axis=(1, 2, 3))


def opposite_cosine_similarity(real, fake):
    """metrics.py:108-111 -> (B,): .5 * (1 + keras cosine_similarity) with keras' sign (= -cos)."""
    return (0.5 * (1.0 - _cos(real, fake).astype(np.float64))).mean(axis=(1, 2, 3))


def log_spectral_distance(real, fake):
    """metrics.py:121-137 -> (B,).  rfft2d runs over the two innermost axes of the (B,T,H,W,C) tensor: (W, C)."""
    eps = 1e-7
    pr = np.abs(np.fft.rfft2(np.asarray(real, np.float64), axes=(-2, -1))) ** 2
    pf = np.abs(np.fft.rfft2(np.asarray(fake, np.float64), axes=(-2, -1))) ** 2
    ratio = (pr + eps) / (pf + eps)
    res = (10 * np.log10(ratio)) ** 2
    lsd = np.sqrt(res.mean(axis=(1, 2, 3, 4)))
    return np.where(np.isnan(lsd), 0.0, lsd)


KS_POINTS = np.linspace(-30., 30., 100).astype(F32)      # metrics.py:156, compared in the samples' dtype


def spatially_convolved_ks_stat(real, fake, patch_size=None):
    """metrics.py:155-187 -> (H-P+1, W-P+1) mean KS image over (time, channel, sample); P = W // 10 by default."""
    real, fake = np.asarray(real, F32), np.asarray(fake, F32)
    B, T, H, W, C = real.shape
    P = patch_size or W // 10
    Ho, Wo = H - P + 1, W - P + 1
    acc = np.zeros((Ho, Wo), np.float64)
    for t in range(T):
        for ch in range(C):
            # cdf difference at every point via cumulative box sums of the indicator images
            for b in range(B):
                ir = (real[b, t, :, :, ch][None] <= KS_POINTS[:, None, None]).astype(np.int32)
                jf = (fake[b, t, :, :, ch][None] <= KS_POINTS[:, None, None]).astype(np.int32)
                d = (ir - jf).astype(np.int64)
                cs = np.zeros((100, H + 1, W + 1), np.int64)
                cs[:, 1:, 1:] = d.cumsum(1).cumsum(2)
                win = cs[:, P:, P:] - cs[:, :-P, P:] - cs[:, P:, :-P] + cs[:, :-P, :-P]
                ks = (np.abs(win).max(0).astype(F32) / F32(P * P)).astype(np.float64)
                acc += ks
    return acc / (T * C * B)


def tanh_wind_speed_weighted_rmse(u, v, u_hat, v_hat):
    """metrics.py:48-60 on plain arrays (the reference takes xarray datasets with U_10M/V_10M and u10/v10)."""
    est, rea = np.sqrt(u_hat ** 2 + v_hat ** 2), np.sqrt(u ** 2 + v ** 2)
    beta = (4 + rea) / (4 + est)
    tau = np.where(est >= rea, 0.425, 1 - 0.425)
    w = tau * ((u_hat - beta * u) ** 2 + (v_hat - beta * v) ** 2)
    m = (np.mean(w) + np.quantile(w, 0.5)) / 2
    return np.tanh(w / m)
