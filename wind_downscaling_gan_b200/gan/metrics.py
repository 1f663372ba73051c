"""Evaluation metrics with the reference's names and call signatures (`gan/metrics.py`), computed by the CUDA kernels
of csrc/wdg_metrics.cu.  Inputs are (B, T, H, W, C) float32 arrays / CUDA tensors; per-sample metrics return (B,)
float32 numpy arrays, `spatially_convolved_ks_stat` returns the mean KS image.  The `*_from_xarray` helpers of the
reference are host-side numpy one-liners on xarray datasets; they are mirrored on `GridDataset` / mappings of arrays.
There is no CPU fallback for the device metrics.
"""
import ctypes as C

import numpy as np

from .. import _lib


def _dev(x):
    import torch
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(x, np.float32)))
    t = t.to(device="cuda", dtype=torch.float32).contiguous()
    if t.dim() != 5:
        raise ValueError(f"expected a (B, T, H, W, C) tensor, got shape {tuple(t.shape)}")
    return t


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


_POINTWISE = ("ws_weighted_rmse", "ws_rmse", "acd", "opposite_cosine_similarity", "extreme_rmse")


def pointwise_metrics(real_output, fake_output):
    """All five fused per-sample reductions in one pass: dict name -> (B,) float32."""
    import torch
    r, f = _dev(real_output), _dev(fake_output)
    if r.shape != f.shape:
        raise ValueError("real and fake outputs differ in shape")
    B, T, H, W, Cc = r.shape
    nb = C.c_size_t()
    _lib.check(_lib.lib().wdg_metrics_pointwise_scratch(B, C.byref(nb)))
    scratch = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    out = torch.empty((5, B), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().wdg_metrics_pointwise(_p(r), _p(f), B, T * H * W, Cc, _p(out), _p(scratch), _stream()))
    o = out.cpu().numpy()
    return {n: o[i] for i, n in enumerate(_POINTWISE)}


def wind_speed_weighted_rmse(real_output, fake_output):       # metrics.py:32
    return pointwise_metrics(real_output, fake_output)["ws_weighted_rmse"]


def wind_speed_rmse(real_output, fake_output):                # metrics.py:81
    return pointwise_metrics(real_output, fake_output)["ws_rmse"]


def extreme_weighted_rmse(real_output, fake_output):          # metrics.py:66
    return pointwise_metrics(real_output, fake_output)["extreme_rmse"]


def angular_cosine_distance(real_output, fake_output):        # metrics.py:97
    return pointwise_metrics(real_output, fake_output)["acd"]


def opposite_cosine_similarity(real_output, fake_output):     # metrics.py:108
    return pointwise_metrics(real_output, fake_output)["opposite_cosine_similarity"]


def log_spectral_distance(real_output, fake_output):          # metrics.py:121
    import torch
    r, f = _dev(real_output), _dev(fake_output)
    B, T, H, W, Cc = r.shape
    scratch = torch.empty(B * T * H, dtype=torch.float64, device="cuda")
    out = torch.empty(B, dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().wdg_metric_lsd(_p(r), _p(f), B, T, H, W, Cc, _p(out), _p(scratch), _stream()))
    return out.cpu().numpy()


def spatially_convolved_ks_stat(real_output, fake_output, patch_size=None):   # metrics.py:165
    import torch
    r, f = _dev(real_output), _dev(fake_output)
    B, T, H, W, Cc = r.shape
    P = patch_size or W // 10
    out = torch.empty((H - P + 1, W - P + 1), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().wdg_metric_spatial_ks(_p(r), _p(f), B, T, H, W, Cc, P, _p(out), _stream()))
    return out.cpu().numpy()


class MeanMetricWrapper:
    """tfa.metrics.MeanMetricWrapper stand-in: running mean of fn(y_true, y_pred) over every sample seen."""

    def __init__(self, fn, name=None):
        self.fn, self.name = fn, name or fn.__name__
        self.reset_state()

    def reset_state(self):
        self.total, self.count = 0.0, 0

    reset_states = reset_state

    def update_state(self, y_true, y_pred, sample_weight=None):
        v = np.asarray(self.fn(y_true, y_pred), np.float64).ravel()
        self.total += float(v.sum())
        self.count += v.size

    def result(self):
        return self.total / self.count if self.count else 0.0


class _ScoreMean(MeanMetricWrapper):
    def __init__(self, pick, name):
        super().__init__(lambda real, fake: np.asarray(pick(real, fake), np.float64), name=name)


def discriminator_score_real(name="d_real"):                  # metrics.py:8-15
    return _ScoreMean(lambda real, fake: real, name)


def discriminator_score_fake(name="d_fake"):                  # metrics.py:18-25
    return _ScoreMean(lambda real, fake: fake, name)


WindSpeedWeightedRMSE = lambda: MeanMetricWrapper(wind_speed_weighted_rmse, name="ws_weighted_rmse")   # noqa: E731
WeightedRMSEForExtremes = lambda: MeanMetricWrapper(extreme_weighted_rmse, name="extreme_rmse")        # noqa: E731
WindSpeedRMSE = lambda: MeanMetricWrapper(wind_speed_rmse, name="ws_rmse")                              # noqa: E731
AngularCosineDistance = lambda: MeanMetricWrapper(angular_cosine_distance, name="acd")                  # noqa: E731
LogSpectralDistance = lambda: MeanMetricWrapper(log_spectral_distance, name="lsd")                      # noqa: E731
SpatialKS = lambda: MeanMetricWrapper(spatially_convolved_ks_stat, name="spatial_ks")                   # noqa: E731


# ---- host-side helpers on gridded datasets (the reference's *_from_xarray functions; plain numpy, no device work)
def _var(ds, *names):
    for n in names:
        try:
            return np.asarray(ds[n])
        except (KeyError, IndexError, TypeError):
            continue
    raise KeyError(names)


def tanh_wind_speed_weighted_rmse_from_xarray(real_output, fake_output):      # metrics.py:48-60
    u, v = _var(real_output, "U_10M"), _var(real_output, "V_10M")
    uh, vh = _var(fake_output, "u10"), _var(fake_output, "v10")
    est, rea = np.sqrt(uh ** 2 + vh ** 2), np.sqrt(u ** 2 + v ** 2)
    beta = (4 + rea) / (4 + est)
    tau = np.where(est >= rea, 0.425, 1 - 0.425)
    w = tau * ((uh - beta * u) ** 2 + (vh - beta * v) ** 2)
    m = (np.mean(w) + np.quantile(w, 0.5)) / 2
    return np.tanh(w / m)


def cosine_similarity_from_xarray(real_output, fake_output):                  # metrics.py:114-119
    u, v = _var(real_output, "U_10M"), _var(real_output, "V_10M")
    uh, vh = _var(fake_output, "u10"), _var(fake_output, "v10")
    return (u * uh + v * vh) / (np.sqrt(u ** 2 + v ** 2) * np.sqrt(uh ** 2 + vh ** 2))


def log_spectral_distance_from_xarray(real_output, fake_output):              # metrics.py:143-152
    fake = np.stack([_var(fake_output, "u10"), _var(fake_output, "v10")])
    real = np.stack([_var(real_output, "U_10M"), _var(real_output, "V_10M")])
    eps = 1e-7
    pr, pf = np.abs(np.fft.fft2(real)) ** 2, np.abs(np.fft.fft2(fake)) ** 2
    return np.mean((10 * np.log10((pr + eps) / (pf + eps))) ** 2, axis=0)


def rmse_from_xarray(real_output, fake_output):                               # metrics.py:192-197
    u, v = real_output[..., 0], real_output[..., 1]
    uh, vh = fake_output[..., 0], fake_output[..., 1]
    return np.sqrt(np.mean((u - uh) ** 2 + (v - vh) ** 2, axis=(1, 2, 3)))
