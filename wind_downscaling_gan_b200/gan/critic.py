"""`make_discriminator` product object (reference `gan/models.py:76-142`): weight bookkeeping with the checkpoint's
variable names and a forward pass on the fp32 CUDA training kernels (train/nets.py)."""
import os

import numpy as np

from .models import _glorot_uniform, _orthogonal

LW = "layer_with_weights-%d/"


def critic_weight_shapes(size, lr_ch, hr_ch, F):
    """Variable name -> shape of the graph `make_discriminator` builds (current code: no shortcut branch)."""
    from ..train.nets import critic_plan
    convs, dense_idx, flat = critic_plan(size, F)
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
    for e in convs:
        s[(LW % e["idx"]) + "layer/w"] = (e["k"], e["k"], e["cin"], e["cout"])
        s[(LW % e["idx"]) + "layer/layer/bias"] = (e["cout"],)
        s[(LW % e["idx"]) + "layer/sn_u"] = (1, e["cout"])
        s[(LW % e["ln"]) + "gamma"] = (e["cout"],)
        s[(LW % e["ln"]) + "beta"] = (e["cout"],)
    s[(LW % dense_idx) + "layer/kernel"] = (flat, 1)
    s[(LW % dense_idx) + "layer/bias"] = (1,)
    return s


class Critic:
    """Drop-in for the Keras critic `Model` of models.py:142."""

    name = "discriminator"

    def __init__(self, low_res_size, high_res_size, low_res_channels, high_res_channels, n_timesteps, batch_size=None,
                 feature_channels=16, seed=None):
        if low_res_size != high_res_size:
            raise NotImplementedError("The discriminator assumes that the low res and high res images have the same size."
                                      "Perhaps you should upsample your low res image first?")
        self.image_size, self.low_res_channels, self.high_res_channels = high_res_size, low_res_channels, high_res_channels
        self.n_timesteps, self.feature_channels = n_timesteps, feature_channels
        self.input_names = ["low_resolution_image", "high_resolution_image"]
        self.output_names = ["score"]
        self.optimizer = self.compiled_loss = self.compiled_metrics = None
        self.metrics = []
        self._shapes = critic_weight_shapes(high_res_size, low_res_channels, high_res_channels, feature_channels)
        self._dev = None
        self._before_read = None     # hooks of an owning GAN, see gan/models.py
        self._after_write = None
        rng = np.random.default_rng(seed)
        w = {}
        for name, shp in self._shapes.items():
            leaf = name.rsplit("/", 1)[1]
            if leaf == "recurrent_kernel":
                a = _orthogonal(rng, shp)
            elif leaf in ("w", "kernel"):
                a = _glorot_uniform(rng, shp)
            elif leaf == "bias":
                a = np.zeros(shp, np.float32)
                if "cell" in name:
                    F = shp[0] // 4
                    a[F:2 * F] = 1.0
            elif leaf == "sn_u":
                a = np.clip(rng.standard_normal(shp) * 0.02, -0.04, 0.04).astype(np.float32)
            elif leaf == "gamma":
                a = np.ones(shp, np.float32)
            else:
                a = np.zeros(shp, np.float32)
            w[name] = a
        self._w = w

    def weight_names(self):
        return list(self._shapes)

    @property
    def trainable_weights(self):
        return [n for n in self._shapes if not n.endswith("sn_u")]

    def set_weights(self, weights, _from_state=False):
        if not _from_state:
            if self._before_read is not None:
                self._before_read()
            if self._after_write is not None:
                self._after_write()
        for name, arr in weights.items():
            if name not in self._shapes:
                raise KeyError(name)
            a = np.ascontiguousarray(np.asarray(arr, np.float32))
            if tuple(a.shape) != tuple(self._shapes[name]):
                raise ValueError(f"{name}: expected shape {self._shapes[name]}, got {tuple(a.shape)}")
            self._w[name] = a
        self._dev = None

    def get_weights(self):
        if self._before_read is not None:
            self._before_read()
        return {k: v.copy() for k, v in self._w.items()}

    def save_weights(self, filepath, *args, **kwargs):
        filepath = str(filepath)
        os.makedirs(os.path.dirname(filepath) or ".", exist_ok=True)
        np.savez(filepath + ".npz", **self.get_weights())

    def load_weights(self, filepath, *args, **kwargs):
        filepath = str(filepath)
        if os.path.exists(filepath + ".npz"):
            with np.load(filepath + ".npz") as z:
                self.set_weights({k: z[k] for k in z.files})
            return
        if os.path.exists(filepath + ".index"):
            from ..tf_checkpoint import read_bundle, select_model_variables
            self.set_weights(select_model_variables(read_bundle(filepath), self._shapes, "discriminator"))
            return
        raise FileNotFoundError(filepath)

    def compile(self, optimizer=None, loss=None, metrics=None, **kwargs):
        self.optimizer, self.compiled_loss, self.compiled_metrics = optimizer, loss, metrics
        self.metrics = list(metrics or [])

    def __call__(self, inputs, training=False, mask=None):
        """`discriminator([low_res, high_res], training=False)` -> score (B, 1) CUDA tensor."""
        from ..train.nets import CriticNet, to_device
        from ..train.step import _dev
        from ..train import ops as _ops
        _ops.use_current_stream()
        if self._before_read is not None:
            self._before_read()
        if self._dev is None:
            self._dev = to_device(self._w)
        low_res, high_res = _dev(inputs[0]), _dev(inputs[1])
        if tuple(low_res.shape[:-1]) != tuple(high_res.shape[:-1]):
            raise ValueError("low_resolution_image and high_resolution_image must share (B, T, H, W)")
        score = CriticNet(self._dev, self.image_size).forward(low_res, high_res, training=bool(training))
        if training:   # the spectral-norm power iteration mutated w / sn_u in place, as the Keras wrapper does
            self._w = {k: v.cpu().numpy() for k, v in self._dev.items()}
            if self._after_write is not None:
                self._after_write()
        return score

    call = __call__

    def predict(self, inputs, **kwargs):
        return self(inputs, training=False).cpu().numpy()
